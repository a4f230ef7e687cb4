/* oracle/shim/mkl_lapacke.h — TEST INFRASTRUCTURE ONLY. See mkl.h. */
#include "mkl.h"
