// explicit instantiation of the streaming GEMM kernel for 13 column groups per tile (104 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<13>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<13>(bool);
} }
