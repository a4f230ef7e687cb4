/* oracle/shim/mkl.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Minimal stand-in for Intel MKL's <mkl.h> so the UNMODIFIED reference sources under
 * /root/reference/multi_core_mkl_code{,_64bit}/ compile with gcc and link against the image's
 * LP64 OpenBLAS 0.3.15 (CBLAS + LAPACKE, unprefixed symbols).  Declares exactly the routines the
 * reference calls. */
#ifndef ORACLE_SHIM_MKL_H
#define ORACLE_SHIM_MKL_H
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <stdint.h>

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

void cblas_dgemm(int layout, int transA, int transB, int M, int N, int K, double alpha,
                 const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
void cblas_dgemv(int layout, int trans, int M, int N, double alpha, const double *A, int lda,
                 const double *x, int incx, double beta, double *y, int incy);
void cblas_dtrsm(int layout, int side, int uplo, int trans, int diag, int M, int N, double alpha,
                 const double *A, int lda, double *B, int ldb);
void cblas_dtrsv(int layout, int uplo, int trans, int diag, int N, const double *A, int lda,
                 double *x, int incx);

int LAPACKE_dgeqrf(int layout, int m, int n, double *a, int lda, double *tau);
int LAPACKE_dorgqr(int layout, int m, int n, int k, double *a, int lda, const double *tau);
int LAPACKE_dgeqp3(int layout, int m, int n, double *a, int lda, int *jpvt, double *tau);
int LAPACKE_dgesvd(int layout, char jobu, char jobvt, int m, int n, double *a, int lda, double *s,
                   double *u, int ldu, double *vt, int ldvt, double *superb);
int LAPACKE_dsyev(int layout, char jobz, char uplo, int n, double *a, int lda, double *w);
int LAPACKE_dgesv(int layout, int n, int nrhs, double *a, int lda, int *ipiv, double *b, int ldb);
int LAPACKE_dtrtri(int layout, char uplo, char diag, int n, double *a, int lda);
#endif
