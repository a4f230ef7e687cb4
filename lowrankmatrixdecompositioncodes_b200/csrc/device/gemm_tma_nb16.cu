// explicit instantiation of the streaming GEMM kernel for 16 column groups per tile (128 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<16>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<16>(bool);
} }
