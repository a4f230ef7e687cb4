// explicit instantiation of the streaming GEMM kernel for 8 column groups per tile (64 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<8>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<8>(bool);
} }
