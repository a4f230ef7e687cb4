// explicit instantiation of the streaming GEMM kernel for 4 column groups per tile (32 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<4>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<4>(bool);
} }
