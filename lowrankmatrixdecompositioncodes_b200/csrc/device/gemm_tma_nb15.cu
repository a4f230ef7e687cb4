// explicit instantiation of the streaming GEMM kernel for 15 column groups per tile (120 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<15>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<15>(bool);
} }
