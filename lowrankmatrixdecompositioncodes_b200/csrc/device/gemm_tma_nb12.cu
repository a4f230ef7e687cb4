// explicit instantiation of the streaming GEMM kernel for 12 column groups per tile (96 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<12>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<12>(bool);
} }
