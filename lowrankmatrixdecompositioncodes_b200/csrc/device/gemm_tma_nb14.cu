// explicit instantiation of the streaming GEMM kernel for 14 column groups per tile (112 columns)
#include "gemm_tma_kernel.cuh"
namespace rsvd { namespace tma {
template bool launch_tma<14>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
template int max_sketch_clusters<14>(bool);
} }
