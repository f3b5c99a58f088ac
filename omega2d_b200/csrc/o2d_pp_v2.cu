// o2d_pp_v2.cu -- pair-sum kernels for the CORE_V2 core (one TU per core: parallel builds).
#include "o2d_pp_impl.cuh"
namespace o2d {
template cudaError_t launch_pp_core<CORE_V2>(int, const PPArgs&, int, cudaStream_t, bool);
template int pp_occupancy_core<CORE_V2>(int, bool);
}  // namespace o2d
