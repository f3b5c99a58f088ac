// o2d_pp_rm.cu -- pair-sum kernels for the CORE_RM core (one TU per core: parallel builds).
#include "o2d_pp_impl.cuh"
namespace o2d {
template cudaError_t launch_pp_core<CORE_RM>(int, const PPArgs&, int, cudaStream_t, bool);
template int pp_occupancy_core<CORE_RM>(int, bool);
}  // namespace o2d
