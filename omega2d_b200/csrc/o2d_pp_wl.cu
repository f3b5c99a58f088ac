// o2d_pp_wl.cu -- pair-sum kernels for the CORE_WL core (one TU per core: parallel builds).
#include "o2d_pp_impl.cuh"
namespace o2d {
template cudaError_t launch_pp_core<CORE_WL>(int, const PPArgs&, int, cudaStream_t, bool);
template int pp_occupancy_core<CORE_WL>(int, bool);
}  // namespace o2d
