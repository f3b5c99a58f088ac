// o2d_status.cu -- status-file reductions over a particle collection.
//
// Simulation::write_status (reference src/Simulation.cpp:765-790) sums, per particle collection,
//   ElementBase::get_total_circ   (src/ElementBase.h:374-394)  sum s_i        (double accumulator, float result)
//   Points::get_total_impulse     (src/Points.h:270-285)       (-sum s_i y_i, +sum s_i x_i)   (float accumulators)
// and Points::update_max_str uses ElementBase::get_max_str (src/ElementBase.h:362-372) max |s_i|.
// With the particle state resident on the GPU (o2d_advect) these would each force a download of every array;
// here one pass reduces all four in float64, in a fixed order (per-CTA partials, then one CTA), so the result
// is bit-reproducible.  The impulse is therefore MORE accurate than the reference's serial float sum, whose
// own rounding error grows like 1e-7 sqrt(N).
#include "o2d_context.h"

#include <algorithm>

using namespace o2d;

namespace {

constexpr int kStatThreads = 256;
constexpr int kStatBlocks = 592;   // 4 per SM on a 148-SM part; any fixed number keeps the order fixed

__device__ __forceinline__ void block_reduce(double (&v)[4], double* smem) {
  // v[0..2] summed, v[3] max-reduced
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    v[3] = fmax(v[3], __shfl_down_sync(0xffffffffu, v[3], off));
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) smem[warp * 4 + k] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = lane < nw ? smem[lane * 4 + k] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
      v[3] = fmax(v[3], __shfl_down_sync(0xffffffffu, v[3], off));
    }
  }
}

__global__ void __launch_bounds__(kStatThreads)
status_partial_kernel(int n, const float* __restrict__ x, const float* __restrict__ y,
                      const float* __restrict__ s, double* __restrict__ partial) {
  __shared__ double smem[4 * (kStatThreads / 32)];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = blockIdx.x * kStatThreads + threadIdx.x; i < n; i += gridDim.x * kStatThreads) {
    const double si = (double)s[i];
    v[0] += si;
    v[1] -= si * (double)y[i];
    v[2] += si * (double)x[i];
    v[3] = fmax(v[3], fabs(si));
  }
  block_reduce(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = v[k];
  }
}

__global__ void __launch_bounds__(kStatThreads)
status_final_kernel(int nblocks, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double smem[4 * (kStatThreads / 32)];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += kStatThreads) {
    v[0] += partial[b * 4 + 0];
    v[1] += partial[b * 4 + 1];
    v[2] += partial[b * 4 + 2];
    v[3] = fmax(v[3], partial[b * 4 + 3]);
  }
  block_reduce(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = v[k];
  }
}

}  // namespace

extern "C" int o2d_points_status(o2d_context* c, int mem, int n, const float* x, const float* y, const float* s,
                                 float* total_circ, float* impulse_x, float* impulse_y, float* max_str) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (n < 0) return fail(O2D_ERR_INVALID, "negative count");
  double h[4] = {0.0, 0.0, 0.0, 0.0};
  if (n > 0) {
    if (!x || !y || !s) return fail(O2D_ERR_INVALID, "NULL array");
    const float *dx = x, *dy = y, *ds = s;
    if (mem == O2D_MEM_HOST) {
      CU(c->hsrc.reserve(3 * padded_floats(n)));
      Stager S(c, &c->hsrc);
      float *tx = S.take(n), *ty = S.take(n), *ts = S.take(n);
      int rc;
      if ((rc = o2d_h2d(c, tx, x, n)) || (rc = o2d_h2d(c, ty, y, n)) || (rc = o2d_h2d(c, ts, s, n))) return rc;
      dx = tx; dy = ty; ds = ts;
    }
    CU(c->counter.reserve((kStatBlocks + 1) * 4 * sizeof(double) + 64));
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(c->counter.p) + 64);
    double* out = partial + kStatBlocks * 4;
    const int nblocks = std::min(kStatBlocks, (n + kStatThreads - 1) / kStatThreads);
    status_partial_kernel<<<nblocks, kStatThreads, 0, c->stream>>>(n, dx, dy, ds, partial);
    CU(cudaGetLastError());
    status_final_kernel<<<1, kStatThreads, 0, c->stream>>>(nblocks, partial, out);
    CU(cudaGetLastError());
    c->launches += 2;
    CU(cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (int rc_ = o2d_finish_host(c)) return rc_;
  }
  if (total_circ) *total_circ = (float)h[0];
  if (impulse_x) *impulse_x = (float)h[1];
  if (impulse_y) *impulse_y = (float)h[2];
  // get_max_str returns 1.0 for collections without strengths; an empty one has none (src/ElementBase.h:362-372)
  if (max_str) *max_str = n > 0 ? (float)h[3] : 1.0f;
  return O2D_OK;
}
