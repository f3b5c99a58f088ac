// tools/pipe_mix.cu -- development micro-benchmark: how well do the FMA pipe (FFMA2) and the XU pipe
// (MUFU.RCP) co-issue on one SM sub-partition?  The RM velocity pair costs 8 FP32 lane-ops + 1 MUFU
// (4 FFMA2 : 1 MUFU per lane pair), which saturates both pipes at 16 pairs/clk/SM.  This measures the
// fraction of that cap a dependency-free instruction stream with the same mix reaches.
// Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_mix.cu -o build/pipe_mix
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float mufu_rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// NF FFMA2 and NM MUFU per "group"; G independent groups per thread; dependency only across iterations
template <int NF, int NM, int G>
__global__ void __launch_bounds__(256) mix_kernel(int iters, float* sink, float seed) {
  float2 a[G][NF > 0 ? NF : 1];
  float m[G][NM > 0 ? NM : 1];
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int k = 0; k < NF; ++k) a[g][k] = make_float2(seed + g + k + threadIdx.x, seed - g - k);
#pragma unroll
    for (int k = 0; k < NM; ++k) m[g][k] = seed + 1.5f + g + k + threadIdx.x;
  }
  const float2 c1 = make_float2(1.0000001f, 0.9999999f), c2 = make_float2(1e-9f, -1e-9f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int k = 0; k < NF; ++k) a[g][k] = __ffma2_rn(a[g][k], c1, c2);
#pragma unroll
      for (int k = 0; k < NM; ++k) m[g][k] = mufu_rcp(m[g][k]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int k = 0; k < NF; ++k) s += a[g][k].x + a[g][k].y;
#pragma unroll
    for (int k = 0; k < NM; ++k) s += m[g][k];
  }
  if (s == 123.456f) sink[0] = s;
}

template <int NF, int NM, int G>
void run(const char* name, int sms, int threads, int ctas_per_sm, float* sink) {
  const int iters = 20000;
  auto k = mix_kernel<NF, NM, G>;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<<<sms * ctas_per_sm, threads>>>(100, sink, 1.0f);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    k<<<sms * ctas_per_sm, threads>>>(iters, sink, 1.0f);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double clk = 1.965e9;
  const double thr = (double)sms * ctas_per_sm * threads;
  const double ffma2 = thr * iters * G * NF, mufu = thr * iters * G * NM;
  const double t = best * 1e-3;
  // per SM per clock: lane-ops (2 per FFMA2 thread-instr) and MUFU thread-ops
  const double lane_per_clk = 2 * ffma2 / t / clk / sms, mufu_per_clk = mufu / t / clk / sms;
  printf("%-26s thr=%d cta/sm=%d  %8.3f ms  FMA lane-ops/clk/SM %6.1f (%.1f%% of 128)  MUFU/clk/SM %5.2f (%.1f%% of 16)\n",
         name, threads, ctas_per_sm, best, lane_per_clk, lane_per_clk / 1.28, mufu_per_clk, mufu_per_clk / 0.16);
  fflush(stdout);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float* sink; CK(cudaMalloc(&sink, 4));
  printf("device %s, %d SMs\n", prop.name, sms);
  for (int cps : {1, 2, 3, 4}) {
    run<8, 0, 4>("FFMA2 only", sms, 256, cps, sink);
    run<0, 2, 4>("MUFU only", sms, 256, cps, sink);
    run<8, 2, 4>("4 FFMA2 : 1 MUFU", sms, 256, cps, sink);
    run<7, 2, 4>("3.5 FFMA2 : 1 MUFU", sms, 256, cps, sink);
    run<9, 2, 4>("4.5 FFMA2 : 1 MUFU", sms, 256, cps, sink);
    run<10, 2, 4>("5 FFMA2 : 1 MUFU", sms, 256, cps, sink);
    run<6, 2, 4>("3 FFMA2 : 1 MUFU", sms, 256, cps, sink);
  }
  return 0;
}
