// tools/pipe_ops.cu -- development micro-benchmark: do FADD2 / FMUL2 issue at the rate of FFMA2 on sm_100a, and what
// does the instruction mix of the velocity+gradient pair (7 FFMA2 + 4 FADD2 + 4 FMUL2 + 2 MUFU per two pairs) reach
// as a dependency-free stream?  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_ops.cu -o build/pipe_ops
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// per group: NFMA FFMA2, NADD FADD2, NMUL FMUL2, NM MUFU; G independent groups per thread
template <int NFMA, int NADD, int NMUL, int NM, int G>
__global__ void __launch_bounds__(256) ops_kernel(int iters, float* sink, float seed) {
  constexpr int N = NFMA + NADD + NMUL;
  float2 a[G][N > 0 ? N : 1];
  float m[G][NM > 0 ? NM : 1];
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[g][k] = make_float2(seed + g + k + threadIdx.x, seed - g - k);
#pragma unroll
    for (int k = 0; k < NM; ++k) m[g][k] = seed + 1.5f + g + k + threadIdx.x;
  }
  const float2 c1 = make_float2(1.0000001f, 0.9999999f), c2 = make_float2(1e-9f, -1e-9f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int k = 0; k < NFMA; ++k) a[g][k] = __ffma2_rn(a[g][k], c1, c2);
#pragma unroll
      for (int k = NFMA; k < NFMA + NADD; ++k) a[g][k] = __fadd2_rn(a[g][k], c2);
#pragma unroll
      for (int k = NFMA + NADD; k < N; ++k) a[g][k] = __fmul2_rn(a[g][k], c1);
#pragma unroll
      for (int k = 0; k < NM; ++k) m[g][k] = mufu_rcp(m[g][k]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int k = 0; k < N; ++k) s += a[g][k].x + a[g][k].y;
#pragma unroll
    for (int k = 0; k < NM; ++k) s += m[g][k];
  }
  if (s == 123.456f) sink[0] = s;
}

template <int NFMA, int NADD, int NMUL, int NM, int G>
void run(const char* name, int sms, int ctas_per_sm, float* sink) {
  const int iters = 20000, threads = 256;
  auto k = ops_kernel<NFMA, NADD, NMUL, NM, G>;
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
  const double clk = 1.965e9, thr = (double)sms * ctas_per_sm * threads, t = best * 1e-3;
  const double packed = thr * iters * G * (NFMA + NADD + NMUL), mufu = thr * iters * G * NM;
  const double lane = 2 * packed / t / clk / sms;
  printf("%-44s cta/sm=%d %8.3f ms  packed lane-ops/clk/SM %6.1f (%.1f%% of 128)  MUFU/clk/SM %5.2f\n",
         name, ctas_per_sm, best, lane, lane / 1.28, mufu / t / clk / sms);
  fflush(stdout);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float* sink; CK(cudaMalloc(&sink, 4));
  printf("device %s, %d SMs\n", prop.name, sms);
  for (int cps : {2, 3}) {
    run<8, 0, 0, 0, 4>("FFMA2 only", sms, cps, sink);
    run<0, 8, 0, 0, 4>("FADD2 only", sms, cps, sink);
    run<0, 0, 8, 0, 4>("FMUL2 only", sms, cps, sink);
    run<4, 2, 2, 0, 4>("4 FFMA2 + 2 FADD2 + 2 FMUL2", sms, cps, sink);
    run<7, 4, 4, 0, 2>("7 FFMA2 + 4 FADD2 + 4 FMUL2 (UG, no MUFU)", sms, cps, sink);
    run<7, 4, 4, 2, 2>("7 FFMA2 + 4 FADD2 + 4 FMUL2 + 2 MUFU (UG)", sms, cps, sink);
    run<15, 0, 0, 2, 2>("15 FFMA2 + 2 MUFU (UG as all-FMA)", sms, cps, sink);
    run<4, 3, 1, 2, 4>("4 FFMA2 + 3 FADD2 + 1 FMUL2 + 2 MUFU (RM vel)", sms, cps, sink);
    run<8, 0, 0, 2, 4>("8 FFMA2 + 2 MUFU (RM vel as all-FMA)", sms, cps, sink);
  }
  return 0;
}
