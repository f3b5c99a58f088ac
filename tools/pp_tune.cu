// tools/pp_tune.cu -- development harness: times variants (targets per thread, block size, unroll) of the
// pair-sum kernel on synthetic data.  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I omega2d_b200/csrc tools/pp_tune.cu -o build/pp_tune
#include "o2d_pp_impl.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace o2d;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void pack_kernel(int ns, const float* sx, const float* sy, const float* ss, const float* sr, float4* packed, int npairs) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  int j0 = min(2 * p, ns - 1), j1 = min(2 * p + 1, ns - 1);
  packed[2 * p] = make_float4(-sx[j0], -sx[j1], -sy[j0], -sy[j1]);
  packed[2 * p + 1] = make_float4(2 * p < ns ? ss[j0] : 0.f, 2 * p + 1 < ns ? ss[j1] : 0.f, sr[j0] * sr[j0], sr[j1] * sr[j1]);
}

template <int CORE, int RES, int THREADS, int R, int UNROLL>
double run_variant(const char* name, const PPArgs& base, int n, int sms, double* d_partial, float* d_u, int reps) {
  auto kern = pp_kernel<CORE, RES, THREADS, kTile, R, UNROLL>;
  size_t smem = 2 * kTile * 16 + 64;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  int nb = (n + THREADS * R - 1) / (THREADS * R);
  int ntiles = base.tile_end;
  int slots = sms * occ;
  // pick chunks so the last wave is full
  int best = 1; double beff = 0;
  for (int nc = 1; nc <= 64 && nc <= ntiles; ++nc) {
    int tpc = (ntiles + nc - 1) / nc; if ((ntiles + tpc - 1) / tpc != nc) continue;
    long long items = (long long)nb * nc; long long waves = (items + slots - 1) / slots;
    double eff = (double)nb * ntiles / ((double)waves * tpc * slots) - 0.0005 * nc;
    if (eff > beff) { beff = eff; best = nc; }
    if (beff > 0.985) break;
  }
  PPArgs a = base;
  a.tiles_per_chunk = (ntiles + best - 1) / best;
  a.partial = d_partial;
  dim3 grid(nb, best);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<grid, THREADS, smem>>>(a);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    kern<<<grid, THREADS, smem>>>(a);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best_ms = fminf(best_ms, ms);
  }
  double pairs = (double)n * (double)ntiles * kTile;
  double rate = pairs / (best_ms * 1e-3);
  // checksum of chunk 0 partial U
  std::vector<double> h(16);
  CK(cudaMemcpy(h.data(), d_partial, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  printf("%-34s regs=%3d occ=%d chunks=%2d eff=%.3f  %8.3f ms  %.3e pairs/s  %.2f pairs/clk/SM  chk=%.9e\n", name, fa.numRegs, occ,
         best, beff, best_ms, rate, rate / (sms * 1.965e9), h[3]);
  fflush(stdout);
  return rate;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 262144;
  int reps = argc > 2 ? atoi(argv[2]) : 3;
  const bool uniform = argc > 3 && atoi(argv[3]) != 0;   // same radius for every particle
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, n=%d, %s radii\n", prop.name, sms, n, uniform ? "uniform" : "per-particle");
  std::vector<float> x(n), y(n), s(n), r(n);
  srand(1);
  for (int i = 0; i < n; ++i) {
    float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = rand() / (float)RAND_MAX;
    float g = sqrtf(-2 * logf(u1));
    x[i] = 0.25f * g * cosf(6.2831853f * u2); y[i] = 0.25f * g * sinf(6.2831853f * u2);
    s[i] = 1.0f / n; r[i] = 1.5f / sqrtf((float)n) * (uniform ? 1.0f : 0.8f + 0.4f * (rand() / (float)RAND_MAX));
  }
  float *dx, *dy, *ds, *dr, *du, *dv; float4* packed; double* partial;
  int ntiles = (n + kTile - 1) / kTile;
  CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&dy, n * 4)); CK(cudaMalloc(&ds, n * 4)); CK(cudaMalloc(&dr, n * 4));
  CK(cudaMalloc(&du, n * 4)); CK(cudaMalloc(&dv, n * 4));
  CK(cudaMalloc(&packed, (size_t)ntiles * kTile * 16));
  CK(cudaMalloc(&partial, (size_t)64 * 6 * n * sizeof(double)));
  CK(cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, y.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ds, s.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dr, r.data(), n * 4, cudaMemcpyHostToDevice));
  int npairs = ntiles * kTile / 2;
  pack_kernel<<<(npairs + 255) / 256, 256>>>(n, dx, dy, ds, dr, packed, npairs);
  CK(cudaDeviceSynchronize());
  PPArgs a{};
  a.packed = packed; a.tile_begin = 0; a.tile_end = ntiles; a.skip_begin = 1 << 30; a.skip_len = 0; a.chunk_offset = 0;
  a.tx = dx; a.ty = dy; a.tr = dr; a.nt = n; a.out.p[0] = du; a.out.p[1] = dv;
#define V(CORE, RES, T, R, U) run_variant<CORE, RES, T, R, U>(#CORE " " #RES " T" #T " R" #R " U" #U, a, n, sms, partial, du, reps)
  V(CORE_RM, RES_U, 256, 4, 4);
  V(CORE_RM, RES_U, 256, 2, 4);
  V(CORE_RM, RES_U, 256, 8, 2);
  V(CORE_RM, RES_U, 128, 4, 4);
  V(CORE_RM, RES_U, 128, 8, 2);
  V(CORE_RM, RES_U, 128, 2, 4);
  V(CORE_RM, RES_U, 256, 4, 2);
  V(CORE_RM, RES_U, 256, 4, 8);
  V(CORE_RM, RES_U, 256, 6, 2);
  V(CORE_RM, RES_U, 512, 2, 4);
  V(CORE_RM, RES_U, 512, 4, 2);
  V(CORE_V2, RES_U, 256, 4, 4);
  V(CORE_V2, RES_U, 256, 2, 4);
  V(CORE_V2, RES_U, 256, 8, 2);
  V(CORE_WL, RES_U, 256, 4, 4);
  V(CORE_WL, RES_U, 256, 2, 4);
  V(CORE_RM, RES_UG, 256, 2, 4);
  V(CORE_RM, RES_UG, 256, 4, 2);
  V(CORE_RM, RES_UG, 256, 1, 4);
  V(CORE_RM, RES_UG, 128, 2, 4);
  V(CORE_RM, RES_UW, 256, 2, 4);
  V(CORE_RM, RES_UW, 256, 4, 2);
  return 0;
}
