// o2d_panels.cu -- constant-strength panel kernels.
//
//   panels_on_points  replaces panels_affect_points<S,A>   (reference src/Influence.h:548-602)
//   points_on_panels  replaces points_affect_panels<S,A>   (reference src/Influence.h:719-763)
// Pair kernels: kernelu_1vs_0p / kernelu_1v_0p (reference src/Kernels.h:362-405, :312-354) with
// get_ustar (src/Kernels.h:262-268).
//
// Per pair the reference evaluates
//     ustar = atan2(dx1,dy1) - atan2(dx0,dy0)   wrapped into [-pi,pi]
//     vstar = 0.5 log(rij2/rij12)
// and rotates (ustar,vstar) by the unit panel tangent (px,py), scaled by the sheet strengths.
// Here the same two numbers are formed without the cancellations of that form:
//     ustar = atan2(cross, dot),  cross = py*dx0 - px*dy0 (= dx1*dy0 - dy1*dx0),  dot = dx0*dx1 + dy0*dy1
//       -- the angle between the two node vectors; the +-pi wrap is implicit;
//     vstar = atanh((rij2-rij12)/(rij2+rij12)),  rij2-rij12 = px*(dx0+dx1) + py*(dy0+dy1)
// with (px,py) = node1 - node0; both evaluated with FMA-pipe polynomials (atan2_fast, panel_pair below).  The rotation and the strengths are linear, so they are hoisted:
//   panels_on_points:  u += ustar*a - vstar*b ,  v += ustar*b + vstar*a
//                      a = (vs*px + ss*py)/len , b = (vs*py - ss*px)/len   (per panel, packed once)
//   points_on_panels:  A1 = sum s_j ustar_j , A2 = sum s_j vstar_j   (per panel), then
//                      U = (A1*px - A2*py)/len , V = (A1*py + A2*px)/len ,  pu -= U/area.
// Accumulation: float partial sums per tile, float64 across tiles, as in the particle kernels.
#include "o2d_device.cuh"
#include "o2d_internal.h"

namespace o2d {

// ---- packing: gather node coordinates through idx once (the reference gathers per pair,
// src/Influence.h:561-562) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_panels_kernel(int np, const float* __restrict__ nx, const float* __restrict__ ny,
                   const uint32_t* __restrict__ idx, const float* __restrict__ vs,
                   const float* __restrict__ ss, float4* __restrict__ packed, int np_padded) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np_padded) return;
  const int jj = j < np ? j : np - 1;
  const uint32_t i0 = idx[2 * jj], i1 = idx[2 * jj + 1];
  const float x0 = nx[i0], y0 = ny[i0];
  const float px = nx[i1] - x0, py = ny[i1] - y0;
  const float mult = 1.0f / sqrtf(px * px + py * py);   // my_rsqrt(panl), src/Kernels.h:385
  const float v = j < np ? vs[jj] : 0.0f;
  const float s = (j < np && ss != nullptr) ? ss[jj] : 0.0f;
  const float a = (v * px + s * py) * mult;
  const float b = (v * py - s * px) * mult;
  packed[2 * j]     = make_float4(x0, y0, px, py);
  packed[2 * j + 1] = make_float4(a, b, 0.0f, 0.0f);
}

cudaError_t launch_pack_panels(int np, const float* nx, const float* ny, const uint32_t* idx,
                               const float* vs, const float* ss, float4* packed, int np_padded,
                               cudaStream_t st) {
  if (np_padded <= 0) return cudaSuccess;
  pack_panels_kernel<<<(np_padded + 255) / 256, 256, 0, st>>>(np, nx, ny, idx, vs, ss, packed, np_padded);
  return cudaGetLastError();
}

// ---- panels on points --------------------------------------------------------------------------------
constexpr int kPanThreads = 128;

__global__ void __launch_bounds__(kPanThreads)
panels_on_points_kernel(const PanelArgs a) {
  __shared__ float4 s[2 * kPanelTile];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kPanThreads + tid;
  const bool ok = i < a.nt;
  const float tx = ok ? a.tx[i] : 1.0e6f;
  const float ty = ok ? a.ty[i] : 1.0e6f;
  const int ntiles_all = a.npanels_padded / kPanelTile;
  const int tile_begin = blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(tile_begin + a.tiles_per_chunk, ntiles_all);
  double accu = 0.0, accv = 0.0;
  for (int t = tile_begin; t < tile_end; ++t) {
    __syncthreads();
    const float4* g = a.packed + (size_t)t * (2 * kPanelTile);
    for (int k = tid; k < 2 * kPanelTile; k += kPanThreads) s[k] = g[k];
    __syncthreads();
    float pu = 0.0f, pv = 0.0f;
#pragma unroll 2
    for (int j = 0; j < kPanelTile; ++j) {
      const float4 G = s[2 * j];
      const float4 H = s[2 * j + 1];
      float ustar, vstar;
      panel_pair(G.x, G.y, G.z, G.w, tx, ty, ustar, vstar);
      pu = fmaf(ustar, H.x, pu);  pu = fmaf(-vstar, H.y, pu);
      pv = fmaf(ustar, H.y, pv);  pv = fmaf(vstar, H.x, pv);
    }
    accu += (double)pu;
    accv += (double)pv;
  }
  if (!ok) return;
  if (a.partial == nullptr) {
    a.out.p[0][i] = (float)((double)a.out.p[0][i] + accu);
    a.out.p[1][i] = (float)((double)a.out.p[1][i] + accv);
  } else {
    double* dst = a.partial + ((size_t)blockIdx.y * 2) * (size_t)a.nt + i;
    dst[0] = accu;
    dst[a.nt] = accv;
  }
}

__global__ void __launch_bounds__(256)
panels_combine_kernel(const double* __restrict__ partial, int nchunks, int nt, const OutPtrs out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  double u = 0.0, v = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const double* src = partial + ((size_t)c * 2) * (size_t)nt + i;
    u += src[0];
    v += src[nt];
  }
  out.p[0][i] = (float)((double)out.p[0][i] + u);
  out.p[1][i] = (float)((double)out.p[1][i] + v);
}

cudaError_t launch_panels_on_points(const PanelArgs& a, int nchunks, cudaStream_t st) {
  dim3 grid((a.nt + kPanThreads - 1) / kPanThreads, nchunks, 1);
  panels_on_points_kernel<<<grid, kPanThreads, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_panels_combine(const double* partial, int nchunks, int nt, const OutPtrs& out, cudaStream_t st) {
  panels_combine_kernel<<<(nt + 255) / 256, 256, 0, st>>>(partial, nchunks, nt, out);
  return cudaGetLastError();
}
int panels_occupancy() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, panels_on_points_kernel, kPanThreads, 0) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

// ---- points on panels (reciprocity) ----------------------------------------------------------------------
constexpr int kPtPanThreads = 64;

__global__ void __launch_bounds__(kPtPanThreads)
points_on_panels_kernel(const PtPanArgs a) {
  __shared__ float4 s[kTile];   // one tile of packed particle pairs (16 KiB)
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kPtPanThreads + tid;
  const bool ok = i < a.np;
  float x0 = 1.0e6f, y0 = 1.0e6f, px = 1.0f, py = 0.0f;
  if (ok) {
    const uint32_t i0 = a.idx[2 * i], i1 = a.idx[2 * i + 1];
    x0 = a.nx[i0]; y0 = a.ny[i0];
    px = a.nx[i1] - x0; py = a.ny[i1] - y0;
  }
  const int tile_begin = blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(tile_begin + a.tiles_per_chunk, a.ntiles);
  double acc1 = 0.0, acc2 = 0.0;
  for (int t = tile_begin; t < tile_end; ++t) {
    __syncthreads();
    const float4* g = a.packed_src + (size_t)t * kTile;
    for (int k = tid; k < kTile; k += kPtPanThreads) s[k] = g[k];
    __syncthreads();
    float p1 = 0.0f, p2 = 0.0f;
#pragma unroll 2
    for (int p = 0; p < kTile / 2; ++p) {
      const float4 A = s[2 * p];      // -x0 -x1 -y0 -y1 of two particles
      const float4 B = s[2 * p + 1];  // s0 s1 . .
      float ustar, vstar;
      panel_pair(x0, y0, px, py, -A.x, -A.z, ustar, vstar);
      p1 = fmaf(B.x, ustar, p1);  p2 = fmaf(B.x, vstar, p2);
      panel_pair(x0, y0, px, py, -A.y, -A.w, ustar, vstar);
      p1 = fmaf(B.y, ustar, p1);  p2 = fmaf(B.y, vstar, p2);
    }
    acc1 += (double)p1;
    acc2 += (double)p2;
  }
  if (!ok) return;
  if (a.partial == nullptr) {
    const double mult = 1.0 / sqrt((double)px * px + (double)py * py);
    const double U = (acc1 * px - acc2 * py) * mult;
    const double V = (acc1 * py + acc2 * px) * mult;
    const double plen = 1.0 / (double)a.area[i];       // src/Influence.h:727
    a.pu[i] = (float)((double)a.pu[i] - plen * U);      // src/Influence.h:758-759
    a.pv[i] = (float)((double)a.pv[i] - plen * V);
  } else {
    double* dst = a.partial + ((size_t)blockIdx.y * 2) * (size_t)a.np + i;
    dst[0] = acc1;
    dst[a.np] = acc2;
  }
}

__global__ void __launch_bounds__(256)
ptpan_combine_kernel(const double* __restrict__ partial, int nchunks, int np,
                     const float* __restrict__ nx, const float* __restrict__ ny,
                     const uint32_t* __restrict__ idx, const float* __restrict__ area,
                     float* pu, float* pv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  double acc1 = 0.0, acc2 = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const double* src = partial + ((size_t)c * 2) * (size_t)np + i;
    acc1 += src[0];
    acc2 += src[np];
  }
  const uint32_t i0 = idx[2 * i], i1 = idx[2 * i + 1];
  const float px = nx[i1] - nx[i0], py = ny[i1] - ny[i0];
  const double mult = 1.0 / sqrt((double)px * px + (double)py * py);
  const double U = (acc1 * px - acc2 * py) * mult;
  const double V = (acc1 * py + acc2 * px) * mult;
  const double plen = 1.0 / (double)area[i];
  pu[i] = (float)((double)pu[i] - plen * U);
  pv[i] = (float)((double)pv[i] - plen * V);
}

cudaError_t launch_points_on_panels(const PtPanArgs& a, int nchunks, cudaStream_t st) {
  dim3 grid((a.np + kPtPanThreads - 1) / kPtPanThreads, nchunks, 1);
  points_on_panels_kernel<<<grid, kPtPanThreads, 0, st>>>(a);
  return cudaGetLastError();
}

// combine needs the panel geometry again; it is passed through PtPanArgs by the caller
cudaError_t launch_ptpan_combine_geom(const double* partial, int nchunks, const PtPanArgs& a, cudaStream_t st) {
  ptpan_combine_kernel<<<(a.np + 255) / 256, 256, 0, st>>>(partial, nchunks, a.np, a.nx, a.ny, a.idx, a.area, a.pu, a.pv);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256)
ptpan_reduce_chunks_kernel(const double* __restrict__ partial, int nchunks, int np, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * np) return;
  double acc = 0.0;
  for (int c = 0; c < nchunks; ++c) acc += partial[(size_t)c * 2 * (size_t)np + i];
  out[i] = acc;
}
cudaError_t launch_ptpan_reduce_chunks(const double* partial, int nchunks, int np, double* out, cudaStream_t st) {
  ptpan_reduce_chunks_kernel<<<(2 * np + 255) / 256, 256, 0, st>>>(partial, nchunks, np, out);
  return cudaGetLastError();
}
int ptpan_occupancy() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, points_on_panels_kernel, kPtPanThreads, 0) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

}  // namespace o2d
