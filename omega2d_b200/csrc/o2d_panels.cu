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
// One thread per PAIR of panels (A = 2q, B = 2q+1); 4 float4 per pair so that every operand of the
// packed-lane arithmetic is one 8-byte half of an LDS.128:
//   {x0A x0B y0A y0B} {HxA HxB HyA HyB} {aA aB bA bB} {thrA thrB thr2A thr2B}
// H = -p/2 (p = node1 - node0), a,b = the strength-weighted unit tangent of the header comment times -2
// (panel_far2 returns -ustar/2, -vstar/2), thr = |H|^2 / kFarZeta2: a target with |t - centre|^2 <= thr takes
// the general panel_pair(); thr2 = |H|^2 / kVeryFarZeta2: beyond it the shorter series of panel_far2v is exact enough.
__global__ void __launch_bounds__(256)
pack_panels_kernel(int np, int j0, const float* __restrict__ nx, const float* __restrict__ ny,
                   const uint32_t* __restrict__ idx, const float* __restrict__ vs,
                   const float* __restrict__ ss, float4* __restrict__ packed, int np_padded) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * q >= np_padded) return;
  float x0[2], y0[2], hx[2], hy[2], a[2], b[2], thr[2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const int j = j0 + 2 * q + l;   // j >= np: padding = the last panel with zero strengths
    const int jj = j < np ? j : np - 1;
    const uint32_t i0 = idx[2 * jj], i1 = idx[2 * jj + 1];
    x0[l] = nx[i0]; y0[l] = ny[i0];
    const float px = nx[i1] - x0[l], py = ny[i1] - y0[l];
    const float mult = 1.0f / sqrtf(px * px + py * py);   // my_rsqrt(panl), src/Kernels.h:385
    const float v = j < np ? vs[jj] : 0.0f;
    const float s = (j < np && ss != nullptr) ? ss[jj] : 0.0f;
    a[l] = -2.0f * ((v * px + s * py) * mult);
    b[l] = -2.0f * ((v * py - s * px) * mult);
    hx[l] = -0.5f * px; hy[l] = -0.5f * py;
    thr[l] = (hx[l] * hx[l] + hy[l] * hy[l]) * (1.0f / kFarZeta2);
  }
  packed[4 * q]     = make_float4(x0[0], x0[1], y0[0], y0[1]);
  packed[4 * q + 1] = make_float4(hx[0], hx[1], hy[0], hy[1]);
  packed[4 * q + 2] = make_float4(a[0], a[1], b[0], b[1]);
  packed[4 * q + 3] = make_float4(thr[0], thr[1], thr[0] * (kFarZeta2 / kVeryFarZeta2), thr[1] * (kFarZeta2 / kVeryFarZeta2));
}

cudaError_t launch_pack_panels(int np, const float* nx, const float* ny, const uint32_t* idx,
                               const float* vs, const float* ss, float4* packed, int np_padded,
                               cudaStream_t st, int j0) {
  if (np_padded <= 0) return cudaSuccess;
  const int nq = np_padded / 2;
  pack_panels_kernel<<<(nq + 255) / 256, 256, 0, st>>>(np, j0, nx, ny, idx, vs, ss, packed, np_padded);
  return cudaGetLastError();
}

// ---- panels on points --------------------------------------------------------------------------------
// A thread owns R targets; a packed lane pair is TWO PANELS against one target.
constexpr int kPanThreads = 128;

// general (near-field) pair for the lanes that need it; result in panel_far2's scaling
__device__ __forceinline__ void near_fix(bool nx_, bool ny_, float x0a, float y0a, float hxa, float hya,
                                         float x0b, float y0b, float hxb, float hyb, float tx, float ty,
                                         float2& rr, float2& ri) {
  float us, vs;
  panel_pair(x0a, y0a, -2.0f * hxa, -2.0f * hya, tx, ty, us, vs);
  if (nx_) { ri.x = -0.5f * us; rr.x = -0.5f * vs; }
  panel_pair(x0b, y0b, -2.0f * hxb, -2.0f * hyb, tx, ty, us, vs);
  if (ny_) { ri.y = -0.5f * us; rr.y = -0.5f * vs; }
}

template <int R>
__global__ void __launch_bounds__(kPanThreads)
panels_on_points_kernel(const PanelArgs a) {
  __shared__ float4 s[2 * kPanelTile];   // kPanelTile/2 panel pairs x 4 float4
  const int tid = threadIdx.x;
  const int base = blockIdx.x * (kPanThreads * R) + tid;
  float tx[R], ty[R];
  double accu[R], accv[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = base + k * kPanThreads;
    const bool ok = i < a.nt;
    tx[k] = ok ? a.tx[i] : 1.0e6f;
    ty[k] = ok ? a.ty[i] : 1.0e6f;
    accu[k] = 0.0; accv[k] = 0.0;
  }
  const int ntiles_all = a.npanels_padded / kPanelTile;
  const int tile_begin = blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(tile_begin + a.tiles_per_chunk, ntiles_all);
  for (int t = tile_begin; t < tile_end; ++t) {
    __syncthreads();
    const float4* g = a.packed + (size_t)t * (2 * kPanelTile);
    for (int k = tid; k < 2 * kPanelTile; k += kPanThreads) s[k] = g[k];
    __syncthreads();
    float2 pu[R], pv[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { pu[k] = f2(0.0f); pv[k] = f2(0.0f); }
#pragma unroll 2
    for (int q = 0; q < kPanelTile / 2; ++q) {
      const float4 P = s[4 * q];        // x0A x0B y0A y0B
      const float4 H = s[4 * q + 1];    // HxA HxB HyA HyB
      const float4 S = s[4 * q + 2];    // aA aB bA bB
      const float4 T = s[4 * q + 3];    // thrA thrB
      const float2 Hx = f2(H.x, H.y), Hy = f2(H.z, H.w);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        // e = (t - node0) + H: the subtraction first, so it is exact for nearby targets as in the reference
        const float2 ex = add2(add2(f2(tx[k]), neg2(f2(P.x, P.y))), Hx);
        const float2 ey = add2(add2(f2(ty[k]), neg2(f2(P.z, P.w))), Hy);
        float2 rr, ri, e2;
        panel_far2v(ex, ey, Hx, Hy, f2(T.z, T.w), rr, ri, e2);
        const bool na = !(e2.x > T.x), nb = !(e2.y > T.y);   // also true for NaN
        if (__any_sync(0xffffffffu, na || nb))
          near_fix(na, nb, P.x, P.z, H.x, H.z, P.y, P.w, H.y, H.w, tx[k], ty[k], rr, ri);
        const float2 A = f2(S.x, S.y), B = f2(S.z, S.w);
        pu[k] = fma2(ri, A, pu[k]);  pu[k] = fma2(neg2(rr), B, pu[k]);
        pv[k] = fma2(ri, B, pv[k]);  pv[k] = fma2(rr, A, pv[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      accu[k] += (double)pu[k].x + (double)pu[k].y;
      accv[k] += (double)pv[k].x + (double)pv[k].y;
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = base + k * kPanThreads;
    if (i >= a.nt) continue;
    if (a.partial == nullptr) {
      a.out.p[0][i] = (float)((double)a.out.p[0][i] + accu[k]);
      a.out.p[1][i] = (float)((double)a.out.p[1][i] + accv[k]);
    } else {
      double* dst = a.partial + ((size_t)blockIdx.y * 2) * (size_t)a.nt + i;
      dst[0] = accu[k];
      dst[a.nt] = accv[k];
    }
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

cudaError_t launch_panels_on_points(const PanelArgs& a, int nchunks, int rpt, cudaStream_t st) {
  dim3 grid((a.nt + kPanThreads * rpt - 1) / (kPanThreads * rpt), nchunks, 1);
  if (rpt == 2) panels_on_points_kernel<2><<<grid, kPanThreads, 0, st>>>(a);
  else          panels_on_points_kernel<1><<<grid, kPanThreads, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_panels_combine(const double* partial, int nchunks, int nt, const OutPtrs& out, cudaStream_t st) {
  panels_combine_kernel<<<(nt + 255) / 256, 256, 0, st>>>(partial, nchunks, nt, out);
  return cudaGetLastError();
}
int panels_occupancy(int rpt) {
  int nb = 0;
  const cudaError_t e = rpt == 2
      ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, panels_on_points_kernel<2>, kPanThreads, 0)
      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, panels_on_points_kernel<1>, kPanThreads, 0);
  if (e != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

// ---- points on panels (reciprocity) ----------------------------------------------------------------------
// 128 threads x 2 panels per thread: a staged particle pair is read once (two LDS.128) for two panels, and the two
// independent series evaluations per thread keep the FMA pipe fed (one panel per thread on 64-thread CTAs held 16
// warps per SM at 76 % FMA-pipe activity).
constexpr int kPtPanThreads = 128;
constexpr int kPtPanPPT = 2;

__global__ void __launch_bounds__(kPtPanThreads)
points_on_panels_kernel(const PtPanArgs a) {
  __shared__ float4 s[kTile];   // one tile of packed particle pairs (16 KiB)
  const int tid = threadIdx.x;
  float x0[kPtPanPPT], y0[kPtPanPPT], px[kPtPanPPT], py[kPtPanPPT], thr[kPtPanPPT], thr2[kPtPanPPT];
  float2 Hx[kPtPanPPT], Hy[kPtPanPPT];
  bool ok[kPtPanPPT];
#pragma unroll
  for (int k = 0; k < kPtPanPPT; ++k) {
    const int i = (blockIdx.x * kPtPanPPT + k) * kPtPanThreads + tid;
    ok[k] = i < a.np;
    x0[k] = 1.0e6f; y0[k] = 1.0e6f; px[k] = 1.0f; py[k] = 0.0f;
    if (ok[k]) {
      const uint32_t i0 = a.idx[2 * i], i1 = a.idx[2 * i + 1];
      x0[k] = a.nx[i0]; y0[k] = a.ny[i0];
      px[k] = a.nx[i1] - x0[k]; py[k] = a.ny[i1] - y0[k];
    }
    // far-field form (panel_far2) with the roles mirrored: e' = centre - t = (node0 - t) + p/2 and H' = +p/2 give
    // the same zeta'; a packed lane pair is TWO PARTICLES against this thread's panel
    Hx[k] = f2(0.5f * px[k]); Hy[k] = f2(0.5f * py[k]);
    thr[k] = (Hx[k].x * Hx[k].x + Hy[k].x * Hy[k].x) * (1.0f / kFarZeta2);
    thr2[k] = thr[k] * (kFarZeta2 / kVeryFarZeta2);
  }
  const int tile_begin = blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(tile_begin + a.tiles_per_chunk, a.ntiles);
  double acc1[kPtPanPPT], acc2[kPtPanPPT];
#pragma unroll
  for (int k = 0; k < kPtPanPPT; ++k) { acc1[k] = 0.0; acc2[k] = 0.0; }
  for (int t = tile_begin; t < tile_end; ++t) {
    __syncthreads();
    const float4* g = a.packed_src + (size_t)t * kTile;
    for (int k = tid; k < kTile; k += kPtPanThreads) s[k] = g[k];
    __syncthreads();
    float2 p1[kPtPanPPT], p2[kPtPanPPT];
#pragma unroll
    for (int k = 0; k < kPtPanPPT; ++k) { p1[k] = f2(0.0f); p2[k] = f2(0.0f); }
#pragma unroll 2
    for (int p = 0; p < kTile / 2; ++p) {
      const float4 A = s[2 * p];      // -x0 -x1 -y0 -y1 of two particles
      const float4 B = s[2 * p + 1];  // s0 s1 . .
      const float2 S = f2(B.x, B.y);
#pragma unroll
      for (int k = 0; k < kPtPanPPT; ++k) {
        const float2 ex = add2(add2(f2(A.x, A.y), f2(x0[k])), Hx[k]);
        const float2 ey = add2(add2(f2(A.z, A.w), f2(y0[k])), Hy[k]);
        float2 rr, ri, e2;
        panel_far2v(ex, ey, Hx[k], Hy[k], f2(thr2[k]), rr, ri, e2);
        const bool na = !(e2.x > thr[k]), nb = !(e2.y > thr[k]);
        if (__any_sync(0xffffffffu, na || nb)) {
          float us, vs;
          panel_pair(x0[k], y0[k], px[k], py[k], -A.x, -A.z, us, vs);
          if (na) { ri.x = -0.5f * us; rr.x = -0.5f * vs; }
          panel_pair(x0[k], y0[k], px[k], py[k], -A.y, -A.w, us, vs);
          if (nb) { ri.y = -0.5f * us; rr.y = -0.5f * vs; }
        }
        p1[k] = fma2(S, ri, p1[k]);
        p2[k] = fma2(S, rr, p2[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kPtPanPPT; ++k) {
      acc1[k] += -2.0 * ((double)p1[k].x + (double)p1[k].y);
      acc2[k] += -2.0 * ((double)p2[k].x + (double)p2[k].y);
    }
  }
#pragma unroll
  for (int k = 0; k < kPtPanPPT; ++k) {
    if (!ok[k]) continue;
    const int i = (blockIdx.x * kPtPanPPT + k) * kPtPanThreads + tid;
    if (a.partial == nullptr) {
      const double mult = 1.0 / sqrt((double)px[k] * px[k] + (double)py[k] * py[k]);
      const double U = (acc1[k] * px[k] - acc2[k] * py[k]) * mult;
      const double V = (acc1[k] * py[k] + acc2[k] * px[k]) * mult;
      const double plen = 1.0 / (double)a.area[i];       // src/Influence.h:727
      a.pu[i] = (float)((double)a.pu[i] - plen * U);      // src/Influence.h:758-759
      a.pv[i] = (float)((double)a.pv[i] - plen * V);
    } else {
      double* dst = a.partial + ((size_t)blockIdx.y * 2) * (size_t)a.np + i;
      dst[0] = acc1[k];
      dst[a.np] = acc2[k];
    }
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
  dim3 grid((a.np + kPtPanThreads * kPtPanPPT - 1) / (kPtPanThreads * kPtPanPPT), nchunks, 1);
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
int ptpan_block_panels() { return kPtPanThreads * kPtPanPPT; }
int ptpan_occupancy() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, points_on_panels_kernel, kPtPanThreads, 0) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

}  // namespace o2d
