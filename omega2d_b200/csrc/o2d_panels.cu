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
// with (px,py) = node1 - node0; both evaluated with FMA-pipe polynomials (atan2_fast, panel_pair below) -- that is the
// NEAR form (targets within 3.2 panel lengths of the panel centre).  Everywhere else the two numbers are the real and
// imaginary part of one complex series in g = conj(1/(t - centre)) whose coefficients absorb the half panel vector, the
// rotation by the tangent and (panels_on_points) the strengths: o2d_device.cuh, "far field as a Laurent series".
//   panels_on_points:  u + i v += g (D1 + D3 g^2 [+ D5 g^4 [+ D7 g^6]])        D_k per panel, packed once (float64 -> float)
//   points_on_panels:  per panel  sum_j s_j g_j  and  sum_j s_j g_j^3  (two-term tier; the coefficients are applied once
//                      at the end in float64), the 3- and 4-term tiers and the near form summed per pair as
//                      A1 = sum s_j ustar_j , A2 = sum s_j vstar_j; then
//                      U = (A1*px - A2*py)/len , V = (A1*py + A2*px)/len ,  pu -= U/area.
// Accumulation: float partial sums per tile, float64 across tiles, as in the particle kernels.
#include "o2d_device.cuh"
#include "o2d_internal.h"

namespace o2d {

// ---- packing: gather node coordinates through idx once (the reference gathers per pair,
// src/Influence.h:561-562) ---------------------------------------------------------------------------
// One thread per PAIR of panels (A = 2q, B = 2q+1); 2*kPanelRecF4 = 8 float4 per pair so that every operand of the
// packed-lane arithmetic is one 8-byte half of an LDS.128:
//   {x0A x0B y0A y0B} {HxA HxB HyA HyB} {D1x D1x D1y D1y} {D3..} {D5..} {D7..} {thrA thrB thr2A thr2B} {thr3A thr3B . .}
// H = -p/2 (p = node1 - node0); D_k = m conj(H)^k / k with m = i (a + i b), a,b = the strength-weighted unit tangent of
// the header comment times -2 (the far series gives -ustar/2, -vstar/2), formed in float64 and rounded once
// (panel_far_coeffs, o2d_device.cuh); thr = |H|^2 / kFarZeta2: a target with |t - centre|^2 <= thr takes the general
// panel_pair(); thr2, thr3 = |H|^2 / kVeryFarZeta2, / kUltraFarZeta2: beyond them 3 and 2 series terms are exact enough.
__global__ void __launch_bounds__(256)
pack_panels_kernel(int np, int j0, const float* __restrict__ nx, const float* __restrict__ ny,
                   const uint32_t* __restrict__ idx, const float* __restrict__ vs,
                   const float* __restrict__ ss, float4* __restrict__ packed, int np_padded) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * q >= np_padded) return;
  float x0[2], y0[2], hx[2], hy[2], thr[2], dx[2][4], dy[2][4];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const int j = j0 + 2 * q + l;   // j >= np: padding = the last panel with zero strengths
    const int jj = j < np ? j : np - 1;
    const uint32_t i0 = idx[2 * jj], i1 = idx[2 * jj + 1];
    x0[l] = nx[i0]; y0[l] = ny[i0];
    const float px = nx[i1] - x0[l], py = ny[i1] - y0[l];
    const double mult = 1.0 / sqrt((double)px * px + (double)py * py);   // my_rsqrt(panl), src/Kernels.h:385
    const double v = j < np ? vs[jj] : 0.0f;
    const double s = (j < np && ss != nullptr) ? ss[jj] : 0.0f;
    const double a = -2.0 * ((v * px + s * py) * mult);
    const double b = -2.0 * ((v * py - s * px) * mult);
    hx[l] = -0.5f * px; hy[l] = -0.5f * py;
    panel_far_coeffs(hx[l], hy[l], -b, a, dx[l], dy[l]);
    thr[l] = (hx[l] * hx[l] + hy[l] * hy[l]) * (1.0f / kFarZeta2);
  }
  float4* o = packed + (size_t)(2 * kPanelRecF4) * q;
  o[0] = make_float4(x0[0], x0[1], y0[0], y0[1]);
  o[1] = make_float4(hx[0], hx[1], hy[0], hy[1]);
#pragma unroll
  for (int k = 0; k < 4; ++k) o[2 + k] = make_float4(dx[0][k], dx[1][k], dy[0][k], dy[1][k]);
  o[6] = make_float4(thr[0], thr[1], thr[0] * (kFarZeta2 / kVeryFarZeta2), thr[1] * (kFarZeta2 / kVeryFarZeta2));
  o[7] = make_float4(thr[0] * (kFarZeta2 / kUltraFarZeta2), thr[1] * (kFarZeta2 / kUltraFarZeta2), 0.0f, 0.0f);
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
#ifndef O2D_PAN_MINB
#define O2D_PAN_MINB 1
#endif
constexpr int kPanThreads = 128;

// general (near-field) pair for one lane that needs it: the panel's contribution to (u, v).  The strength-weighted
// tangent comes back out of the first series coefficient, m = D1 / conj(H)  (a = Im m, b = -Re m).
__device__ __forceinline__ void near_pair(float x0, float y0, float hx, float hy, float d1x, float d1y,
                                          float tx, float ty, float& du, float& dv) {
  float us, vs;
  panel_pair(x0, y0, -2.0f * hx, -2.0f * hy, tx, ty, us, vs);
  const float h2 = hx * hx + hy * hy;
  const float ih2 = h2 > 0.0f ? 1.0f / h2 : 0.0f;
  const float a = (d1x * hy + d1y * hx) * ih2;
  const float b = -(d1x * hx - d1y * hy) * ih2;
  const float ri = -0.5f * us, rr = -0.5f * vs;
  du = ri * a - rr * b;
  dv = ri * b + rr * a;
}

template <int R>
__global__ void __launch_bounds__(kPanThreads, O2D_PAN_MINB)
panels_on_points_kernel(const PanelArgs a) {
  __shared__ float4 s[kPanelRecF4 * kPanelTile];   // kPanelTile/2 panel pairs x 8 float4 (16 KiB)
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
    const float4* g = a.packed + (size_t)t * (kPanelRecF4 * kPanelTile);
    const int npairs = min(kPanelTile, a.npanels_even - t * kPanelTile) / 2;   // the last tile may be partial
    for (int k = tid; k < 2 * kPanelRecF4 * npairs; k += kPanThreads) s[k] = g[k];
    __syncthreads();
    float2 pu[R], pv[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { pu[k] = f2(0.0f); pv[k] = f2(0.0f); }
#pragma unroll 2
    for (int q = 0; q < npairs; ++q) {
      const float4* rec = s + 8 * q;
      const float4 P = rec[0];        // x0A x0B y0A y0B
      const float4 H = rec[1];        // HxA HxB HyA HyB
      const float4 D1 = rec[2], D3 = rec[3];
      const float2 T3 = *reinterpret_cast<const float2*>(rec + 7);   // thr3A thr3B
      const float2 Hx = f2(H.x, H.y), Hy = f2(H.z, H.w);
      float2 gx[R], gy[R], Gx[R], Gy[R], e2[R];
      bool ultra = true;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        // e = (t - node0) + H: the subtraction first, so it is exact for nearby targets as in the reference
        const float2 ex = add2(add2(f2(tx[k]), neg2(f2(P.x, P.y))), Hx);
        const float2 ey = add2(add2(f2(ty[k]), neg2(f2(P.z, P.w))), Hy);
        panel_far_g(ex, ey, gx[k], gy[k], Gx[k], Gy[k], e2[k]);
        ultra = ultra && e2[k].x > T3.x && e2[k].y > T3.y;
      }
      if (__all_sync(0xffffffffu, ultra)) {
        // every pair of this warp beyond 23 panel lengths (the normal case): two terms, no further tests, and the R
        // independent evaluations interleave
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const float2 tr = fma2(f2(D3.x, D3.y), Gx[k], fma2(neg2(f2(D3.z, D3.w)), Gy[k], f2(D1.x, D1.y)));
          const float2 ti = fma2(f2(D3.x, D3.y), Gy[k], fma2(f2(D3.z, D3.w), Gx[k], f2(D1.z, D1.w)));
          pu[k] = fma2(tr, gx[k], pu[k]);  pu[k] = fma2(neg2(ti), gy[k], pu[k]);   // u + i v += t g
          pv[k] = fma2(tr, gy[k], pv[k]);  pv[k] = fma2(ti, gx[k], pv[k]);
        }
      } else {
        const float4 D5 = rec[4];
        const float4 T = rec[6];      // thrA thrB thr2A thr2B
#pragma unroll
        for (int k = 0; k < R; ++k) {
          float2 tr, ti;
          if (__all_sync(0xffffffffu, e2[k].x > T.z && e2[k].y > T.w)) {
            tr = f2(D5.x, D5.y); ti = f2(D5.z, D5.w);
          } else {
            const float4 D7 = rec[5];
            tr = f2(D7.x, D7.y); ti = f2(D7.z, D7.w);
            chorner(tr, ti, Gx[k], Gy[k], f2(D5.x, D5.y), f2(D5.z, D5.w));
          }
          chorner(tr, ti, Gx[k], Gy[k], f2(D3.x, D3.y), f2(D3.z, D3.w));
          chorner(tr, ti, Gx[k], Gy[k], f2(D1.x, D1.y), f2(D1.z, D1.w));
          const bool na = !(e2[k].x > T.x), nb = !(e2[k].y > T.y);   // also true for NaN
          if (__any_sync(0xffffffffu, na || nb)) {
            float du, dv;
            if (na) {
              near_pair(P.x, P.z, H.x, H.z, D1.x, D1.z, tx[k], ty[k], du, dv);
              pu[k].x += du; pv[k].x += dv; tr.x = 0.0f; ti.x = 0.0f; gx[k].x = 0.0f; gy[k].x = 0.0f;
            }
            if (nb) {
              near_pair(P.y, P.w, H.y, H.w, D1.y, D1.w, tx[k], ty[k], du, dv);
              pu[k].y += du; pv[k].y += dv; tr.y = 0.0f; ti.y = 0.0f; gx[k].y = 0.0f; gy[k].y = 0.0f;
            }
          }
          pu[k] = fma2(tr, gx[k], pu[k]);  pu[k] = fma2(neg2(ti), gy[k], pu[k]);
          pv[k] = fma2(tr, gy[k], pv[k]);  pv[k] = fma2(ti, gx[k], pv[k]);
        }
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
  if (rpt == kPanBigR) panels_on_points_kernel<kPanBigR><<<grid, kPanThreads, 0, st>>>(a);
  else          panels_on_points_kernel<1><<<grid, kPanThreads, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_panels_combine(const double* partial, int nchunks, int nt, const OutPtrs& out, cudaStream_t st) {
  panels_combine_kernel<<<(nt + 255) / 256, 256, 0, st>>>(partial, nchunks, nt, out);
  return cudaGetLastError();
}
int panels_occupancy(int rpt) {
  int nb = 0;
  const cudaError_t e = rpt == kPanBigR
      ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, panels_on_points_kernel<kPanBigR>, kPanThreads, 0)
      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, panels_on_points_kernel<1>, kPanThreads, 0);
  if (e != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

// ---- points on panels (reciprocity) ----------------------------------------------------------------------
// A thread owns kPtPanPPT panels; a packed lane pair is TWO PARTICLES against one panel.  Round 2 measured 64-thread
// CTAs x 1 panel (round 1: 16 warps per SM, 24.05 ms at 20 000 panels x 1M particles), 128 x 2 with the zeta series
// (20.7 ms) and, with the Laurent/moment form below, 128 x 2 (139 registers, 12 warps: 14.8 ms) against 128 x 1
// (80 registers, 24 warps: 14.7 ms, and 0.40 instead of 0.45 ms at 143 panels) -- one panel per thread it is.
#ifndef O2D_PTPAN_PPT
#define O2D_PTPAN_PPT 1
#endif
#ifndef O2D_PTPAN_MINB
#define O2D_PTPAN_MINB 1
#endif
constexpr int kPtPanThreads = 128;
constexpr int kPtPanPPT = O2D_PTPAN_PPT;

__global__ void __launch_bounds__(kPtPanThreads, O2D_PTPAN_MINB)
points_on_panels_kernel(const PtPanArgs a) {
  __shared__ float4 s[kTile];   // one tile of packed particle pairs (16 KiB)
  const int tid = threadIdx.x;
  float x0[kPtPanPPT], y0[kPtPanPPT], px[kPtPanPPT], py[kPtPanPPT], thr[kPtPanPPT], thr2[kPtPanPPT], thr3[kPtPanPPT];
  float2 Hx[kPtPanPPT], Hy[kPtPanPPT];
  float Ex[kPtPanPPT][4], Ey[kPtPanPPT][4];   // scalars: the packed instructions take a broadcast 32-bit operand
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
    // far-field series (panel_far_g, o2d_device.cuh) with the roles mirrored: e' = centre - t = (node0 - t) + p/2 and
    // H' = +p/2 give the same zeta' = H'/e'; a packed lane pair is TWO PARTICLES against this thread's panel, and the
    // series coefficients E_k = conj(H')^k / k are this thread's constants
    const float hx = 0.5f * px[k], hy = 0.5f * py[k];
    Hx[k] = f2(hx); Hy[k] = f2(hy);
    panel_far_coeffs(hx, hy, 1.0, 0.0, Ex[k], Ey[k]);
    thr[k] = (hx * hx + hy * hy) * (1.0f / kFarZeta2);
    thr2[k] = thr[k] * (kFarZeta2 / kVeryFarZeta2);
    thr3[k] = thr[k] * (kFarZeta2 / kUltraFarZeta2);
  }
  bool any_ok = false;
#pragma unroll
  for (int k = 0; k < kPtPanPPT; ++k) any_ok = any_ok || ok[k];
  const bool warp_works = __any_sync(0xffffffffu, any_ok);
  const int tile_begin = blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(tile_begin + a.tiles_per_chunk, a.ntiles);
  double acc1[kPtPanPPT], acc2[kPtPanPPT], M1x[kPtPanPPT], M1y[kPtPanPPT], M3x[kPtPanPPT], M3y[kPtPanPPT];
#pragma unroll
  for (int k = 0; k < kPtPanPPT; ++k) { acc1[k] = 0.0; acc2[k] = 0.0; M1x[k] = 0.0; M1y[k] = 0.0; M3x[k] = 0.0; M3y[k] = 0.0; }
  for (int t = tile_begin; t < tile_end; ++t) {
    __syncthreads();
    const float4* g = a.packed_src + (size_t)t * kTile;
    for (int k = tid; k < kTile; k += kPtPanThreads) s[k] = g[k];
    __syncthreads();
    float2 pi[kPtPanPPT], pr[kPtPanPPT];   // sums of s_j Im / Re conj(atanh zeta') = s_j ustar_j / 2, -s_j vstar_j / 2
    float2 m1x[kPtPanPPT], m1y[kPtPanPPT], m3x[kPtPanPPT], m3y[kPtPanPPT];   // moments sum s_j g_j, sum s_j g_j^3
#pragma unroll
    for (int k = 0; k < kPtPanPPT; ++k) {
      pi[k] = f2(0.0f); pr[k] = f2(0.0f);
      m1x[k] = f2(0.0f); m1y[k] = f2(0.0f); m3x[k] = f2(0.0f); m3y[k] = f2(0.0f);
    }
    if (!warp_works) continue;   // a warp past the last panel only helps staging (143 panels: 3 of the 8 warps)
#pragma unroll 2
    for (int p = 0; p < kTile / 2; ++p) {
      const float4 A = s[2 * p];      // -x0 -x1 -y0 -y1 of two particles
      const float4 B = s[2 * p + 1];  // s0 s1 . .
      const float2 S = f2(B.x, B.y);
      float2 gx[kPtPanPPT], gy[kPtPanPPT], Gx[kPtPanPPT], Gy[kPtPanPPT], e2[kPtPanPPT];
      bool ultra = true;
#pragma unroll
      for (int k = 0; k < kPtPanPPT; ++k) {
        const float2 ex = add2(add2(f2(A.x, A.y), f2(x0[k])), Hx[k]);
        const float2 ey = add2(add2(f2(A.z, A.w), f2(y0[k])), Hy[k]);
        panel_far_g(ex, ey, gx[k], gy[k], Gx[k], Gy[k], e2[k]);
        ultra = ultra && e2[k].x > thr3[k] && e2[k].y > thr3[k];
      }
      if (__all_sync(0xffffffffu, ultra)) {
        // every pair of this warp beyond 23 panel lengths (the normal case): two series terms, and because the
        // coefficients are this thread's constants only the moments sum s g and sum s g^3 are accumulated here
#pragma unroll
        for (int k = 0; k < kPtPanPPT; ++k) {
          const float2 sgx = mul2(S, gx[k]), sgy = mul2(S, gy[k]);
          m1x[k] = add2(m1x[k], sgx);
          m1y[k] = add2(m1y[k], sgy);
          m3x[k] = fma2(sgx, Gx[k], m3x[k]);  m3x[k] = fma2(neg2(sgy), Gy[k], m3x[k]);
          m3y[k] = fma2(sgx, Gy[k], m3y[k]);  m3y[k] = fma2(sgy, Gx[k], m3y[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < kPtPanPPT; ++k) {
          float2 tr, ti;
          if (__all_sync(0xffffffffu, e2[k].x > thr2[k] && e2[k].y > thr2[k])) {
            tr = f2(Ex[k][2]); ti = f2(Ey[k][2]);
          } else {
            tr = f2(Ex[k][3]); ti = f2(Ey[k][3]);
            chorner(tr, ti, Gx[k], Gy[k], f2(Ex[k][2]), f2(Ey[k][2]));
          }
          chorner(tr, ti, Gx[k], Gy[k], f2(Ex[k][1]), f2(Ey[k][1]));
          chorner(tr, ti, Gx[k], Gy[k], f2(Ex[k][0]), f2(Ey[k][0]));
          float2 Lr = fma2(tr, gx[k], neg2(mul2(ti, gy[k])));
          float2 Li = fma2(tr, gy[k], mul2(ti, gx[k]));
          const bool na = !(e2[k].x > thr[k]), nb = !(e2[k].y > thr[k]);
          if (__any_sync(0xffffffffu, na || nb)) {
            float us, vs;
            panel_pair(x0[k], y0[k], px[k], py[k], -A.x, -A.z, us, vs);
            if (na) { Li.x = 0.5f * us; Lr.x = -0.5f * vs; }
            panel_pair(x0[k], y0[k], px[k], py[k], -A.y, -A.w, us, vs);
            if (nb) { Li.y = 0.5f * us; Lr.y = -0.5f * vs; }
          }
          pi[k] = fma2(S, Li, pi[k]);
          pr[k] = fma2(S, Lr, pr[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kPtPanPPT; ++k) {
      acc1[k] += 2.0 * ((double)pi[k].x + (double)pi[k].y);
      acc2[k] += -2.0 * ((double)pr[k].x + (double)pr[k].y);
      M1x[k] += (double)m1x[k].x + (double)m1x[k].y;  M1y[k] += (double)m1y[k].x + (double)m1y[k].y;
      M3x[k] += (double)m3x[k].x + (double)m3x[k].y;  M3y[k] += (double)m3y[k].x + (double)m3y[k].y;
    }
  }
  // the far moments times this panel's coefficients: sum s conj(atanh zeta') = E1 M1 + E3 M3 (float64, E_k = conj(H')^k / k)
#pragma unroll
  for (int k = 0; k < kPtPanPPT; ++k) {
    const double cx = 0.5 * (double)px[k], cy = -0.5 * (double)py[k];
    const double c2x = cx * cx - cy * cy, c2y = 2.0 * cx * cy;
    const double e3x = (cx * c2x - cy * c2y) * (1.0 / 3.0), e3y = (cx * c2y + cy * c2x) * (1.0 / 3.0);
    const double Lr = cx * M1x[k] - cy * M1y[k] + e3x * M3x[k] - e3y * M3y[k];
    const double Li = cx * M1y[k] + cy * M1x[k] + e3x * M3y[k] + e3y * M3x[k];
    acc1[k] += 2.0 * Li;
    acc2[k] += -2.0 * Lr;
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
