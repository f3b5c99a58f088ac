// o2d_device.cuh -- device-side building blocks shared by the pair-sum kernels (sm_100a only).
//
//  * MUFU wrappers (rcp/rsqrt/ex2/lg2 .approx.ftz): one SFU instruction each, <= 2^-22 relative
//    error, two orders of magnitude inside the 1e-5 parity tolerance.
//  * Blackwell packed-FP32 helpers: FADD2/FMUL2/FFMA2 process two independent float lanes per
//    issue slot.  A lane pair here is always TWO SOURCES against ONE target.
//  * TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers used to stage source
//    tiles into shared memory.
//  * The five core functions of reference src/CoreFunc.h, written on packed lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "o2d_internal.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "omega2d_b200 kernels are written for sm_100a only"
#endif

namespace o2d {

// ---- SFU -----------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_rcp(float x)   { float r; asm("rcp.approx.ftz.f32 %0, %1;"   : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x)   { float r; asm("ex2.approx.ftz.f32 %0, %1;"   : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x)   { float r; asm("lg2.approx.ftz.f32 %0, %1;"   : "=f"(r) : "f"(x)); return r; }

// ---- packed FP32 (FADD2 / FMUL2 / FFMA2) ------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a)          { return make_float2(a, a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b)           { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b)           { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 rcp2(float2 a)   { return make_float2(mufu_rcp(a.x), mufu_rcp(a.y)); }
__device__ __forceinline__ float2 rsqrt2(float2 a) { return make_float2(mufu_rsqrt(a.x), mufu_rsqrt(a.y)); }

// ---- mbarrier + TMA bulk copy ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy through the TMA unit; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- constant-strength panel pair (shared by o2d_panels.cu and o2d_bem.cu) ----------------------------
// atan2(c, d) on the FMA pipe + one MUFU.RCP (libm's atan2f is a ~60-instruction branchy sequence).
// Octant reduction to |x| <= tan(pi/8) (x = min/max, or (min-max)/(min+max) with a pi/4 offset), then
// the Cephes atanf minimax polynomial; measured <= 2.3e-7 relative to float64 atan2 over all
// quadrants, the same as libm's float result, and relative (not absolute) for small angles.
__device__ __forceinline__ float atan2_fast(float c, float d) {
  const float ax = fabsf(d), ay = fabsf(c);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const bool big = mn > 0.41421356f * mx;
  const float nn = big ? mn - mx : mn;
  const float dd = big ? mn + mx : mx;
  const float x = nn * mufu_rcp(dd);
  const float z = x * x;
  float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
  p = fmaf(p, z, 1.99777106478e-1f);
  p = fmaf(p, z, -3.33329491539e-1f);
  p = fmaf(p * z, x, x);
  float r = big ? p + 0.78539816339744831f : p;
  r = ay > ax ? 1.57079632679489662f - r : r;
  r = d < 0.0f ? 3.14159265358979324f - r : r;
  r = mx == 0.0f ? 0.0f : r;                    // atan2(0,0) = 0, as std::atan2
  return copysignf(r, c);
}

// ustar, vstar of one (panel, point) pair.  vstar = 0.5 log(rij2/rij12) = atanh(z) with
// z = (rij2 - rij12)/(rij2 + rij12): the numerator is formed without cancellation from the panel
// vector, and for |z| < 0.2 (point further than ~2.5 panel lengths away: almost every pair) the odd
// series z + z^3/3 + ... + z^9/9 is exact to 6e-8 relative.  Closer in, one MUFU.LG2 of the ratio is accurate
// enough in relative terms because |vstar| >= 0.2; that path is skipped warp-wide when no lane needs it.
__device__ __forceinline__ void panel_pair(float x0, float y0, float px, float py, float tx, float ty,
                                           float& ustar, float& vstar) {
  const float dx0 = tx - x0;
  const float dy0 = ty - y0;
  const float dx1 = dx0 - px;
  const float dy1 = dy0 - py;
  const float rij2 = fmaf(dx0, dx0, dy0 * dy0);
  const float rij12 = fmaf(dx1, dx1, dy1 * dy1);
  const float cross = fmaf(py, dx0, -(px * dy0));
  const float dot = fmaf(dx0, dx1, dy0 * dy1);
  ustar = atan2_fast(cross, dot);
  // a target exactly ON the panel, between its nodes: the angle is +-pi and the reference's difference of two atan2
  // (src/Kernels.h:262-268, arguments (dx, dy)) lands on -pi when the panel points to +x (or straight down), +pi otherwise
  if (cross == 0.0f && dot < 0.0f) ustar = (px > 0.0f || (px == 0.0f && py < 0.0f)) ? -3.14159265358979f : 3.14159265358979f;
  const float num = fmaf(px, dx0 + dx1, py * (dy0 + dy1));   // = rij2 - rij12
  const float z = num * mufu_rcp(rij2 + rij12);
  const float w = z * z;
  float q = fmaf(1.0f / 9.0f, w, 1.0f / 7.0f);
  q = fmaf(q, w, 0.2f);
  q = fmaf(q, w, 1.0f / 3.0f);
  vstar = fmaf(z * w, q, z);
  const bool near = !(fabsf(z) < 0.2f);          // also catches NaN (point on a node)
  if (__any_sync(__activemask(), near)) {
    // 0.5 ln 2 * lg2 of the RATIO: the difference of two lg2 of small squared distances (lg2 ~ -17 for millimetre panels)
    // would carry the absolute rounding of each, 1.6e-6 of vstar at |z| = 0.2
    const float lg = 0.34657359027997264f * mufu_lg2(rij2 * mufu_rcp(rij12));
    vstar = near ? lg : vstar;
  }
}

// ---- far-field panel pair on packed lanes ---------------------------------------------------------------
// With d0 = t - node0, d1 = t - node1 (complex numbers x + i y) the reference's two outputs are the real and
// imaginary part of ONE complex logarithm:  vstar + i ustar = log(d0/d1)  (ustar = arg d0 - arg d1 is
// get_ustar's wrapped atan2 difference, src/Kernels.h:262-268; vstar = 0.5 log(rij2/rij12), :338).  With
// e = t - centre = d0 - p/2 and  zeta = (d0-d1)/(d0+d1) = (p/2)/e:   log(d0/d1) = 2 atanh(zeta)
//   = 2 zeta (1 + zeta^2/3 + zeta^4/5 + zeta^6/7 + ...),
// a series in the complex zeta that needs no arctangent, no logarithm and no range reduction (the general panel_pair
// above is ~55 instructions + 2-4 MUFU), and it has no cancellation (<= 4.5e-7 relative to float64; the reference's own
// float expression is off by up to 4e-4 relative for short panels).  Truncated after zeta^6/7 the relative error is
// |zeta|^8/9 <= 4.3e-8 for |zeta|^2 < kFarZeta2, i.e. targets further than 3.16 panel lengths from the panel centre --
// all but a handful of pairs; the callers fall back to panel_pair() for the rest (warp-uniform test).
// Sign convention kept from round 1's per-pair zeta evaluation: H = -p/2 (or e = centre - t and H = +p/2), so the series
// gives  (rr, ri) = (-vstar/2, -ustar/2).
constexpr float kFarZeta2 = 0.025f;
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }   // folds into FFMA2 operand modifiers

// ---- far field as a Laurent series with per-panel coefficients (round 2) -----------------------------------
// The strengths and the rotation by the panel tangent are linear, and H is a per-panel constant, so everything but
// 1/e folds into coefficients:  with g = conj(1/e) = e/|e|^2 and G = g^2
//   conj(atanh(H/e)) = g (E1 + E3 G + E5 G^2 + E7 G^3),   E_k = conj(H)^k / k ,
// and for panels acting on a point  u + i v += g (D1 + D3 G + D5 G^2 + D7 G^3),  D_k = m E_k  with m = i (a + i b) the
// strength-weighted tangent (pack_panels_kernel) -- no per-pair zeta, no per-pair rotation: e (4 lane-ops, the exact
// (t - node0) + H), |e|^2 (2), g (2), G (4), one complex Horner level (4) per term after the first, and the product with g
// accumulated by four FMAs: 20 / 24 / 28 lane-ops + one MUFU.RCP for 2 / 3 / 4 terms (round 1's zeta form: 34 for 4).
// Series length per warp vote, relative truncation error <= 4.3e-8 in each tier:
//   |zeta|^2 < kFarZeta2 (3.2 panel lengths)   4 terms     next term |zeta|^8/9
//   |zeta|^2 < kVeryFarZeta2 (6.2 lengths)     3 terms     |zeta|^6/7
//   |zeta|^2 < kUltraFarZeta2 (23 lengths)     2 terms     |zeta|^4/5
constexpr float kVeryFarZeta2 = 0.0066f;
constexpr float kUltraFarZeta2 = 0.00046f;

__device__ __forceinline__ void panel_far_g(float2 ex, float2 ey, float2& gx, float2& gy, float2& Gx, float2& Gy, float2& e2) {
  e2 = fma2(ex, ex, mul2(ey, ey));
  const float2 r = rcp2(e2);
  gx = mul2(ex, r); gy = mul2(ey, r);
  Gx = fma2(gx, gx, neg2(mul2(gy, gy)));
  Gy = mul2(add2(gx, gx), gy);
}
// one Horner level in the complex G:  t <- t G + D
__device__ __forceinline__ void chorner(float2& tr, float2& ti, float2 Gx, float2 Gy, float2 Dx, float2 Dy) {
  const float2 nr = fma2(tr, Gx, fma2(neg2(ti), Gy, Dx));
  const float2 ni = fma2(tr, Gy, fma2(ti, Gx, Dy));
  tr = nr; ti = ni;
}
// E_k = conj(H)^k / k, k = 1,3,5,7, times the complex m, in float64 (pack time / once per thread)
__device__ __forceinline__ void panel_far_coeffs(double hx, double hy, double mx, double my, float* dx, float* dy) {
  const double cx = hx, cy = -hy;                        // conj(H)
  const double c2x = cx * cx - cy * cy, c2y = 2.0 * cx * cy;
  double px = mx * cx - my * cy, py = mx * cy + my * cx; // m conj(H)^k
  const double inv[4] = {1.0, 1.0 / 3.0, 0.2, 1.0 / 7.0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    dx[k] = (float)(px * inv[k]); dy[k] = (float)(py * inv[k]);
    const double qx = px * c2x - py * c2y, qy = px * c2y + py * c2x;
    px = qx; py = qy;
  }
}

// ---- radius terms ----------------------------------------------------------------------------
// Per-element radius term the cores add to the distance term: r^2 for RM/EXP/WL, r^4 for V2,
// r^6 for V3, formed with the reference's own float products (s2 = sr*sr; s2*s2; s2*s2*s2,
// reference src/CoreFunc.h:48,147,199,247-249,292-294).  Inert targets use 0, which makes the
// blob-target ("0b") expression equal the inert-target ("0p") one.
template <int CORE>
__host__ __device__ __forceinline__ float radius_term(float r) {
  const float r2 = r * r;
  if (CORE == CORE_V2) return r2 * r2;
  if (CORE == CORE_V3) return r2 * r2 * r2;
  return r2;
}

// ---- compact-exponential core, one lane -------------------------------------------------------
// reference src/CoreFunc.h:118-130 (exp_cond), :144-150, :170-179:
//   ood2 = 1/d2, corefac = 1/(sr^2+tr^2), x = reld2 = corefac/ood2
//   K   = ood2 (x>16) | corefac (x<0.001) | ood2 (1-exp(-x))
//   bbb = -2 K ood2 + 2 corefac ood2 exp(-x)                     (NaN at d2 == 0, as the reference)
// The reference forms 1-exp(-x) and the bbb difference in float with heavy cancellation for
// x << 1 (its own result carries ~6e-8/x relative noise there).  Here both are evaluated in
// cancellation-free form so the only disagreement left is the reference's own rounding:
//   1-exp(-x) = x P(x)                         (Taylor, x < 0.5)
//   bbb = 2 corefac ood2 g(x),  g(x) = exp(-x) - (1-exp(-x))/x = sum_n (-1)^n n x^n/(n+1)!   (0.001<=x<0.5)
// while keeping the reference's three branches (including its x<0.001 branch, where K is
// replaced by corefac inside the bbb expression too).
template <bool GRAD>
__device__ __forceinline__ void exp_core_lane(float d2, float sig2, float& K, float& B) {
  const float ood2 = mufu_rcp(d2);
  const float corefac = mufu_rcp(sig2);
  const float x = d2 * corefac;
  const float e = mufu_ex2(-1.4426950408889634f * x);
  float em1;   // 1 - exp(-x)
  if (x < 0.5f) {
    float p = -1.0f / 40320.0f;
    p = fmaf(p, x, 1.0f / 5040.0f);
    p = fmaf(p, x, -1.0f / 720.0f);
    p = fmaf(p, x, 1.0f / 120.0f);
    p = fmaf(p, x, -1.0f / 24.0f);
    p = fmaf(p, x, 1.0f / 6.0f);
    p = fmaf(p, x, -0.5f);
    p = fmaf(p, x, 1.0f);
    em1 = x * p;
  } else {
    em1 = 1.0f - e;
  }
  K = x > 16.0f ? ood2 : (x < 0.001f ? corefac : ood2 * em1);
  if (GRAD) {
    float g;   // bbb = 2 ood2 * g
    if (x > 16.0f) {
      g = fmaf(corefac, e, -ood2);
    } else if (x < 0.001f) {
      g = -corefac * em1;
    } else if (x < 0.5f) {
      float q = 1.0f / 45360.0f;
      q = fmaf(q, x, -1.0f / 5760.0f);
      q = fmaf(q, x, 1.0f / 840.0f);
      q = fmaf(q, x, -1.0f / 144.0f);
      q = fmaf(q, x, 1.0f / 30.0f);
      q = fmaf(q, x, -0.125f);
      q = fmaf(q, x, 1.0f / 3.0f);
      q = fmaf(q, x, -0.5f);
      g = corefac * (x * q);
    } else {
      g = fmaf(corefac, e, -K);
    }
    B = 2.0f * ood2 * g;
  }
}

// The same function for the velocity-only kernels, both lanes at once and without branches: when one pair of a
// warp's 64 is close enough to need the exponential, all of them run this, so its cost -- not its frequency
// (3.6 % of the warp steps of a 1M-particle blob) -- is what the kernel pays for.  6 MUFU + 14 packed FP32 ops.
__device__ __forceinline__ float2 exp_core_pair(float2 d2, float2 sig2) {
  const float2 ood2 = rcp2(d2);
  const float2 corefac = rcp2(sig2);
  const float2 x = mul2(d2, corefac);
  const float2 t = mul2(f2(-1.4426950408889634f), x);
  const float2 e = make_float2(mufu_ex2(t.x), mufu_ex2(t.y));
  float2 p = fma2(f2(-1.0f / 40320.0f), x, f2(1.0f / 5040.0f));     // 1-exp(-x) = x P(x), x < 0.5 (as exp_core_lane)
  p = fma2(p, x, f2(-1.0f / 720.0f));
  p = fma2(p, x, f2(1.0f / 120.0f));
  p = fma2(p, x, f2(-1.0f / 24.0f));
  p = fma2(p, x, f2(1.0f / 6.0f));
  p = fma2(p, x, f2(-0.5f));
  p = fma2(p, x, f2(1.0f));
  const float2 ser = mul2(x, p);
  const float2 dir = add2(f2(1.0f), neg2(e));
  const float2 em1 = make_float2(x.x < 0.5f ? ser.x : dir.x, x.y < 0.5f ? ser.y : dir.y);
  const float2 mid = mul2(ood2, em1);
  return make_float2(x.x > 16.0f ? ood2.x : (x.x < 0.001f ? corefac.x : mid.x),
                     x.y > 16.0f ? ood2.y : (x.y < 0.001f ? corefac.y : mid.y));
}

// ---- core functions on packed lanes ----------------------------------------------------------
// Inputs per lane: dx,dy = target - source, rho = source radius term, tau = target radius term.
// Outputs: K (the reference's core_func return / *r2), d2 = dx^2+dy^2 (only valid when GRAD or
// when the core needs it), and, when GRAD, B (the reference's *bbb).
// Strength multiplication is left to the caller.
// UNI: every target of the thread has the same radius term (inert targets, tau = 0, or the usual
// uniform-radius particle set); `rho` then already holds rho + tau, formed once per source pair and
// shared by the thread's R targets, which removes one FP32 op per pair (RM velocity: 8 -> 7.25
// lane-ops, moving the kernel off the FMA pipe and onto the 16/clk/SM MUFU bound alone).
//
// Deferred gradient factor (kBbDeferred): for RM and V2 the reference's bbb is -2 K times something cheap
// (RM: bbb = -2 K K;  V2: bbb = -2 K (K^2 d2)), so those cores return B = bbb / (-2 K) and the caller forms
// bb = (ss K) B from the r2 = ss K it needs anyway; the factor -2 is applied once per tile when the float partial
// sums are folded into float64 (pp_kernel).  Saves two FP32 ops per pair on the gradient-type results.
template <int CORE> struct CoreTraits { static constexpr bool kBbDeferred = (CORE == CORE_RM || CORE == CORE_V2); };

// NEED_D2 = false: the caller does not read d2 (velocity + gradient accumulates bb dx dy, bb dx^2, bb dy^2 only),
// which lets RM fold the squares straight into q.
template <int CORE, bool GRAD, bool UNI, bool NEED_D2 = true>
__device__ __forceinline__ void core_eval(float2 dx, float2 dy, float2 rho, float2 tau,
                                          float2& K, float2& B, float2& d2, float lim = 0.0f) {
  if (CORE == CORE_RM) {
    // src/CoreFunc.h:54-58,74-80: K = 1/(d2 + sr^2 + tr^2), bbb = -2 K^2
    float2 q;
    if (GRAD && NEED_D2) {
      d2 = fma2(dy, dy, mul2(dx, dx));
      q = UNI ? add2(d2, rho) : add2(add2(d2, tau), rho);
    } else if (UNI) {
      q = fma2(dy, dy, fma2(dx, dx, rho));               // 2 FP32 ops
    } else {
      q = add2(fma2(dy, dy, fma2(dx, dx, tau)), rho);   // 3 FP32 ops, no separate d2
    }
    K = rcp2(q);
    if (GRAD) B = K;                                     // deferred: bbb / (-2 K)
  } else if (CORE == CORE_V2) {
    // src/CoreFunc.h:245-250,266-273: K = rsqrt(d2^2 + sr^4 + tr^4), bbb = -2 K^3 d2
    d2 = fma2(dy, dy, mul2(dx, dx));
    const float2 q = UNI ? fma2(d2, d2, rho) : add2(fma2(d2, d2, tau), rho);
    K = rsqrt2(q);
    if (GRAD) B = mul2(mul2(K, K), d2);                  // deferred: bbb / (-2 K)
  } else if (CORE == CORE_V3) {
    // src/CoreFunc.h:290-295,312-320: K = cbrt(d2^3 + sr^6 + tr^6)^-1, bbb = -2 d2^2 K / (d2^3+..)
    d2 = fma2(dy, dy, mul2(dx, dx));
    const float2 q = UNI ? fma2(mul2(d2, d2), d2, rho) : add2(fma2(mul2(d2, d2), d2, tau), rho);
    // q^(-1/3) = 2^(-lg2(q)/3) on the SFU.  The float result of MUFU.LG2 is rounded at |lg2 q| <= 64, so the
    // seed carries up to ~4e-7 relative error for the closest pairs (1e-7 typical): that is what the
    // velocity-only kernel uses (random in sign; the sums agree with the reference to ~1e-7).  The
    // gradient-type results also feed K into B, so they take one Newton step  k <- k + k (1 - q k^3)/3
    // on the FMA pipe.
    const float2 k = make_float2(mufu_ex2(-0.33333334f * mufu_lg2(q.x)), mufu_ex2(-0.33333334f * mufu_lg2(q.y)));
    if (GRAD) {
      const float2 k3 = mul2(mul2(k, k), k);
      const float2 e = fma2(neg2(q), k3, f2(1.0f));
      K = fma2(mul2(k, f2(0.33333334f)), e, k);
    } else {
      K = k;
    }
    // q == 0 (coincident zero-radius pair) gives inf*0 above; the reference returns inf there
    if (q.x == 0.0f) K.x = __int_as_float(0x7f800000);
    if (q.y == 0.0f) K.y = __int_as_float(0x7f800000);
    if (GRAD) B = mul2(mul2(mul2(f2(-2.0f), mul2(d2, d2)), K), rcp2(q));
  } else if (CORE == CORE_WL) {
    // src/CoreFunc.h:197-202,220-228: s2 = sr^2+tr^2, D = d2+s2, K = (D+s2)/D^2 = r + s2 r^2
    // with r = 1/D;  bbb = (2 - 4 K D)/D^2 = (2 - 4 K D) r^2
    const float2 s2 = UNI ? rho : add2(rho, tau);
    if (GRAD) d2 = fma2(dy, dy, mul2(dx, dx));
    const float2 D = GRAD ? add2(d2, s2) : fma2(dy, dy, fma2(dx, dx, s2));
    const float2 r = rcp2(D);
    K = fma2(mul2(s2, r), r, r);
    if (GRAD) B = mul2(fma2(mul2(f2(-4.0f), K), D, f2(2.0f)), mul2(r, r));
  } else {  // CORE_EXP
    d2 = fma2(dy, dy, mul2(dx, dx));
    // The reference's first branch (src/CoreFunc.h:118-130: d2/sig2 > 16 -> K = 1/d2) is what almost every
    // pair of a large particle set takes, and it needs one MUFU instead of three.  Taken warp-wide (the
    // callers are convergent: every thread of the CTA walks the same source pairs) on a test that costs no
    // arithmetic: `lim` = 16 x (largest source radius term of the staged tile + this target's), so d2 > lim
    // implies d2/sig2 > 16 for the pair; pairs between their own limit and the tile's take the general path,
    // which evaluates the reference's branches per lane.  The gradient factor keeps a corefac*exp(-x) term in
    // the far branch (x exp(-x) = 1.8e-6 at x = 16): it is dropped only beyond x = 24 (lim is 24 x ... there),
    // where it is below half an ulp of the 1/d2 term.
    const bool far = (d2.x > lim) && (d2.y > lim);
    if (__all_sync(0xffffffffu, far)) {
      K = rcp2(d2);
      if (GRAD) B = mul2(mul2(f2(-2.0f), K), K);
    } else if (!GRAD) {
      K = exp_core_pair(d2, UNI ? rho : add2(rho, tau));
    } else {
      const float2 sig2 = UNI ? rho : add2(rho, tau);
      exp_core_lane<GRAD>(d2.x, sig2.x, K.x, B.x);
      exp_core_lane<GRAD>(d2.y, sig2.y, K.y, B.y);
    }
  }
}

}  // namespace o2d
