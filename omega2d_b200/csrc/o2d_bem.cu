// o2d_bem.cu -- BEM influence-coefficient assembly.
//
// Replaces the O(Np^2) loop nest of panels_on_panels_coeff<S> (reference src/Coefficients.h:155-427):
// for every source panel j (columns) and target panel i (rows)
//   1. velocity at the centre of panel i induced by panel j carrying unit vortex (and unit source)
//      sheet strength: kernelu_1vos_0p / kernelu_1v_0p (src/Kernels.h:413-457, :312-354);
//   2. the same with the roles swapped (panel i on the centre of panel j), scaled by
//      area_j/area_i -- the "use_two_way" average (src/Coefficients.h:353-364, :380-387);
//   3. projection on panel i's tangent (and normal), times 0.5, times 1/(2 pi);
//   4. self influence (source surface IS the target surface, i == j): pi -> 0.5 after scaling
//      (src/Coefficients.h:392-404).
// Output is the reference's column-major block: column 2j = vortex strength of panel j, column
// 2j+1 = its source strength, rows 2i / 2i+1 = tangential / normal condition on panel i (one row and
// one column per panel when only vortex strengths are unknown).  The caller picks the leading
// dimension, so the block can be written straight into an augmented matrix.
//
// Mapping: one thread per target panel (rows are contiguous in memory -> coalesced 8-byte stores),
// source panels staged 64 at a time in shared memory and broadcast.  Each pair costs two panel_pair()
// evaluations (the Laurent far-field series of o2d_device.cuh, two terms for nearly every pair; panel_pair() close in)
// and writes 16 (or 4) bytes: for Np = 20k that is 4e8 pairs, 6.4 GB of coefficients -- the store stream is the bound.
#include "o2d_context.h"
#include "o2d_device.cuh"

using namespace o2d;

namespace {

constexpr int kCoefThreads = 128;
constexpr int kCoefSrcTile = 64;

// per panel: {x0, y0, Hx, Hy} {cx, cy, area, thr} {D1x, D1y, D3x, D3y} {D5x, D5y, D7x, D7y} {Ux, Uy, thr2, thr3}
//   H = -p/2; thr, thr2, thr3 = |H|^2 / kFarZeta2, / kVeryFarZeta2, / kUltraFarZeta2 (4, 3, 2 series terms; inside thr the
//   general panel_pair); D_k = m conj(H)^k / k with m = i (Ux + i Uy), U = -2 p/|p|: the Laurent coefficients of the
//   velocity of this panel carrying UNIT vortex sheet strength (o2d_device.cuh, "far field as a Laurent series"); U itself
//   is what the near form multiplies its (-ustar/2, -vstar/2) with
constexpr int kCoefRec = 5;
__global__ void __launch_bounds__(256)
pack_coeff_panels_kernel(int np, const float* __restrict__ nx, const float* __restrict__ ny,
                         const uint32_t* __restrict__ idx, const float* __restrict__ area,
                         float4* __restrict__ packed) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  const uint32_t i0 = idx[2 * j], i1 = idx[2 * j + 1];
  const float x0 = nx[i0], y0 = ny[i0], x1 = nx[i1], y1 = ny[i1];
  const float px = x1 - x0, py = y1 - y0;
  const float hx = -0.5f * px, hy = -0.5f * py;
  const double rlen = 1.0 / sqrt((double)px * px + (double)py * py);
  const double ux = -2.0 * (px * rlen), uy = -2.0 * (py * rlen);
  float dx[4], dy[4];
  panel_far_coeffs(hx, hy, -uy, ux, dx, dy);
  const float thr = (hx * hx + hy * hy) * (1.0f / kFarZeta2);
  packed[kCoefRec * j] = make_float4(x0, y0, hx, hy);
  // collocation point as the reference forms it: 0.5*(x1+x0) (src/Coefficients.h:327-328,356)
  packed[kCoefRec * j + 1] = make_float4(0.5f * (x1 + x0), 0.5f * (y1 + y0), area[j], thr);
  packed[kCoefRec * j + 2] = make_float4(dx[0], dy[0], dx[1], dy[1]);
  packed[kCoefRec * j + 3] = make_float4(dx[2], dy[2], dx[3], dy[3]);
  packed[kCoefRec * j + 4] = make_float4((float)ux, (float)uy, thr * (kFarZeta2 / kVeryFarZeta2), thr * (kFarZeta2 / kUltraFarZeta2));
}

struct CoefArgs {
  const float4* spack;
  const float4* tpack;
  const float* ttx; const float* tty; const float* tnx; const float* tny;
  int nsrc, ntarg;
  int same;            // source surface is the target surface
  size_t ld;           // leading dimension (rows) of the column-major output
  float* coeffs;
};

// One (target panel i, source panel j) pair needs TWO unit-strength panel velocities -- panel j on the centre of
// panel i and, for the two-way average, panel i on the centre of panel j -- and they are the two lanes of the
// packed arithmetic: lane x = (vu, vv) of j on i, lane y = (wu, wv) of i on j.  The unit source velocity is
// (vv, -vu) (src/Kernels.h:447-456 with vs = ss = 1).
template <bool BOTH>
__global__ void __launch_bounds__(kCoefThreads)
coeff_kernel(const CoefArgs a) {
  __shared__ float4 s[kCoefRec * kCoefSrcTile];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kCoefThreads + tid;
  const int j0 = blockIdx.y * kCoefSrcTile;
  const int nj = min(kCoefSrcTile, a.nsrc - j0);
  for (int k = tid; k < kCoefRec * nj; k += kCoefThreads) s[k] = a.spack[kCoefRec * (size_t)j0 + k];
  __syncthreads();
  // threads past the last target keep computing on a copy of it (the warp votes in the loop need every lane)
  const int ii = min(i, a.ntarg - 1);
  const float4 T = a.tpack[kCoefRec * (size_t)ii];
  const float4 Tc = a.tpack[kCoefRec * (size_t)ii + 1];
  const float4 Td = a.tpack[kCoefRec * (size_t)ii + 2];
  const float4 Te = a.tpack[kCoefRec * (size_t)ii + 3];
  const float4 Tu = a.tpack[kCoefRec * (size_t)ii + 4];
  const float ttx = a.ttx[ii], tty = a.tty[ii];
  const float2 T1 = f2(ttx, BOTH ? a.tnx[ii] : 0.0f);     // {tangent.x, normal.x}
  const float2 T2 = f2(tty, BOTH ? a.tny[ii] : 0.0f);     // {tangent.y, normal.y}
  const float rta = 1.0f / Tc.z;
  const float fac = (float)(1.0 / (2.0 * 3.14159265358979323846));   // "const S fac = 1.0/(2.0*M_PI)"
  const float hfac = 0.5f * fac;                                     // exact scaling: (0.5 c) fac == c (0.5 fac)
  const float pifac = 3.14159265358979323846f * fac;
  for (int jj = 0; jj < nj; ++jj) {
    const int j = j0 + jj;
    const float4 S = s[kCoefRec * jj];
    const float4 Sc = s[kCoefRec * jj + 1];
    const float4 Sd = s[kCoefRec * jj + 2];
    const float4 Su = s[kCoefRec * jj + 4];
    const float2 Hx = f2(S.z, T.z), Hy = f2(S.w, T.w);
    const float2 ex = add2(add2(f2(Tc.x, Sc.x), neg2(f2(S.x, T.x))), Hx);
    const float2 ey = add2(add2(f2(Tc.y, Sc.y), neg2(f2(S.y, T.y))), Hy);
    float2 gx, gy, Gx, Gy, e2;
    panel_far_g(ex, ey, gx, gy, Gx, Gy, e2);
    float2 tr, ti;
    bool near = false;
    if (__all_sync(0xffffffffu, e2.x > Su.w && e2.y > Tu.w)) {
      // both directions beyond 23 panel lengths (nearly every pair of a large body): two terms, no further tests
      tr = fma2(f2(Sd.z, Td.z), Gx, fma2(neg2(f2(Sd.w, Td.w)), Gy, f2(Sd.x, Td.x)));
      ti = fma2(f2(Sd.z, Td.z), Gy, fma2(f2(Sd.w, Td.w), Gx, f2(Sd.y, Td.y)));
    } else {
      const float4 Se = s[kCoefRec * jj + 3];
      if (__all_sync(0xffffffffu, e2.x > Su.z && e2.y > Tu.z)) {
        tr = f2(Se.x, Te.x); ti = f2(Se.y, Te.y);
      } else {
        tr = f2(Se.z, Te.z); ti = f2(Se.w, Te.w);
        chorner(tr, ti, Gx, Gy, f2(Se.x, Te.x), f2(Se.y, Te.y));
      }
      chorner(tr, ti, Gx, Gy, f2(Sd.z, Td.z), f2(Sd.w, Td.w));
      chorner(tr, ti, Gx, Gy, f2(Sd.x, Td.x), f2(Sd.y, Td.y));
      near = true;
    }
    float2 vu2 = fma2(tr, gx, neg2(mul2(ti, gy)));   // {vu, wu}: velocity of the unit vortex sheet = t g
    float2 vv2 = fma2(tr, gy, mul2(ti, gx));         // {vv, wv}
    if (near) {
      const bool na = !(e2.x > Sc.w), nb = !(e2.y > Tc.w);
      if (__any_sync(0xffffffffu, na || nb)) {
        float us, vs;
        panel_pair(S.x, S.y, -2.0f * S.z, -2.0f * S.w, Tc.x, Tc.y, us, vs);
        if (na) { vu2.x = -0.5f * (us * Su.x - vs * Su.y); vv2.x = -0.5f * (us * Su.y + vs * Su.x); }
        panel_pair(T.x, T.y, -2.0f * T.z, -2.0f * T.w, Sc.x, Sc.y, us, vs);
        if (nb) { vu2.y = -0.5f * (us * Tu.x - vs * Tu.y); vv2.y = -0.5f * (us * Tu.y + vs * Tu.x); }
      }
    }
    const float fact = Sc.z * rta;                         // sa[j]/ta[i]
    const bool self = a.same && (i == j);
    if (i >= a.ntarg) continue;
    if (BOTH) {
      // {c1e1, c1e2} = vortex column on {tangent, normal}; {c2e1, c2e2} = source column
      float2 c1 = fma2(f2(vu2.x), T1, mul2(f2(vv2.x), T2));
      float2 c2 = fma2(f2(vv2.x), T1, neg2(mul2(f2(vu2.x), T2)));
      const float2 w1 = fma2(f2(vu2.y), T1, mul2(f2(vv2.y), T2));
      const float2 w2 = fma2(f2(vv2.y), T1, neg2(mul2(f2(vu2.y), T2)));
      c1 = mul2(fma2(f2(-fact), w1, c1), f2(hfac));
      c2 = mul2(fma2(f2(-fact), w2, c2), f2(hfac));
      if (self) { c1 = f2(pifac, 0.0f); c2 = f2(0.0f, pifac); }
      float2* col1 = reinterpret_cast<float2*>(a.coeffs + (size_t)(2 * j) * a.ld) + i;
      float2* col2 = reinterpret_cast<float2*>(a.coeffs + (size_t)(2 * j + 1) * a.ld) + i;
      *col1 = c1;
      *col2 = c2;
    } else {
      const float2 t2 = fma2(vu2, f2(ttx), mul2(vv2, f2(tty)));
      float c = hfac * fmaf(-fact, t2.y, t2.x);
      if (self) c = pifac;
      a.coeffs[(size_t)j * a.ld + i] = c;
    }
  }
}

}  // namespace

static int coeff_device(o2d_context* c, int nsrc, const float* snx, const float* sny, const uint32_t* sidx,
                        const float* sarea, int ntarg, const float* tnx_, const float* tny_, const uint32_t* tidx,
                        const float* tarea, const float* ttx, const float* tty, const float* tnx, const float* tny,
                        int both, int same, size_t ld, float* coeffs) {
  CU(c->panel_packed.reserve(((size_t)nsrc + (size_t)ntarg) * (16 * kCoefRec)));
  float4* spack = c->panel_packed.as<float4>();
  float4* tpack = spack + kCoefRec * (size_t)nsrc;
  pack_coeff_panels_kernel<<<(nsrc + 255) / 256, 256, 0, c->stream>>>(nsrc, snx, sny, sidx, sarea, spack);
  CU(cudaGetLastError());
  pack_coeff_panels_kernel<<<(ntarg + 255) / 256, 256, 0, c->stream>>>(ntarg, tnx_, tny_, tidx, tarea, tpack);
  CU(cudaGetLastError());
  c->launches += 2;
  CoefArgs a;
  a.spack = spack; a.tpack = tpack;
  a.ttx = ttx; a.tty = tty; a.tnx = tnx; a.tny = tny;
  a.nsrc = nsrc; a.ntarg = ntarg; a.same = same; a.ld = ld; a.coeffs = coeffs;
  dim3 grid((ntarg + kCoefThreads - 1) / kCoefThreads, (nsrc + kCoefSrcTile - 1) / kCoefSrcTile, 1);
  {
    PairTimer timer(c);
    if (both) coeff_kernel<true><<<grid, kCoefThreads, 0, c->stream>>>(a);
    else      coeff_kernel<false><<<grid, kCoefThreads, 0, c->stream>>>(a);
    CU(cudaGetLastError());
    c->launches++;
    c->pair_launches++;
  }
  return O2D_OK;
}

extern "C" int o2d_panels_on_panels_coeff(o2d_context* c, int mem,
                                          int nsrc, int nsnodes, const float* snx, const float* sny,
                                          const uint32_t* sidx, const float* sarea,
                                          int ntarg, int ntnodes, const float* tnx_, const float* tny_,
                                          const uint32_t* tidx, const float* tarea,
                                          const float* ttx, const float* tty, const float* tnx, const float* tny,
                                          int both_strengths, int same_surface, long long ld, float* coeffs) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (nsrc < 0 || ntarg < 0 || nsnodes < 0 || ntnodes < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (nsrc == 0 || ntarg == 0) return O2D_OK;
  const int unk = both_strengths ? 2 : 1;
  if (ld < (long long)unk * ntarg) return fail(O2D_ERR_INVALID, "leading dimension %lld < %d rows", ld, unk * ntarg);
  if (!snx || !sny || !sidx || !sarea || !tnx_ || !tny_ || !tidx || !tarea || !ttx || !tty || !coeffs)
    return fail(O2D_ERR_INVALID, "NULL array");
  if (both_strengths && (!tnx || !tny)) return fail(O2D_ERR_INVALID, "panel normals are required when both strengths are unknown");
  if (mem == O2D_MEM_DEVICE) {
    if (both_strengths && ((ld & 1) || (reinterpret_cast<uintptr_t>(coeffs) & 7)))
      return fail(O2D_ERR_INVALID, "device-memory coeffs need an even leading dimension and 8-byte alignment");
    return coeff_device(c, nsrc, snx, sny, sidx, sarea, ntarg, tnx_, tny_, tidx, tarea, ttx, tty, tnx, tny,
                        both_strengths, same_surface, (size_t)ld, coeffs);
  }
  for (int k = 0; k < 2 * nsrc; ++k)
    if (sidx[k] >= (uint32_t)nsnodes) return fail(O2D_ERR_INVALID, "source panel node index %u out of range", sidx[k]);
  for (int k = 0; k < 2 * ntarg; ++k)
    if (tidx[k] >= (uint32_t)ntnodes) return fail(O2D_ERR_INVALID, "target panel node index %u out of range", tidx[k]);
  // host arrays: mirror geometry, assemble into a device matrix, copy the block back column by column
  CU(c->hsrc.reserve(2 * padded_floats(nsnodes) + padded_floats(nsrc) + 2 * padded_floats(ntnodes) + 5 * padded_floats(ntarg)));
  CU(c->hidx.reserve(((size_t)2 * nsrc + (size_t)2 * ntarg) * sizeof(uint32_t)));
  const size_t rows = (size_t)unk * ntarg, cols = (size_t)unk * nsrc;
  const size_t dld = (rows + 1) & ~(size_t)1;   // device block: even leading dimension
  CU(c->coef.reserve(dld * cols * sizeof(float)));
  Stager S(c, &c->hsrc);
  float *dsnx = S.take(nsnodes), *dsny = S.take(nsnodes), *dsa = S.take(nsrc);
  float *dtnx = S.take(ntnodes), *dtny = S.take(ntnodes), *dta = S.take(ntarg);
  float *dttx = S.take(ntarg), *dtty = S.take(ntarg), *dnnx = S.take(ntarg), *dnny = S.take(ntarg);
  int rc;
  if ((rc = o2d_h2d(c, dsnx, snx, nsnodes)) || (rc = o2d_h2d(c, dsny, sny, nsnodes)) || (rc = o2d_h2d(c, dsa, sarea, nsrc)) ||
      (rc = o2d_h2d(c, dtnx, tnx_, ntnodes)) || (rc = o2d_h2d(c, dtny, tny_, ntnodes)) || (rc = o2d_h2d(c, dta, tarea, ntarg)) ||
      (rc = o2d_h2d(c, dttx, ttx, ntarg)) || (rc = o2d_h2d(c, dtty, tty, ntarg)))
    return rc;
  if (both_strengths && ((rc = o2d_h2d(c, dnnx, tnx, ntarg)) || (rc = o2d_h2d(c, dnny, tny, ntarg)))) return rc;
  uint32_t* dsi = c->hidx.as<uint32_t>();
  uint32_t* dti = dsi + 2 * (size_t)nsrc;
  if (int rc_ = o2d_h2d_bytes(c, dsi, sidx, (size_t)2 * nsrc * sizeof(uint32_t))) return rc_;
  if (int rc_ = o2d_h2d_bytes(c, dti, tidx, (size_t)2 * ntarg * sizeof(uint32_t))) return rc_;
  if ((rc = coeff_device(c, nsrc, dsnx, dsny, dsi, dsa, ntarg, dtnx, dtny, dti, dta, dttx, dtty, dnnx, dnny,
                         both_strengths, same_surface, dld, c->coef.as<float>())))
    return rc;
  CU(cudaMemcpy2DAsync(coeffs, (size_t)ld * sizeof(float), c->coef.p, dld * sizeof(float), rows * sizeof(float), cols,
                       cudaMemcpyDeviceToHost, c->stream));
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}
