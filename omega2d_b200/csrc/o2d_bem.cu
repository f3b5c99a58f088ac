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
// evaluations (FMA-pipe atan2/atanh, see o2d_device.cuh) and writes 16 (or 4) bytes: for Np = 20k that is
// 4e8 pairs, 6.4 GB of coefficients -- the store stream and the ~130 instructions per pair are of
// the same order, so the kernel is neither purely HBM- nor FP32-bound.
#include "o2d_context.h"
#include "o2d_device.cuh"

using namespace o2d;

namespace {

constexpr int kCoefThreads = 128;
constexpr int kCoefSrcTile = 64;

// per panel: {x0, y0, px, py} {cx, cy, area, 1/len}
__global__ void __launch_bounds__(256)
pack_coeff_panels_kernel(int np, const float* __restrict__ nx, const float* __restrict__ ny,
                         const uint32_t* __restrict__ idx, const float* __restrict__ area,
                         float4* __restrict__ packed) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  const uint32_t i0 = idx[2 * j], i1 = idx[2 * j + 1];
  const float x0 = nx[i0], y0 = ny[i0], x1 = nx[i1], y1 = ny[i1];
  const float px = x1 - x0, py = y1 - y0;
  packed[2 * j] = make_float4(x0, y0, px, py);
  // collocation point as the reference forms it: 0.5*(x1+x0) (src/Coefficients.h:327-328,356)
  packed[2 * j + 1] = make_float4(0.5f * (x1 + x0), 0.5f * (y1 + y0), area[j], 1.0f / sqrtf(px * px + py * py));
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

// unit-strength vortex velocity (vu, vv) of panel {x0,y0,px,py,1/len} at (tx,ty);
// the unit source velocity is (vv, -vu) (src/Kernels.h:447-456 with vs = ss = 1)
__device__ __forceinline__ void unit_panel_vel(const float4 G, float rlen, float tx, float ty, float& vu, float& vv) {
  float ustar, vstar;
  panel_pair(G.x, G.y, G.z, G.w, tx, ty, ustar, vstar);
  const float ux = G.z * rlen, uy = G.w * rlen;
  vu = ustar * ux - vstar * uy;
  vv = ustar * uy + vstar * ux;
}

template <bool BOTH>
__global__ void __launch_bounds__(kCoefThreads)
coeff_kernel(const CoefArgs a) {
  __shared__ float4 s[2 * kCoefSrcTile];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kCoefThreads + tid;
  const int j0 = blockIdx.y * kCoefSrcTile;
  const int nj = min(kCoefSrcTile, a.nsrc - j0);
  for (int k = tid; k < 2 * nj; k += kCoefThreads) s[k] = a.spack[2 * (size_t)j0 + k];
  __syncthreads();
  if (i >= a.ntarg) return;
  const float4 T = a.tpack[2 * (size_t)i];
  const float4 Tc = a.tpack[2 * (size_t)i + 1];
  const float ttx = a.ttx[i], tty = a.tty[i];
  const float tnx = BOTH ? a.tnx[i] : 0.0f, tny = BOTH ? a.tny[i] : 0.0f;
  const float rta = 1.0f / Tc.z;
  const float fac = (float)(1.0 / (2.0 * 3.14159265358979323846));   // "const S fac = 1.0/(2.0*M_PI)"
  for (int jj = 0; jj < nj; ++jj) {
    const int j = j0 + jj;
    const float4 S = s[2 * jj];
    const float4 Sc = s[2 * jj + 1];
    float vu, vv, wu, wv;
    unit_panel_vel(S, Sc.w, Tc.x, Tc.y, vu, vv);      // panel j on the centre of panel i
    unit_panel_vel(T, Tc.w, Sc.x, Sc.y, wu, wv);      // panel i on the centre of panel j ("two way")
    const float fact = Sc.z * rta;                    // sa[j]/ta[i]
    const bool self = a.same && (i == j);
    if (BOTH) {
      float c1e1 = vu * ttx + vv * tty;
      float c1e2 = vu * tnx + vv * tny;
      float c2e1 = vv * ttx - vu * tty;               // source velocity = (vv, -vu)
      float c2e2 = vv * tnx - vu * tny;
      c1e1 = 0.5f * (c1e1 - fact * (wu * ttx + wv * tty));
      c1e2 = 0.5f * (c1e2 - fact * (wu * tnx + wv * tny));
      c2e1 = 0.5f * (c2e1 - fact * (wv * ttx - wu * tty));
      c2e2 = 0.5f * (c2e2 - fact * (wv * tnx - wu * tny));
      if (self) { c1e1 = 3.14159265358979323846f; c1e2 = 0.0f; c2e1 = 0.0f; c2e2 = 3.14159265358979323846f; }
      float2* col1 = reinterpret_cast<float2*>(a.coeffs + (size_t)(2 * j) * a.ld) + i;
      float2* col2 = reinterpret_cast<float2*>(a.coeffs + (size_t)(2 * j + 1) * a.ld) + i;
      *col1 = make_float2(c1e1 * fac, c1e2 * fac);
      *col2 = make_float2(c2e1 * fac, c2e2 * fac);
    } else {
      float c = 0.5f * ((vu * ttx + vv * tty) - fact * (wu * ttx + wv * tty));
      if (self) c = 3.14159265358979323846f;
      a.coeffs[(size_t)j * a.ld + i] = c * fac;
    }
  }
}

}  // namespace

static int coeff_device(o2d_context* c, int nsrc, const float* snx, const float* sny, const uint32_t* sidx,
                        const float* sarea, int ntarg, const float* tnx_, const float* tny_, const uint32_t* tidx,
                        const float* tarea, const float* ttx, const float* tty, const float* tnx, const float* tny,
                        int both, int same, size_t ld, float* coeffs) {
  CU(c->panel_packed.reserve(((size_t)nsrc + (size_t)ntarg) * 32));
  float4* spack = c->panel_packed.as<float4>();
  float4* tpack = spack + 2 * (size_t)nsrc;
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
  CU(cudaMemcpyAsync(dsi, sidx, (size_t)2 * nsrc * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(dti, tidx, (size_t)2 * ntarg * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  if ((rc = coeff_device(c, nsrc, dsnx, dsny, dsi, dsa, ntarg, dtnx, dtny, dti, dta, dttx, dtty, dnnx, dnny,
                         both_strengths, same_surface, dld, c->coef.as<float>())))
    return rc;
  CU(cudaMemcpy2DAsync(coeffs, (size_t)ld * sizeof(float), c->coef.p, dld * sizeof(float), rows * sizeof(float), cols,
                       cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return O2D_OK;
}
