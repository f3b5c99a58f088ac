// o2d_reflect.cu -- clear_inner_layer ("push" method) and reflect_interior.
//
// Replaces clear_inner_panp2<S> with _method == 1 (reference src/Reflect.h:490-733) as driven by
// clear_inner_layer<S> (:735-768), which Convection runs on every Lagrangian collection after every
// move (src/Convection.h:303-307, :334, :342, :388-391): for each particle find the closest panel or
// node of the boundary (panel_point_distance, :108-170; exact ties are averaged, :571-593), form the
// mean normal and contact point, and if the particle sits less than cutoff_mult*ips above the surface
// push it out along the normal; repeat until nothing moves.  O(N * Np) per sweep.
//
// One warp per particle.  The search is exact but not brute force: consecutive panels (kCilGroup = 32 of them)
// share a bounding circle, and a group is opened only if it can still hold the closest element:
//   pass A   ub = min over groups of (|x - centre| + radius)^2   -- some panel point IS that close;
//   pass B   groups in panel order; skip one when (|x - centre| - radius)^2 > min(ub, best so far).
// A skipped group holds only elements strictly farther than the final minimum, so the set of closest elements,
// the order in which ties are accumulated, and therefore every output bit are those of the reference's full scan
// (the reference's intermediate "new best" resets differ, its final state does not).  1M particles x 20 000
// panels: 104 ms as a full scan -> see DESIGN 3.5.  The reference's "while any untested" sweeps only ever read
// the panels, so every particle is iterated to completion independently.  The C++ promotion rules
// the reference's expressions imply (double pow, double contact point, double/float nextafter
// overloads -- restated in oracle/o2d_oracle.c) are followed so that the closest-hit logic, which is
// discontinuous, sees the same numbers: "new best" is distsq < mindist, a tie is distsq == mindist.
//
// reflect_interior<S> -> reflect_panp2<S> (src/Reflect.h:174-350, driver :357-380; Diffusion::step calls it
// before VRM, src/Diffusion.h:301) is the same closest-element search run ONCE per particle with a
// different ending: a particle on the inner side of its mean contact normal is mirrored to
// cp + dist*normal.  It is the same kernel, template parameter MODE.
#include "o2d_context.h"

#include <cfloat>

using namespace o2d;

namespace {

constexpr int kCilThreads = 128;
constexpr int kCilGroup = 32;   // consecutive panels under one bounding circle

// node normals = sum of the adjacent panel normals (src/Reflect.h:521-530).  A node of a closed body
// belongs to two panels, for which the float sum does not depend on the order of the two atomics.
__global__ void __launch_bounds__(256)
node_normals_kernel(int np, const uint32_t* __restrict__ idx, const float* __restrict__ snx,
                    const float* __restrict__ sny, float* nnx, float* nny) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  const uint32_t i0 = idx[2 * j], i1 = idx[2 * j + 1];
  atomicAdd(&nnx[i0], snx[j]); atomicAdd(&nny[i0], sny[j]);
  atomicAdd(&nnx[i1], snx[j]); atomicAdd(&nny[i1], sny[j]);
}

__global__ void __launch_bounds__(256)
pack_cil_panels_kernel(int np, const float* __restrict__ nx, const float* __restrict__ ny,
                       const uint32_t* __restrict__ idx, const float* __restrict__ snx,
                       const float* __restrict__ sny, const float* __restrict__ nnx,
                       const float* __restrict__ nny, float4* __restrict__ packed) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  const uint32_t i0 = idx[2 * j], i1 = idx[2 * j + 1];
  packed[3 * j]     = make_float4(nx[i0], ny[i0], nx[i1], ny[i1]);
  packed[3 * j + 1] = make_float4(snx[j], sny[j], 0.0f, 0.0f);
  packed[3 * j + 2] = make_float4(nnx[i0], nny[i0], nnx[i1], nny[i1]);
}

// bounding circle of the end nodes of panels [g*kCilGroup, (g+1)*kCilGroup): centre of their bounding box, radius to
// the farthest node, widened by a few ulp so that rounding can only make it larger
__global__ void __launch_bounds__(128)
cil_group_bounds_kernel(int np, const float4* __restrict__ packed, float4* __restrict__ bounds) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = g * kCilGroup;
  if (j0 >= np) return;
  const int j1 = min(j0 + kCilGroup, np);
  float xlo = FLT_MAX, xhi = -FLT_MAX, ylo = FLT_MAX, yhi = -FLT_MAX;
  for (int j = j0; j < j1; ++j) {
    const float4 P = packed[3 * (size_t)j];
    xlo = fminf(xlo, fminf(P.x, P.z)); xhi = fmaxf(xhi, fmaxf(P.x, P.z));
    ylo = fminf(ylo, fminf(P.y, P.w)); yhi = fmaxf(yhi, fmaxf(P.y, P.w));
  }
  const float cx = 0.5f * (xlo + xhi), cy = 0.5f * (ylo + yhi);
  float r2 = 0.0f;
  for (int j = j0; j < j1; ++j) {
    const float4 P = packed[3 * (size_t)j];
    r2 = fmaxf(r2, fmaxf((P.x - cx) * (P.x - cx) + (P.y - cy) * (P.y - cy), (P.z - cx) * (P.z - cx) + (P.w - cy) * (P.w - cy)));
  }
  bounds[g] = make_float4(cx, cy, sqrtf(r2) * 1.00001f + FLT_MIN, 0.0f);
}

struct CilArgs {
  const float4* packed;
  const float4* bounds;
  int np;
  int ngroups;
  float* tx; float* ty;
  int n;
  float cutoff;        // _cutoff_mult * _ips  (float product, src/Reflect.h:655)
  float ips;
  int max_iter;
  unsigned long long* moved;
};

enum { MODE_PUSH = 1, MODE_REFLECT = 2 };

// One WARP per particle.  Pass A and the group tests of pass B are spread over the lanes (group = round*32 + lane);
// a group that passes is then tested by the whole warp, one panel per lane, and folded into the running state in
// panel order: the group's minimum by shuffle reduction, its ties by walking the ballot mask from lane 0 up -- the
// same elements, added in the same order, as the reference's serial scan.
template <int MODE>
__global__ void __launch_bounds__(kCilThreads)
clear_inner_kernel(const CilArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kCilThreads / 32) + (threadIdx.x >> 5);
  if (i >= a.n) return;
  float x = a.tx[i];
  float y = a.ty[i];
  unsigned int pushes = 0;
  bool untested = true;
  const int nrounds = (a.ngroups + 31) / 32;
  for (int iter = 0; iter < a.max_iter && untested; ++iter) {
    // push: "0.1*numeric_limits<S>::max()" (src/Reflect.h:566); reflect: numeric_limits<S>::max() (:214)
    float mindist = MODE == MODE_PUSH ? (float)(0.1 * (double)FLT_MAX) : FLT_MAX;
    // pass A: an upper bound on the squared distance to the closest element (rounded up)
    float ub = mindist;
    for (int g = lane; g < a.ngroups; g += 32) {
      const float4 B = __ldg(&a.bounds[g]);
      const float dx = x - B.x, dy = y - B.y;
      const float u = sqrtf(dx * dx + dy * dy) + B.z;
      ub = fminf(ub, u * u * 1.00001f);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ub = fminf(ub, __shfl_xor_sync(FULL, ub, off));
    int nhits = 0;
    float normx = 0.0f, normy = 0.0f, cpx = 0.0f, cpy = 0.0f;
    float n0x = 0.0f, n0y = 0.0f, c0x = 0.0f, c0y = 0.0f;   // hits[0] (reflect's zero-normal fallback, :305-317)
    // pass B: the reference's scan in panel order, minus the groups that cannot hold a closest element
    for (int r = 0; r < nrounds; ++r) {
      const int g = r * 32 + lane;
      float lb2 = FLT_MAX;   // squared lower bound of this lane's group (shrunk a little: rounding can only open more groups)
      if (g < a.ngroups) {
        const float4 B = __ldg(&a.bounds[g]);
        const float dx = x - B.x, dy = y - B.y;
        const float lb = sqrtf(dx * dx + dy * dy) - B.z;
        lb2 = lb > 0.0f ? lb * lb * 0.9998f : 0.0f;
      }
      unsigned open = __ballot_sync(FULL, g < a.ngroups && !(lb2 > ub));
      while (open) {
        const int b = __ffs(open) - 1;
        open &= open - 1;
        if (__shfl_sync(FULL, lb2, b) > mindist) continue;   // the running minimum has closed it meanwhile
        const int j = (r * 32 + b) * kCilGroup + lane;
        float distsq = FLT_MAX;
        int kind = 0;
        float t = 0.0f;
        float4 P = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (j < a.np) {
          P = __ldg(&a.packed[3 * (size_t)j]);
          // panel_point_distance (src/Reflect.h:108-170)
          const float bx = P.z - P.x, by = P.w - P.y;
          const float blensq = bx * bx + by * by;
          const float ax = x - P.x, ay = y - P.y;
          t = __fdiv_rn(ax * bx + ay * by, blensq);
          if (t > 0.0f && t < 1.0f) {
            const float c = ay * bx - ax * by;
            distsq = (float)(((double)c * (double)c) / (double)blensq);
            kind = 0;
          } else {
            const float rij2 = (float)((double)ax * (double)ax + (double)ay * (double)ay);
            const float cx = x - P.z, cy = y - P.w;
            const float rij12 = (float)((double)cx * (double)cx + (double)cy * (double)cy);
            if (rij2 < rij12) { distsq = rij2; kind = 1; } else { distsq = rij12; kind = 2; }
          }
        }
        float gmin = distsq;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gmin = fminf(gmin, __shfl_xor_sync(FULL, gmin, off));
        if (gmin > mindist) continue;
        if (gmin < mindist) { mindist = gmin; nhits = 0; normx = normy = cpx = cpy = 0.0f; }
        // this panel's contribution if it is one of the closest elements
        const bool hit = j < a.np && distsq == mindist;
        float hnx = 0.0f, hny = 0.0f, hcx = 0.0f, hcy = 0.0f;
        if (hit) {
          if (kind == 0) {
            const float4 Np = __ldg(&a.packed[3 * (size_t)j + 1]);
            hnx = Np.x; hny = Np.y;
            // "(1.0-t)*sx0 + t*sx1" (:141-142): (1.0-t)*sx0 is a double product, t*sx1 a FLOAT product promoted for the sum
            hcx = (float)((1.0 - (double)t) * (double)P.x + (double)__fmul_rn(t, P.z));
            hcy = (float)((1.0 - (double)t) * (double)P.y + (double)__fmul_rn(t, P.w));
          } else {
            const float4 Nn = __ldg(&a.packed[3 * (size_t)j + 2]);
            if (kind == 1) { hnx = Nn.x; hny = Nn.y; hcx = P.x; hcy = P.y; }
            else           { hnx = Nn.z; hny = Nn.w; hcx = P.z; hcy = P.w; }
          }
        }
        unsigned hits = __ballot_sync(FULL, hit);
        while (hits) {   // ties in panel order (float sums are order dependent from the third term on)
          const int h = __ffs(hits) - 1;
          hits &= hits - 1;
          const float vnx = __shfl_sync(FULL, hnx, h), vny = __shfl_sync(FULL, hny, h);
          const float vcx = __shfl_sync(FULL, hcx, h), vcy = __shfl_sync(FULL, hcy, h);
          normx += vnx; normy += vny; cpx += vcx; cpy += vcy;
          if (MODE == MODE_REFLECT && nhits == 0) { n0x = vnx; n0y = vny; c0x = vcx; c0y = vcy; }
          nhits++;
        }
      }
    }
    // from here on every lane holds the same state and takes the same branches
    if (MODE == MODE_REFLECT) {
      // reflect_panp2's ending (src/Reflect.h:291-342): one pass, no re-test
      if (nhits > 0) {
        if (normx * normx + normy * normy == 0.0f) {   // opposing normals cancel: fall back to hits[0]
          normx += n0x; normy += n0y; cpx = c0x; cpy = c0y;
        }
        const float normilen = (float)(1.0 / (double)sqrtf(normx * normx + normy * normy));
        normx *= normilen; normy *= normilen;
        cpx = __fdiv_rn(cpx, (float)nhits); cpy = __fdiv_rn(cpy, (float)nhits);
        const float dotp = normx * (x - cpx) + normy * (y - cpy);
        if (dotp < 0.0f) {
          const float dist = sqrtf(mindist);           // hits[0].distsq == mindist
          x = cpx + dist * normx;
          y = cpy + dist * normy;
          pushes++;
        }
      }
      untested = false;
    } else {
      if (nhits == 0) {
        untested = false;   // "this messes up long runs, ignore it" (src/Reflect.h:620-622)
      } else {
        const float normilen = (float)(1.0 / (double)sqrtf(normx * normx + normy * normy));
        normx *= normilen; normy *= normilen;
        cpx = __fdiv_rn(cpx, (float)nhits); cpy = __fdiv_rn(cpy, (float)nhits);
        const float dotp = normx * (x - cpx) + normy * (y - cpy) - a.cutoff;
        if ((double)dotp < -0.001 * (double)a.ips) {
          x -= dotp * normx;
          y -= dotp * normy;
          pushes++;
        } else {
          untested = false;
        }
      }
    }
  }
  if (pushes > 0 && lane == 0) {
    a.tx[i] = x;
    a.ty[i] = y;
    if (a.moved) atomicAdd(a.moved, (unsigned long long)pushes);
  }
}

}  // namespace

// panel-side preparation of a clear/reflect pass (node normals, packed panels, group bounds) into c->cil; the panels
// of a body do not change between the Runge-Kutta stages of a step, so o2d_advect_bodies prepares once and runs per stage
int o2d_cil_prepare(o2d_context* c, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                    const float* snx, const float* sny) {
  // scratch: node normals (2*nnodes floats) + packed panels (3 float4 each) + group bounds (1 float4 per 32 panels)
  const size_t nn_bytes = ((2 * (size_t)nnodes * sizeof(float) + 255) / 256) * 256;
  const int ngroups = (np + kCilGroup - 1) / kCilGroup;
  CU(c->cil.reserve(nn_bytes + (size_t)np * 48 + (size_t)ngroups * 16));
  float* nnx = c->cil.as<float>();
  float* nny = nnx + nnodes;
  float4* packed = reinterpret_cast<float4*>(reinterpret_cast<char*>(c->cil.p) + nn_bytes);
  float4* bounds = packed + 3 * (size_t)np;
  CU(cudaMemsetAsync(nnx, 0, 2 * (size_t)nnodes * sizeof(float), c->stream));
  node_normals_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(np, idx, snx, sny, nnx, nny);
  CU(cudaGetLastError());
  pack_cil_panels_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(np, nx, ny, idx, snx, sny, nnx, nny, packed);
  CU(cudaGetLastError());
  cil_group_bounds_kernel<<<(ngroups + 127) / 128, 128, 0, c->stream>>>(np, packed, bounds);
  CU(cudaGetLastError());
  c->launches += 3;
  c->cil_np = np;
  c->cil_nn_bytes = nn_bytes;
  return O2D_OK;
}

// the particle-side pass over the panels prepared last (mode 1 = clear_inner_layer push, 2 = reflect_interior)
int o2d_cil_run(o2d_context* c, int mode, int n, float* tx, float* ty, float cutoff_mult, float ips,
                unsigned long long* d_moved) {
  const int np = c->cil_np;
  float4* packed = reinterpret_cast<float4*>(reinterpret_cast<char*>(c->cil.p) + c->cil_nn_bytes);
  CilArgs a;
  a.packed = packed; a.bounds = packed + 3 * (size_t)np; a.np = np; a.ngroups = (np + kCilGroup - 1) / kCilGroup;
  a.tx = tx; a.ty = ty; a.n = n;
  a.cutoff = cutoff_mult * ips; a.ips = ips; a.max_iter = 1000; a.moved = d_moved;
  {
    PairTimer timer(c);
    const int per_cta = kCilThreads / 32;   // one warp per particle
    const int nb = (n + per_cta - 1) / per_cta;
    if (mode == MODE_PUSH) clear_inner_kernel<MODE_PUSH><<<nb, kCilThreads, 0, c->stream>>>(a);
    else                   clear_inner_kernel<MODE_REFLECT><<<nb, kCilThreads, 0, c->stream>>>(a);
    CU(cudaGetLastError());
  }
  c->launches++;
  c->pair_launches++;
  return O2D_OK;
}

int o2d_cil_device(o2d_context* c, int mode, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                   const float* snx, const float* sny, int n, float* tx, float* ty, float cutoff_mult, float ips,
                   unsigned long long* d_moved) {
  if (int rc = o2d_cil_prepare(c, np, nnodes, nx, ny, idx, snx, sny)) return rc;
  return o2d_cil_run(c, mode, n, tx, ty, cutoff_mult, ips, d_moved);
}

static int cil_entry(o2d_context* c, int mode, int mem,
                     int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                     const float* snx, const float* sny,
                     int n, float* tx, float* ty, float cutoff_mult, float ips,
                     long long* num_moved) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (np < 0 || nnodes < 0 || n < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (num_moved) *num_moved = 0;
  if (np == 0 || n == 0) return O2D_OK;
  if (!nx || !ny || !idx || !snx || !sny || !tx || !ty) return fail(O2D_ERR_INVALID, "NULL array");
  CU(c->counter.reserve(sizeof(unsigned long long)));
  unsigned long long* d_moved = c->counter.as<unsigned long long>();
  CU(cudaMemsetAsync(d_moved, 0, sizeof(unsigned long long), c->stream));
  int rc;
  if (mem == O2D_MEM_DEVICE) {
    if ((rc = o2d_cil_device(c, mode, np, nnodes, nx, ny, idx, snx, sny, n, tx, ty, cutoff_mult, ips, d_moved))) return rc;
    if (num_moved) {   // the count is the only thing that needs the host: synchronous only when asked for
      unsigned long long h = 0;
      CU(cudaMemcpyAsync(&h, d_moved, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
      if (int rc_ = o2d_finish_host(c)) return rc_;
      *num_moved = (long long)h;
    }
    return O2D_OK;
  }
  for (int k = 0; k < 2 * np; ++k)
    if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
  CU(c->hpan.reserve(2 * padded_floats(nnodes) + 2 * padded_floats(np)));
  CU(c->hidx.reserve((size_t)2 * np * sizeof(uint32_t)));
  CU(c->htgt.reserve(3 * padded_floats(n)));
  Stager S(c, &c->hpan), T(c, &c->htgt);
  float *dnx = S.take(nnodes), *dny = S.take(nnodes), *dsnx = S.take(np), *dsny = S.take(np);
  float *dtx = T.take(n), *dty = T.take(n);
  if ((rc = o2d_h2d(c, dnx, nx, nnodes)) || (rc = o2d_h2d(c, dny, ny, nnodes)) || (rc = o2d_h2d(c, dsnx, snx, np)) ||
      (rc = o2d_h2d(c, dsny, sny, np)) || (rc = o2d_h2d(c, dtx, tx, n)) || (rc = o2d_h2d(c, dty, ty, n)))
    return rc;
  if (int rc_ = o2d_h2d_bytes(c, c->hidx.p, idx, (size_t)2 * np * sizeof(uint32_t))) return rc_;
  if ((rc = o2d_cil_device(c, mode, np, nnodes, dnx, dny, c->hidx.as<uint32_t>(), dsnx, dsny, n, dtx, dty, cutoff_mult, ips, d_moved)))
    return rc;
  unsigned long long h = 0;
  if ((rc = o2d_d2h(c, tx, dtx, n)) || (rc = o2d_d2h(c, ty, dty, n))) return rc;
  CU(cudaMemcpyAsync(&h, d_moved, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  if (int rc_ = o2d_finish_host(c)) return rc_;
  if (num_moved) *num_moved = (long long)h;
  return O2D_OK;
}

extern "C" int o2d_clear_inner_layer(o2d_context* c, int method, int mem,
                                     int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                     const float* snx, const float* sny,
                                     int n, float* tx, float* ty, float cutoff_mult, float ips,
                                     long long* num_moved) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (method != 1)
    return fail(O2D_ERR_UNSUPPORTED, "only the push method (1) is built; method 0 (strength cropping with the cut "
                                     "tables, reference src/Reflect.h:392-480,668-690) is never called by the reference");
  return cil_entry(c, MODE_PUSH, mem, np, nnodes, nx, ny, idx, snx, sny, n, tx, ty, cutoff_mult, ips, num_moved);
}

extern "C" int o2d_reflect_interior(o2d_context* c, int mem,
                                    int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                    const float* snx, const float* sny,
                                    int n, float* tx, float* ty, long long* num_reflected) {
  return cil_entry(c, MODE_REFLECT, mem, np, nnodes, nx, ny, idx, snx, sny, n, tx, ty, 0.0f, 0.0f, num_reflected);
}
