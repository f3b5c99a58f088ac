// o2d_reflect.cu -- clear_inner_layer ("push" method) and reflect_interior.
//
// Replaces clear_inner_panp2<S> with _method == 1 (reference src/Reflect.h:490-733) as driven by
// clear_inner_layer<S> (:735-768), which Convection runs on every Lagrangian collection after every
// move (src/Convection.h:303-307, :334, :342, :388-391): for each particle find the closest panel or
// node of the boundary (panel_point_distance, :108-170; exact ties are averaged, :571-593), form the
// mean normal and contact point, and if the particle sits less than cutoff_mult*ips above the surface
// push it out along the normal; repeat until nothing moves.  O(N * Np) per sweep.
//
// One thread per particle, panels staged in shared memory 128 at a time (48 B per panel: end nodes,
// panel normal, the two node normals -- so a hit never needs a second lookup).  The reference's
// "while any untested" sweeps only ever read the panels, so every particle is iterated to completion
// independently; a CTA keeps sweeping while any of its particles still moves.  The C++ promotion rules
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
constexpr int kCilTile = 128;

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

struct CilArgs {
  const float4* packed;
  int np;
  float* tx; float* ty;
  int n;
  float cutoff;        // _cutoff_mult * _ips  (float product, src/Reflect.h:655)
  float ips;
  int max_iter;
  unsigned long long* moved;
};

enum { MODE_PUSH = 1, MODE_REFLECT = 2 };

template <int MODE>
__global__ void __launch_bounds__(kCilThreads)
clear_inner_kernel(const CilArgs a) {
  __shared__ float4 s[3 * kCilTile];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kCilThreads + tid;
  bool untested = i < a.n;
  float x = untested ? a.tx[i] : 0.0f;
  float y = untested ? a.ty[i] : 0.0f;
  unsigned int pushes = 0;
  for (int iter = 0; iter < a.max_iter; ++iter) {
    if (!__syncthreads_or(untested)) break;
    // push: "0.1*numeric_limits<S>::max()" (src/Reflect.h:566); reflect: numeric_limits<S>::max() (:214)
    float mindist = MODE == MODE_PUSH ? (float)(0.1 * (double)FLT_MAX) : FLT_MAX;
    int nhits = 0;
    float normx = 0.0f, normy = 0.0f, cpx = 0.0f, cpy = 0.0f;
    float n0x = 0.0f, n0y = 0.0f, c0x = 0.0f, c0y = 0.0f;   // hits[0] (reflect's zero-normal fallback, :305-317)
    for (int j0 = 0; j0 < a.np; j0 += kCilTile) {
      const int nj = min(kCilTile, a.np - j0);
      __syncthreads();
      for (int k = tid; k < 3 * nj; k += kCilThreads) s[k] = a.packed[3 * (size_t)j0 + k];
      __syncthreads();
      if (!untested) continue;
      for (int j = 0; j < nj; ++j) {
        const float4 P = s[3 * j];
        // panel_point_distance (src/Reflect.h:108-170)
        const float bx = P.z - P.x, by = P.w - P.y;
        const float blensq = bx * bx + by * by;
        const float ax = x - P.x, ay = y - P.y;
        const float t = __fdiv_rn(ax * bx + ay * by, blensq);
        float distsq;
        int kind;   // 0 panel, 1 node 0, 2 node 1
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
        const bool newbest = distsq < mindist;
        if (newbest || distsq == mindist) {
          if (newbest) { mindist = distsq; nhits = 0; normx = normy = cpx = cpy = 0.0f; }
          const float4 Nn = s[3 * j + 2];
          float hnx, hny, hcx, hcy;
          if (kind == 0) {
            const float4 Np = s[3 * j + 1];
            hnx = Np.x; hny = Np.y;
            hcx = (float)((1.0 - (double)t) * (double)P.x + (double)t * (double)P.z);
            hcy = (float)((1.0 - (double)t) * (double)P.y + (double)t * (double)P.w);
          } else if (kind == 1) {
            hnx = Nn.x; hny = Nn.y; hcx = P.x; hcy = P.y;
          } else {
            hnx = Nn.z; hny = Nn.w; hcx = P.z; hcy = P.w;
          }
          normx += hnx; normy += hny; cpx += hcx; cpy += hcy;
          if (MODE == MODE_REFLECT && nhits == 0) { n0x = hnx; n0y = hny; c0x = hcx; c0y = hcy; }
          nhits++;
        }
      }
    }
    if (MODE == MODE_REFLECT) {
      // reflect_panp2's ending (src/Reflect.h:291-342): one pass, no re-test
      if (untested && nhits > 0) {
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
    } else if (untested) {
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
  if (pushes > 0) {
    a.tx[i] = x;
    a.ty[i] = y;
    if (a.moved) atomicAdd(a.moved, (unsigned long long)pushes);
  }
}

}  // namespace

static int cil_device(o2d_context* c, int mode, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                      const float* snx, const float* sny, int n, float* tx, float* ty, float cutoff_mult, float ips,
                      unsigned long long* d_moved) {
  // scratch: node normals (2*nnodes floats) + packed panels (3 float4 each)
  const size_t nn_bytes = ((2 * (size_t)nnodes * sizeof(float) + 255) / 256) * 256;
  CU(c->panel_packed.reserve(nn_bytes + (size_t)np * 48));
  float* nnx = c->panel_packed.as<float>();
  float* nny = nnx + nnodes;
  float4* packed = reinterpret_cast<float4*>(reinterpret_cast<char*>(c->panel_packed.p) + nn_bytes);
  CU(cudaMemsetAsync(nnx, 0, 2 * (size_t)nnodes * sizeof(float), c->stream));
  node_normals_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(np, idx, snx, sny, nnx, nny);
  CU(cudaGetLastError());
  pack_cil_panels_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(np, nx, ny, idx, snx, sny, nnx, nny, packed);
  CU(cudaGetLastError());
  CilArgs a;
  a.packed = packed; a.np = np; a.tx = tx; a.ty = ty; a.n = n;
  a.cutoff = cutoff_mult * ips; a.ips = ips; a.max_iter = 1000; a.moved = d_moved;
  {
    PairTimer timer(c);
    const int nb = (n + kCilThreads - 1) / kCilThreads;
    if (mode == MODE_PUSH) clear_inner_kernel<MODE_PUSH><<<nb, kCilThreads, 0, c->stream>>>(a);
    else                   clear_inner_kernel<MODE_REFLECT><<<nb, kCilThreads, 0, c->stream>>>(a);
    CU(cudaGetLastError());
  }
  c->launches += 3;
  c->pair_launches++;
  return O2D_OK;
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
    if ((rc = cil_device(c, mode, np, nnodes, nx, ny, idx, snx, sny, n, tx, ty, cutoff_mult, ips, d_moved))) return rc;
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
  if ((rc = cil_device(c, mode, np, nnodes, dnx, dny, c->hidx.as<uint32_t>(), dsnx, dsny, n, dtx, dty, cutoff_mult, ips, d_moved)))
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
