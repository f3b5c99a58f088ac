// o2d_substep.cu -- device-resident Runge-Kutta substeps.
//
// The reference evaluates velocities 1-3 times per time step (Convection::advect_1st,
// advect_2nd_ralston, advect_3rd; reference src/Convection.h:277-595), each evaluation being
// zero_vels -> influence sums over every source collection -> finalize_vels
// (Convection::find_vels, src/Convection.h:172-223), with ElementBase::move
// (src/ElementBase.h:276-359) in between.  Called through o2d_points_on_points each evaluation
// costs a host->device copy of every array and a device->host copy of the results.  The two
// entry points here keep the particle state on the GPU for a whole evaluation (o2d_find_vels) or a
// whole time step (o2d_advect): one upload, one download.
#include "o2d_context.h"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace o2d;

namespace {

// ElementBase::move, one weight (src/ElementBase.h:276-296):
//   x[d][i] += (S)_dt * _wt1 * _u1.u[d][i]     -- (float)dt * double wt -> double; * float u -> double;
//   the float position is updated through a double sum.  xo may alias xi (in-place move).
__global__ void __launch_bounds__(256)
move1_kernel(int n, const float* __restrict__ xi, const float* __restrict__ yi, float* xo, float* yo,
             const float* __restrict__ u, const float* __restrict__ v, float dt, double wt1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = (double)dt * wt1;
  xo[i] = (float)((double)xi[i] + f * (double)u[i]);
  yo[i] = (float)((double)yi[i] + f * (double)v[i]);
}

// ElementBase::move, two / three weights (src/ElementBase.h:299-359):
//   u[d][i] = _wt1*_u1.u[d][i] + _wt2*_u2.u[d][i] (+ _wt3*_u3.u[d][i])   double math, float store
//   x[d][i] += (S)_dt * u[d][i]                                            float math
// u1,v1 belong to the moved collection itself and receive the combined velocity.
__global__ void __launch_bounds__(256)
move23_kernel(int n, float* x, float* y, float* u1, float* v1, const float* __restrict__ u2,
              const float* __restrict__ v2, const float* __restrict__ u3, const float* __restrict__ v3,
              float dt, double wt1, double wt2, double wt3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double du = wt1 * (double)u1[i] + wt2 * (double)u2[i];
  double dv = wt1 * (double)v1[i] + wt2 * (double)v2[i];
  if (u3 != nullptr) {
    du += wt3 * (double)u3[i];
    dv += wt3 * (double)v3[i];
  }
  const float fu = (float)du, fv = (float)dv;
  u1[i] = fu;
  v1[i] = fv;
  x[i] = x[i] + dt * fu;
  y[i] = y[i] + dt * fv;
}

// Surfaces::add_rot_strengths_base (reference src/Surfaces.h:571-620): sheet strengths that stand for the
// solid-body rotation of the body the panels belong to, ADDED to vs/ss.  Same float expression order.
__host__ __device__ inline void rot_strength_one(int i, const float* ux, const float* uy, const uint32_t* idx,
                                                 const float* area, float cx, float cy, float rotvel,
                                                 float* vs, float* ss) {
  const uint32_t j = idx[2 * i], jp1 = idx[2 * i + 1];
  const float dx = 0.5f * (ux[j] + ux[jp1]) - cx;
  const float dy = 0.5f * (uy[j] + uy[jp1]) - cy;
  const float ui = -rotvel * dy;
  const float vi = rotvel * dx;
  float panelx = ux[jp1] - ux[j];
  float panely = uy[jp1] - uy[j];
  const float oopanell = 1.0f / area[i];
  panelx *= oopanell;
  panely *= oopanell;
  vs[i] += -1.0f * (ui * panelx + vi * panely);
  ss[i] += -1.0f * (ui * panely - vi * panelx);
}

__global__ void __launch_bounds__(256)
rot_strengths_kernel(int np, const float* __restrict__ ux, const float* __restrict__ uy,
                     const uint32_t* __restrict__ idx, const float* __restrict__ area, float cx, float cy,
                     float rotvel, float* vs, float* ss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) rot_strength_one(i, ux, uy, idx, area, cx, cy, rotvel, vs, ss);
}

struct Panels {
  int np = 0;
  const float* nx = nullptr;
  const float* ny = nullptr;
  const uint32_t* idx = nullptr;
  const float* vs = nullptr;
  const float* ss = nullptr;
};

// Convection::find_vels for one particle collection + one surface collection on one target
// collection, all in device memory: zero -> particles -> panels -> finalize.
// nv_block > 0: the particles are this rank's block of a sharded set (o2d_points_on_points_sharded).
int find_vels_device(o2d_context* c, int core, int res, int nv, const float* vx, const float* vy, const float* vs,
                     const float* vr, const Panels& pan, int nt, const float* tx, const float* ty, const float* tr,
                     float* tu, float* tv, float* tq, double fsx, double fsy, int nv_block = 0) {
  if (nt <= 0 && nv_block <= 0) return O2D_OK;
  float* w = res == RES_UW ? tq : nullptr;
  float* e = res == RES_UE ? tq : nullptr;
  if (nt > 0) {
    CU(launch_zero(nt, tu, tv, w, e, c->stream));
    c->launches++;
  }
  OutPtrs o{};
  o.p[0] = tu; o.p[1] = tv; o.p[2] = tq;
  if (nv_block > 0) {
    if (int rc = o2d_pp_sharded_device(c, core, res, nv, nv_block, vx, vy, vs, vr, nt, tx, ty, tr, o)) return rc;
  } else {
    if (int rc = o2d_pp_device(c, core, res, nv, vx, vy, vs, vr, nt, tx, ty, tr, o)) return rc;
  }
  if (nt <= 0) return O2D_OK;
  if (pan.np > 0)
    if (int rc = o2d_panels_on_points_device(c, pan.np, pan.nx, pan.ny, pan.idx, pan.vs, pan.ss, nt, tx, ty, tu, tv))
      return rc;
  CU(launch_finalize(nt, tu, tv, w, e, fsx, fsy, c->stream));
  c->launches++;
  return O2D_OK;
}

// upload the panel description (host -> context mirrors); returns device pointers in `d`
int stage_panels(o2d_context* c, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                 const float* vs, const float* ss, Panels* d) {
  d->np = np;
  if (np <= 0) return O2D_OK;
  for (int k = 0; k < 2 * np; ++k)
    if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
  CU(c->hpan.reserve(2 * padded_floats(nnodes) + 2 * padded_floats(np)));
  CU(c->hidx.reserve((size_t)2 * np * sizeof(uint32_t)));
  Stager S(c, &c->hpan);
  float *dnx = S.take(nnodes), *dny = S.take(nnodes), *dvs = S.take(np), *dss = ss ? S.take(np) : nullptr;
  HostBatch bp(c);
  bp.add(dnx, nx, nnodes); bp.add(dny, ny, nnodes); bp.add(dvs, vs, np);
  if (ss) bp.add(dss, ss, np);
  if (int rc_ = bp.upload()) return rc_;
  if (int rc_ = o2d_h2d_bytes(c, c->hidx.p, idx, (size_t)2 * np * sizeof(uint32_t))) return rc_;
  d->nx = dnx; d->ny = dny; d->idx = c->hidx.as<uint32_t>(); d->vs = dvs; d->ss = dss;
  return O2D_OK;
}

int check_panels(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx, const float* vs) {
  if (np < 0 || nnodes < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (np > 0 && (!nx || !ny || !idx || !vs)) return fail(O2D_ERR_INVALID, "NULL panel array");
  return O2D_OK;
}

}  // namespace

extern "C" int o2d_add_rot_strengths(o2d_context* c, int mem, int np, int nnodes, const float* ux, const float* uy,
                                     const uint32_t* idx, const float* area, float cx, float cy, float rotvel,
                                     float* vs, float* ss) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (np < 0 || nnodes < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (np == 0) return O2D_OK;
  if (!ux || !uy || !idx || !area || !vs || !ss) return fail(O2D_ERR_INVALID, "NULL array");
  if (mem == O2D_MEM_HOST) {
    // O(Np) on arrays that live on the host: the reference's own loop, nothing worth a round trip
    for (int k = 0; k < 2 * np; ++k)
      if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
    for (int i = 0; i < np; ++i) rot_strength_one(i, ux, uy, idx, area, cx, cy, rotvel, vs, ss);
    return O2D_OK;
  }
  rot_strengths_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(np, ux, uy, idx, area, cx, cy, rotvel, vs, ss);
  CU(cudaGetLastError());
  c->launches++;
  return O2D_OK;
}

extern "C" int o2d_find_vels(o2d_context* c, int core, int results, int mem,
                             int nv, const float* vx, const float* vy, const float* vs, const float* vr,
                             int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                             const float* pvs, const float* pss,
                             int nt, const float* tx, const float* ty, const float* tr,
                             float* tu, float* tv, float* tq, double fsx, double fsy) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (results == O2D_PSIONLY) return fail(O2D_ERR_UNSUPPORTED, "streamfunction results are not supported");
  if (results == O2D_VELANDGRAD)
    return fail(O2D_ERR_UNSUPPORTED, "find_vels has no velocity-gradient results (reference src/Convection.h:172-223)");
  const int res = o2d_res_from_results(results);
  if (res < 0) return fail(O2D_ERR_INVALID, "unknown results type %d", results);
  if (nv < 0 || nt < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (int rc = check_panels(np, nnodes, nx, ny, idx, pvs)) return rc;
  if (nt == 0) return O2D_OK;
  if (nv > 0 && (!vx || !vy || !vs || !vr)) return fail(O2D_ERR_INVALID, "NULL particle array");
  if (!tx || !ty || !tu || !tv) return fail(O2D_ERR_INVALID, "NULL target array");
  if (res != RES_U && !tq) return fail(O2D_ERR_INVALID, "tq must not be NULL for vorticity/shear results");
  Panels pan;
  if (mem == O2D_MEM_DEVICE) {
    pan.np = np; pan.nx = nx; pan.ny = ny; pan.idx = idx; pan.vs = pvs; pan.ss = pss;
    return find_vels_device(c, core, res, nv, vx, vy, vs, vr, pan, nt, tx, ty, tr, tu, tv, tq, fsx, fsy);
  }
  // host arrays: one upload of sources and target positions, one download of the results
  // (the outputs are overwritten, so unlike o2d_points_on_points they are not uploaded)
  CU(c->hsrc.reserve(4 * padded_floats(nv) + 256));
  CU(c->htgt.reserve(3 * padded_floats(nt)));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt)));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dvx = S.take(nv), *dvy = S.take(nv), *dvs = S.take(nv), *dvr = S.take(nv);
  float *dtx = T.take(nt), *dty = T.take(nt), *dtr = tr ? T.take(nt) : nullptr;
  float *dtu = O.take(nt), *dtv = O.take(nt), *dtq = tq ? O.take(nt) : nullptr;
  int rc;
  HostBatch bs(c), bt(c), bo(c);
  bs.add(dvx, vx, nv); bs.add(dvy, vy, nv); bs.add(dvs, vs, nv); bs.add(dvr, vr, nv);
  bt.add(dtx, tx, nt); bt.add(dty, ty, nt);
  if (tr) bt.add(dtr, tr, nt);
  if ((rc = bs.upload()) || (rc = bt.upload())) return rc;
  if ((rc = stage_panels(c, np, nnodes, nx, ny, idx, pvs, pss, &pan))) return rc;
  if ((rc = find_vels_device(c, core, res, nv, dvx, dvy, dvs, dvr, pan, nt, dtx, dty, dtr, dtu, dtv, dtq, fsx, fsy)))
    return rc;
  bo.add(dtu, tu, nt); bo.add(dtv, tv, nt);
  if (res != RES_U) bo.add(dtq, tq, nt);
  if ((rc = bo.download())) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

extern "C" int o2d_find_vels_sharded(o2d_context* c, int core, int results, int mem,
                                     int nv_local, int nv_block,
                                     const float* vx, const float* vy, const float* vs, const float* vr,
                                     int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                     const float* pvs, const float* pss,
                                     int nt, const float* tx, const float* ty, const float* tr,
                                     float* tu, float* tv, float* tq, double fsx, double fsy) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (c->nranks > 1 && !c->comm) return fail(O2D_ERR_COMM, "communicator not initialised (o2d_comm_init)");
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (results == O2D_PSIONLY || results == O2D_VELANDGRAD)
    return fail(O2D_ERR_UNSUPPORTED, "find_vels computes vel, vel+vort or vel+shear (reference src/Convection.h:172-223)");
  const int res = o2d_res_from_results(results);
  if (res < 0) return fail(O2D_ERR_INVALID, "unknown results type %d", results);
  if (nv_local < 0 || nt < 0 || nv_block <= 0 || nv_local > nv_block)
    return fail(O2D_ERR_INVALID, "bad counts nv_local=%d nv_block=%d nt=%d", nv_local, nv_block, nt);
  if (int rc = check_panels(np, nnodes, nx, ny, idx, pvs)) return rc;
  if (nv_local > 0 && (!vx || !vy || !vs || !vr)) return fail(O2D_ERR_INVALID, "NULL particle array");
  if (nt > 0 && (!tx || !ty || !tu || !tv)) return fail(O2D_ERR_INVALID, "NULL target array");
  if (nt > 0 && res != RES_U && !tq) return fail(O2D_ERR_INVALID, "tq must not be NULL for vorticity/shear results");
  Panels pan;
  if (mem == O2D_MEM_DEVICE) {
    pan.np = np; pan.nx = nx; pan.ny = ny; pan.idx = idx; pan.vs = pvs; pan.ss = pss;
    return find_vels_device(c, core, res, nv_local, vx, vy, vs, vr, pan, nt, tx, ty, tr, tu, tv, tq, fsx, fsy, nv_block);
  }
  CU(c->hsrc.reserve(4 * padded_floats(nv_local) + 256));
  CU(c->htgt.reserve(3 * padded_floats(nt) + 256));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt) + 256));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dvx = S.take(nv_local), *dvy = S.take(nv_local), *dvs = S.take(nv_local), *dvr = S.take(nv_local);
  float *dtx = T.take(nt), *dty = T.take(nt), *dtr = tr ? T.take(nt) : nullptr;
  float *dtu = O.take(nt), *dtv = O.take(nt), *dtq = tq ? O.take(nt) : nullptr;
  int rc;
  if ((rc = o2d_h2d(c, dvx, vx, nv_local)) || (rc = o2d_h2d(c, dvy, vy, nv_local)) || (rc = o2d_h2d(c, dvs, vs, nv_local)) ||
      (rc = o2d_h2d(c, dvr, vr, nv_local)) || (rc = o2d_h2d(c, dtx, tx, nt)) || (rc = o2d_h2d(c, dty, ty, nt)))
    return rc;
  if (tr && (rc = o2d_h2d(c, dtr, tr, nt))) return rc;
  if ((rc = stage_panels(c, np, nnodes, nx, ny, idx, pvs, pss, &pan))) return rc;
  if ((rc = find_vels_device(c, core, res, nv_local, dvx, dvy, dvs, dvr, pan, nt, dtx, dty, dtr, dtu, dtv, dtq, fsx, fsy,
                             nv_block)))
    return rc;
  if ((rc = o2d_d2h(c, tu, dtu, nt)) || (rc = o2d_d2h(c, tv, dtv, nt))) return rc;
  if (res != RES_U && (rc = o2d_d2h(c, tq, dtq, nt))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

extern "C" int o2d_find_vels_multi(o2d_context* c, int core, int results, int mem,
                                   int nsrc, const o2d_points_view* src, int npan, const o2d_panels_view* pan,
                                   int nt, const float* tx, const float* ty, const float* tr,
                                   float* tu, float* tv, float* tq, double fsx, double fsy) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (results == O2D_PSIONLY || results == O2D_VELANDGRAD)
    return fail(O2D_ERR_UNSUPPORTED, "find_vels computes vel, vel+vort or vel+shear (reference src/Convection.h:172-223)");
  const int res = o2d_res_from_results(results);
  if (res < 0) return fail(O2D_ERR_INVALID, "unknown results type %d", results);
  if (nsrc < 0 || npan < 0 || nt < 0) return fail(O2D_ERR_INVALID, "negative count");
  if ((nsrc > 0 && !src) || (npan > 0 && !pan)) return fail(O2D_ERR_INVALID, "NULL collection list");
  for (int k = 0; k < nsrc; ++k) {
    if (src[k].n < 0) return fail(O2D_ERR_INVALID, "negative count");
    if (src[k].n > 0 && (!src[k].x || !src[k].y || !src[k].s || !src[k].r)) return fail(O2D_ERR_INVALID, "NULL particle array");
  }
  for (int k = 0; k < npan; ++k)
    if (int rc = check_panels(pan[k].np, pan[k].nnodes, pan[k].nx, pan[k].ny, pan[k].idx, pan[k].vs)) return rc;
  if (nt == 0) return O2D_OK;
  if (!tx || !ty || !tu || !tv) return fail(O2D_ERR_INVALID, "NULL target array");
  if (res != RES_U && !tq) return fail(O2D_ERR_INVALID, "tq must not be NULL for vorticity/shear results");
  float* w = res == RES_UW ? tq : nullptr;
  float* e = res == RES_UE ? tq : nullptr;
  auto run = [&](const o2d_points_view* ds, const o2d_panels_view* dp, const float* dtx, const float* dty, const float* dtr,
                 float* dtu, float* dtv, float* dtq) -> int {
    CU(launch_zero(nt, dtu, dtv, res == RES_UW ? dtq : nullptr, res == RES_UE ? dtq : nullptr, c->stream));
    c->launches++;
    OutPtrs o{};
    o.p[0] = dtu; o.p[1] = dtv; o.p[2] = dtq;
    if (int rc = o2d_pp_multi_device(c, core, res, nsrc, ds, nt, dtx, dty, dtr, o)) return rc;
    if (int rc = o2d_panels_multi_device(c, npan, dp, nt, dtx, dty, dtu, dtv)) return rc;
    CU(launch_finalize(nt, dtu, dtv, res == RES_UW ? dtq : nullptr, res == RES_UE ? dtq : nullptr, fsx, fsy, c->stream));
    c->launches++;
    return O2D_OK;
  };
  (void)w; (void)e;
  if (mem == O2D_MEM_DEVICE) return run(src, pan, tx, ty, tr, tu, tv, tq);
  // host arrays: every collection goes up once, the target's results come down once
  size_t sbytes = 256, pbytes = 256, ibytes = 256;
  for (int k = 0; k < nsrc; ++k) sbytes += 4 * padded_floats(std::max(src[k].n, 0));
  for (int k = 0; k < npan; ++k) {
    for (int q = 0; q < 2 * pan[k].np; ++q)
      if (pan[k].idx[q] >= (uint32_t)pan[k].nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", pan[k].idx[q]);
    pbytes += 2 * padded_floats(pan[k].nnodes) + 2 * padded_floats(pan[k].np);
    ibytes += ((size_t)2 * pan[k].np * sizeof(uint32_t) + 255) / 256 * 256;
  }
  CU(c->hsrc.reserve(sbytes));
  CU(c->hpan.reserve(pbytes));
  CU(c->hidx.reserve(ibytes));
  CU(c->htgt.reserve(3 * padded_floats(nt)));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt)));
  Stager S(c, &c->hsrc), Pn(c, &c->hpan), I(c, &c->hidx), T(c, &c->htgt), O(c, &c->hout);
  std::vector<o2d_points_view> ds((size_t)nsrc);
  std::vector<o2d_panels_view> dp((size_t)npan);
  int rc;
  HostBatch bs(c), bp(c), bi(c), bt(c), bo(c);
  for (int k = 0; k < nsrc; ++k) {
    const int n = src[k].n;
    ds[k].n = n;
    if (n <= 0) continue;
    float *dx = S.take(n), *dy = S.take(n), *dss = S.take(n), *dr = S.take(n);
    bs.add(dx, src[k].x, n); bs.add(dy, src[k].y, n); bs.add(dss, src[k].s, n); bs.add(dr, src[k].r, n);
    ds[k].x = dx; ds[k].y = dy; ds[k].s = dss; ds[k].r = dr;
  }
  for (int k = 0; k < npan; ++k) {
    const int np = pan[k].np, nn = pan[k].nnodes;
    dp[k] = o2d_panels_view{np, nn, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (np <= 0) continue;
    float *dnx = Pn.take(nn), *dny = Pn.take(nn), *dvs = Pn.take(np), *dss = pan[k].ss ? Pn.take(np) : nullptr;
    uint32_t* di = reinterpret_cast<uint32_t*>(I.take((size_t)2 * np));
    bp.add(dnx, pan[k].nx, nn); bp.add(dny, pan[k].ny, nn); bp.add(dvs, pan[k].vs, np);
    if (dss) bp.add(dss, pan[k].ss, np);
    bi.add((void*)di, (const void*)pan[k].idx, (size_t)2 * np * sizeof(uint32_t));
    dp[k].nx = dnx; dp[k].ny = dny; dp[k].idx = di; dp[k].vs = dvs; dp[k].ss = dss;
  }
  float *dtx = T.take(nt), *dty = T.take(nt), *dtr = tr ? T.take(nt) : nullptr;
  float *dtu = O.take(nt), *dtv = O.take(nt), *dtq = tq ? O.take(nt) : nullptr;
  bt.add(dtx, tx, nt); bt.add(dty, ty, nt);
  if (tr) bt.add(dtr, tr, nt);
  if ((rc = bs.upload()) || (rc = bp.upload()) || (rc = bi.upload()) || (rc = bt.upload())) return rc;
  if ((rc = run(ds.data(), dp.data(), dtx, dty, dtr, dtu, dtv, dtq))) return rc;
  bo.add(dtu, tu, nt); bo.add(dtv, tv, nt);
  if (res != RES_U) bo.add(dtq, tq, nt);
  if ((rc = bo.download())) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

// One convection step of a single Lagrangian particle collection, entirely on the device.
static int advect_device(o2d_context* c, int core, int order, int n, float* x, float* y, const float* s,
                         const float* r, float* u, float* v, const Panels& pan, double fsx, double fsy, double dt) {
  const int nb = (n + 255) / 256;
  const float fdt = (float)dt;   // "(S)_dt" in ElementBase::move
  auto vels = [&](const float* px, const float* py, float* pu, float* pv) {
    return find_vels_device(c, core, RES_U, n, px, py, s, r, pan, n, px, py, r, pu, pv, nullptr, fsx, fsy);
  };
  int rc;
  if ((rc = vels(x, y, u, v))) return rc;                       // find_derivs at the initial state
  if (order == 1) {
    // advect_1st (src/Convection.h:277-308): elem.move(_time, _dt, 1.0, elem)
    move1_kernel<<<nb, 256, 0, c->stream>>>(n, x, y, x, y, u, v, fdt, 1.0);
    CU(cudaGetLastError());
    c->launches++;
    return O2D_OK;
  }
  // scratch: intermediate positions and velocities (interim_vort / vort1, vort2)
  CU(c->adv.reserve(8 * padded_floats(n)));
  Stager A(c, &c->adv);
  float *x1 = A.take(n), *y1 = A.take(n), *u1 = A.take(n), *v1 = A.take(n);
  float *x2 = A.take(n), *y2 = A.take(n), *u2 = A.take(n), *v2 = A.take(n);
  if (order == 2) {
    // advect_2nd_ralston (src/Convection.h:314-392)
    const double twothirds = 2.0 / 3.0;
    move1_kernel<<<nb, 256, 0, c->stream>>>(n, x, y, x1, y1, u, v, (float)(twothirds * dt), 1.0);
    CU(cudaGetLastError());
    c->launches++;
    if ((rc = vels(x1, y1, u1, v1))) return rc;
    move23_kernel<<<nb, 256, 0, c->stream>>>(n, x, y, u, v, u1, v1, nullptr, nullptr, fdt, 0.25, 0.75, 0.0);
    CU(cudaGetLastError());
    c->launches++;
    return O2D_OK;
  }
  // advect_3rd (src/Convection.h:472-595)
  move1_kernel<<<nb, 256, 0, c->stream>>>(n, x, y, x1, y1, u, v, (float)(0.5 * dt), 1.0);
  CU(cudaGetLastError());
  if ((rc = vels(x1, y1, u1, v1))) return rc;
  move1_kernel<<<nb, 256, 0, c->stream>>>(n, x, y, x2, y2, u1, v1, (float)(0.75 * dt), 1.0);
  CU(cudaGetLastError());
  if ((rc = vels(x2, y2, u2, v2))) return rc;
  move23_kernel<<<nb, 256, 0, c->stream>>>(n, x, y, u, v, u1, v1, u2, v2, fdt, 2.0 / 9.0, 3.0 / 9.0, 4.0 / 9.0);
  CU(cudaGetLastError());
  c->launches += 3;
  return O2D_OK;
}

extern "C" int o2d_advect_steps(o2d_context* c, int core, int order, int mem,
                                int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                                int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                const float* pvs, const float* pss, double fsx, double fsy, double dt, int nsteps) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (order < 1 || order > 3)
    return fail(O2D_ERR_INVALID, "Convection integrator orders over 3 unsupported (reference src/Convection.h:259)");
  if (n < 0 || nsteps < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (int rc = check_panels(np, nnodes, nx, ny, idx, pvs)) return rc;
  if (n == 0 || nsteps == 0) return O2D_OK;
  if (!x || !y || !s || !r || !u || !v) return fail(O2D_ERR_INVALID, "NULL particle array");
  Panels pan;
  int rc;
  if (mem == O2D_MEM_DEVICE) {
    pan.np = np; pan.nx = nx; pan.ny = ny; pan.idx = idx; pan.vs = pvs; pan.ss = pss;
    for (int k = 0; k < nsteps; ++k)
      if ((rc = advect_device(c, core, order, n, x, y, s, r, u, v, pan, fsx, fsy, dt))) return rc;
    return O2D_OK;
  }
  CU(c->hsrc.reserve(4 * padded_floats(n) + 256));
  CU(c->hout.reserve(kMaxOut * padded_floats(n)));
  Stager S(c, &c->hsrc), O(c, &c->hout);
  float *dx = S.take(n), *dy = S.take(n), *ds = S.take(n), *dr = S.take(n);
  float *du = O.take(n), *dv = O.take(n);
  if ((rc = o2d_h2d(c, dx, x, n)) || (rc = o2d_h2d(c, dy, y, n)) || (rc = o2d_h2d(c, ds, s, n)) ||
      (rc = o2d_h2d(c, dr, r, n)))
    return rc;
  if ((rc = stage_panels(c, np, nnodes, nx, ny, idx, pvs, pss, &pan))) return rc;
  // the particle state stays on the device between the steps: one upload, one download for the whole run
  for (int k = 0; k < nsteps; ++k)
    if ((rc = advect_device(c, core, order, n, dx, dy, ds, dr, du, dv, pan, fsx, fsy, dt))) return rc;
  if ((rc = o2d_d2h(c, x, dx, n)) || (rc = o2d_d2h(c, y, dy, n)) || (rc = o2d_d2h(c, u, du, n)) ||
      (rc = o2d_d2h(c, v, dv, n)))
    return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

extern "C" int o2d_advect(o2d_context* c, int core, int order, int mem,
                          int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                          int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                          const float* pvs, const float* pss, double fsx, double fsy, double dt) {
  return o2d_advect_steps(c, core, order, mem, n, x, y, s, r, u, v, np, nnodes, nx, ny, idx, pvs, pss, fsx, fsy, dt, 1);
}

// ---- convection step of a flow WITH bodies ---------------------------------------------------------------------
// The reference's find_derivs is  solve_bem -> find_vels  (src/Convection.h:229-243), and after every move of every
// Runge-Kutta stage it runs clear_inner_layer (:303-307,:334,:388-391,:489-578).  solve_bem's linear solve is the
// host application's (Eigen GMRES, src/BEM.h:166-232), but everything around it is on the path: the right-hand side
// is the particle-induced velocity at the panel centres (points_affect_panels + Surfaces::finalize_vels,
// src/BEMHelper.h:88-100), the solved strengths feed panels_affect_points, and the particles are then moved and
// pushed out of the body.  Here the particles stay on the GPU for the whole step; per stage only 2*Np floats go to
// the host (the panel velocities, for the caller's `solve`) and 2*Np come back (the sheet strengths).
namespace {

struct BodyDev {
  Panels pan;                 // geometry + strengths (device)
  const float* area = nullptr;
  const float* snx = nullptr;
  const float* sny = nullptr;
  float* vs = nullptr;        // strengths, rewritten after every solve
  float* ss = nullptr;
  float* pu = nullptr;        // panel-centre velocities
  float* pv = nullptr;
  int nnodes = 0;
};

int stage_derivs(o2d_context* c, int core, int n, const float* x, const float* y, const float* s, const float* r,
                 float* u, float* v, int nf, const float* fx, const float* fy, float* fu, float* fv,
                 BodyDev& b, float* h_pu, float* h_pv, float* h_vs, float* h_ss,
                 double fsx, double fsy, int stage, double time, o2d_bem_solve_fn solve, void* user) {
  const int np = b.pan.np;
  int rc;
  // solve_bem, right-hand side: zero_vels; points_affect_panels; finalize_vels(fs)  (src/BEMHelper.h:88-100)
  CU(launch_zero(np, b.pu, b.pv, nullptr, nullptr, c->stream));
  c->launches++;
  if ((rc = o2d_points_on_panels_device(c, n, x, y, s, np, b.pan.nx, b.pan.ny, b.pan.idx, b.area, b.pu, b.pv))) return rc;
  CU(launch_finalize(np, b.pu, b.pv, nullptr, nullptr, fsx, fsy, c->stream));
  c->launches++;
  {
    HostBatch down(c);
    down.add(b.pu, h_pu, (size_t)np); down.add(b.pv, h_pv, (size_t)np);
    if ((rc = down.download()) || (rc = o2d_finish_host(c))) return rc;   // waits for the stream: the stage's only host sync
  }
  // the caller's BEM: RHS from these velocities, linear solve, new sheet strengths
  if (int urc = solve(user, stage, time, np, h_pu, h_pv, h_vs, h_ss))
    return fail(O2D_ERR_INVALID, "o2d_advect_bodies: the solve callback returned %d at stage %d", urc, stage);
  {
    HostBatch up(c);
    up.add(b.vs, h_vs, (size_t)np);
    if (b.ss) up.add(b.ss, h_ss, (size_t)np);
    if ((rc = up.upload())) return rc;
  }
  // find_vels(fs, vort, bdry, vort); find_vels(fs, vort, bdry, fldpt)   (src/Convection.h:241-242)
  if ((rc = find_vels_device(c, core, RES_U, n, x, y, s, r, b.pan, n, x, y, r, u, v, nullptr, fsx, fsy))) return rc;
  if (nf > 0) return find_vels_device(c, core, RES_U, n, x, y, s, r, b.pan, nf, fx, fy, nullptr, fu, fv, nullptr, fsx, fsy);
  return O2D_OK;
}

}  // namespace

extern "C" int o2d_advect_bodies(o2d_context* c, int core, int order, int mem,
                                 int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                                 int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                 const float* area, const float* snx, const float* sny, float* pvs, float* pss,
                                 double fsx, double fsy, double time, double dt, float cutoff_mult, float ips,
                                 int nf, float* fx, float* fy, float* fu, float* fv,
                                 o2d_bem_solve_fn solve, void* user) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (nf < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (nf > 0 && (!fx || !fy || !fu || !fv)) return fail(O2D_ERR_INVALID, "NULL field-point array");
  if (mem != O2D_MEM_HOST)
    return fail(O2D_ERR_UNSUPPORTED, "o2d_advect_bodies takes host arrays (the solve callback runs on the host every stage)");
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (order < 1 || order > 3)
    return fail(O2D_ERR_INVALID, "Convection integrator orders over 3 unsupported (reference src/Convection.h:259)");
  if (n < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (np <= 0) return fail(O2D_ERR_INVALID, "o2d_advect_bodies needs a surface collection (particle-only flows: o2d_advect)");
  if (int rc = check_panels(np, nnodes, nx, ny, idx, pvs)) return rc;
  if (!area || !snx || !sny || !solve) return fail(O2D_ERR_INVALID, "NULL panel array or callback");
  if (n == 0) return O2D_OK;
  if (!x || !y || !s || !r || !u || !v) return fail(O2D_ERR_INVALID, "NULL particle array");
  for (int k = 0; k < 2 * np; ++k)
    if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
  int rc;
  // ---- one upload: particles and the panel description -------------------------------------------------------------
  CU(c->hsrc.reserve(4 * padded_floats(n) + 256));
  CU(c->hout.reserve(kMaxOut * padded_floats(n)));
  CU(c->adv.reserve(8 * padded_floats(n)));
  CU(c->hpan.reserve(2 * padded_floats(nnodes) + 7 * padded_floats(np)));
  CU(c->hidx.reserve((size_t)2 * np * sizeof(uint32_t)));
  CU(c->htgt.reserve(12 * padded_floats(nf) + 256));
  Stager S(c, &c->hsrc), O(c, &c->hout), A(c, &c->adv), Pn(c, &c->hpan), F(c, &c->htgt);
  // field points (inert Lagrangian tracers): positions, velocities and the intermediate stages
  float *dfx = F.take(nf), *dfy = F.take(nf), *dfu = F.take(nf), *dfv = F.take(nf);
  float *fx1 = F.take(nf), *fy1 = F.take(nf), *fu1 = F.take(nf), *fv1 = F.take(nf);
  float *fx2 = F.take(nf), *fy2 = F.take(nf), *fu2 = F.take(nf), *fv2 = F.take(nf);
  float *dx = S.take(n), *dy = S.take(n), *ds = S.take(n), *dr = S.take(n);
  float *du = O.take(n), *dv = O.take(n);
  float *x1 = A.take(n), *y1 = A.take(n), *u1 = A.take(n), *v1 = A.take(n);
  float *x2 = A.take(n), *y2 = A.take(n), *u2 = A.take(n), *v2 = A.take(n);
  float *dnx = Pn.take(nnodes), *dny = Pn.take(nnodes);
  float *darea = Pn.take(np), *dsnx = Pn.take(np), *dsny = Pn.take(np);
  BodyDev b;
  b.vs = Pn.take(np); b.ss = pss ? Pn.take(np) : nullptr;
  b.pu = Pn.take(np); b.pv = Pn.take(np);
  {
    HostBatch bp(c), bb(c);
    bp.add(dx, x, n); bp.add(dy, y, n); bp.add(ds, s, n); bp.add(dr, r, n);
    bb.add(dnx, nx, nnodes); bb.add(dny, ny, nnodes); bb.add(darea, area, np); bb.add(dsnx, snx, np); bb.add(dsny, sny, np);
    if ((rc = bp.upload()) || (rc = bb.upload())) return rc;
    if (nf > 0) {
      HostBatch bf(c);
      bf.add(dfx, fx, nf); bf.add(dfy, fy, nf);
      if ((rc = bf.upload())) return rc;
    }
  }
  if ((rc = o2d_h2d_bytes(c, c->hidx.p, idx, (size_t)2 * np * sizeof(uint32_t)))) return rc;
  b.pan.np = np; b.pan.nx = dnx; b.pan.ny = dny; b.pan.idx = c->hidx.as<uint32_t>(); b.pan.vs = b.vs; b.pan.ss = b.ss;
  b.area = darea; b.snx = dsnx; b.sny = dsny; b.nnodes = nnodes;
  // host side of the per-stage exchange: 2*Np floats out (one DMA), 2*Np back (one DMA)
  std::vector<float> hbuf(2 * (size_t)np);
  float* h_pu = hbuf.data();
  float* h_pv = h_pu + np;
  const int nb = (n + 255) / 256, nbf = (nf + 255) / 256;
  const float fdt = (float)dt;   // "(S)_dt" in ElementBase::move
  auto derivs = [&](int stage, double t, const float* px, const float* py, float* pu_, float* pv_,
                    const float* qx, const float* qy, float* qu, float* qv) {
    return stage_derivs(c, core, n, px, py, ds, dr, pu_, pv_, nf, qx, qy, qu, qv, b, h_pu, h_pv, pvs, pss, fsx, fsy, stage, t,
                        solve, user);
  };
  if ((rc = o2d_cil_prepare(c, np, nnodes, dnx, dny, b.pan.idx, dsnx, dsny))) return rc;   // the body is fixed within the step
  auto clear = [&](int m, float* px, float* py) { return m > 0 ? o2d_cil_run(c, 1, m, px, py, cutoff_mult, ips, nullptr) : O2D_OK; };
  // ElementBase::move for the particles and, with the same weights, for the field points
  auto move1 = [&](const float* xi, const float* yi, float* xo, float* yo, const float* u_, const float* v_,
                   const float* fxi, const float* fyi, float* fxo, float* fyo, const float* fu_, const float* fv_, float sdt) {
    move1_kernel<<<nb, 256, 0, c->stream>>>(n, xi, yi, xo, yo, u_, v_, sdt, 1.0);
    c->launches++;
    if (nf > 0) { move1_kernel<<<nbf, 256, 0, c->stream>>>(nf, fxi, fyi, fxo, fyo, fu_, fv_, sdt, 1.0); c->launches++; }
    return cudaGetLastError();
  };
  auto move23 = [&](const float* ua, const float* va, const float* ub, const float* vb, const float* fua, const float* fva,
                    const float* fub, const float* fvb, double w1, double w2, double w3) {
    move23_kernel<<<nb, 256, 0, c->stream>>>(n, dx, dy, du, dv, ua, va, ub, vb, fdt, w1, w2, w3);
    c->launches++;
    if (nf > 0) { move23_kernel<<<nbf, 256, 0, c->stream>>>(nf, dfx, dfy, dfu, dfv, fua, fva, fub, fvb, fdt, w1, w2, w3); c->launches++; }
    return cudaGetLastError();
  };
  if ((rc = derivs(0, time, dx, dy, du, dv, dfx, dfy, dfu, dfv))) return rc;
  if (order == 1) {
    // advect_1st (src/Convection.h:277-308)
    CU(move1(dx, dy, dx, dy, du, dv, dfx, dfy, dfx, dfy, dfu, dfv, fdt));
    if ((rc = clear(n, dx, dy)) || (rc = clear(nf, dfx, dfy))) return rc;
  } else if (order == 2) {
    // advect_2nd_ralston (src/Convection.h:314-392)
    const double twothirds = 2.0 / 3.0;
    CU(move1(dx, dy, x1, y1, du, dv, dfx, dfy, fx1, fy1, dfu, dfv, (float)(twothirds * dt)));
    if ((rc = clear(n, x1, y1)) || (rc = clear(nf, fx1, fy1))) return rc;
    if ((rc = derivs(1, time + twothirds * dt, x1, y1, u1, v1, fx1, fy1, fu1, fv1))) return rc;
    CU(move23(u1, v1, nullptr, nullptr, fu1, fv1, nullptr, nullptr, 0.25, 0.75, 0.0));
    if ((rc = clear(n, dx, dy)) || (rc = clear(nf, dfx, dfy))) return rc;
  } else {
    // advect_3rd (src/Convection.h:472-595)
    CU(move1(dx, dy, x1, y1, du, dv, dfx, dfy, fx1, fy1, dfu, dfv, (float)(0.5 * dt)));
    if ((rc = clear(n, x1, y1)) || (rc = clear(nf, fx1, fy1))) return rc;
    if ((rc = derivs(1, time + 0.5 * dt, x1, y1, u1, v1, fx1, fy1, fu1, fv1))) return rc;
    CU(move1(dx, dy, x2, y2, u1, v1, dfx, dfy, fx2, fy2, fu1, fv1, (float)(0.75 * dt)));
    if ((rc = clear(n, x2, y2)) || (rc = clear(nf, fx2, fy2))) return rc;
    if ((rc = derivs(2, time + 0.75 * dt, x2, y2, u2, v2, fx2, fy2, fu2, fv2))) return rc;
    CU(move23(u1, v1, u2, v2, fu1, fv1, fu2, fv2, 2.0 / 9.0, 3.0 / 9.0, 4.0 / 9.0));
    if ((rc = clear(n, dx, dy)) || (rc = clear(nf, dfx, dfy))) return rc;
  }
  // ---- one download -------------------------------------------------------------------------------------------------------
  {
    HostBatch bx(c), bu(c);
    bx.add(dx, x, n); bx.add(dy, y, n);
    bu.add(du, u, n); bu.add(dv, v, n);
    if ((rc = bx.download()) || (rc = bu.download())) return rc;
    if (nf > 0) {
      HostBatch bf(c);
      bf.add(dfx, fx, nf); bf.add(dfy, fy, nf); bf.add(dfu, fu, nf); bf.add(dfv, fv, nf);
      if ((rc = bf.download())) return rc;
    }
  }
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}
