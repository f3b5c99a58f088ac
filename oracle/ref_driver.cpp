// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" shim around the UNMODIFIED Omega2D reference headers, built by
// oracle/build_ref.sh into oracle/_ref/libo2dref_<core>.so (git-ignored).  One
// shared object per core function because the reference selects the core with
// an unguarded #define (reference src/CoreFunc.h:35-39).
//
// Every entry point builds the reference's own containers (Points<float>,
// Surfaces<float>) from raw SoA arrays, calls the reference's own templates
// with <S=float, A=double> (the USE_VC-off configuration, reference
// src/Simulation.h:44-50), and copies the results back with the reference's
// accumulate ("+=") semantics.  Nothing here restates reference arithmetic
// except the two loops the reference never wraps itself:
//   * velocity+gradient: kernelug_0v_0b is never called by any reference loop
//     (SURVEY F5), so the points_affect_points loop nest (src/Influence.h:347-361)
//     is restated around the reference's own kernelug_0v_0b (src/Kernels.h:112-133);
//   * vortex-only panels: Surfaces always engages the source strengths
//     (src/Surfaces.h:87-92), so the have_source_strengths==false branch of
//     panels_affect_points (src/Influence.h:576-594) is restated around the
//     reference's own kernelu_1v_0p (src/Kernels.h:312-354).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load this library.

// Collection.h must come first: Influence.h is not self-contained.
#include "Collection.h"
#include "Influence.h"
#include "InfluenceVort.h"
#include "Coefficients.h"
#include "RHS.h"
#include "Reflect.h"

#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <iostream>
#include <unistd.h>
#include <vector>

namespace {

// The reference narrates on std::cout and printf from inside the hot functions
// (src/Influence.h:59,115,412-415).  Silence fd 1 for the duration of a call so
// that a harness printing one JSON line on stdout stays clean.
struct StdoutMute {
  int saved = -1;
  std::streambuf* cout_buf = nullptr;
  StdoutMute() {
    fflush(stdout);
    std::cout.flush();
    saved = dup(1);
    int devnull = open("/dev/null", O_WRONLY);
    if (devnull >= 0) { dup2(devnull, 1); close(devnull); }
  }
  ~StdoutMute() {
    fflush(stdout);
    std::cout.flush();
    if (saved >= 0) { dup2(saved, 1); close(saved); }
  }
};

Points<float> make_points(int n, const float* x, const float* y, const float* s, const float* r) {
  // ElementPacket carries interleaved xy (src/ElementPacket.h:166-171)
  std::vector<float> px(2 * (size_t)n);
  for (int i = 0; i < n; ++i) { px[2*i] = x[i]; px[2*i+1] = y[i]; }
  if (s == nullptr) {
    // inert field points: no strength, no radius (src/Points.h:93-95)
    ElementPacket<float> pk(px, std::vector<Int>(), std::vector<float>(), (size_t)n, 0);
    return Points<float>(pk, inert, fixed, nullptr, 0.0f);
  }
  std::vector<float> val(s, s + n);
  ElementPacket<float> pk(px, std::vector<Int>(), val, (size_t)n, 0);
  Points<float> pts(pk, active, lagrangian, nullptr, 1.0f);
  if (r != nullptr) {
    Vector<float>& rad = pts.get_rad();
    for (int i = 0; i < n; ++i) rad[i] = r[i];
  }
  return pts;
}

Surfaces<float> make_surfaces(int np, int nnodes, const float* nx, const float* ny,
                              const uint32_t* idx, const float* vs, const float* ss) {
  std::vector<float> px(2 * (size_t)nnodes);
  for (int i = 0; i < nnodes; ++i) { px[2*i] = nx[i]; px[2*i+1] = ny[i]; }
  std::vector<Int> pidx(idx, idx + 2 * (size_t)np);
  std::vector<float> val(2 * (size_t)np);
  for (int i = 0; i < np; ++i) { val[2*i] = vs ? vs[i] : 0.0f; val[2*i+1] = ss ? ss[i] : 0.0f; }
  ElementPacket<float> pk(px, pidx, val, (size_t)np, 1);
  return Surfaces<float>(pk, active, fixed, nullptr);
}

}  // namespace

extern "C" {

// which core this object was built with: 0 RM, 1 EXP, 2 WL, 3 V2, 4 V3
int o2dref_core_id() {
#if defined(USE_RM_KERNEL)
  return 0;
#elif defined(USE_EXPONENTIAL_KERNEL)
  return 1;
#elif defined(USE_WL_KERNEL)
  return 2;
#elif defined(USE_V2_KERNEL)
  return 3;
#elif defined(USE_V3_KERNEL)
  return 4;
#else
  return -1;
#endif
}

// reference per-pair flop counters (src/Kernels.h flops*_...(), src/CoreFunc.h flops_t*_*())
// which: 0 u_0p, 1 u_0b, 2 uw_0p, 3 uw_0b, 4 ue_0p, 5 ue_0b, 6 ug_0b, 7 u_1v_0p, 8 u_1vs_0p
int o2dref_flops(int which) {
  switch (which) {
    case 0: return (int)flopsu_0v_0p<float,double>();
    case 1: return (int)flopsu_0v_0b<float,double>();
    case 2: return (int)flopsuw_0v_0p<float,double>();
    case 3: return (int)flopsuw_0v_0b<float,double>();
    case 4: return (int)flopsue_0v_0p<float,double>();
    case 5: return (int)flopsue_0v_0b<float,double>();
    case 6: return (int)flopsug_0v_0b<float,double>();
    case 7: return (int)flopsu_1v_0p<float,double>();
    case 8: return (int)flopsu_1vs_0p<float,double>();
  }
  return -1;
}

// reference points_affect_points<float,double> (src/Influence.h:56-416).
// results_type uses the reference's results_t values (src/ResultsType.h:25-31):
// 1 velonly, 4 velandvort, 5 velandshear.  tr==NULL => inert targets (0p kernels).
// tu,tv (and tq for vort/shear) are in/out, accumulated into like the reference does.
int o2dref_points_on_points(int results_type,
                            int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                            int nt, const float* tx, const float* ty, const float* tr,
                            float* tu, float* tv, float* tq) {
  if (results_type != velonly && results_type != velandvort && results_type != velandshear) return -1;
  StdoutMute mute;
  Points<float> src = make_points(ns, sx, sy, ss, sr);
  Points<float> targ = tr ? make_points(nt, tx, ty, tx /*dummy strengths*/, tr)
                          : make_points(nt, tx, ty, nullptr, nullptr);
  auto& u = targ.get_vel();
  for (int i = 0; i < nt; ++i) { u[0][i] = tu[i]; u[1][i] = tv[i]; }
  if (results_type == velandvort) { auto& w = targ.get_vort();  for (int i = 0; i < nt; ++i) w[i] = tq[i]; }
  if (results_type == velandshear) { auto& e = targ.get_shear(); for (int i = 0; i < nt; ++i) e[i] = tq[i]; }
  ExecEnv env(true, direct, cpu_x86);
  points_affect_points<float,double>(src, targ, ResultsType((results_t)results_type), env);
  for (int i = 0; i < nt; ++i) { tu[i] = u[0][i]; tv[i] = u[1][i]; }
  if (results_type == velandvort) { auto& w = targ.get_vort();  for (int i = 0; i < nt; ++i) tq[i] = w[i]; }
  if (results_type == velandshear) { auto& e = targ.get_shear(); for (int i = 0; i < nt; ++i) tq[i] = e[i]; }
  return 0;
}

// velocity + 2x2 gradient: restated loop (shape of src/Influence.h:347-361) around the
// reference's own kernelug_0v_0b (src/Kernels.h:112-133).  skip_self!=0 drops i==j when
// source and target arrays are the same set (needed for the EXP core).
int o2dref_points_on_points_grad(int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                 int nt, const float* tx, const float* ty, const float* tr,
                                 float* tu, float* tv, float* tux, float* tvx, float* tuy, float* tvy,
                                 int skip_self) {
  #pragma omp parallel for
  for (int32_t i = 0; i < (int32_t)nt; ++i) {
    double au = 0.0, av = 0.0, aux = 0.0, avx = 0.0, auy = 0.0, avy = 0.0;
    for (int j = 0; j < ns; ++j) {
      if (skip_self && j == i) continue;
      kernelug_0v_0b<float,double>(sx[j], sy[j], sr[j], ss[j], tx[i], ty[i], tr[i],
                                   &au, &av, &aux, &avx, &auy, &avy);
    }
    tu[i] += au;  tv[i] += av;
    tux[i] += aux; tvx[i] += avx; tuy[i] += auy; tvy[i] += avy;
  }
  return 0;
}

// reference panels_affect_points<float,double> (src/Influence.h:422-614); always the
// vortex+source kernel kernelu_1vs_0p because Surfaces engages ps[1].
int o2dref_panels_on_points(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                            const float* vs, const float* ss,
                            int nt, const float* tx, const float* ty, float* tu, float* tv) {
  StdoutMute mute;
  Surfaces<float> src = make_surfaces(np, nnodes, nx, ny, idx, vs, ss);
  Points<float> targ = make_points(nt, tx, ty, nullptr, nullptr);
  auto& u = targ.get_vel();
  for (int i = 0; i < nt; ++i) { u[0][i] = tu[i]; u[1][i] = tv[i]; }
  ExecEnv env(true, direct, cpu_x86);
  panels_affect_points<float,double>(src, targ, ResultsType(velonly), env);
  for (int i = 0; i < nt; ++i) { tu[i] = u[0][i]; tv[i] = u[1][i]; }
  return 0;
}

// vortex-only panels: restated have_source_strengths==false branch (src/Influence.h:576-594)
// around the reference's own kernelu_1v_0p (src/Kernels.h:312-354).
int o2dref_panels_on_points_vortex(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                   const float* vs,
                                   int nt, const float* tx, const float* ty, float* tu, float* tv) {
  (void)nnodes;
  #pragma omp parallel for
  for (int32_t i = 0; i < (int32_t)nt; ++i) {
    double au = 0.0, av = 0.0, ru = 0.0, rv = 0.0;
    for (int j = 0; j < np; ++j) {
      const size_t j0 = idx[2*j], j1 = idx[2*j+1];
      kernelu_1v_0p<float,double>(nx[j0], ny[j0], nx[j1], ny[j1], vs[j], tx[i], ty[i], &ru, &rv);
      au += ru; av += rv;
    }
    tu[i] += au; tv[i] += av;
  }
  return 0;
}

// reference points_affect_panels<float,double> (src/Influence.h:633-771): particle
// velocities on panel centres by reciprocity, pu -= (1/area) * sum.  Panel areas are
// the ones Surfaces::compute_bases derives (src/Surfaces.h:682-726) and are returned
// through `area` (may be NULL) so a caller can feed the same values elsewhere.
int o2dref_points_on_panels(int ns, const float* sx, const float* sy, const float* ss,
                            int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                            float* pu, float* pv, float* area) {
  StdoutMute mute;
  std::vector<float> rad((size_t)ns, 1.0f);
  Points<float> src = make_points(ns, sx, sy, ss, rad.data());
  Surfaces<float> targ = make_surfaces(np, nnodes, nx, ny, idx, nullptr, nullptr);
  auto& u = targ.get_vel();
  for (int i = 0; i < np; ++i) { u[0][i] = pu[i]; u[1][i] = pv[i]; }
  ExecEnv env(true, direct, cpu_x86);
  points_affect_panels<float,double>(src, targ, ResultsType(velonly), env);
  for (int i = 0; i < np; ++i) { pu[i] = u[0][i]; pv[i] = u[1][i]; }
  if (area) { const auto& a = targ.get_area(); for (int i = 0; i < np; ++i) area[i] = a[i]; }
  return 0;
}

// reference points_affect_points_vorticity<float,double> (src/InfluenceVort.h:47-126)
int o2dref_points_vorticity(int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                            int nt, const float* tx, const float* ty, float* tw) {
  StdoutMute mute;
  Points<float> src = make_points(ns, sx, sy, ss, sr);
  Points<float> targ = make_points(nt, tx, ty, nullptr, nullptr);
  auto& w = targ.get_vort();
  for (int i = 0; i < nt; ++i) w[i] = tw[i];
  ExecEnv env(true, direct, cpu_x86);
  points_affect_points_vorticity<float,double>(src, targ, env);
  for (int i = 0; i < nt; ++i) tw[i] = w[i];
  return 0;
}

// reference ElementBase::finalize_vels (src/ElementBase.h:215-234) through Points
int o2dref_finalize_vels(int n, float* u, float* v, float* w, float* e, double fsx, double fsy) {
  StdoutMute mute;
  std::vector<float> zeros((size_t)n, 0.0f);
  Points<float> p = make_points(n, zeros.data(), zeros.data(), nullptr, nullptr);
  auto& vel = p.get_vel();
  for (int i = 0; i < n; ++i) { vel[0][i] = u[i]; vel[1][i] = v[i]; }
  if (w) { auto& pw = p.get_vort();  for (int i = 0; i < n; ++i) pw[i] = w[i]; }
  if (e) { auto& pe = p.get_shear(); for (int i = 0; i < n; ++i) pe[i] = e[i]; }
  p.finalize_vels(std::array<double,Dimensions>({fsx, fsy}));
  for (int i = 0; i < n; ++i) { u[i] = vel[0][i]; v[i] = vel[1][i]; }
  if (w) { auto& pw = p.get_vort();  for (int i = 0; i < n; ++i) w[i] = pw[i]; }
  if (e) { auto& pe = p.get_shear(); for (int i = 0; i < n; ++i) e[i] = pe[i]; }
  return 0;
}


// ---- find_vels and the convection step --------------------------------------------------------
// Convection.h itself cannot be compiled here (it pulls in BEM.h -> <Eigen/Dense>), so
// the ORCHESTRATION of Convection::find_vels (src/Convection.h:172-223) and of advect_1st /
// advect_2nd_ralston / advect_3rd (:277-595) is restated below -- around the reference's own
// InfluenceVisitor (src/Influence.h:885-900), its own zero_vels / finalize_vels and its own
// Points::move (src/Points.h:214-252 -> src/ElementBase.h:276-359), operating on the reference's own
// std::vector<Collection>.  Left out, as in the product entry point: solve_bem (no Eigen) and
// clear_inner_layer; both are no-ops for a particle-only flow.

static void ref_find_vels(const std::array<double,Dimensions>& fs, std::vector<Collection>& vort,
                          std::vector<Collection>& bdry, std::vector<Collection>& targets,
                          const results_t results) {
  InfluenceVisitor<float,double> visitor = {ResultsType(results), ExecEnv(true, direct, cpu_x86)};
  for (auto& src : bdry) std::visit([=](auto& elem) { elem.add_solved_rot_strengths(1.0); }, src);
  for (auto& targ : targets) {
    std::visit([=](auto& elem) { elem.zero_vels(); }, targ);
    for (auto& src : vort) std::visit(visitor, src, targ);
    for (auto& src : bdry) std::visit(visitor, src, targ);
    std::visit([=](auto& elem) { elem.finalize_vels(fs); }, targ);
  }
  for (auto& src : bdry) std::visit([=](auto& elem) { elem.add_solved_rot_strengths(-1.0); }, src);
}

static void get_xu(Collection& c, int n, float* x, float* y, float* u, float* v) {
  Points<float>& p = std::get<Points<float>>(c);
  const auto& px = p.get_pos();
  const auto& pu = p.get_vel();
  for (int i = 0; i < n; ++i) {
    if (x) { x[i] = px[0][i]; y[i] = px[1][i]; }
    if (u) { u[i] = pu[0][i]; v[i] = pu[1][i]; }
  }
}

int o2dref_find_vels(int results,
                     int nv, const float* vx, const float* vy, const float* vs, const float* vr,
                     int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                     const float* pvs, const float* pss,
                     int nt, const float* tx, const float* ty, const float* tr,
                     float* tu, float* tv, float* tq, double fsx, double fsy) {
  if (results != velonly && results != velandvort && results != velandshear) return -1;
  StdoutMute mute;
  std::vector<Collection> vort, bdry, targ;
  if (nv > 0) vort.push_back(make_points(nv, vx, vy, vs, vr));
  if (np > 0) bdry.push_back(make_surfaces(np, nnodes, nx, ny, idx, pvs, pss));
  targ.push_back(tr ? make_points(nt, tx, ty, tx /*dummy strengths*/, tr) : make_points(nt, tx, ty, nullptr, nullptr));
  Points<float>& t = std::get<Points<float>>(targ[0]);
  if (results == velandvort) (void)t.get_vort();    // allocate the optional result arrays
  if (results == velandshear) (void)t.get_shear();
  ref_find_vels({fsx, fsy}, vort, bdry, targ, (results_t)results);
  get_xu(targ[0], nt, nullptr, nullptr, tu, tv);
  if (results == velandvort) { auto& w = t.get_vort();  for (int i = 0; i < nt; ++i) tq[i] = w[i]; }
  if (results == velandshear) { auto& e = t.get_shear(); for (int i = 0; i < nt; ++i) tq[i] = e[i]; }
  return 0;
}

int o2dref_advect(int order, int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                  int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                  const float* pvs, const float* pss, double fsx, double fsy, double dt) {
  if (order < 1 || order > 3) return -1;
  StdoutMute mute;
  const std::array<double,Dimensions> fs = {fsx, fsy};
  const double time = 0.0;
  std::vector<Collection> vort, bdry;
  vort.push_back(make_points(n, x, y, s, r));
  if (np > 0) bdry.push_back(make_surfaces(np, nnodes, nx, ny, idx, pvs, pss));

  ref_find_vels(fs, vort, bdry, vort, velonly);
  if (order == 1) {
    for (auto& coll : vort) std::visit([=](auto& elem) { elem.move(time, dt, 1.0, elem); }, coll);
  } else if (order == 2) {
    const double twothirds = 2.0/3.0;
    std::vector<Collection> interim = vort;
    for (auto& coll : interim) std::visit([=](auto& elem) { elem.move(time, twothirds*dt, 1.0, elem); }, coll);
    ref_find_vels(fs, interim, bdry, interim, velonly);
    Points<float>& p1 = std::get<Points<float>>(vort[0]);
    Points<float>& p2 = std::get<Points<float>>(interim[0]);
    p1.move(time, dt, 0.25, p1, 0.75, p2);
  } else {
    std::vector<Collection> vort1 = vort;
    for (auto& coll : vort1) std::visit([=](auto& elem) { elem.move(time, 0.5*dt, 1.0, elem); }, coll);
    ref_find_vels(fs, vort1, bdry, vort1, velonly);
    std::vector<Collection> vort2 = vort;
    {
      Points<float>& p1 = std::get<Points<float>>(vort1[0]);
      Points<float>& p2 = std::get<Points<float>>(vort2[0]);
      p2.move(time, 0.75*dt, 1.0, p1);
    }
    ref_find_vels(fs, vort2, bdry, vort2, velonly);
    Points<float>& p0 = std::get<Points<float>>(vort[0]);
    Points<float>& p1 = std::get<Points<float>>(vort1[0]);
    Points<float>& p2 = std::get<Points<float>>(vort2[0]);
    p0.move(time, dt, 2.0/9.0, p0, 3.0/9.0, p1, 4.0/9.0, p2);
  }
  get_xu(vort[0], n, x, y, u, v);
  return 0;
}


// ---- the convection step of a flow with bodies ---------------------------------------------------------------
// advect_1st / advect_2nd_ralston / advect_3rd (src/Convection.h:277-595) WITH the per-stage solve_bem and
// clear_inner_layer.  Restated orchestration (Convection.h / BEMHelper.h need Eigen) around the reference's own
// pieces: for solve_bem (src/BEMHelper.h:43-260) the right-hand-side part -- zero_vels, InfluenceVisitor
// (points_affect_panels), Surfaces::finalize_vels -- runs on the reference's own Surfaces; the linear solve itself
// (BEM::solve, Eigen) is replaced by the caller's `solve` callback, which receives those panel velocities and returns
// the sheet strengths that the reference would copy back with set_str (:237-254); then the reference's find_vels
// orchestration, Points::move and clear_inner_layer<float>(1, ...) (src/Reflect.h:735-768).
typedef int (*bem_solve_fn)(void* user, int stage, double time, int np, const float* pu, const float* pv, float* vs, float* ss);

static int ref_derivs(const std::array<double,Dimensions>& fs, std::vector<Collection>& vort, std::vector<Collection>& bdry,
                      std::vector<Collection>& fldpt, int stage, double time, bem_solve_fn solve, void* user, float* vs, float* ss) {
  InfluenceVisitor<float,double> ivisitor = {ResultsType(velonly), ExecEnv(true, direct, cpu_x86)};
  Surfaces<float>& surf = std::get<Surfaces<float>>(bdry[0]);
  const int np = (int)surf.get_npanels();
  surf.zero_vels();
  for (auto& src : vort) std::visit(ivisitor, src, bdry[0]);
  surf.finalize_vels(fs);
  const auto& pu = surf.get_vel();
  std::vector<float> hu(pu[0].begin(), pu[0].end()), hv(pu[1].begin(), pu[1].end());
  if (int rc = solve(user, stage, time, np, hu.data(), hv.data(), vs, ss)) return rc;
  auto& gv = surf.get_vort_str();
  for (int i = 0; i < np; ++i) gv[i] = vs[i];
  if (ss) { auto& gs = surf.get_src_str(); for (int i = 0; i < np; ++i) gs[i] = ss[i]; }
  ref_find_vels(fs, vort, bdry, vort, velonly);       // find_derivs, src/Convection.h:241-242
  ref_find_vels(fs, vort, bdry, fldpt, velonly);
  return 0;
}

// nf > 0: a collection of inert Lagrangian field points rides along, as _fldpt does in the reference's advect_*
int o2dref_advect_bodies(int order, int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx, float* vs, float* ss,
                         double fsx, double fsy, double time, double dt, float ips,
                         int nf, float* fx, float* fy, float* fu, float* fv, bem_solve_fn solve, void* user) {
  if (order < 1 || order > 3 || np <= 0) return -1;
  StdoutMute mute;
  const std::array<double,Dimensions> fs = {fsx, fsy};
  const float cutoff = 0.5/std::sqrt(2.0*M_PI);
  std::vector<Collection> vort, bdry, fldpt;
  vort.push_back(make_points(n, x, y, s, r));
  bdry.push_back(make_surfaces(np, nnodes, nx, ny, idx, vs, ss));
  if (nf > 0) {
    std::vector<float> pf(2 * (size_t)nf);
    for (int i = 0; i < nf; ++i) { pf[2*i] = fx[i]; pf[2*i+1] = fy[i]; }
    fldpt.push_back(Points<float>(ElementPacket<float>(pf, std::vector<Int>(), std::vector<float>(), (size_t)nf, 0),
                                  inert, lagrangian, nullptr, 0.0f));
  }
  auto move_all = [&](std::vector<Collection>& c, double sdt) {
    for (auto& coll : c) std::visit([=](auto& elem) { elem.move(time, sdt, 1.0, elem); }, coll);
  };
  auto move2 = [&](std::vector<Collection>& c0, std::vector<Collection>& c1, double w0, double w1) {
    for (size_t i = 0; i < c0.size(); ++i) {
      Points<float>& p0 = std::get<Points<float>>(c0[i]);
      Points<float>& p1 = std::get<Points<float>>(c1[i]);
      p0.move(time, dt, w0, p0, w1, p1);
    }
  };
  if (ref_derivs(fs, vort, bdry, fldpt, 0, time, solve, user, vs, ss)) return -2;
  if (order == 1) {
    move_all(vort, dt); move_all(fldpt, dt);
    clear_inner_layer<float>(1, bdry, vort, cutoff, ips);
    clear_inner_layer<float>(1, bdry, fldpt, cutoff, ips);
  } else if (order == 2) {
    const double twothirds = 2.0/3.0;
    std::vector<Collection> interim = vort;
    move_all(interim, twothirds*dt);
    clear_inner_layer<float>(1, bdry, interim, cutoff, ips);
    std::vector<Collection> interim_f = fldpt;
    move_all(interim_f, twothirds*dt);
    clear_inner_layer<float>(1, bdry, interim_f, cutoff, ips);
    if (ref_derivs(fs, interim, bdry, interim_f, 1, time + twothirds*dt, solve, user, vs, ss)) return -2;
    move2(vort, interim, 0.25, 0.75);
    move2(fldpt, interim_f, 0.25, 0.75);
    clear_inner_layer<float>(1, bdry, vort, cutoff, ips);
    clear_inner_layer<float>(1, bdry, fldpt, cutoff, ips);
  } else {
    std::vector<Collection> vort1 = vort, fld1 = fldpt;
    move_all(vort1, 0.5*dt);
    clear_inner_layer<float>(1, bdry, vort1, cutoff, ips);
    move_all(fld1, 0.5*dt);
    clear_inner_layer<float>(1, bdry, fld1, cutoff, ips);
    if (ref_derivs(fs, vort1, bdry, fld1, 1, time + 0.5*dt, solve, user, vs, ss)) return -2;
    std::vector<Collection> vort2 = vort, fld2 = fldpt;
    for (size_t i = 0; i < vort.size(); ++i)
      std::get<Points<float>>(vort2[i]).move(time, 0.75*dt, 1.0, std::get<Points<float>>(vort1[i]));
    clear_inner_layer<float>(1, bdry, vort2, cutoff, ips);
    for (size_t i = 0; i < fldpt.size(); ++i)
      std::get<Points<float>>(fld2[i]).move(time, 0.75*dt, 1.0, std::get<Points<float>>(fld1[i]));
    clear_inner_layer<float>(1, bdry, fld2, cutoff, ips);
    if (ref_derivs(fs, vort2, bdry, fld2, 2, time + 0.75*dt, solve, user, vs, ss)) return -2;
    for (size_t i = 0; i < vort.size(); ++i) {
      Points<float>& p0 = std::get<Points<float>>(vort[i]);
      p0.move(time, dt, 2.0/9.0, p0, 3.0/9.0, std::get<Points<float>>(vort1[i]), 4.0/9.0, std::get<Points<float>>(vort2[i]));
    }
    for (size_t i = 0; i < fldpt.size(); ++i) {
      Points<float>& p0 = std::get<Points<float>>(fldpt[i]);
      p0.move(time, dt, 2.0/9.0, p0, 3.0/9.0, std::get<Points<float>>(fld1[i]), 4.0/9.0, std::get<Points<float>>(fld2[i]));
    }
    clear_inner_layer<float>(1, bdry, vort, cutoff, ips);
    clear_inner_layer<float>(1, bdry, fldpt, cutoff, ips);
  }
  get_xu(vort[0], n, x, y, u, v);
  if (nf > 0) get_xu(fldpt[0], nf, fx, fy, fu, fv);
  return 0;
}


// ---- the adapter routes (src/Influence.h:774-795, :810-869) ---------------------------------------------------------
// reference panels_affect_panels<float,double>: collocation points 1e-4 off the target panels
// (Surfaces::represent_as_particles) as temporary Points; their velocities are ADDED to the target's panel velocities.
int o2dref_panels_affect_panels(int nsp, int nsn, const float* snx, const float* sny, const uint32_t* sidx,
                                const float* svs, const float* sss,
                                int ntp, int ntn, const float* tnx, const float* tny, const uint32_t* tidx,
                                float* pu, float* pv) {
  StdoutMute mute;
  std::vector<float> z((size_t)ntp, 0.0f);
  Surfaces<float> src = make_surfaces(nsp, nsn, snx, sny, sidx, svs, sss);
  Surfaces<float> targ = make_surfaces(ntp, ntn, tnx, tny, tidx, z.data(), z.data());
  auto& u = targ.get_vel();
  for (int i = 0; i < ntp; ++i) { u[0][i] = pu[i]; u[1][i] = pv[i]; }
  panels_affect_panels<float,double>(src, targ, ResultsType(velonly), ExecEnv(true, direct, cpu_x86));
  for (int i = 0; i < ntp; ++i) { pu[i] = u[0][i]; pv[i] = u[1][i]; }
  return 0;
}

// reference points_affect_bricks / panels_affect_bricks<float,double> on a Volumes collection of `ne` first-order quads
// over `nn` nodes (bx,by; qidx = 4 node indices per quad): node velocities (and, for particles with velandvort, node
// vorticities) are ADDED to the Volumes' arrays; returned through u,v,(w).
int o2dref_affect_bricks(int from_panels, int results,
                         int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx, const float* pvs, const float* pss,
                         int ne, int nn, const float* bx, const float* by, const uint32_t* qidx,
                         float* u, float* v, float* w) {
  StdoutMute mute;
  std::vector<float> px(2 * (size_t)nn), val((size_t)nn, 0.0f);
  for (int i = 0; i < nn; ++i) { px[2*i] = bx[i]; px[2*i+1] = by[i]; }
  std::vector<Int> qi(qidx, qidx + 4 * (size_t)ne);
  Volumes<float> vol(ElementPacket<float>(px, qi, val, (size_t)ne, 2), inert, fixed, nullptr);
  if (results == velandvort) (void)vol.get_vort();
  const ExecEnv env(true, direct, cpu_x86);
  if (from_panels) {
    Surfaces<float> src = make_surfaces(np, nnodes, nx, ny, idx, pvs, pss);
    panels_affect_bricks<float,double>(src, vol, ResultsType((results_t)results), env);
  } else {
    Points<float> src = make_points(ns, sx, sy, ss, sr);
    points_affect_bricks<float,double>(src, vol, ResultsType((results_t)results), env);
  }
  const auto& vu = vol.get_vel();
  for (int i = 0; i < nn; ++i) { u[i] = vu[0][i]; v[i] = vu[1][i]; }
  if (w && results == velandvort) { const auto& vw = vol.get_vort(); for (int i = 0; i < nn; ++i) w[i] = vw[i]; }
  return 0;
}


// reference Surfaces::represent_as_particles (src/Surfaces.h:830-873): the particle version of the panels that
// panels_affect_panels (offset 1e-4, ips -1) and the diffusion step use.  Returns the number of points written.
int o2dref_represent_as_particles(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx, const float* vs,
                                  float offset, float ips, int cap, float* px, float* py, float* pval) {
  StdoutMute mute;
  std::vector<float> z((size_t)np, 0.0f);
  Surfaces<float> s = make_surfaces(np, nnodes, nx, ny, idx, vs, z.data());
  const ElementPacket<float> pk = s.represent_as_particles(offset, ips);
  const int n = (int)pk.nelem;
  if (n > cap) return -n;
  for (int i = 0; i < n; ++i) { px[i] = pk.x[2*i]; py[i] = pk.x[2*i+1]; pval[i] = pk.val[i]; }
  return n;
}


// ---- BEM coefficients -----------------------------------------------------------------------------
// reference panels_on_panels_coeff<float> (src/Coefficients.h:155-573) called on the reference's
// own Surfaces.  same != 0: one surface acting on itself (self-influence diagonal).  The reference
// always has source_str_is_unknown == true (src/Surfaces.h:69), so this is the 2x2-per-pair layout;
// returns the unaugmented block (no Body attached), column-major with ld = 2*ntarg.
// Also returns the bases the reference derived (tangent, normal, length; compute_bases, :682-726).
int o2dref_panels_on_panels_coeff(int nsrc, int nsnodes, const float* snx, const float* sny, const uint32_t* sidx,
                                  int ntarg, int ntnodes, const float* tnx, const float* tny, const uint32_t* tidx,
                                  int same, float* coeffs, float* tt_x, float* tt_y, float* tn_x, float* tn_y,
                                  float* tarea, float* sarea) {
  StdoutMute mute;
  std::vector<float> zs((size_t)nsrc, 0.0f), zt((size_t)ntarg, 0.0f);
  Surfaces<float> src = make_surfaces(nsrc, nsnodes, snx, sny, sidx, zs.data(), zs.data());
  Vector<float> c;
  if (same) {
    c = panels_on_panels_coeff<float>(src, src);
  } else {
    Surfaces<float> targ = make_surfaces(ntarg, ntnodes, tnx, tny, tidx, zt.data(), zt.data());
    c = panels_on_panels_coeff<float>(src, targ);
  }
  Surfaces<float> tcopy = same ? src : make_surfaces(ntarg, ntnodes, tnx, tny, tidx, zt.data(), zt.data());
  if (c.size() != (size_t)4 * nsrc * ntarg) return -1;
  for (size_t k = 0; k < c.size(); ++k) coeffs[k] = c[k];
  const auto& tt = tcopy.get_tang();
  const auto& tn = tcopy.get_norm();
  const auto& ta = tcopy.get_area();
  for (int i = 0; i < ntarg; ++i) { tt_x[i] = tt[0][i]; tt_y[i] = tt[1][i]; tn_x[i] = tn[0][i]; tn_y[i] = tn[1][i]; tarea[i] = ta[i]; }
  const auto& sa = src.get_area();
  for (int j = 0; j < nsrc; ++j) sarea[j] = sa[j];
  return 0;
}

// vortex-strength-only layout: the reference's else-branch (src/Coefficients.h:372-389) is unreachable
// through Surfaces (source_str_is_unknown is always true), so its loop is restated here around the
// reference's own kernelu_1v_0p<float,float> (src/Kernels.h:312-354).
int o2dref_panels_on_panels_coeff_vortex(int nsrc, const float* snx, const float* sny, const uint32_t* sidx,
                                         const float* sa,
                                         int ntarg, const float* tnx, const float* tny, const uint32_t* tidx,
                                         const float* ta, const float* ttx, const float* tty, int same, float* coeffs) {
  const float fac = 1.0 / (2.0 * M_PI);
  #pragma omp parallel for
  for (int32_t j = 0; j < (int32_t)nsrc; j++) {
    const float sx0 = snx[sidx[2*j]], sy0 = sny[sidx[2*j]], sx1 = snx[sidx[2*j+1]], sy1 = sny[sidx[2*j+1]];
    for (int i = 0; i < ntarg; i++) {
      const float tx0 = tnx[tidx[2*i]], ty0 = tny[tidx[2*i]], tx1 = tnx[tidx[2*i+1]], ty1 = tny[tidx[2*i+1]];
      const float xi = 0.5 * (tx1 + tx0);
      const float yi = 0.5 * (ty1 + ty0);
      float vortu, vortv;
      kernelu_1v_0p<float,float>(sx0, sy0, sx1, sy1, 1.0, xi, yi, &vortu, &vortv);
      float col = vortu*ttx[i] + vortv*tty[i];
      kernelu_1v_0p<float,float>(tx0, ty0, tx1, ty1, 1.0, 0.5*(sx0+sx1), 0.5*(sy0+sy1), &vortu, &vortv);
      col -= (vortu*ttx[i] + vortv*tty[i]) * (sa[j]/ta[i]);
      col *= 0.5;
      if (same && i == j) col = M_PI;
      coeffs[(size_t)j * ntarg + i] = col * fac;
    }
  }
  return 0;
}

// reference vels_to_rhs_panels<float> (src/RHS.h:54-116) on a reactive Surfaces whose panel velocities
// and normal boundary condition are given; returns the number of rhs entries (2 per panel)
int o2dref_vels_to_rhs_panels(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                              const float* pu, const float* pv, const float* nbc, float* rhs) {
  StdoutMute mute;
  std::vector<float> px(2 * (size_t)nnodes);
  for (int i = 0; i < nnodes; ++i) { px[2*i] = nx[i]; px[2*i+1] = ny[i]; }
  std::vector<Int> pidx(idx, idx + 2 * (size_t)np);
  std::vector<float> val(2 * (size_t)np, 0.0f);
  for (int i = 0; i < np; ++i) val[2*i+1] = nbc ? nbc[i] : 0.0f;
  Surfaces<float> s(ElementPacket<float>(px, pidx, val, (size_t)np, 1), reactive, fixed, nullptr);
  auto& u = s.get_vel();
  for (int i = 0; i < np; ++i) { u[0][i] = pu[i]; u[1][i] = pv[i]; }
  std::vector<float> r = vels_to_rhs_panels<float>(s);
  for (size_t k = 0; k < r.size(); ++k) rhs[k] = r[k];
  return (int)r.size();
}


// ---- clear_inner_layer ------------------------------------------------------------------------------
// reference clear_inner_layer<float> (src/Reflect.h:735-768 -> clear_inner_panp2 :490-733) on the
// reference's own Collections: one Surfaces boundary, one Lagrangian Points collection (blobs with
// radius r, or inert field points when s == NULL).  method 1 = push (what Convection calls after
// every move, src/Convection.h:303-307), method 0 = crop with the cut tables.  Positions (and, for
// method 0, strengths) are returned in place.
int o2dref_clear_inner_layer(int method, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                             int n, float* x, float* y, float* s, const float* r, float cutoff_factor, float ips) {
  StdoutMute mute;
  std::vector<float> z((size_t)np, 0.0f);
  std::vector<Collection> bdry, vort;
  bdry.push_back(make_surfaces(np, nnodes, nx, ny, idx, z.data(), z.data()));
  if (s) {
    vort.push_back(make_points(n, x, y, s, r));
  } else {
    std::vector<float> px(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { px[2*i] = x[i]; px[2*i+1] = y[i]; }
    vort.push_back(Points<float>(ElementPacket<float>(px, std::vector<Int>(), std::vector<float>(), (size_t)n, 0),
                                 inert, lagrangian, nullptr, 0.0f));
  }
  clear_inner_layer<float>(method, bdry, vort, cutoff_factor, ips);
  Points<float>& p = std::get<Points<float>>(vort[0]);
  const auto& px = p.get_pos();
  for (int i = 0; i < n; ++i) { x[i] = px[0][i]; y[i] = px[1][i]; }
  if (s) { const auto& ps = p.get_str(); for (int i = 0; i < n; ++i) s[i] = ps[i]; }
  return 0;
}


// reference reflect_interior<float> (src/Reflect.h:357-380 -> reflect_panp2 :174-350) on the reference's own
// Collections: one Surfaces boundary, one Points collection; positions returned in place.
int o2dref_reflect_interior(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                            int n, float* x, float* y, const float* s, const float* r) {
  StdoutMute mute;
  std::vector<float> z((size_t)np, 0.0f);
  std::vector<Collection> bdry, vort;
  bdry.push_back(make_surfaces(np, nnodes, nx, ny, idx, z.data(), z.data()));
  vort.push_back(make_points(n, x, y, s, r));
  reflect_interior<float>(bdry, vort);
  Points<float>& p = std::get<Points<float>>(vort[0]);
  const auto& px = p.get_pos();
  for (int i = 0; i < n; ++i) { x[i] = px[0][i]; y[i] = px[1][i]; }
  return 0;
}


// reference Points<float>::write_vtk (src/Points.h:556-740; VtkXmlWriter.h, cppcodec base64, tinyxml2 printer) on
// the reference's own Points: blobs (s != NULL: circulation + radius arrays) or inert field points; w != NULL
// also engages the vorticity array, e != NULL the shear-rate array.  The file is written into `dir` (the reference writes into the current
// directory); its name (e.g. part_00_00007.vtu) is returned in fname.
int o2dref_write_vtk_points(const char* dir, int index, int frameno, double time,
                            int n, const float* x, const float* y, const float* s, const float* r,
                            const float* u, const float* v, const float* w, const float* e,
                            char* fname, int fname_len) {
  StdoutMute mute;
  Points<float> p = make_points(n, x, y, s, r);
  auto& pu = p.get_vel();
  for (int i = 0; i < n; ++i) { pu[0][i] = u[i]; pu[1][i] = v[i]; }
  if (w) { auto& pw = p.get_vort(); for (int i = 0; i < n; ++i) pw[i] = w[i]; }
  if (e) { auto& pe = p.get_shear(); for (int i = 0; i < n; ++i) pe[i] = e[i]; }
  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd))) return -1;
  if (chdir(dir) != 0) return -2;
  const std::string name = p.write_vtk((size_t)index, (size_t)frameno, time);
  if (chdir(cwd) != 0) return -3;
  snprintf(fname, (size_t)fname_len, "%s", name.c_str());
  return 0;
}


// reference Surfaces::add_rot_strengths_base (src/Surfaces.h:571-620) on a body-bound Surfaces attached to a
// Body (so that the untransformed coordinates and the geometric centre exist, :623-676).  Returns the
// geometric centre the reference derived in (cx, cy) and the updated strengths in vs, ss.
int o2dref_add_rot_strengths(int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                             float rotvel, float* vs, float* ss, float* cx, float* cy) {
  StdoutMute mute;
  std::vector<float> px(2 * (size_t)nnodes);
  for (int i = 0; i < nnodes; ++i) { px[2*i] = nx[i]; px[2*i+1] = ny[i]; }
  std::vector<Int> pidx(idx, idx + 2 * (size_t)np);
  std::vector<float> val(2 * (size_t)np);
  for (int i = 0; i < np; ++i) { val[2*i] = vs[i]; val[2*i+1] = ss[i]; }
  auto body = std::make_shared<Body>();
  Surfaces<float> s(ElementPacket<float>(px, pidx, val, (size_t)np, 1), active, bodybound, body);
  s.transform(0.0);                 // body at rest at the origin: x = ux, tc = utc (src/Surfaces.h:729-752)
  s.add_rot_strengths_base(rotvel);
  const auto& a = s.get_str();
  const auto& b = s.get_src_str();
  for (int i = 0; i < np; ++i) { vs[i] = a[i]; ss[i] = b[i]; }
  const auto tc = s.get_geom_center();
  *cx = tc[0]; *cy = tc[1];
  return 0;
}

}  // extern "C"
