// tests/cpp/ref_dropin.cpp -- TEST INFRASTRUCTURE.  The drop-in claim end to end:
//   * the containers are the REFERENCE's own Points<float> / Surfaces<float> / Collection, built from ElementPackets the
//     way Simulation::add_elements builds them (reference src/Simulation.cpp:826-912);
//   * the influence layer is this repository's InfluenceB200.h / ConvectionB200.h in place of the reference's
//     Influence.h / InfluenceVort.h (the include swap of INTEGRATION.md section 2), so find_vels / find_vort run on
//     the GPU through the C ABI;
//   * the answer is checked against the reference's own CPU loops on the same inputs, taken from
//     oracle/_ref/libo2dref_v2.so (built -Bsymbolic: its calls bind to its own template instantiations, not to the
//     identically named ones this program defines).
// Built only where /root/reference exists (oracle/build_ref.sh -> oracle/_ref/ref_dropin_v2); run on the GPU box by
// tests/test_host_mirror.py.  Prints one line per check and exits non-zero on a mismatch.
#include "Collection.h"
#include "ResultsType.h"
#include "ExecEnv.h"
#define OMEGA2D_REFERENCE_HEADERS
#include "InfluenceB200.h"
#include "ConvectionB200.h"
#include "ReflectB200.h"
#include "CoefficientsB200.h"
#include "VtkB200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <unistd.h>
#include <random>
#include <vector>

extern "C" {
int o2dref_core_id();
int o2dref_find_vels(int results, int nv, const float* vx, const float* vy, const float* vs, const float* vr,
                     int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                     const float* pvs, const float* pss, int nt, const float* tx, const float* ty, const float* tr,
                     float* tu, float* tv, float* tq, double fsx, double fsy);
int o2dref_points_vorticity(int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                            int nt, const float* tx, const float* ty, float* tw);
int o2dref_panels_on_panels_coeff(int nsrc, int nsnodes, const float* snx, const float* sny, const uint32_t* sidx,
                                  int ntarg, int ntnodes, const float* tnx, const float* tny, const uint32_t* tidx,
                                  int same, float* coeffs, float* tt_x, float* tt_y, float* tn_x, float* tn_y,
                                  float* tarea, float* sarea);
int o2dref_clear_inner_layer(int method, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                             int n, float* x, float* y, float* s, const float* r, float cutoff_factor, float ips);
int o2dref_panels_affect_panels(int nsp, int nsn, const float* snx, const float* sny, const uint32_t* sidx,
                                const float* svs, const float* sss,
                                int ntp, int ntn, const float* tnx, const float* tny, const uint32_t* tidx, float* pu, float* pv);
int o2dref_affect_bricks(int from_panels, int results,
                         int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx, const float* pvs, const float* pss,
                         int ne, int nn, const float* bx, const float* by, const uint32_t* qidx, float* u, float* v, float* w);
typedef int (*bem_solve_fn)(void* user, int stage, double time, int np, const float* pu, const float* pv, float* vs, float* ss);
int o2dref_advect_bodies(int order, int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx, float* vs, float* ss,
                         double fsx, double fsy, double time, double dt, float ips,
                         int nf, float* fx, float* fy, float* fu, float* fv, bem_solve_fn solve, void* user);
}

// stand-in for the application's BEM solve (Eigen is not in this image): sheet strengths that cancel the tangential and
// normal velocity at the panel centres.  The same function serves the reference-side step and the GPU-side step.
struct SyntheticBem {
  std::vector<float> tx, ty, nx, ny;
  void operator()(int, double, int np, const float* pu, const float* pv, float* vs, float* ss) const {
    for (int i = 0; i < np; ++i) {
      vs[i] = -(pu[i] * tx[i] + pv[i] * ty[i]);
      if (ss) ss[i] = -(pu[i] * nx[i] + pv[i] * ny[i]);
    }
  }
  static int c_call(void* user, int stage, double time, int np, const float* pu, const float* pv, float* vs, float* ss) {
    (*static_cast<const SyntheticBem*>(user))(stage, time, np, pu, pv, vs, ss);
    return 0;
  }
};

static double rel_err(const std::vector<float>& gu, const std::vector<float>& gv,
                      const std::vector<float>& wu, const std::vector<float>& wv) {
  double num = 0.0, den = 0.0;
  for (size_t i = 0; i < gu.size(); ++i) {
    num = std::max(num, std::hypot((double)gu[i] - wu[i], (double)gv[i] - wv[i]));
    den = std::max(den, std::hypot((double)wu[i], (double)wv[i]));
  }
  return num / den;
}

int main() {
  if (o2dref_core_id() != 3 || o2d_host::selected_core() != O2D_CORE_V2) { std::puts("core mismatch"); return 2; }
  const size_t nv = 6000, np = 143, nf = 700;
  const float vdelta = 0.043818f;
  const std::array<double,Dimensions> fs = {1.0, 0.0};
  std::mt19937 gen(7);
  std::uniform_real_distribution<float> U(0.0f, 1.0f);

  // vortex particles in an annulus around a unit-diameter circle; field points above it; the circle as panels
  std::vector<float> vx(nv), vy(nv), vs(nv), vr(nv, vdelta), px(2 * nv);
  for (size_t i = 0; i < nv; ++i) {
    const float rho = 0.55f + 0.9f * U(gen), th = 6.2831853f * U(gen);
    vx[i] = rho * std::cos(th); vy[i] = rho * std::sin(th); vs[i] = (U(gen) - 0.5f) * 2e-3f;
    px[2*i] = vx[i]; px[2*i+1] = vy[i];
  }
  std::vector<float> fx(nf), fy(nf), pf(2 * nf);
  for (size_t i = 0; i < nf; ++i) { fx[i] = -1.5f + 3.0f * U(gen); fy[i] = 0.8f + U(gen); pf[2*i] = fx[i]; pf[2*i+1] = fy[i]; }
  std::vector<float> nx(np), ny(np), pvs(np), pss(np), pn(2 * np), pval(2 * np);
  std::vector<Int> idx(2 * np);
  for (size_t i = 0; i < np; ++i) {
    const float th = -6.2831853f * (float)i / (float)np;
    nx[i] = 0.5f * std::cos(th); ny[i] = 0.5f * std::sin(th);
    pvs[i] = 2.0f * std::sin(th) + 0.1f; pss[i] = 0.3f * std::cos(th);
    pn[2*i] = nx[i]; pn[2*i+1] = ny[i]; pval[2*i] = pvs[i]; pval[2*i+1] = pss[i];
    idx[2*i] = (Int)i; idx[2*i+1] = (Int)((i + 1) % np);
  }
  // the reference's constructors (src/Points.h:51-112, src/Surfaces.h:57-170)
  std::vector<Collection> vort, bdry, fldpt;
  vort.push_back(Points<float>(ElementPacket<float>(px, std::vector<Int>(), vs, nv, 0), active, lagrangian, nullptr, vdelta));
  bdry.push_back(Surfaces<float>(ElementPacket<float>(pn, idx, pval, np, 1), active, fixed, nullptr));
  fldpt.push_back(Points<float>(ElementPacket<float>(pf, std::vector<Int>(), std::vector<float>(), nf, 0), inert, lagrangian, nullptr, 0.0f));   // tracers: find_vels skips collections that do not move

  int bad = 0;
  auto check = [&](const char* what, double err, double tol) {
    std::printf("  %-58s max rel err %.2e (tolerance %.0e) %s\n", what, err, tol, err < tol ? "ok" : "MISMATCH");
    if (!(err < tol)) ++bad;
  };

  // Convection::find_vels on the particles and on the field points (reference containers, GPU influence layer)
  find_vels<float,double>(fs, vort, bdry, vort);
  find_vels<float,double>(fs, vort, bdry, fldpt);
  {
    std::vector<float> wu(nv), wv(nv);
    o2dref_find_vels(velonly, (int)nv, vx.data(), vy.data(), vs.data(), vr.data(), (int)np, (int)np, nx.data(), ny.data(),
                     idx.data(), pvs.data(), pss.data(), (int)nv, vx.data(), vy.data(), vr.data(), wu.data(), wv.data(), nullptr, fs[0], fs[1]);
    const auto& u = std::get<Points<float>>(vort[0]).get_vel();
    check("find_vels: particles + panels on the particles", rel_err(u[0], u[1], wu, wv), 1e-5);
    std::vector<float> fu(nf), fv(nf);
    o2dref_find_vels(velonly, (int)nv, vx.data(), vy.data(), vs.data(), vr.data(), (int)np, (int)np, nx.data(), ny.data(),
                     idx.data(), pvs.data(), pss.data(), (int)nf, fx.data(), fy.data(), nullptr, fu.data(), fv.data(), nullptr, fs[0], fs[1]);
    const auto& uf = std::get<Points<float>>(fldpt[0]).get_vel();
    check("find_vels: particles + panels on inert field points", rel_err(uf[0], uf[1], fu, fv), 1e-5);
  }
  // velocity + vorticity, forced like write_vtk does (src/Simulation.cpp:408-414)
  find_vels<float,double>(fs, vort, bdry, vort, velandvort, true);
  {
    std::vector<float> wu(nv), wv(nv), ww(nv), zero(nv, 0.0f);
    o2dref_find_vels(velandvort, (int)nv, vx.data(), vy.data(), vs.data(), vr.data(), (int)np, (int)np, nx.data(), ny.data(),
                     idx.data(), pvs.data(), pss.data(), (int)nv, vx.data(), vy.data(), vr.data(), wu.data(), wv.data(), ww.data(), fs[0], fs[1]);
    const Points<float>& p = std::get<Points<float>>(vort[0]);
    check("find_vels(velandvort): velocity", rel_err(p.get_vel()[0], p.get_vel()[1], wu, wv), 1e-5);
    check("find_vels(velandvort): vorticity", rel_err(p.get_vort(), zero, ww, zero), 1e-5);
  }
  // find_vort: Gaussian vorticity on the field points, then 1/(2 pi)
  find_vort<float,double>(vort, bdry, fldpt);
  {
    std::vector<float> ww(nf, 0.0f), zero(nf, 0.0f);
    o2dref_points_vorticity((int)nv, vx.data(), vy.data(), vs.data(), vr.data(), (int)nf, fx.data(), fy.data(), ww.data());
    for (auto& w : ww) w = (float)(w * (0.5 / M_PI));
    check("find_vort on the field points", rel_err(std::get<Points<float>>(fldpt[0]).get_vort(), zero, ww, zero), 1e-5);
  }
  // clear_inner_layer (push) on a copy of the particles pulled towards the body
  {
    std::vector<float> qx(vx), qy(vy), qs(vs), qr(vr), pq(2 * nv);
    for (size_t i = 0; i < nv; ++i) { qx[i] *= 0.9f; qy[i] *= 0.9f; pq[2*i] = qx[i]; pq[2*i+1] = qy[i]; }
    std::vector<Collection> near;
    near.push_back(Points<float>(ElementPacket<float>(pq, std::vector<Int>(), vs, nv, 0), active, lagrangian, nullptr, vdelta));
    const float cutoff = 0.5f / std::sqrt(2.0f * 3.14159265f), ips = 0.02191f;
    clear_inner_layer<float>(1, bdry, near, cutoff, ips);
    o2dref_clear_inner_layer(1, (int)np, (int)np, nx.data(), ny.data(), idx.data(), (int)nv, qx.data(), qy.data(), qs.data(), qr.data(), cutoff, ips);
    const auto& x = std::get<Points<float>>(near[0]).get_pos();
    double d = 0.0; size_t moved = 0;
    for (size_t i = 0; i < nv; ++i) {
      d = std::max(d, std::hypot((double)x[0][i] - qx[i], (double)x[1][i] - qy[i]));
      moved += (x[0][i] != pq[2*i] || x[1][i] != pq[2*i+1]);
    }
    std::printf("  clear_inner_layer moved %zu particles\n", moved);
    check("clear_inner_layer (push): max |dx| against the reference", d, 1e-5);
    if (moved < 10) ++bad;
  }
  // panels_affect_panels (src/Influence.h:774-795) through the visitor: the body acting on a second, smaller body; the
  // collocation points come from the reference's own Surfaces::represent_as_particles on both sides
  {
    const size_t np2 = 60;
    std::vector<float> n2x(np2), n2y(np2), p2(2 * np2), v2(2 * np2, 0.0f);
    std::vector<Int> i2(2 * np2);
    for (size_t i = 0; i < np2; ++i) {
      const float th = -6.2831853f * (float)i / (float)np2;
      n2x[i] = 0.9f + 0.2f * std::cos(th); n2y[i] = 0.3f + 0.2f * std::sin(th);
      p2[2*i] = n2x[i]; p2[2*i+1] = n2y[i];
      i2[2*i] = (Int)i; i2[2*i+1] = (Int)((i + 1) % np2);
    }
    Collection second = Surfaces<float>(ElementPacket<float>(p2, i2, v2, np2, 1), active, fixed, nullptr);
    InfluenceVisitor<float,double> visitor = {ResultsType(velonly), ExecEnv(true, direct, cpu_x86)};
    std::visit([](auto& e) { e.zero_vels(); }, second);
    std::visit(visitor, bdry[0], second);
    std::vector<float> wu(np2, 0.0f), wv(np2, 0.0f);
    o2dref_panels_affect_panels((int)np, (int)np, nx.data(), ny.data(), idx.data(), pvs.data(), pss.data(),
                                (int)np2, (int)np2, n2x.data(), n2y.data(), i2.data(), wu.data(), wv.data());
    const auto& u2 = std::get<Surfaces<float>>(second).get_vel();
    check("panels_affect_panels: body on a second body", rel_err(u2[0], u2[1], wu, wv), 1e-5);
    // ... and the body on itself (collocation points 1e-4 above its own panels: the near-field pair)
    Collection self = bdry[0];
    std::visit([](auto& e) { e.zero_vels(); }, self);
    std::visit(visitor, bdry[0], self);
    std::vector<float> su(np, 0.0f), sv(np, 0.0f);
    o2dref_panels_affect_panels((int)np, (int)np, nx.data(), ny.data(), idx.data(), pvs.data(), pss.data(),
                                (int)np, (int)np, nx.data(), ny.data(), idx.data(), su.data(), sv.data());
    const auto& us = std::get<Surfaces<float>>(self).get_vel();
    check("panels_affect_panels: body on itself", rel_err(us[0], us[1], su, sv), 1e-5);
  }
  // points_affect_bricks / panels_affect_bricks (src/Influence.h:810-869): a Volumes collection of 12 x 8 quads
  {
    const int gx = 13, gy = 9, nn = gx * gy, ne = (gx - 1) * (gy - 1);
    std::vector<float> bx(nn), by(nn), pb(2 * nn), vb(nn, 0.0f);
    for (int j = 0; j < gy; ++j) for (int i = 0; i < gx; ++i) {
      bx[j*gx+i] = 1.2f + 0.05f * i; by[j*gx+i] = -0.2f + 0.05f * j;
      pb[2*(j*gx+i)] = bx[j*gx+i]; pb[2*(j*gx+i)+1] = by[j*gx+i];
    }
    std::vector<Int> qi;
    for (int j = 0; j < gy - 1; ++j) for (int i = 0; i < gx - 1; ++i) {
      qi.push_back((Int)(j*gx+i)); qi.push_back((Int)(j*gx+i+1)); qi.push_back((Int)((j+1)*gx+i+1)); qi.push_back((Int)((j+1)*gx+i));
    }
    InfluenceVisitor<float,double> visitor = {ResultsType(velandvort), ExecEnv(true, direct, cpu_x86)};
    Collection vol = Volumes<float>(ElementPacket<float>(pb, qi, vb, (size_t)ne, 2), inert, fixed, nullptr);
    (void)std::get<Volumes<float>>(vol).get_vort();
    std::visit([](auto& e) { e.zero_vels(); }, vol);
    std::visit(visitor, vort[0], vol);
    std::vector<float> wu(nn), wv(nn), ww(nn), zero(nn, 0.0f);
    o2dref_affect_bricks(0, velandvort, (int)nv, vx.data(), vy.data(), vs.data(), vr.data(), 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                         ne, nn, bx.data(), by.data(), qi.data(), wu.data(), wv.data(), ww.data());
    {
      const Volumes<float>& V = std::get<Volumes<float>>(vol);
      check("points_affect_bricks: node velocities", rel_err(V.get_vel()[0], V.get_vel()[1], wu, wv), 1e-5);
      check("points_affect_bricks: node vorticities", rel_err(V.get_vort(), zero, ww, zero), 1e-5);
    }
    InfluenceVisitor<float,double> vvel = {ResultsType(velonly), ExecEnv(true, direct, cpu_x86)};
    std::visit([](auto& e) { e.zero_vels(); }, vol);
    std::visit(vvel, bdry[0], vol);
    o2dref_affect_bricks(1, velonly, 0, nullptr, nullptr, nullptr, nullptr, (int)np, (int)np, nx.data(), ny.data(), idx.data(), pvs.data(), pss.data(),
                         ne, nn, bx.data(), by.data(), qi.data(), wu.data(), wv.data(), nullptr);
    const Volumes<float>& V = std::get<Volumes<float>>(vol);
    check("panels_affect_bricks: node velocities", rel_err(V.get_vel()[0], V.get_vel()[1], wu, wv), 1e-5);
  }
  // a whole convection step with the body: advect_with_body (GPU: BEM right-hand side, solve callback, find_vels, move,
  // clear_inner_layer per stage, particles resident) against the reference-side step restated around the reference's
  // own pieces (oracle/ref_driver.cpp), same stand-in BEM solve on both sides
  {
    SyntheticBem bemfn;
    bemfn.tx.resize(np); bemfn.ty.resize(np); bemfn.nx.resize(np); bemfn.ny.resize(np);
    {
      const Surfaces<float>& sf = std::get<Surfaces<float>>(bdry[0]);
      for (size_t i = 0; i < np; ++i) { bemfn.tx[i] = sf.get_tang()[0][i]; bemfn.ty[i] = sf.get_tang()[1][i]; bemfn.nx[i] = sf.get_norm()[0][i]; bemfn.ny[i] = sf.get_norm()[1][i]; }
    }
    const float ips = 0.02191f;
    const double dt = 0.02;
    for (int order = 1; order <= 3; ++order) {
      // particles pulled towards the body so that the clear layer has work to do
      std::vector<float> qx(nv), qy(nv), pq(2 * nv);
      for (size_t i = 0; i < nv; ++i) { qx[i] = 0.92f * vx[i]; qy[i] = 0.92f * vy[i]; pq[2*i] = qx[i]; pq[2*i+1] = qy[i]; }
      Points<float> pts(ElementPacket<float>(pq, std::vector<Int>(), vs, nv, 0), active, lagrangian, nullptr, vdelta);
      Surfaces<float> surf = std::get<Surfaces<float>>(bdry[0]);
      // tracers close to the body ride along (inert Lagrangian Points, like the reference's fldpt collections)
      std::vector<float> tfx(nf), tfy(nf), ptf(2 * nf);
      for (size_t i = 0; i < nf; ++i) { tfx[i] = 0.45f * fx[i]; tfy[i] = 0.45f * fy[i] - 0.1f; ptf[2*i] = tfx[i]; ptf[2*i+1] = tfy[i]; }
      Points<float> trc(ElementPacket<float>(ptf, std::vector<Int>(), std::vector<float>(), nf, 0), inert, lagrangian, nullptr, 0.0f);
      advect_with_body<float>(order, 0.0, dt, fs, ips, pts, surf, bemfn, &trc);
      std::vector<float> wu(nv), wv(nv), ws(pvs), wss(pss), tfu(nf), tfv(nf);
      o2dref_advect_bodies(order, (int)nv, qx.data(), qy.data(), vs.data(), vr.data(), wu.data(), wv.data(), (int)np, (int)np,
                           nx.data(), ny.data(), idx.data(), ws.data(), wss.data(), fs[0], fs[1], 0.0, dt, ips,
                           (int)nf, tfx.data(), tfy.data(), tfu.data(), tfv.data(), &SyntheticBem::c_call, &bemfn);
      {
        double dtr = 0.0;
        for (size_t i = 0; i < nf; ++i)
          dtr = std::max(dtr, std::hypot((double)trc.get_pos()[0][i] - tfx[i], (double)trc.get_pos()[1][i] - tfy[i]));
        char w2[96];
        std::snprintf(w2, sizeof(w2), "advect_with_body order %d: tracer velocities", order);
        check(w2, rel_err(trc.get_vel()[0], trc.get_vel()[1], tfu, tfv), 1e-5);
        std::snprintf(w2, sizeof(w2), "advect_with_body order %d: tracer max |dx|", order);
        check(w2, dtr, 0.0011 * ips + 2e-6);
      }
      double d = 0.0; size_t flips = 0;
      for (size_t i = 0; i < nv; ++i) {
        const double di = std::hypot((double)pts.get_pos()[0][i] - qx[i], (double)pts.get_pos()[1][i] - qy[i]);
        if (di > 2e-6) ++flips;     // the push threshold of clear_inner_layer is discontinuous
        d = std::max(d, di);
      }
      char what[96];
      std::snprintf(what, sizeof(what), "advect_with_body order %d: stage velocities", order);
      check(what, rel_err(pts.get_vel()[0], pts.get_vel()[1], wu, wv), 1e-5);
      std::snprintf(what, sizeof(what), "advect_with_body order %d: solved sheet strengths", order);
      check(what, rel_err(surf.get_vort_str(), surf.get_src_str(), ws, wss), 1e-5);
      std::snprintf(what, sizeof(what), "advect_with_body order %d: max |dx| (%zu threshold flips)", order, flips);
      check(what, d, 0.0011 * ips + 2e-6);
      if (flips > nv / 500) ++bad;
    }
  }
  // BEM influence-coefficient block of the body on itself: CoefficientVisitor on the reference's Surfaces
  {
    Vector<float> c = std::visit(CoefficientVisitor<float>(), bdry[0], bdry[0]);
    std::vector<float> want(4 * np * np), t1(np), t2(np), t3(np), t4(np), t5(np), t6(np);
    o2dref_panels_on_panels_coeff((int)np, (int)np, nx.data(), ny.data(), idx.data(), (int)np, (int)np, nx.data(), ny.data(), idx.data(),
                                  1, want.data(), t1.data(), t2.data(), t3.data(), t4.data(), t5.data(), t6.data());
    double num = 0.0, den = 0.0;
    if (c.size() == want.size())
      for (size_t k = 0; k < want.size(); ++k) { num = std::max(num, std::fabs((double)c[k] - want[k])); den = std::max(den, std::fabs((double)want[k])); }
    else { num = 1.0; den = 1.0; }
    check("panels_on_panels_coeff through CoefficientVisitor", num / den, 1e-5);
  }
  // Points::write_vtk: the reference's own member (CPU) and write_vtk_points (GPU) on the same reference object
  {
    char tmpl[] = "/tmp/o2d_dropin_XXXXXX";
    const char* dir = mkdtemp(tmpl);
    char cwd[4096];
    if (dir && getcwd(cwd, sizeof(cwd)) && chdir(dir) == 0) {
      Points<float>& p = std::get<Points<float>>(vort[0]);
      const std::string a = p.write_vtk(0, 7, 0.25);                  // reference: VtkXmlWriter + cppcodec + tinyxml2
      const std::string b = write_vtk_points(p, 0, 8, 0.25);          // this repository: one kernel + skeleton
      auto slurp = [](const std::string& f) { std::ifstream s(f, std::ios::binary); return std::string(std::istreambuf_iterator<char>(s), {}); };
      const std::string fa = slurp(a), fb = slurp(b);
      const bool same = !fa.empty() && fa == fb;
      std::printf("  write_vtk: reference %s (%zu bytes), GPU %s (%zu bytes): %s\n", a.c_str(), fa.size(), b.c_str(), fb.size(),
                  same ? "identical ok" : "DIFFERENT");
      if (!same) ++bad;
      std::remove(a.c_str()); std::remove(b.c_str());
      if (chdir(cwd) != 0) ++bad;
      rmdir(dir);
    } else { std::puts("  write_vtk: no temporary directory"); ++bad; }
  }
  std::printf("ref_dropin: %s\n", bad ? "FAILED" : "all checks passed");
  return bad ? 1 : 0;
}
