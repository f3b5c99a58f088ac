// ConvectionB200.h -- the two callers of the influence path inside Convection, restated as free
// function templates: find_vels (reference src/Convection.h:172-223) and find_vort
// (src/Convection.h:134-166).  The reference's Convection class cannot be instantiated without
// its BEM/Eigen members (src/Convection.h:23-26 -> src/BEM.h:25-27), so the orchestration --
// rotation strengths on/off, lagrangian filter, zero -> accumulate over every source collection
// -> finalize -- is restated here around the same InfluenceVisitor.
#pragma once
#include "InfluenceB200.h"
#include <cmath>
#include <type_traits>
#include <variant>
#include <vector>

namespace o2d_host {
// All Points and Surfaces sources acting on a Points target: the whole zero -> accumulate -> finalize sequence is ONE
// library call -- one upload of every collection, one pair-sum launch over all the particles, one panel launch over all
// the panels, one download -- instead of one call, and one host<->device round trip, per (source, target) pair
// (reference loop: src/Convection.h:205-213).  With one GPU per process this is o2d_find_vels_multi; with a device
// group (O2D_DEVICES) and a single particle + surface collection it is o2d_group_find_vels.
template <class S>
bool fused_find_vels(const std::array<double,Dimensions>& _fs, std::vector<Collection>& _vort,
                     std::vector<Collection>& _bdry, Collection& _targ, const results_t _results) {
  if (_results != velonly && _results != velandvort && _results != velandshear) return false;
  if (!std::holds_alternative<Points<S>>(_targ)) return false;
  std::vector<o2d_points_view> pts;
  std::vector<o2d_panels_view> pans;
  for (auto* list : {&_vort, &_bdry})
    for (auto& src : *list) {
      if (std::holds_alternative<Points<S>>(src)) {
        const Points<S>& p = std::get<Points<S>>(src);
        if (p.is_inert()) return false;                               // no strengths to sum: the reference asserts in get_str()
        pts.push_back(o2d_points_view{(int)p.get_n(), p.get_pos()[0].data(), p.get_pos()[1].data(), p.get_str().data(),
                                      p.get_rad().data()});
      } else if (std::holds_alternative<Surfaces<S>>(src)) {
        const Surfaces<S>& q = std::get<Surfaces<S>>(src);
        if (q.is_inert()) return false;
        pans.push_back(o2d_panels_view{(int)q.get_npanels(), (int)q.get_n(), q.get_pos()[0].data(), q.get_pos()[1].data(),
                                       q.get_idx().data(), q.get_str().data(),
                                       q.have_src_str() ? q.get_src_str().data() : nullptr});
      } else {
        return false;                                                 // Volumes as a source: leave it to the visitor
      }
    }
  Points<S>& targ = std::get<Points<S>>(_targ);
  const auto& tx = targ.get_pos();
  auto& tu = targ.get_vel();
  S* tq = nullptr;
  if (_results == velandvort) tq = targ.get_vort().data();
  if (_results == velandshear) tq = targ.get_shear().data();
  // finalize_vels also scales whichever of w/e the collection carries; keep the one not computed here
  // consistent with the unfused sequence (zeroed, then scaled: stays zero)
  targ.zero_vels();
  const S* tr = targ.is_inert() ? nullptr : targ.get_rad().data();
  if (o2d_group_size(group()) > 1 && pts.size() == 1 && pans.size() <= 1) {
    const o2d_panels_view* q = pans.empty() ? nullptr : &pans[0];
    check(o2d_group_find_vels(group(), selected_core(), (int)_results, pts[0].n, pts[0].x, pts[0].y, pts[0].s, pts[0].r,
                              q ? q->np : 0, q ? q->nnodes : 0, q ? q->nx : nullptr, q ? q->ny : nullptr, q ? q->idx : nullptr,
                              q ? q->vs : nullptr, q ? q->ss : nullptr,
                              (int)targ.get_n(), tx[0].data(), tx[1].data(), tr, tu[0].data(), tu[1].data(), tq, _fs[0], _fs[1]),
          "find_vels");
    return true;
  }
  check(o2d_find_vels_multi(device(), selected_core(), (int)_results, O2D_MEM_HOST, (int)pts.size(), pts.data(),
                            (int)pans.size(), pans.data(), (int)targ.get_n(), tx[0].data(), tx[1].data(), tr,
                            tu[0].data(), tu[1].data(), tq, _fs[0], _fs[1]),
        "find_vels");
  return true;
}
}  // namespace o2d_host

template <class S, class A>
void find_vels(const std::array<double,Dimensions>& _fs,
               std::vector<Collection>& _vort, std::vector<Collection>& _bdry, std::vector<Collection>& _targets,
               const results_t _results = velonly, const bool _force = false, const ExecEnv& _env = ExecEnv()) {
  InfluenceVisitor<S,A> visitor = {ResultsType(_results), _env};
  // add vortex and source strengths to account for rotating bodies
  for (auto& src : _bdry) std::visit([=](auto& elem) { elem.add_solved_rot_strengths(1.0); }, src);
  for (auto& targ : _targets) {
    // only for Lagrangian items, or if told to (as when exporting to vtk)
    const move_t tmt = std::visit([=](auto& elem) { return elem.get_movet(); }, targ);
    if (not (_force or tmt == lagrangian)) continue;
    if (o2d_host::fused_find_vels<S>(_fs, _vort, _bdry, targ, _results)) continue;
    std::visit([=](auto& elem) { elem.zero_vels(); }, targ);
    for (auto& src : _vort) std::visit(visitor, src, targ);
    for (auto& src : _bdry) std::visit(visitor, src, targ);
    // add freestream and divide by 2 pi
    std::visit([=](auto& elem) { elem.finalize_vels(_fs); }, targ);
  }
  for (auto& src : _bdry) std::visit([=](auto& elem) { elem.add_solved_rot_strengths(-1.0); }, src);
}

// One convection step of a particle-only flow, all Runge-Kutta stages on the GPU (o2d_advect):
// Convection::advect_1st / advect_2nd_ralston / advect_3rd (reference src/Convection.h:277-595)
// for `_vort` = one Lagrangian Points collection and no boundaries (then solve_bem and
// clear_inner_layer are no-ops).  Positions move in place; the velocity arrays end up holding what
// the reference leaves there (the weighted stage velocities).
template <class S>
void advect_particles(const int _order, const double _dt, const std::array<double,Dimensions>& _fs, Points<S>& _pts,
                      const int _nsteps = 1) {
  // _nsteps > 1: that many steps with the particles resident on the GPU in between (an inviscid particle-only run,
  // e.g. the 10 steps of examples/speedtest20kx10.json, is nothing else): one upload, one download
  auto& x = _pts.get_pos();
  auto& u = _pts.get_vel();
  o2d_host::check(o2d_advect_steps(o2d_host::device(), o2d_host::selected_core(), _order, O2D_MEM_HOST, (int)_pts.get_n(),
                                   x[0].data(), x[1].data(), _pts.get_str().data(), _pts.get_rad().data(),
                                   u[0].data(), u[1].data(), 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   _fs[0], _fs[1], _dt, _nsteps),
                  "advect_particles");
}

// One convection step of a flow with a body, all Runge-Kutta stages on the GPU (o2d_advect_bodies):
// Convection::advect_1st / advect_2nd_ralston / advect_3rd (reference src/Convection.h:277-595) for one Lagrangian
// Points collection and one Surfaces boundary, INCLUDING what the reference does between the stages: the right-hand
// side of solve_bem (points_affect_panels + finalize_vels on the panels, src/BEMHelper.h:88-100), find_vels with the
// newly solved strengths, Points::move and clear_inner_layer(1, ...).  The linear solve stays the application's:
//   _solve(stage, time, pu, pv, vs, ss)   pu,pv: finalised particle-induced velocities at the panel centres
//                                          (Surfaces::get_vel as vels_to_rhs_panels reads it); fills the new sheet
//                                          strengths vs (and ss when the surface carries source strengths)
// e.g. a lambda around vels_to_rhs_panels + BEM::set_rhs/solve/get_str.  The particles are uploaded once and
// downloaded once per step; per stage 2*Np floats travel each way.  _surf's strength arrays end up holding the
// strengths of the last stage, as after the reference's last solve_bem.
// _fldpt: optional inert Lagrangian field points (tracers); evaluated, moved and cleared alongside the particles, as the
// reference does with its _fldpt collections in every stage.
template <class S, class SolveFn>
void advect_with_body(const int _order, const double _time, const double _dt, const std::array<double,Dimensions>& _fs,
                      const S _ips, Points<S>& _pts, Surfaces<S>& _surf, SolveFn&& _solve, Points<S>* _fldpt = nullptr) {
  struct Tramp {
    static int call(void* user, int stage, double time, int np, const float* pu, const float* pv, float* vs, float* ss) {
      try {
        (*static_cast<std::remove_reference_t<SolveFn>*>(user))(stage, time, np, pu, pv, vs, ss);
        return 0;
      } catch (...) {
        return 1;   // never let an exception cross the C boundary
      }
    }
  };
  static_assert(std::is_same<S, float>::value, "the B200 path stores float, like the reference (STORE=float)");
  auto& x = _pts.get_pos();
  auto& u = _pts.get_vel();
  const auto& px = _surf.get_pos();
  const auto& nrm = _surf.get_norm();
  o2d_host::check(o2d_advect_bodies(o2d_host::device(), o2d_host::selected_core(), _order, O2D_MEM_HOST, (int)_pts.get_n(),
                                    x[0].data(), x[1].data(), _pts.get_str().data(), _pts.get_rad().data(),
                                    u[0].data(), u[1].data(), (int)_surf.get_npanels(), (int)_surf.get_n(),
                                    px[0].data(), px[1].data(), _surf.get_idx().data(), _surf.get_area().data(),
                                    nrm[0].data(), nrm[1].data(), _surf.get_str().data(),
                                    _surf.have_src_str() ? _surf.get_src_str().data() : nullptr,
                                    _fs[0], _fs[1], _time, _dt, (float)(0.5 / std::sqrt(2.0 * M_PI)), (float)_ips,
                                    _fldpt ? (int)_fldpt->get_n() : 0,
                                    _fldpt ? _fldpt->get_pos()[0].data() : nullptr, _fldpt ? _fldpt->get_pos()[1].data() : nullptr,
                                    _fldpt ? _fldpt->get_vel()[0].data() : nullptr, _fldpt ? _fldpt->get_vel()[1].data() : nullptr,
                                    &Tramp::call, &_solve),
                  "advect_with_body");
}

template <class S, class A>
void find_vort(std::vector<Collection>& _vort, std::vector<Collection>& _bdry, std::vector<Collection>& _targets,
               const ExecEnv& _env = ExecEnv()) {
  for (auto& targ : _targets) if (std::holds_alternative<Points<S>>(targ)) {
    Points<S>& ptarg = std::get<Points<S>>(targ);
    ptarg.zero_vels();
    (void)ptarg.get_vort();
    ptarg.zero_vels();
    for (auto& src : _vort) if (std::holds_alternative<Points<S>>(src))
      points_affect_points_vorticity<S,A>(std::get<Points<S>>(src), ptarg, _env);
    for (auto& src : _bdry) if (std::holds_alternative<Points<S>>(src))
      points_affect_points_vorticity<S,A>(std::get<Points<S>>(src), ptarg, _env);
    ptarg.finalize_vels(std::array<double,Dimensions>({0.0, 0.0}));
  }
}
