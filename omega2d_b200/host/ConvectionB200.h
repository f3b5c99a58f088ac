// ConvectionB200.h -- the two callers of the influence path inside Convection, restated as free
// function templates: find_vels (reference src/Convection.h:172-223) and find_vort
// (src/Convection.h:134-166).  The reference's Convection class cannot be instantiated without
// its BEM/Eigen members (src/Convection.h:23-26 -> src/BEM.h:25-27), so the orchestration --
// rotation strengths on/off, lagrangian filter, zero -> accumulate over every source collection
// -> finalize -- is restated here around the same InfluenceVisitor.
#pragma once
#include "InfluenceB200.h"
#include <variant>
#include <vector>

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
    std::visit([=](auto& elem) { elem.zero_vels(); }, targ);
    for (auto& src : _vort) std::visit(visitor, src, targ);
    for (auto& src : _bdry) std::visit(visitor, src, targ);
    // add freestream and divide by 2 pi
    std::visit([=](auto& elem) { elem.finalize_vels(_fs); }, targ);
  }
  for (auto& src : _bdry) std::visit([=](auto& elem) { elem.add_solved_rot_strengths(-1.0); }, src);
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
