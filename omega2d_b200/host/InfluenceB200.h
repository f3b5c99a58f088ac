// InfluenceB200.h -- drop-in replacement for the reference's src/Influence.h + src/InfluenceVort.h.
//
// Same function templates, same argument meaning, same in-place "+=" semantics on the target
// collection's arrays, same InfluenceVisitor -- but every loop nest is one call into the C ABI of
// libomega2d_b200.so (include/omega2d_b200.h).  The bodies only touch the containers through
// the accessors the reference's own loops use (get_pos/get_rad/get_str/get_vel/get_vort/
// get_shear/get_idx/get_area/get_src_str/have_src_str/is_inert/get_n/get_npanels), so this
// header works with
//   (a) the host containers shipped next to it (Points.h, Surfaces.h, Collection.h), or
//   (b) the reference's own containers: include the reference's "Collection.h" first and
//       #define OMEGA2D_REFERENCE_HEADERS before including this file instead of "Influence.h".
//
// The core function is picked like the reference does, by macro (reference src/CoreFunc.h:35-39):
// USE_RM_KERNEL / USE_EXPONENTIAL_KERNEL / USE_WL_KERNEL / USE_V2_KERNEL / USE_V3_KERNEL,
// default V2.  There is no CPU path in here: if the library or the GPU is missing the process
// stops with a message (the reference's own failure mode for unsupported cases is assert()).
#pragma once

#include "../../include/omega2d_b200.h"
#ifndef OMEGA2D_REFERENCE_HEADERS
#include "Collection.h"
#endif

#include <cassert>
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <cstdlib>

namespace o2d_host {

constexpr int selected_core() {
#if defined(USE_RM_KERNEL)
  return O2D_CORE_RM;
#elif defined(USE_EXPONENTIAL_KERNEL)
  return O2D_CORE_EXP;
#elif defined(USE_WL_KERNEL)
  return O2D_CORE_WL;
#elif defined(USE_V3_KERNEL)
  return O2D_CORE_V3;
#else
  return O2D_CORE_V2;
#endif
}

// The reference's influence functions return void and assert; a library failure (no sm_100 GPU, CUDA error, bad
// argument) is reported by an exception instead of taking the process down, so a GUI session can show it.  A host
// program that does not catch it terminates with the message, as an assert would.
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// process-wide device group, created on first use: one GPU -- device $O2D_DEVICE, default 0 -- unless
// $O2D_DEVICES=N|all asks for several; the reference is a single-process OpenMP program, so this is how it
// reaches all GPUs of a box (include/omega2d_b200.h, "all GPUs of the box from ONE process")
inline o2d_group* group() {
  static o2d_group* grp = [] {
    o2d_group* g = nullptr;
    int rc;
    if (const char* e = std::getenv("O2D_DEVICES")) {
      const bool all = (e[0] == 'a' || e[0] == 'A');
      rc = o2d_group_create(&g, all ? 0 : std::atoi(e), nullptr);
    } else {
      int dev = 0;
      if (const char* e1 = std::getenv("O2D_DEVICE")) dev = std::atoi(e1);
      rc = o2d_group_create(&g, 1, &dev);
    }
    if (rc != O2D_OK) throw Error(std::string("omega2d_b200: ") + o2d_last_error());   // no GPU: no CPU fallback either
    return g;
  }();
  return grp;
}
// the group's first context: everything that is not a big influence sum runs there
inline o2d_context* device() { return o2d_group_context(group(), 0); }

inline void check(int rc, const char* what) {
  if (rc != O2D_OK) throw Error(std::string("omega2d_b200: ") + what + " failed: " + o2d_last_error());
}

inline bool verbose() {
  static const bool v = std::getenv("O2D_VERBOSE") != nullptr;
  return v;
}

// the reference prints "[seconds] at GFlop/s" after every influence call (src/Influence.h:412-415)
struct CallTimer {
  const char* name;
  double flops;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  CallTimer(const char* n, double f) : name(n), flops(f) {}
  ~CallTimer() {
    if (!verbose()) return;
    const std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
    std::printf("    %s: [%.4f] seconds at %.3f GFlop/s\n", name, (float)dt.count(), (float)(1.e-9 * flops / dt.count()));
  }
};

}  // namespace o2d_host

//
// Points/Particles affecting Points/Particles -- reference src/Influence.h:56-416
//
template <class S, class A>
void points_affect_points(const Points<S>& src, Points<S>& targ, const ResultsType& restype, const ExecEnv& env) {
  static_assert(sizeof(S) == sizeof(float), "omega2d_b200 stores float32, like the reference's STORE");
  (void)env;
  assert(!restype.compute_psi() && "Point elements cannot compute streamfunction yet.");
  assert(!restype.compute_grad() && "Point elements cannot compute velocity gradients yet.");
  if (!restype.compute_vel()) return;

  const auto& sx = src.get_pos();
  const auto& sr = src.get_rad();
  const auto& ss = src.get_str();
  const auto& tx = targ.get_pos();
  auto& tu = targ.get_vel();
  const S* tr = targ.is_inert() ? nullptr : targ.get_rad().data();
  S* tq = nullptr;
  if (restype.get_type() == velandvort) tq = targ.get_vort().data();
  if (restype.get_type() == velandshear) tq = targ.get_shear().data();

  const int core = o2d_host::selected_core();
  const double per_pair = o2d_flops_per_pair(0, core, (int)restype.get_type(), tr != nullptr);
  o2d_host::CallTimer timer("points_affect_points", (double)targ.get_n() * (2.0 + per_pair * (double)src.get_n()));
  o2d_host::check(o2d_group_points_on_points(o2d_host::group(), core, (int)restype.get_type(),
                                       (int)src.get_n(), sx[0].data(), sx[1].data(), ss.data(), sr.data(),
                                       (int)targ.get_n(), tx[0].data(), tx[1].data(), tr,
                                       tu[0].data(), tu[1].data(), tq, nullptr, nullptr, nullptr, nullptr),
                  "points_affect_points");
}

//
// velocity + 2x2 gradient at blob targets: the loop the reference never wrote around its own
// kernelug_0v_0b (src/Kernels.h:112-133).  grads = {du/dx, dv/dx, du/dy, dv/dy}, accumulated into.
//
template <class S, class A>
void points_affect_points_grad(const Points<S>& src, Points<S>& targ, std::array<Vector<S>,4>& grads, const ExecEnv& env) {
  (void)env;
  for (auto& g : grads) if (g.size() != targ.get_n()) g.assign(targ.get_n(), S(0));
  const auto& sx = src.get_pos();
  const auto& tx = targ.get_pos();
  auto& tu = targ.get_vel();
  const S* tr = targ.is_inert() ? nullptr : targ.get_rad().data();
  o2d_host::check(o2d_group_points_on_points(o2d_host::group(), o2d_host::selected_core(), O2D_VELANDGRAD,
                                       (int)src.get_n(), sx[0].data(), sx[1].data(), src.get_str().data(), src.get_rad().data(),
                                       (int)targ.get_n(), tx[0].data(), tx[1].data(), tr, tu[0].data(), tu[1].data(), nullptr,
                                       grads[0].data(), grads[1].data(), grads[2].data(), grads[3].data()),
                  "points_affect_points_grad");
}

//
// Panels/Surfaces affecting Points/Particles -- reference src/Influence.h:422-614
//
template <class S, class A>
void panels_affect_points(const Surfaces<S>& src, Points<S>& targ, const ResultsType& restype, const ExecEnv& env) {
  (void)env;
  assert(!restype.compute_psi() && "Surface elements cannot compute streamfunction yet.");
  assert(!restype.compute_grad() && "Surface elements cannot compute velocity gradients yet.");
  if (!restype.compute_vel()) return;
  const auto& sx = src.get_pos();
  const auto& si = src.get_idx();
  const auto& vs = src.get_str();
  const S* ss = src.have_src_str() ? src.get_src_str().data() : nullptr;
  const auto& tx = targ.get_pos();
  auto& tu = targ.get_vel();
  o2d_host::CallTimer timer("panels_affect_points",
                            (double)targ.get_n() * (2.0 + (double)o2d_flops_per_pair(ss ? 2 : 1, 0, 1, 0) * (double)src.get_npanels()));
  o2d_host::check(o2d_group_panels_on_points(o2d_host::group(), (int)src.get_npanels(), (int)src.get_n(),
                                       sx[0].data(), sx[1].data(), si.data(), vs.data(), ss,
                                       (int)targ.get_n(), tx[0].data(), tx[1].data(), tu[0].data(), tu[1].data()),
                  "panels_affect_points");
}

//
// Points/Particles affecting Panels/Surfaces -- reference src/Influence.h:633-771
//
template <class S, class A>
void points_affect_panels(const Points<S>& src, Surfaces<S>& targ, const ResultsType& restype, const ExecEnv& env) {
  (void)env;
  assert(!restype.compute_psi() && "Point elements cannot compute streamfunction yet.");
  assert(!restype.compute_grad() && "Point elements cannot compute velocity gradients yet.");
  if (!restype.compute_vel()) return;
  const auto& sx = src.get_pos();
  const auto& vs = src.get_str();
  const auto& tx = targ.get_pos();
  const auto& ti = targ.get_idx();
  const auto& ta = targ.get_area();
  auto& tu = targ.get_vel();
  o2d_host::CallTimer timer("points_affect_panels",
                            (double)targ.get_npanels() * (11.0 + (double)o2d_flops_per_pair(1, 0, 1, 0) * (double)src.get_n()));
  o2d_host::check(o2d_group_points_on_panels(o2d_host::group(), (int)src.get_n(), sx[0].data(), sx[1].data(), vs.data(),
                                       (int)targ.get_npanels(), (int)targ.get_n(), tx[0].data(), tx[1].data(), ti.data(),
                                       ta.data(), tu[0].data(), tu[1].data()),
                  "points_affect_panels");
}

//
// Panels on panels -- reference src/Influence.h:774-795: collocation points 1e-4 off the target
// panels are treated as temporary inert-radius Points and their velocities added to the panels.
//
template <class S, class A>
void panels_affect_panels(const Surfaces<S>& src, Surfaces<S>& targ, const ResultsType& restype, const ExecEnv& env) {
  assert(!restype.compute_psi() && "Surface elements cannot compute streamfunction yet.");
  assert(!restype.compute_grad() && "Surface elements cannot compute velocity gradients yet.");
  ElementPacket<float> surfaspts = targ.represent_as_particles(0.0001, -1.0);
#ifdef OMEGA2D_REFERENCE_HEADERS
  Points<S> temppts(surfaspts, active, lagrangian, nullptr, 0.0001);
#else
  Points<S> temppts(surfaspts, active, lagrangian, 0.0001f);
#endif
  panels_affect_points<S,A>(src, temppts, restype, env);
  const auto& fromvel = temppts.get_vel();
  auto& tovel = targ.get_vel();
  for (size_t d = 0; d < Dimensions; ++d)
    for (size_t i = 0; i < tovel[d].size(); ++i) tovel[d][i] += fromvel[d][i];
}

//
// Gaussian vorticity of particles at points -- reference src/InfluenceVort.h:47-126
//
template <class S, class A>
void points_affect_points_vorticity(const Points<S>& src, Points<S>& targ, const ExecEnv& env) {
  (void)env;
  const auto& sx = src.get_pos();
  const auto& tx = targ.get_pos();
  auto& tw = targ.get_vort();
  o2d_host::CallTimer timer("points_affect_points_vorticity", (double)targ.get_n() * (1.0 + 13.0 * (double)src.get_n()));
  o2d_host::check(o2d_points_vorticity(o2d_host::device(), O2D_MEM_HOST, (int)src.get_n(), sx[0].data(), sx[1].data(),
                                       src.get_str().data(), src.get_rad().data(),
                                       (int)targ.get_n(), tx[0].data(), tx[1].data(), tw.data()),
                  "points_affect_points_vorticity");
}

#ifdef OMEGA2D_REFERENCE_HEADERS
// Volumes only enter through node-to-Points adapters (reference src/Influence.h:810-869)
template <class S, class A>
void bricks_affect_points(const Volumes<S>&, Points<S>&, const ResultsType&, const ExecEnv&) { assert(false && "Volume elements cannot affect points yet."); }
template <class S, class A>
void bricks_affect_panels(const Volumes<S>&, Surfaces<S>&, const ResultsType&, const ExecEnv&) { assert(false && "Volume elements cannot affect panels yet."); }
template <class S, class A>
void bricks_affect_bricks(const Volumes<S>&, Volumes<S>&, const ResultsType&, const ExecEnv&) { assert(false && "Volume elements cannot affect themselves yet."); }
template <class S, class A>
void points_affect_bricks(const Points<S>& src, Volumes<S>& targ, const ResultsType& restype, const ExecEnv& env) {
  ElementPacket<float> nodesaspts = targ.represent_nodes_as_particles(true);
  Points<S> volsaspts(nodesaspts, inert, fixed, nullptr, 0.0f);
  points_affect_points<S,A>(src, volsaspts, restype, env);
  const auto& fromvel = volsaspts.get_vel();
  auto& tovel = targ.get_vel();
  for (size_t d = 0; d < Dimensions; ++d)
    for (size_t i = 0; i < tovel[d].size(); ++i) tovel[d][i] += fromvel[d][i];
  if (restype.compute_vort()) {
    const auto& fromvort = volsaspts.get_vort();
    auto& tovort = targ.get_vort();
    for (size_t i = 0; i < tovort.size(); ++i) tovort[i] += fromvort[i];
  }
}
template <class S, class A>
void panels_affect_bricks(const Surfaces<S>& src, Volumes<S>& targ, const ResultsType& restype, const ExecEnv& env) {
  ElementPacket<float> nodesaspts = targ.represent_nodes_as_particles(true);
  Points<S> volsaspts(nodesaspts, inert, fixed, nullptr, 0.0f);
  panels_affect_points<S,A>(src, volsaspts, restype, env);
  const auto& fromvel = volsaspts.get_vel();
  auto& tovel = targ.get_vel();
  for (size_t d = 0; d < Dimensions; ++d)
    for (size_t i = 0; i < tovel[d].size(); ++i) tovel[d][i] += fromvel[d][i];
}
#endif

//
// helper struct for dispatching through a variant -- reference src/Influence.h:885-900
//
template <class S, class A>
struct InfluenceVisitor {
  void operator()(const Points<S>& src,   Points<S>& targ)   { points_affect_points<S,A>(src, targ, restype, env); }
  void operator()(const Surfaces<S>& src, Points<S>& targ)   { panels_affect_points<S,A>(src, targ, restype, env); }
  void operator()(const Points<S>& src,   Surfaces<S>& targ) { points_affect_panels<S,A>(src, targ, restype, env); }
  void operator()(const Surfaces<S>& src, Surfaces<S>& targ) { panels_affect_panels<S,A>(src, targ, restype, env); }
#ifdef OMEGA2D_REFERENCE_HEADERS
  void operator()(const Volumes<S>& src,  Points<S>& targ)   { bricks_affect_points<S,A>(src, targ, restype, env); }
  void operator()(const Volumes<S>& src,  Surfaces<S>& targ) { bricks_affect_panels<S,A>(src, targ, restype, env); }
  void operator()(const Points<S>& src,   Volumes<S>& targ)  { points_affect_bricks<S,A>(src, targ, restype, env); }
  void operator()(const Surfaces<S>& src, Volumes<S>& targ)  { panels_affect_bricks<S,A>(src, targ, restype, env); }
  void operator()(const Volumes<S>& src,  Volumes<S>& targ)  { bricks_affect_bricks<S,A>(src, targ, restype, env); }
#endif
  ResultsType restype;
  ExecEnv env;
};
