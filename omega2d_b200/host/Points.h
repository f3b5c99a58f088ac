// Points.h / ElementBase -- SoA particle container with the accessor surface the influence
// path uses, mirrored from the reference (src/ElementBase.h:39-444, src/Points.h:48-193).
// Only what find_vels / find_vort and the *_affect_* functions touch is provided: storage,
// the ElementPacket constructor, zero_vels / finalize_vels.  Motion, bodies, VTK and
// diffusion hooks are out of scope.
#pragma once
#ifndef OMEGA2D_REFERENCE_HEADERS
#include "ElementPacket.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <optional>

template <class S>
class ElementBase {
public:
  ElementBase(const size_t _n, const elem_t _e, const move_t _m) : E(_e), M(_m), n(_n) {}

  size_t get_n() const { return n; }
  bool is_inert() const { return E == inert; }
  elem_t get_elemt() const { return E; }
  move_t get_movet() const { return M; }
  const std::array<Vector<S>,Dimensions>& get_pos() const { return x; }
  std::array<Vector<S>,Dimensions>&       get_pos()       { return x; }
  const Vector<S>& get_str() const { return *s; }
  Vector<S>&       get_str()       { return *s; }
  const std::array<Vector<S>,Dimensions>& get_vel() const { return u; }
  std::array<Vector<S>,Dimensions>&       get_vel()       { return u; }

  // vorticity / shear arrays are created on first non-const access (reference src/ElementBase.h:67-112)
  bool has_vort() const { return (bool)w; }
  const Vector<S>& get_vort() const { return *w; }
  Vector<S>& get_vort() {
    if (!w) w = Vector<S>(n, S(0));
    else if (w->size() != n) w->resize(n);
    return *w;
  }
  bool has_shear() const { return (bool)e; }
  const Vector<S>& get_shear() const { return *e; }
  Vector<S>& get_shear() {
    if (!e) e = Vector<S>(n, S(0));
    else if (e->size() != n) e->resize(n);
    return *e;
  }

  // reference src/ElementBase.h:200-212
  void zero_vels() {
    for (size_t d = 0; d < Dimensions; ++d) std::fill(u[d].begin(), u[d].end(), S(0));
    if (w) { w->resize(n); std::fill(w->begin(), w->end(), S(0)); }
    if (e) { e->resize(n); std::fill(e->begin(), e->end(), S(0)); }
  }
  // reference src/ElementBase.h:215-234: u = fs + u/(2 pi) in double, stored as S
  void finalize_vels(const std::array<double,Dimensions>& _fs) {
    const double factor = 0.5 / M_PI;
    for (size_t d = 0; d < Dimensions; ++d)
      for (size_t i = 0; i < n; ++i) u[d][i] = _fs[d] + u[d][i] * factor;
    if (w) for (size_t i = 0; i < n; ++i) (*w)[i] *= factor;
    if (e) for (size_t i = 0; i < n; ++i) (*e)[i] *= factor;
  }
  void add_solved_rot_strengths(const S) {}   // no attached bodies in this layer

protected:
  elem_t E;
  move_t M;
  size_t n;
  std::array<Vector<S>,Dimensions> x;
  std::optional<Vector<S>> s;
  std::array<Vector<S>,Dimensions> u;
  std::optional<Vector<S>> w;
  std::optional<Vector<S>> e;
};

// 0-D elements: vortex particles (active) or field points / tracers (inert)
template <class S>
class Points : public ElementBase<S> {
public:
  // reference src/Points.h:51-112; _vd is the core radius given to every particle
  Points(const ElementPacket<S>& _in, const elem_t _e, const move_t _m, const float _vd)
    : ElementBase<S>(_in.nelem, _e, _m) {
    assert(_in.idx.size() == 0 && _in.ndim == 0 && "Input ElementPacket is not Points");
    if (_e == inert) assert(_in.val.size() == 0 && "Input ElementPacket with fldpts has val array");
    else assert(_in.val.size() == _in.nelem && "Input ElementPacket with vortons has incorrect size val array");
    for (size_t d = 0; d < Dimensions; ++d) {
      this->x[d].resize(this->n);
      for (size_t i = 0; i < this->n; ++i) this->x[d][i] = _in.x[Dimensions*i+d];
      this->u[d].assign(this->n, S(0));
    }
    if (_e != inert) {
      r.assign(this->n, _vd);
      this->s = Vector<S>(_in.val.begin(), _in.val.end());
    }
  }
  const Vector<S>& get_rad() const { return r; }
  Vector<S>&       get_rad()       { return r; }
protected:
  Vector<S> r;
};
#endif
