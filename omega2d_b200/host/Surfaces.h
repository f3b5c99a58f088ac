// Surfaces.h -- 1-D elements (straight two-node panels) with the accessor surface of the
// reference container (src/Surfaces.h:59-270, :682-726, :755-774), restricted to what the
// influence path reads and writes: node positions, panel node indices, panel lengths
// ("area"), vortex and source sheet strengths, panel-centre velocities.
#pragma once
#include <algorithm>
#ifndef OMEGA2D_REFERENCE_HEADERS
#include "Points.h"

template <class S>
class Surfaces : public ElementBase<S> {
public:
  // value per panel: 1 (vortex strength) or 2 (vortex, source) entries in _elems.val
  Surfaces(const ElementPacket<float>& _elems, const elem_t _e, const move_t _m)
    : ElementBase<S>(_elems.x.size() / Dimensions, _e, _m), np(_elems.nelem) {
    assert(_elems.ndim == 1 && "Input elements are not (1D) surfaces");
    for (size_t d = 0; d < Dimensions; ++d) {
      this->x[d].resize(this->n);
      for (size_t i = 0; i < this->n; ++i) this->x[d][i] = _elems.x[Dimensions*i+d];
      this->u[d].assign(this->n, S(0));
      pu[d].assign(np, S(0));
    }
    idx.assign(_elems.idx.begin(), _elems.idx.end());
    assert(idx.size() == 2 * np && "this layer handles two-node panels");
    for (auto i : idx) { assert(i < this->n && "Some indicies are bad"); (void)i; }
    compute_bases();
    if (_e != inert) {
      // like the reference, both sheet strengths always exist for non-inert surfaces
      // (src/Surfaces.h:87-92), so panels_affect_points takes the vortex+source kernel
      ps[0] = Vector<S>(np, S(0));
      ps[1] = Vector<S>(np, S(0));
      const size_t nvalper = np ? _elems.val.size() / np : 0;
      for (size_t i = 0; i < np; ++i) {
        if (nvalper >= 1) (*ps[0])[i] = _elems.val[nvalper*i];
        if (nvalper >= 2) (*ps[1])[i] = _elems.val[nvalper*i+1];
      }
    }
  }

  size_t get_npanels() const { return np; }
  const std::vector<Int>& get_idx() const { return idx; }
  const Vector<S>& get_area() const { return area; }
  // panel-centre velocities shadow the node velocities (reference src/Surfaces.h:241-243)
  const std::array<Vector<S>,Dimensions>& get_vel() const { return pu; }
  std::array<Vector<S>,Dimensions>&       get_vel()       { return pu; }
  const Vector<S>& get_str() const { return *ps[0]; }
  Vector<S>&       get_str()       { return *ps[0]; }
  bool have_src_str() const { return (bool)ps[1]; }
  const Vector<S>& get_src_str() const { return *ps[1]; }
  Vector<S>&       get_src_str()       { return *ps[1]; }

  // basis vectors and boundary conditions the BEM assembly reads (reference src/Surfaces.h:237-262)
  const std::array<Vector<S>,Dimensions>& get_tang() const { return b[0]; }
  const std::array<Vector<S>,Dimensions>& get_norm() const { return b[1]; }
  const Vector<S>& get_norm_bcs() const { return nbc; }
  Vector<S>&       get_norm_bcs()       { return nbc; }
  Int num_unknowns_per_panel() const { return 2; }     // source_str_is_unknown is always true (src/Surfaces.h:69)
  bool is_augmented() const { return false; }          // no attached Body in this layer (src/Surfaces.h:304-313)
  S get_vol() const { return S(0); }

  // Solid-body rotation of the parent body.  The reference reads the rate from its Body object and the centre
  // from set_geom_center (src/Surfaces.h:540-566,623-676); this layer has no Body class, so both are given.
  void set_body_rotation(const S _cx, const S _cy, const S _rotvel) { rot_c[0] = _cx; rot_c[1] = _cy; rot_vel = _rotvel; }
  // reference src/Surfaces.h:556-620: add factor * (sheet strengths equivalent to the body's rotation)
  void add_solved_rot_strengths(const S _factor) {
    if (rot_vel == S(0) || !ps[0] || !ps[1]) return;
    const S rotvel = _factor * rot_vel;
    for (size_t i = 0; i < np; i++) {
      const size_t j = idx[2*i], jp1 = idx[2*i+1];
      const S dx = 0.5 * (this->x[0][j] + this->x[0][jp1]) - rot_c[0];
      const S dy = 0.5 * (this->x[1][j] + this->x[1][jp1]) - rot_c[1];
      const S ui = -rotvel * dy;
      const S vi =  rotvel * dx;
      S panelx = this->x[0][jp1] - this->x[0][j];
      S panely = this->x[1][jp1] - this->x[1][j];
      const S oopanell = 1.0 / area[i];
      panelx *= oopanell;
      panely *= oopanell;
      (*ps[0])[i] += -1.0 * (ui*panelx + vi*panely);
      (*ps[1])[i] += -1.0 * (ui*panely - vi*panelx);
    }
  }

  // panel length, unit tangent and normal (reference src/Surfaces.h:682-726); the normal points into
  // the fluid, to the left when walking from node 0 to node 1
  void compute_bases() {
    area.resize(np);
    nbc.assign(np, S(0));
    for (size_t i = 0; i < Dimensions; ++i) for (size_t j = 0; j < Dimensions; ++j) b[i][j].resize(np);
    for (size_t i = 0; i < np; ++i) {
      const size_t id0 = idx[2*i], id1 = idx[2*i+1];
      S x1[2] = { this->x[0][id1] - this->x[0][id0], this->x[1][id1] - this->x[1][id0] };
      const S base = std::sqrt(x1[0]*x1[0] + x1[1]*x1[1]);
      for (size_t j = 0; j < Dimensions; ++j) x1[j] *= (1.0/base);
      area[i] = base;
      b[0][0][i] = x1[0];  b[0][1][i] = x1[1];
      b[1][0][i] = -x1[1]; b[1][1][i] = x1[0];
    }
  }

  // Particle version of the panels (reference src/Surfaces.h:830-873): per panel nshed = 1 + int(area/(1.2 ips)) points
  // spread along the panel and pushed _offset along its stored normal, each carrying 1/nshed of the panel's
  // circulation; _ips < 0 means "use the first panel's length" (what panels_affect_panels passes).  Same expressions,
  // and therefore the same float/double promotions, as the reference.
  ElementPacket<float> represent_as_particles(const S _offset, const S _ips) const {
    std::vector<float> px, vals;
    px.reserve(Dimensions * np);
    vals.reserve(np);
    const std::array<Vector<S>,Dimensions>& norm = b[1];
    S thisips = _ips;
    if (_ips < 0.0) thisips = area[0];
    for (size_t i = 0; i < np; ++i) {
      const Int id0 = idx[2*i];
      const Int id1 = idx[2*i+1];
      float this_str = ps[0] ? (*ps[0])[i] : 0.0f;
      this_str *= area[i];
      const size_t nshed = 1 + std::max(0, (int)(area[i] / (1.2*thisips)));
      for (size_t j = 0; j < nshed; ++j) {
        const S wgt = (1+2*j) / (S)(2*nshed);
        px.emplace_back(wgt*this->x[0][id1] + (1.0-wgt)*this->x[0][id0] + _offset * norm[0][i]);
        px.emplace_back(wgt*this->x[1][id1] + (1.0-wgt)*this->x[1][id0] + _offset * norm[1][i]);
        vals.emplace_back(this_str / (S)nshed);
      }
    }
    return ElementPacket<float>(px, std::vector<Int>(), vals, vals.size(), 0);
  }

  void zero_vels() {
    for (size_t d = 0; d < Dimensions; ++d) std::fill(pu[d].begin(), pu[d].end(), S(0));
    ElementBase<S>::zero_vels();
  }
  void finalize_vels(const std::array<double,Dimensions>& _fs) {
    const double factor = 0.5 / M_PI;
    for (size_t d = 0; d < Dimensions; ++d)
      for (size_t i = 0; i < np; ++i) pu[d][i] = _fs[d] + pu[d][i] * factor;
    ElementBase<S>::finalize_vels(_fs);
  }

protected:
  size_t np;
  std::vector<Int> idx;
  Vector<S> area;
  std::array<std::array<Vector<S>,Dimensions>,Dimensions> b;   // b[0] tangent, b[1] normal
  Vector<S> nbc;                                               // normal-velocity boundary condition per panel
  std::array<Vector<S>,Dimensions> pu;
  std::array<std::optional<Vector<S>>,2> ps;
  S rot_c[2] = {S(0), S(0)};
  S rot_vel = S(0);
};
#endif
