// ReflectB200.h -- drop-in replacement for the part of the reference's src/Reflect.h that the
// convection step calls: clear_inner_layer<S> (reference src/Reflect.h:735-768) with _method == 1, the
// "push" variant Convection runs on every Lagrangian collection after every move
// (src/Convection.h:303-307, :334, :342, :388-391).  Same signature and collection filtering; the
// O(N * Np) closest-panel search of clear_inner_panp2 (:490-733) is one call into libomega2d_b200.so.
// _method == 0 (strength cropping with the cut tables, :392-480) has no caller in the reference and is not
// part of this layer.  Also here: reflect_interior<S> (:357-380, reflect_panp2 :174-350), which
// Diffusion::step runs before VRM (src/Diffusion.h:301) -- the same search with a mirror ending.
#pragma once
#include "InfluenceB200.h"

#include <variant>
#include <vector>

template <class S>
S clear_inner_panp2(const int _method, Surfaces<S> const& _src, Points<S>& _targ, const S _cutoff_mult, const S _ips) {
  static_assert(sizeof(S) == sizeof(float), "omega2d_b200 stores float32, like the reference's STORE");
  assert(_method == 1 && "only the push method is available on the GPU path");
  const auto& sx = _src.get_pos();
  const auto& sn = _src.get_norm();
  auto& tx = _targ.get_pos();
  long long moved = 0;
  o2d_host::check(o2d_clear_inner_layer(o2d_host::device(), _method, O2D_MEM_HOST,
                                        (int)_src.get_npanels(), (int)_src.get_n(), sx[0].data(), sx[1].data(),
                                        _src.get_idx().data(), sn[0].data(), sn[1].data(),
                                        (int)_targ.get_n(), tx[0].data(), tx[1].data(), _cutoff_mult, _ips, &moved),
                  "clear_inner_layer");
  if (o2d_host::verbose()) std::printf("    pushed %lld particles\n", moved);
  return S(0);   // method 1 keeps all strength: no circulation is reabsorbed (src/Reflect.h:692-706)
}

template <class S>
void clear_inner_layer(const int _method, std::vector<Collection>& _bdry, std::vector<Collection>& _vort,
                       const S _cutoff_factor, const S _ips) {
  for (auto& targ : _vort) {
    if (std::holds_alternative<Points<S>>(targ)) {
      Points<S>& pts = std::get<Points<S>>(targ);
      // don't push anything if they don't move themselves
      if (pts.get_movet() == lagrangian) {
        for (auto& src : _bdry) {
          if (std::holds_alternative<Surfaces<S>>(src)) {
            Surfaces<S>& surf = std::get<Surfaces<S>>(src);
            (void)clear_inner_panp2<S>(_method, surf, pts, _cutoff_factor, _ips);
          }
        }
      }
    }
  }
}

template <class S>
void reflect_panp2(Surfaces<S> const& _src, Points<S>& _targ) {
  static_assert(sizeof(S) == sizeof(float), "omega2d_b200 stores float32, like the reference's STORE");
  const auto& sx = _src.get_pos();
  const auto& sn = _src.get_norm();
  auto& tx = _targ.get_pos();
  long long moved = 0;
  o2d_host::check(o2d_reflect_interior(o2d_host::device(), O2D_MEM_HOST,
                                       (int)_src.get_npanels(), (int)_src.get_n(), sx[0].data(), sx[1].data(),
                                       _src.get_idx().data(), sn[0].data(), sn[1].data(),
                                       (int)_targ.get_n(), tx[0].data(), tx[1].data(), &moved),
                  "reflect_interior");
  if (o2d_host::verbose()) std::printf("    reflected %lld particles\n", moved);
}

template <class S>
void reflect_interior(std::vector<Collection> const& _bdry, std::vector<Collection>& _vort) {
  for (auto& targ : _vort) {
    if (std::holds_alternative<Points<S>>(targ)) {
      Points<S>& pts = std::get<Points<S>>(targ);
      for (auto& src : _bdry) {
        if (std::holds_alternative<Surfaces<S>>(src)) {
          (void)reflect_panp2<S>(std::get<Surfaces<S>>(src), pts);
        }
      }
    }
  }
}
