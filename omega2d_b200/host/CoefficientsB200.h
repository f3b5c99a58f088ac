// CoefficientsB200.h -- drop-in replacement for the reference's src/Coefficients.h and src/RHS.h.
//
// panels_on_panels_coeff<S> keeps the reference's signature and return value -- the column-major
// influence-coefficient block as a Vector<S>, scaled by 1/(2 pi), augmented with the extra row /
// column when either surface is augmented (reference src/Coefficients.h:155-573) -- but the O(Np^2)
// pair loop is one call into libomega2d_b200.so (o2d_panels_on_panels_coeff), which writes the block
// straight into the augmented matrix (leading dimension = augmented row count).  The O(Np)
// augmentation entries and vels_to_rhs_panels (src/RHS.h:54-116) stay on the host, as plain loops.
// Like InfluenceB200.h this header works with the host containers shipped next to it or, with
// OMEGA2D_REFERENCE_HEADERS defined, with the reference's own.
#pragma once
#include "InfluenceB200.h"

#include <cmath>
#include <vector>

template <class S>
Vector<S> panels_on_panels_coeff(Surfaces<S> const& src, Surfaces<S>& targ) {
  static_assert(sizeof(S) == sizeof(float), "omega2d_b200 stores float32, like the reference's STORE");
  const size_t nsrc = src.get_npanels();
  const size_t ntarg = targ.get_npanels();
  const bool both = (src.num_unknowns_per_panel() == 2) and (targ.num_unknowns_per_panel() == 2);
  const size_t oldncols = nsrc * src.num_unknowns_per_panel();
  const size_t oldnrows = ntarg * targ.num_unknowns_per_panel();
  const size_t nrows = oldnrows + (targ.is_augmented() ? 1 : 0);
  const size_t ncols = oldncols + (src.is_augmented() ? 1 : 0);
  const bool same = ((const void*)&src == (const void*)&targ);

  Vector<S> coeffs;
  coeffs.resize(nrows * ncols);
  const auto& sx = src.get_pos();
  const auto& tx = targ.get_pos();
  const auto& tt = targ.get_tang();
  const auto& tn = targ.get_norm();
  const Vector<S>& sa = src.get_area();
  o2d_host::CallTimer timer("panels_on_panels_coeff",
                            (double)nsrc * (double)ntarg * 2.0 * (4.0 + (both ? 45.0 : 39.0) + (both ? 12.0 : 3.0)));
  o2d_host::check(o2d_panels_on_panels_coeff(o2d_host::device(), O2D_MEM_HOST,
                                             (int)nsrc, (int)src.get_n(), sx[0].data(), sx[1].data(), src.get_idx().data(), sa.data(),
                                             (int)ntarg, (int)targ.get_n(), tx[0].data(), tx[1].data(), targ.get_idx().data(),
                                             targ.get_area().data(), tt[0].data(), tt[1].data(), tn[0].data(), tn[1].data(),
                                             both ? 1 : 0, same ? 1 : 0, (long long)nrows, coeffs.data()),
                  "panels_on_panels_coeff");
  if (not targ.is_augmented() and not src.is_augmented()) return coeffs;

  // bottom row: panel lengths of the source body (src/Coefficients.h:466-492)
  if (targ.is_augmented()) {
    for (size_t j = 0; j < oldncols; j++) {
      S& last = coeffs[j * nrows + oldnrows];
      if (same) {
        if (targ.num_unknowns_per_panel() == 1) last = sa[j];
        else if (j % 2 == 0) last = sa[j / 2];
        else last = 0.0;
      } else {
        last = 0.0;
      }
    }
  }
  // last column: velocity of the source body at unit rotation rate on the target panels, and the
  // circulation at unit rotation in the corner (src/Coefficients.h:507-535)
  if (src.is_augmented()) {
    S* a_iter = coeffs.data() + oldncols * nrows;
    const auto& vel = targ.get_vel();
    const bool targ_have_src = (targ.num_unknowns_per_panel() == 2);
    for (size_t i = 0; i < ntarg; ++i) {
      *a_iter++ = vel[0][i] * tt[0][i] + vel[1][i] * tt[1][i];
      if (targ_have_src) *a_iter++ = vel[0][i] * tn[0][i] + vel[1][i] * tn[1][i];
    }
    if (targ.is_augmented()) *a_iter = same ? (S)(2.0 * src.get_vol()) : (S)0.0;
  }
  return coeffs;
}

// the reference's other coefficient functions are stubs returning empty blocks (src/Coefficients.h:40-153,574-609)
template <class S> Vector<S> points_on_points_coeff(Points<S> const&, Points<S>&) { return Vector<S>(); }
template <class S> Vector<S> panels_on_points_coeff(Surfaces<S> const&, Points<S>&) { return Vector<S>(); }
template <class S> Vector<S> points_on_panels_coeff(Points<S> const&, Surfaces<S>&) { return Vector<S>(); }

// helper struct for dispatching through a variant -- reference src/Coefficients.h:611-624
template <class S>
struct CoefficientVisitor {
  Vector<S> operator()(Points<S> const& src,   Points<S>& targ)   { return points_on_points_coeff<S>(src, targ); }
  Vector<S> operator()(Surfaces<S> const& src, Points<S>& targ)   { return panels_on_points_coeff<S>(src, targ); }
  Vector<S> operator()(Points<S> const& src,   Surfaces<S>& targ) { return points_on_panels_coeff<S>(src, targ); }
  Vector<S> operator()(Surfaces<S> const& src, Surfaces<S>& targ) { return panels_on_panels_coeff<S>(src, targ); }
#ifdef OMEGA2D_REFERENCE_HEADERS
  Vector<S> operator()(Volumes<S> const&,  Points<S>&)   { return Vector<S>(); }
  Vector<S> operator()(Volumes<S> const&,  Surfaces<S>&) { return Vector<S>(); }
  Vector<S> operator()(Points<S> const&,   Volumes<S>&)  { return Vector<S>(); }
  Vector<S> operator()(Surfaces<S> const&, Volumes<S>&)  { return Vector<S>(); }
  Vector<S> operator()(Volumes<S> const&,  Volumes<S>&)  { return Vector<S>(); }
#endif
};

// velocities at the panel centres -> right-hand side of the BEM system -- reference src/RHS.h:54-116
template <class S>
std::vector<S> vels_to_rhs_panels(Surfaces<S> const& targ) {
  const auto& tu = targ.get_vel();
  const auto& tt = targ.get_tang();
  const auto& tn = targ.get_norm();
  const Vector<S>& tnbc = targ.get_norm_bcs();
  const size_t ntarg = targ.get_npanels();
  const size_t nunk = targ.num_unknowns_per_panel();
  std::vector<S> rhs;
  rhs.resize(ntarg * nunk);
  if (nunk == 1) {
    for (size_t i = 0; i < ntarg; i++) rhs[i] = -(tu[0][i]*tt[0][i] + tu[1][i]*tt[1][i]);
  } else {
    for (size_t i = 0; i < ntarg; i++) {
      rhs[2*i]   = -(tu[0][i]*tt[0][i] + tu[1][i]*tt[1][i]);
      rhs[2*i+1] = -(tu[0][i]*tn[0][i] + tu[1][i]*tn[1][i]);
      rhs[2*i+1] += tnbc[i];
    }
  }
  return rhs;
}
template <class S> std::vector<S> vels_to_rhs_points(Points<S> const&) { return std::vector<S>(); }

// reference src/RHS.h:133-139
template <class S>
struct RHSVisitor {
  std::vector<S> operator()(Points<S> const& targ)   { return vels_to_rhs_points<S>(targ); }
  std::vector<S> operator()(Surfaces<S> const& targ) { return vels_to_rhs_panels<S>(targ); }
#ifdef OMEGA2D_REFERENCE_HEADERS
  std::vector<S> operator()(Volumes<S> const&) { return std::vector<S>(); }
#endif
};
