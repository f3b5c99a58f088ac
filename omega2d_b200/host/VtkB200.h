// VtkB200.h -- drop-in replacement for Points<S>::write_vtk (reference src/Points.h:556-740).
//
// Same file name ("part_II_FFFFF.vtu" for vortex particles, "fldpt_..." for inert field points), same return
// value (the name), same bytes in the file -- but the (x,y,0) interleave, the connectivity/offsets/types arrays
// and the base64 text of every DataArray are produced by ONE kernel launch of libomega2d_b200.so
// (o2d_write_vtu_points) instead of VtkXmlWriter + cppcodec + tinyxml2's character-wise printer.
// A free function, because the reference's member lives inside its Points class: call
//     write_vtk_points(pts, index, frameno, time)     where the reference calls   pts.write_vtk(index, frameno, time)
// (src/Simulation.cpp:416-428).  Works with the mirror containers next to this file or, with
// OMEGA2D_REFERENCE_HEADERS defined, with the reference's own Points<S>.
#pragma once
#include "InfluenceB200.h"

#include <cstdio>
#include <string>

template <class S>
std::string write_vtk_points(const Points<S>& _pts, const size_t _index, const size_t _frameno, const double _time,
                             const std::string& _dir = std::string()) {
  static_assert(sizeof(S) == sizeof(float), "omega2d_b200 stores float32, like the reference's STORE");
  assert(_pts.get_n() > 0 && "Inside write_vtk with no points");
  const bool inert_pts = _pts.is_inert();          // no strengths, no radii, other prefix (src/Points.h:561-568)
  char name[64];
  std::snprintf(name, sizeof(name), "%s%02zu_%05zu.vtu", inert_pts ? "fldpt_" : "part_", _index, _frameno);
  const std::string path = _dir.empty() ? std::string(name) : _dir + "/" + name;
  const auto& x = _pts.get_pos();
  const auto& u = _pts.get_vel();
  o2d_host::check(o2d_write_vtu_points(o2d_host::device(), O2D_MEM_HOST, path.c_str(), (int)_pts.get_n(),
                                       x[0].data(), x[1].data(),
                                       inert_pts ? nullptr : _pts.get_str().data(),
                                       inert_pts ? nullptr : _pts.get_rad().data(),
                                       _pts.has_vort() ? _pts.get_vort().data() : nullptr,
                                       _pts.has_shear() ? _pts.get_shear().data() : nullptr,
                                       u[0].data(), u[1].data(), _time),
                  "write_vtu_points");
  if (o2d_host::verbose()) std::printf("Wrote %zu points to %s\n", (size_t)_pts.get_n(), name);
  return std::string(name);
}
