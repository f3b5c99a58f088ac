// tests/cpp/test_find_vels.cpp -- C++ consumer of the host mirror: builds collections from
// ElementPackets exactly like Simulation::add_elements does (reference src/Simulation.cpp:826-912),
// runs find_vels / find_vort through InfluenceVisitor -> C ABI -> CUDA, and dumps inputs and
// results as raw float32 for tests/test_host_mirror.py to check against the oracle.
#include "ConvectionB200.h"
#include "CoefficientsB200.h"
#include "ReflectB200.h"
#include "VtkB200.h"

#include <cstdio>
#include <cstdint>
#include <vector>

static uint32_t lcg_state = 12345u;
static float urand() { lcg_state = lcg_state * 1664525u + 1013904223u; return (lcg_state >> 8) * (1.0f / 16777216.0f); }

static void dump(FILE* f, const std::vector<float>& v) {
  const uint64_t n = v.size();
  fwrite(&n, sizeof(n), 1, f);
  fwrite(v.data(), sizeof(float), v.size(), f);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s out.bin [nvort] [npanels] [nfld]\n", argv[0]); return 2; }
  const size_t nv = argc > 2 ? atoi(argv[2]) : 3000;
  const size_t np = argc > 3 ? atoi(argv[3]) : 143;
  const size_t nf = argc > 4 ? atoi(argv[4]) : 500;
  const float vdelta = 0.043818f;

  // vortex particles in an annulus around a unit-diameter circle
  std::vector<float> px(2 * nv), pval(nv);
  for (size_t i = 0; i < nv; ++i) {
    const float rho = 0.55f + 0.9f * urand(), th = 6.2831853f * urand();
    px[2*i] = rho * std::cos(th); px[2*i+1] = rho * std::sin(th);
    pval[i] = (urand() - 0.5f) * 2e-3f;
  }
  Points<float> vortons(ElementPacket<float>(px, {}, pval, nv, 0), active, lagrangian, vdelta);

  // closed circle of panels, clockwise, vortex + source sheet strengths
  std::vector<float> nx(2 * np), sval(2 * np);
  std::vector<Int> idx(2 * np);
  for (size_t i = 0; i < np; ++i) {
    const float th = -6.2831853f * (float)i / (float)np;
    nx[2*i] = 0.5f * std::cos(th); nx[2*i+1] = 0.5f * std::sin(th);
    idx[2*i] = (Int)i; idx[2*i+1] = (Int)((i + 1) % np);
    sval[2*i] = 2.0f * std::sin(th) + 0.1f; sval[2*i+1] = 0.3f * std::cos(th);
  }
  Surfaces<float> body(ElementPacket<float>(nx, idx, sval, np, 1), active, fixed);

  // tracers
  std::vector<float> fx(2 * nf);
  for (size_t i = 0; i < nf; ++i) { fx[2*i] = -1.5f + 3.0f * urand(); fx[2*i+1] = 0.8f + urand(); }
  Points<float> tracers(ElementPacket<float>(fx, {}, {}, nf, 0), inert, lagrangian, 0.0f);

  std::vector<Collection> vort, bdry, fldpt;
  vort.push_back(vortons);
  bdry.push_back(body);
  fldpt.push_back(tracers);
  const std::array<double,Dimensions> fs = {1.0, 0.0};

  // what Convection::find_derivs does after the BEM solve (reference src/Convection.h:241-242)
  find_vels<float,double>(fs, vort, bdry, vort);
  find_vels<float,double>(fs, vort, bdry, fldpt);
  // velocity at the panels themselves, forced like write_vtk does (reference src/Simulation.cpp:408-414)
  find_vels<float,double>(fs, vort, bdry, bdry, velonly, true);

  FILE* f = fopen(argv[1], "wb");
  if (!f) return 3;
  const Points<float>& v = std::get<Points<float>>(vort[0]);
  const Points<float>& t = std::get<Points<float>>(fldpt[0]);
  const Surfaces<float>& b = std::get<Surfaces<float>>(bdry[0]);
  dump(f, v.get_pos()[0]); dump(f, v.get_pos()[1]); dump(f, v.get_str()); dump(f, v.get_rad());
  dump(f, b.get_pos()[0]); dump(f, b.get_pos()[1]); dump(f, b.get_str()); dump(f, b.get_src_str()); dump(f, b.get_area());
  dump(f, t.get_pos()[0]); dump(f, t.get_pos()[1]);
  dump(f, v.get_vel()[0]); dump(f, v.get_vel()[1]);
  dump(f, t.get_vel()[0]); dump(f, t.get_vel()[1]);
  dump(f, b.get_vel()[0]); dump(f, b.get_vel()[1]);
  {   // the collocation points panels_affect_panels used for that (src/Influence.h:783)
    const ElementPacket<float> cp = b.represent_as_particles(0.0001, -1.0);
    std::vector<float> cx(cp.nelem), cy(cp.nelem);
    for (size_t i = 0; i < cp.nelem; ++i) { cx[i] = cp.x[2*i]; cy[i] = cp.x[2*i+1]; }
    dump(f, cx); dump(f, cy);
  }

  // vel + vort on the particles (forced results type) and Gaussian vorticity on the tracers
  find_vels<float,double>(fs, vort, bdry, vort, velandvort, true);
  const Points<float>& v2 = std::get<Points<float>>(vort[0]);
  dump(f, v2.get_vel()[0]); dump(f, v2.get_vel()[1]); dump(f, v2.get_vort());
  find_vort<float,double>(vort, bdry, fldpt);
  dump(f, std::get<Points<float>>(fldpt[0]).get_vort());

  // one particle-only Ralston RK2 step and one RK3 step, all stages on the device
  for (int order = 2; order <= 3; ++order) {
    Points<float> mover = vortons;
    advect_particles<float>(order, 0.01, fs, mover);
    dump(f, mover.get_pos()[0]); dump(f, mover.get_pos()[1]); dump(f, mover.get_vel()[0]); dump(f, mover.get_vel()[1]);
  }

  // BEM block of the body on itself through CoefficientVisitor, and the RHS from the panel velocities
  // found above (what solve_bem does, reference src/BEMHelper.h:96-112,190-200)
  {
    Vector<float> coeffs = std::visit(CoefficientVisitor<float>(), bdry[0], bdry[0]);
    dump(f, coeffs);
    Surfaces<float>& bb = std::get<Surfaces<float>>(bdry[0]);
    for (size_t i = 0; i < np; ++i) bb.get_norm_bcs()[i] = 0.01f * (float)(i % 7);
    dump(f, std::visit(RHSVisitor<float>(), bdry[0]));
  }
  // clear_inner_layer (push) on a copy of the particles shifted towards the body
  {
    std::vector<Collection> near;
    Points<float> shifted = vortons;
    for (size_t i = 0; i < nv; ++i) { shifted.get_pos()[0][i] *= 0.9f; shifted.get_pos()[1][i] *= 0.9f; }
    dump(f, shifted.get_pos()[0]); dump(f, shifted.get_pos()[1]);
    near.push_back(shifted);
    clear_inner_layer<float>(1, bdry, near, 0.5f / std::sqrt(2.0f * 3.14159265f), 0.02191f);
    const Points<float>& out = std::get<Points<float>>(near[0]);
    dump(f, out.get_pos()[0]); dump(f, out.get_pos()[1]);
  }
  // reflect_interior on a copy of the particles pulled inside the body
  {
    std::vector<Collection> inside;
    Points<float> shrunk = vortons;
    for (size_t i = 0; i < nv; ++i) { shrunk.get_pos()[0][i] *= 0.6f; shrunk.get_pos()[1][i] *= 0.6f; }
    dump(f, shrunk.get_pos()[0]); dump(f, shrunk.get_pos()[1]);
    inside.push_back(shrunk);
    reflect_interior<float>(bdry, inside);
    const Points<float>& out = std::get<Points<float>>(inside[0]);
    dump(f, out.get_pos()[0]); dump(f, out.get_pos()[1]);
  }
  // .vtu output of the particles (with vorticity) and of the tracers, next to the dump file
  {
    std::string dir(argv[1]);
    const size_t slash = dir.find_last_of('/');
    dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
    const std::string a = write_vtk_points(std::get<Points<float>>(vort[0]), 0, 7, 0.25, dir);
    const std::string b = write_vtk_points(std::get<Points<float>>(fldpt[0]), 1, 7, 0.25, dir);
    if (a != "part_00_00007.vtu" || b != "fldpt_01_00007.vtu") { fprintf(stderr, "bad vtu names %s %s\n", a.c_str(), b.c_str()); return 4; }
  }
  fclose(f);
  printf("test_find_vels: wrote %s\n", argv[1]);
  return 0;
}
