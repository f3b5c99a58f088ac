// examples/speedtest20kx10.cpp -- the reference's timing scene examples/speedtest20kx10.json through the host mirror.
//
// Omega2D's own run of that file (Omega2Dbatch.bin examples/speedtest20kx10.json) needs its JSON/feature/BEM layers,
// which are outside this repository's path; what the scene DOES is small: 20 000 vortex particles ("block of random",
// centre (0,0), size 2 x 2 -> x,y ~ U[-2,2]; circulation ~ U[-1,1]), every particle with the core radius
// vdelta = 2 sqrt(6) sqrt(dt/Re) = 0.034641 (Re 100, dt 0.005), no bodies, "viscous": "none", 10 steps of the default
// 2nd-order (Ralston) convection with the default Vatistas-2 core.  This program builds that collection with the
// mirror's Points<float> (same constructor arguments as Simulation::add_elements passes) and runs it two ways:
//   per step : one advect_particles() call per step  -- one upload + download per step, like a time loop would
//   resident : advect_particles(..., nsteps = 10)     -- the particles stay on the GPU between steps
// and prints the reference's own style of timing line for each.  Optional argument: a file that receives the
// initial and final positions as raw float32 (tests/test_host_mirror.py checks them against the oracle).
//   g++ -std=c++17 -O2 -I omega2d_b200/host examples/speedtest20kx10.cpp -L omega2d_b200 -lomega2d_b200 -o build/speedtest20kx10
#include "ConvectionB200.h"

#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

int main(int argc, char** argv) {
  const size_t n = 20000;
  const int nsteps = 10;
  const double dt = 0.005, re = 100.0;
  const float vdelta = (float)(2.0 * std::sqrt(6.0) * std::sqrt(dt / re));
  const std::array<double,Dimensions> fs = {0.0, 0.0};

  std::mt19937 gen(12345);
  std::uniform_real_distribution<float> pos(-2.0f, 2.0f), str(-1.0f, 1.0f);
  std::vector<float> x(2 * n), s(n);
  for (size_t i = 0; i < n; ++i) { x[2*i] = pos(gen); x[2*i+1] = pos(gen); s[i] = str(gen); }
  const Points<float> initial(ElementPacket<float>(x, {}, s, n, 0), active, lagrangian, vdelta);

  const double pairs_per_step = 2.0 * (double)n * (double)n;                          // two evaluations per RK2 step
  const double flops_per_step = pairs_per_step * o2d_flops_per_pair(0, O2D_CORE_V2, O2D_VELONLY, 1);
  auto report = [&](const char* what, double seconds) {
    std::printf("  %-9s %d steps of %zu particles: [%.6f] seconds, %.3f ms per step, %.3e interactions/s, %.1f GFlop/s (reference flops)\n",
                what, nsteps, n, seconds, 1e3 * seconds / nsteps, nsteps * pairs_per_step / seconds,
                1e-9 * nsteps * flops_per_step / seconds);
  };

  Points<float> warm = initial;
  advect_particles<float>(2, dt, fs, warm);                                            // context creation, first-launch costs

  Points<float> a = initial;
  auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < nsteps; ++k) advect_particles<float>(2, dt, fs, a);
  report("per step", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

  Points<float> b = initial;
  t0 = std::chrono::steady_clock::now();
  advect_particles<float>(2, dt, fs, b, nsteps);
  report("resident", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

  bool same = true;
  for (size_t d = 0; d < Dimensions; ++d)
    for (size_t i = 0; i < n; ++i) same = same && (a.get_pos()[d][i] == b.get_pos()[d][i]);
  std::printf("  per-step and resident runs agree bit for bit: %s\n", same ? "yes" : "NO");

  // the numbers Simulation::dump_stats_to_status would write for this collection
  float circ = 0, ix = 0, iy = 0, smax = 0;
  o2d_host::check(o2d_points_status(o2d_host::device(), O2D_MEM_HOST, (int)n, b.get_pos()[0].data(), b.get_pos()[1].data(),
                                    b.get_str().data(), &circ, &ix, &iy, &smax), "points_status");
  std::printf("  status: total circulation %.6g, impulse (%.6g, %.6g), max |strength| %.6g\n", circ, ix, iy, smax);

  if (argc > 1) {
    FILE* f = std::fopen(argv[1], "wb");
    if (!f) return 3;
    const Points<float>* both[2] = {&initial, &b};
    for (const Points<float>* p : both)
      for (size_t d = 0; d < Dimensions; ++d) std::fwrite(p->get_pos()[d].data(), sizeof(float), n, f);
    std::fwrite(initial.get_str().data(), sizeof(float), n, f);
    std::fclose(f);
  }
  return same ? 0 : 1;
}
