// examples/flow_over_circle_step.cpp -- the convection step of the reference's examples/flow_over_circle.json through the
// host mirror: a unit-diameter circle of 143 panels (ips = sqrt(6) sqrt(0.02/250) = 0.02191) in a
// freestream (1,0), N vortex particles around it, default Vatistas-2 core, default 2nd-order (Ralston) convection.
//
// Omega2D's own step of that scene is  Convection::advect_2nd_ralston  (reference src/Convection.h:390-466): per stage
// solve_bem (right-hand side from the particles, Eigen GMRES solve), find_vels, move, clear_inner_layer.  Here the whole
// step is ONE advect_with_body() call: the particles are uploaded once and downloaded once, the per-stage work runs on
// the GPU, and the only host work left is the application's linear solve, reached through a callback that sees 2*Np
// floats per stage.  Eigen is not part of this repository's path, so the callback below is a stand-in (sheet strengths
// that cancel the slip at the panel centres) -- an application passes a lambda around its BEM::solve instead.
//   g++ -std=c++17 -O2 -I omega2d_b200/host examples/flow_over_circle_step.cpp -L omega2d_b200 -lomega2d_b200 -o build/flow_over_circle_step
//   build/flow_over_circle_step [N=10000] [steps=20]
#include "ConvectionB200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 10000;
  const int nsteps = argc > 2 ? atoi(argv[2]) : 20;
  const size_t np = 143;
  const double dt = 0.02, re = 250.0;
  const float ips = (float)(std::sqrt(6.0) * std::sqrt(dt / re));
  const float vdelta = 2.0f * ips;
  const std::array<double,Dimensions> fs = {1.0, 0.0};

  std::mt19937 gen(5);
  std::uniform_real_distribution<float> U(0.0f, 1.0f);
  std::vector<float> px(2 * n), ps(n);
  for (size_t i = 0; i < n; ++i) {
    const float rho = std::sqrt(0.27f + (2.25f - 0.27f) * U(gen)), th = 6.2831853f * U(gen);
    px[2*i] = rho * std::cos(th); px[2*i+1] = rho * std::sin(th); ps[i] = (U(gen) - 0.5f) * 2e-3f;
  }
  Points<float> pts(ElementPacket<float>(px, {}, ps, n, 0), active, lagrangian, vdelta);
  std::vector<float> nx(2 * np), sval(2 * np, 0.0f);
  std::vector<Int> idx(2 * np);
  for (size_t i = 0; i < np; ++i) {
    const float th = -6.2831853f * (float)i / (float)np;      // clockwise: fluid on the left of node0 -> node1
    nx[2*i] = 0.5f * std::cos(th); nx[2*i+1] = 0.5f * std::sin(th);
    idx[2*i] = (Int)i; idx[2*i+1] = (Int)((i + 1) % np);
  }
  Surfaces<float> body(ElementPacket<float>(nx, idx, sval, np, 1), active, fixed);

  const auto& tang = body.get_tang();
  const auto& norm = body.get_norm();
  int solves = 0;
  auto solve = [&](int /*stage*/, double /*time*/, int npan, const float* pu, const float* pv, float* vs, float* ss) {
    ++solves;   // an application builds its right-hand side from pu,pv here and calls its BEM solve
    for (int i = 0; i < npan; ++i) {
      vs[i] = -(pu[i] * tang[0][i] + pv[i] * tang[1][i]);
      if (ss) ss[i] = -(pu[i] * norm[0][i] + pv[i] * norm[1][i]);
    }
  };

  advect_with_body<float>(2, 0.0, dt, fs, ips, pts, body, solve);     // context creation, first-launch costs
  solves = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < nsteps; ++k) advect_with_body<float>(2, (k + 1) * dt, dt, fs, ips, pts, body, solve);
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const double pairs = 2.0 * (double)n * ((double)n + 2.0 * (double)np);
  std::printf("flow_over_circle step: %zu particles + %zu panels, Ralston RK2 with BEM right-hand side, solve callback, find_vels, move and\n"
              "  clear_inner_layer per stage: %d steps in [%.6f] seconds = %.3f ms per step (%d solves), %.3e interactions/s\n",
              n, np, nsteps, sec, 1e3 * sec / nsteps, solves, nsteps * pairs / sec);
  return 0;
}
