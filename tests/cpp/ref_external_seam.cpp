// tests/cpp/ref_external_seam.cpp -- TEST INFRASTRUCTURE.  INTEGRATION.md section 1 end to end: the reference's
// UNMODIFIED src/Influence.h compiled with -DEXTERNAL_VEL_SOLVE (what CMake sets for OMEGA2D_USE_ONBODY,
// reference CMakeLists.txt:113-117).  Its points_affect_points<float,double> then leaves through the two extern "C"
// symbols it declares (src/Influence.h:31-35, call sites :82-88) -- which this program resolves from
// libomega2d_b200.so instead of onbody's treecode.  The result is compared with the reference's internal CPU loop on
// the same inputs (oracle/_ref/libo2dref_v2.so, linked -Bsymbolic).
// Built only where /root/reference exists (oracle/build_ref.sh -> oracle/_ref/ref_external_seam_v2).
#include "Collection.h"
#include "Influence.h"

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#ifndef EXTERNAL_VEL_SOLVE
#error "compile with -DEXTERNAL_VEL_SOLVE: the point is the reference's own external branch"
#endif

extern "C" int o2dref_points_on_points(int results_type, int ns, const float* sx, const float* sy, const float* ss,
                                       const float* sr, int nt, const float* tx, const float* ty, const float* tr,
                                       float* tu, float* tv, float* tq);

int main() {
  const size_t ns = 5000, nt = 1800;
  const float vdelta = 0.043818f;
  std::mt19937 gen(11);
  std::uniform_real_distribution<float> U(-1.0f, 1.0f);
  std::vector<float> sx(ns), sy(ns), ss(ns), sr(ns, vdelta), ps(2 * ns), tx(nt), ty(nt), tr(nt, vdelta), pt(2 * nt);
  for (size_t i = 0; i < ns; ++i) { sx[i] = U(gen); sy[i] = U(gen); ss[i] = U(gen) * 1e-3f; ps[2*i] = sx[i]; ps[2*i+1] = sy[i]; }
  for (size_t i = 0; i < nt; ++i) { tx[i] = 1.5f * U(gen); ty[i] = 1.5f * U(gen); pt[2*i] = tx[i]; pt[2*i+1] = ty[i]; }
  const Points<float> src(ElementPacket<float>(ps, std::vector<Int>(), ss, ns, 0), active, lagrangian, nullptr, vdelta);
  const ExecEnv external(false, direct, cpu_x86);       // "not internal": the branch that calls the plugin symbols
  int bad = 0;
  for (int inert_targets = 0; inert_targets < 2; ++inert_targets) {
    Points<float> targ = inert_targets
        ? Points<float>(ElementPacket<float>(pt, std::vector<Int>(), std::vector<float>(), nt, 0), inert, lagrangian, nullptr, 0.0f)
        : Points<float>(ElementPacket<float>(pt, std::vector<Int>(), std::vector<float>(nt, 0.0f), nt, 0), active, lagrangian, nullptr, vdelta);
    targ.zero_vels();
    points_affect_points<float,double>(src, targ, ResultsType(velonly), external);   // the reference's own code
    std::vector<float> wu(nt, 0.0f), wv(nt, 0.0f);
    o2dref_points_on_points(velonly, (int)ns, sx.data(), sy.data(), ss.data(), sr.data(), (int)nt, tx.data(), ty.data(),
                            inert_targets ? nullptr : tr.data(), wu.data(), wv.data(), nullptr);
    const auto& u = targ.get_vel();
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < nt; ++i) {
      num = std::max(num, std::hypot((double)u[0][i] - wu[i], (double)u[1][i] - wv[i]));
      den = std::max(den, std::hypot((double)wu[i], (double)wv[i]));
    }
    const double err = num / den;
    std::printf("  reference points_affect_points -> %s -> libomega2d_b200: max rel err %.2e %s\n",
                inert_targets ? "external_vel_solver_f_   " : "external_vel_solver_tr_f_", err, err < 1e-5 ? "ok" : "MISMATCH");
    if (!(err < 1e-5) || !(den > 0.0)) ++bad;
  }
  std::printf("ref_external_seam: %s\n", bad ? "FAILED" : "all checks passed");
  return bad ? 1 : 0;
}
