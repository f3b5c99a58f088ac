#!/usr/bin/env bash
# Compile checks of the C++ host mirror (no GPU needed):
#  1. tests/cpp/test_find_vels.cpp against the mirror containers, linked with libomega2d_b200.so
#     -> build/test_find_vels (run by tests/test_host_mirror.py on the GPU box);
#  2. where /root/reference exists: InfluenceB200.h against the REFERENCE's own containers
#     (-fsyntax-only), proving the header is a drop-in for src/Influence.h + src/InfluenceVort.h.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${O2D_REFERENCE:-/root/reference}"
mkdir -p "$ROOT/build"
g++ -std=c++17 -O2 -Wall -I"$HERE" "$ROOT/tests/cpp/test_find_vels.cpp" \
    -L"$ROOT/omega2d_b200" -lomega2d_b200 -Wl,-rpath,'$ORIGIN/../omega2d_b200' -o "$ROOT/build/test_find_vels"
echo "built build/test_find_vels"
g++ -std=c++17 -O2 -Wall -I"$HERE" "$ROOT/examples/speedtest20kx10.cpp" \
    -L"$ROOT/omega2d_b200" -lomega2d_b200 -Wl,-rpath,'$ORIGIN/../omega2d_b200' -o "$ROOT/build/speedtest20kx10"
echo "built build/speedtest20kx10"
g++ -std=c++17 -O2 -Wall -I"$HERE" "$ROOT/examples/flow_over_circle_step.cpp" \
    -L"$ROOT/omega2d_b200" -lomega2d_b200 -Wl,-rpath,'$ORIGIN/../omega2d_b200' -o "$ROOT/build/flow_over_circle_step"
echo "built build/flow_over_circle_step"
if [ -d "$REF/src" ]; then
  cat > "$ROOT/build/ref_containers_check.cpp" <<'CPP'
// the reference's containers + this repository's influence header instead of Influence.h
#include "Collection.h"
#include "ResultsType.h"
#include "ExecEnv.h"
#define OMEGA2D_REFERENCE_HEADERS
#include "InfluenceB200.h"
#include "ConvectionB200.h"
#include "CoefficientsB200.h"
#include "ReflectB200.h"
#include "VtkB200.h"
std::string vtk(const Points<float>& p) { return write_vtk_points(p, 0, 1, 0.5); }
void clear(std::vector<Collection>& b, std::vector<Collection>& v) { clear_inner_layer<float>(1, b, v, 0.2f, 0.02f); }
void reflect(std::vector<Collection>& b, std::vector<Collection>& v) { reflect_interior<float>(b, v); }
Vector<float> coeff_block(Collection& a, Collection& b) { return std::visit(CoefficientVisitor<float>(), a, b); }
std::vector<float> rhs_block(Collection& a) { return std::visit(RHSVisitor<float>(), a); }
void instantiate(std::vector<Collection>& a, std::vector<Collection>& b, std::vector<Collection>& c) {
  find_vels<float,double>({0.0, 0.0}, a, b, c, velandvort, true);
  find_vort<float,double>(a, b, c);
}
CPP
  g++ -std=c++17 -fsyntax-only -w -I"$REF/src" -I"$REF/lib" -I"$REF/lib/tinyexpr" -I"$REF/lib/tinyxml2" -I"$HERE" \
      "$ROOT/build/ref_containers_check.cpp"
  echo "InfluenceB200.h compiles against the reference's own Points/Surfaces/Volumes/Collection (InfluenceVisitor, CoefficientVisitor, RHSVisitor)"
fi
