// tests/cpp/mirror_repr.cpp -- CPU only: the host mirror's Surfaces::represent_as_particles on a body with NON-UNIFORM
// panel lengths (some panels long enough to shed several points), dumped as text for tests/test_host_containers.py, which
// compares it with the reference's own output for the same body (tests/golden/ref_golden_repr.npz).
//   g++ -std=c++17 -O2 -I omega2d_b200/host tests/cpp/mirror_repr.cpp -o build/mirror_repr   (no library needed)
#include "Surfaces.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s body.txt offset ips\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "r");
  if (!f) return 3;
  int np = 0;
  if (std::fscanf(f, "%d", &np) != 1) return 3;
  std::vector<float> x(2 * np), val(2 * np);
  std::vector<Int> idx(2 * np);
  for (int i = 0; i < np; ++i) {
    if (std::fscanf(f, "%f %f %f", &x[2*i], &x[2*i+1], &val[2*i]) != 3) return 3;
    val[2*i+1] = 0.0f;
    idx[2*i] = (Int)i; idx[2*i+1] = (Int)((i + 1) % np);
  }
  std::fclose(f);
  Surfaces<float> body(ElementPacket<float>(x, idx, val, (size_t)np, 1), active, fixed);
  const ElementPacket<float> pk = body.represent_as_particles((float)std::atof(argv[2]), (float)std::atof(argv[3]));
  std::printf("%zu\n", pk.nelem);
  for (size_t i = 0; i < pk.nelem; ++i) std::printf("%.9g %.9g %.9g\n", pk.x[2*i], pk.x[2*i+1], pk.val[i]);
  return 0;
}
