// Test harness: autopas_b200/csrc/vtk_format.cuh (the formatter of the device-side checkpoint writer) compiled for the
// host, so that tests/test_vtk.py can compare it with the C library's printf over many values without a GPU.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "vtk_format.cuh"

extern "C" {
int fmt_g(double v, int P, char *out) { return apbFormatG(v, P, out); }
int fmt_u64(uint64_t v, char *out) { return apbFormatU64(v, out); }
int position_precision(double position, double border) {
  static ApbVtkTables t;
  static bool built = false;
  if (!built) {
    apbVtkBuildTables(t);
    built = true;
  }
  return apbVtkPositionPrecision(t, position, border);
}
static int64_t checkOne(double v, int P, char *firstBad) {
  char mine[64], want[64];
  const int n = apbFormatG(v, P, mine);
  mine[n] = 0;
  std::snprintf(want, sizeof want, "%.*g", P, v);
  if (std::strcmp(mine, want) != 0) {
    if (firstBad && !firstBad[0]) std::snprintf(firstBad, 200, "P=%d v=%a mine=%s want=%s", P, v, mine, want);
    return 1;
  }
  return 0;
}
// mode 0: random bit patterns (all exponents, subnormals); 1: MD-like magnitudes 1e-8 .. 1e8; 2: values with few decimal
// digits (exact ties of the rounding: k / 2^j); 3: neighbours of powers of ten. Returns the number of mismatches.
int64_t check_random(uint64_t seed, int64_t count, int mode, char *firstBad) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  if (firstBad) firstBad[0] = 0;
  for (int64_t i = 0; i < count; ++i) {
    double v;
    if (mode == 0) {
      uint64_t b = rng();
      if (((b >> 52) & 0x7ff) == 0x7ff) b &= ~(1ull << 62);
      std::memcpy(&v, &b, 8);
    } else if (mode == 1) {
      v = std::ldexp(static_cast<double>(rng() >> 11), -53) * std::pow(10., static_cast<int>(rng() % 17) - 8);
      if (rng() & 1) v = -v;
    } else if (mode == 2) {
      v = std::ldexp(static_cast<double>(rng() % 20000001), -static_cast<int>(rng() % 24)) * std::pow(10., static_cast<int>(rng() % 7) - 3);
    } else {
      const int n = static_cast<int>(rng() % 640) - 320;
      char text[16];
      std::snprintf(text, sizeof text, "%de%d", 1 + static_cast<int>(rng() % 9), n);
      v = std::strtod(text, nullptr);
      const int steps = static_cast<int>(rng() % 9) - 4;
      for (int s = 0; s < std::abs(steps); ++s) v = std::nextafter(v, steps > 0 ? 1e309 : 0.);
    }
    for (int P = 6; P <= 17; ++P) bad += checkOne(v, P, firstBad);
    bad += checkOne(v, 1, firstBad);
    bad += checkOne(v, 3, firstBad);
  }
  return bad;
}
}
