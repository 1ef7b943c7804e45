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

extern "C" {
double parse_double(const char *s, int len, int *status) { return apbParseDouble(s, len, *status); }
uint64_t parse_u64(const char *s, int len, int *status) { return apbParseU64(s, len, *status); }
static int64_t parseOne(const char *text, char *firstBad) {
  int status = 0;
  const double mine = apbParseDouble(text, static_cast<int>(std::strlen(text)), status);
  const double want = std::strtod(text, nullptr);
  uint64_t a, b;
  std::memcpy(&a, &mine, 8);
  std::memcpy(&b, &want, 8);
  if (status != 0 || a != b) {
    if (firstBad && !firstBad[0]) std::snprintf(firstBad, 200, "text=%s status=%d mine=%a want=%a", text, status, mine, want);
    return 1;
  }
  return 0;
}
// mode 0: "%.{6..17}g" of random bit patterns; 1: random digit strings (1 ... 40 digits, point anywhere, exponent -360 ... 330);
// 2: exact ties and their neighbours (odd integers above 2^53 scaled by powers of two / ten, subnormal borders)
int64_t check_parse(uint64_t seed, int64_t count, int mode, char *firstBad) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  if (firstBad) firstBad[0] = 0;
  char text[128];
  for (int64_t i = 0; i < count; ++i) {
    if (mode == 0) {
      uint64_t b = rng();
      if (((b >> 52) & 0x7ff) == 0x7ff) b &= ~(1ull << 62);
      double v;
      std::memcpy(&v, &b, 8);
      if (rng() & 1) v = std::ldexp(static_cast<double>(rng() >> 11), -53) * std::pow(10., static_cast<int>(rng() % 17) - 8);
      for (int P = 6; P <= 17; ++P) {
        std::snprintf(text, sizeof text, "%.*g", P, v);
        bad += parseOne(text, firstBad);
      }
    } else if (mode == 1) {
      const int nd = 1 + static_cast<int>(rng() % 40), point = static_cast<int>(rng() % (nd + 2)) - 1;
      int len = 0;
      if (rng() % 3 == 0) text[len++] = (rng() & 1) ? '-' : '+';
      for (int d = 0; d < nd; ++d) {
        if (d == point) text[len++] = '.';
        text[len++] = static_cast<char>('0' + rng() % 10);
      }
      if (rng() % 4) len += std::snprintf(text + len, 16, "%c%d", (rng() & 1) ? 'e' : 'E', static_cast<int>(rng() % 691) - 360);
      text[len] = 0;
      bad += parseOne(text, firstBad);
    } else {
      // a 54-bit odd integer is exactly half way between two doubles; printed exactly (and with a digit changed) at
      // several binary / decimal scales
      const uint64_t odd = ((1ull << 53) | (rng() >> 11)) | 1ull;
      const int p2 = static_cast<int>(rng() % 40);
      const __uint128_t big = static_cast<__uint128_t>(odd) << p2;
      char digits[64];
      int n = 0;
      for (__uint128_t t = big; t; t /= 10) digits[n++] = static_cast<char>('0' + static_cast<int>(t % 10));
      for (int variant = 0; variant < 3; ++variant) {
        int len = 0;
        for (int d = n - 1; d >= 0; --d) text[len++] = digits[d];
        if (variant == 1) text[len - 1] = text[len - 1] == '9' ? '8' : static_cast<char>(text[len - 1] + 1);
        if (variant == 2) { text[len++] = '.'; text[len++] = '0'; text[len++] = '0'; text[len++] = '1'; }
        len += std::snprintf(text + len, 16, "e%d", static_cast<int>(rng() % 600) - 340);
        text[len] = 0;
        bad += parseOne(text, firstBad);
      }
      // around the smallest subnormal and the largest finite value
      std::snprintf(text, sizeof text, "%d.%de-324", static_cast<int>(rng() % 10), static_cast<int>(rng() % 100000));
      bad += parseOne(text, firstBad);
      std::snprintf(text, sizeof text, "1.797693134862315%de308", static_cast<int>(rng() % 1000));
      bad += parseOne(text, firstBad);
    }
  }
  return bad;
}
}
