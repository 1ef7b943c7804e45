// Decimal text of doubles and integers, byte for byte what a default-constructed std::ostream (= printf "%.{P}g" / "%lu")
// prints, for the device-side checkpoint writer (vtk.cu; examples/md-flexible/src/ParallelVtkWriter.cpp:79-166 streams
// every value through `timestepFile << value`).
//
// "%.{P}g" needs the first P significant decimal digits of the binary value, correctly rounded (glibc rounds the exact
// value, ties to even). A double is m * 2^e with a 53-bit m, so the scaled value m * 2^e * 10^k, k = P - 1 - floor(log10 v),
// is computed exactly in a small multi-word integer (at most 37 words of 32 bits for the whole double range; the
// reading direction below needs up to 46): powers of
// ten enter by multiplications / divisions with 10^9, powers of two by shifts, and what falls off the low end decides
// the rounding (above / exactly / below one half). No floating-point arithmetic is involved, so host and device agree
// by construction; tests/test_vtk.py compiles this header for the host and compares it with printf over tens of
// millions of values (all exponents, all precisions the writer uses, exact ties).
#pragma once
#include <stdint.h>

#include <limits>

#ifdef __CUDACC__
#define APB_HD __host__ __device__ __forceinline__
#else
#define APB_HD static inline
#endif

#define APB_VTK_BIG_WORDS 48

struct ApbBig {
  uint32_t w[APB_VTK_BIG_WORDS];  // little endian
  int n;                          // words in use (w[n - 1] != 0 unless the value is 0 and n == 1)
};

APB_HD void apbBigSet(ApbBig &b, uint64_t v) {
  b.w[0] = static_cast<uint32_t>(v);
  b.w[1] = static_cast<uint32_t>(v >> 32);
  b.n = b.w[1] ? 2 : 1;
}
APB_HD void apbBigMulSmall(ApbBig &b, uint32_t c) {
  uint64_t carry = 0;
  for (int i = 0; i < b.n; ++i) {
    const uint64_t t = static_cast<uint64_t>(b.w[i]) * c + carry;
    b.w[i] = static_cast<uint32_t>(t);
    carry = t >> 32;
  }
  if (carry) b.w[b.n++] = static_cast<uint32_t>(carry);
}
// b /= d, returns the remainder
APB_HD uint32_t apbBigDivSmall(ApbBig &b, uint32_t d) {
  uint64_t rem = 0;
  for (int i = b.n - 1; i >= 0; --i) {
    const uint64_t t = (rem << 32) | b.w[i];
    b.w[i] = static_cast<uint32_t>(t / d);
    rem = t % d;
  }
  while (b.n > 1 && b.w[b.n - 1] == 0) --b.n;
  return static_cast<uint32_t>(rem);
}
APB_HD void apbBigShl(ApbBig &b, int s) {
  const int ws = s >> 5, bs = s & 31;
  if (bs) {
    uint32_t carry = 0;
    for (int i = 0; i < b.n; ++i) {
      const uint32_t v = b.w[i];
      b.w[i] = (v << bs) | carry;
      carry = v >> (32 - bs);
    }
    if (carry) b.w[b.n++] = carry;
  }
  if (ws) {
    for (int i = b.n - 1; i >= 0; --i) b.w[i + ws] = b.w[i];
    for (int i = 0; i < ws; ++i) b.w[i] = 0;
    b.n += ws;
  }
}
// bit i of b (0 beyond the top)
APB_HD uint32_t apbBigBit(const ApbBig &b, int i) { return (i >> 5) < b.n ? (b.w[i >> 5] >> (i & 31)) & 1u : 0u; }
// any bit below position s set?
APB_HD bool apbBigAnyBelow(const ApbBig &b, int s) {
  const int ws = s >> 5, bs = s & 31;
  for (int i = 0; i < ws && i < b.n; ++i)
    if (b.w[i]) return true;
  return ws < b.n && bs && (b.w[ws] & ((1u << bs) - 1u));
}
// floor(b / 2^s) as 64 bits (the caller guarantees it fits)
APB_HD uint64_t apbBigShr64(const ApbBig &b, int s) {
  const int ws = s >> 5, bs = s & 31;
  uint64_t lo = 0, hi = 0, top = 0;
  if (ws < b.n) lo = b.w[ws];
  if (ws + 1 < b.n) hi = b.w[ws + 1];
  if (ws + 2 < b.n) top = b.w[ws + 2];
  uint64_t v = lo | (hi << 32);
  if (bs) v = (v >> bs) | (top << (64 - bs));
  return v;
}

APB_HD uint64_t apbPow10u64(int k) {  // k <= 19
  uint64_t p = 1;
  for (int i = 0; i < k; ++i) p *= 10u;
  return p;
}
APB_HD int apbCountLeadingZeros64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __clzll(static_cast<long long>(v));
#else
  return __builtin_clzll(v);
#endif
}

// The P (1 <= P <= 17) leading decimal digits of |v| (finite, non-zero), correctly rounded to nearest / even, as an
// integer in [10^(P-1), 10^P), and the decimal exponent x of the first digit: |v| ~ digits * 10^(x - P + 1).
APB_HD uint64_t apbDecimalDigits(uint64_t bits, int P, int &xOut) {
  const int biased = static_cast<int>((bits >> 52) & 0x7ff);
  const uint64_t frac = bits & 0xfffffffffffffull;
  const uint64_t m = biased ? (frac | (1ull << 52)) : frac;
  const int e = biased ? biased - 1075 : -1074;
  const int b2 = 63 - apbCountLeadingZeros64(m) + e;  // floor(log2 |v|)
  // floor(b2 log10(2)) is floor(log10 |v|) or one less (|v| in [2^b2, 2^(b2+1)), log10(2) < 1); the product is
  // never within 1e-4 of an integer for |b2| <= 1100, so the double rounding cannot cross one. Integer arithmetic:
  // 1292913986 / 2^32 = log10(2) - 1.6e-11.
  int x = static_cast<int>((static_cast<int64_t>(b2) * 1292913986ll) >> 32);
  const uint64_t limit = apbPow10u64(P);
  uint64_t Q = 0;
  int cmp = -1;  // what was cut off against one half: -1 below, 0 exactly half, +1 above
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int k = P - 1 - x;
    ApbBig big;
    cmp = -1;
    if (k >= 0) {
      apbBigSet(big, m);
      int kk = k;
      for (; kk >= 9; kk -= 9) apbBigMulSmall(big, 1000000000u);
      if (kk) apbBigMulSmall(big, static_cast<uint32_t>(apbPow10u64(kk)));
      if (e >= 0) {
        apbBigShl(big, e);
        Q = apbBigShr64(big, 0);
      } else {
        const int s = -e;
        Q = apbBigShr64(big, s);
        if (apbBigBit(big, s - 1)) cmp = apbBigAnyBelow(big, s - 1) ? 1 : 0;
        else cmp = -1;
      }
    } else {
      bool sticky = false;
      if (e >= 0) {
        apbBigSet(big, m);
        apbBigShl(big, e);
      } else {  // |v| >= 10^P > 1: -e <= 52
        const int s = -e;
        sticky = (m & ((1ull << s) - 1ull)) != 0;
        apbBigSet(big, m >> s);
      }
      int j = -k;
      for (; j > 9; j -= 9) sticky |= apbBigDivSmall(big, 1000000000u) != 0;
      const uint32_t d = static_cast<uint32_t>(apbPow10u64(j));
      const uint32_t r = apbBigDivSmall(big, d);
      if (r > d / 2) cmp = 1;
      else if (r == d / 2) cmp = sticky ? 1 : 0;
      else cmp = -1;
      Q = apbBigShr64(big, 0);
    }
    if (Q < limit) break;
    ++x;  // the estimate was one short: P + 1 digits came out
  }
  if (cmp > 0 || (cmp == 0 && (Q & 1ull))) ++Q;
  if (Q == limit) {
    Q = limit / 10u;
    ++x;
  }
  xOut = x;
  return Q;
}

// printf("%.{P}g", v) into out (at most 24 + P characters), returns the length; P >= 1.
APB_HD int apbFormatG(double v, int P, char *out) {
  uint64_t bits;
#ifdef __CUDA_ARCH__
  bits = static_cast<uint64_t>(__double_as_longlong(v));
#else
  __builtin_memcpy(&bits, &v, 8);
#endif
  int len = 0;
  if (bits >> 63) out[len++] = '-';
  const uint64_t mag = bits & 0x7fffffffffffffffull;
  if (mag >= 0x7ff0000000000000ull) {
    const bool isInf = mag == 0x7ff0000000000000ull;
    out[len++] = isInf ? 'i' : 'n';
    out[len++] = isInf ? 'n' : 'a';
    out[len++] = isInf ? 'f' : 'n';
    return len;
  }
  if (mag == 0) {
    out[len++] = '0';
    return len;
  }
  if (P > 17) P = 17;
  int x;
  uint64_t Q = apbDecimalDigits(mag, P, x);
  // significant digits without trailing zeros (no '#' flag)
  int nd = P;
  while (nd > 1 && Q % 10u == 0) {
    Q /= 10u;
    --nd;
  }
  char dig[20];
  for (int i = nd - 1; i >= 0; --i) {
    dig[i] = static_cast<char>('0' + Q % 10u);
    Q /= 10u;
  }
  if (x < -4 || x >= P) {
    out[len++] = dig[0];
    if (nd > 1) {
      out[len++] = '.';
      for (int i = 1; i < nd; ++i) out[len++] = dig[i];
    }
    out[len++] = 'e';
    int ax = x;
    if (x < 0) {
      out[len++] = '-';
      ax = -x;
    } else {
      out[len++] = '+';
    }
    if (ax >= 100) out[len++] = static_cast<char>('0' + ax / 100);
    out[len++] = static_cast<char>('0' + (ax / 10) % 10);
    out[len++] = static_cast<char>('0' + ax % 10);
  } else if (x >= 0) {
    for (int i = 0; i <= x; ++i) out[len++] = i < nd ? dig[i] : '0';
    if (nd > x + 1) {
      out[len++] = '.';
      for (int i = x + 1; i < nd; ++i) out[len++] = dig[i];
    }
  } else {
    out[len++] = '0';
    out[len++] = '.';
    for (int i = 0; i < -x - 1; ++i) out[len++] = '0';
    for (int i = 0; i < nd; ++i) out[len++] = dig[i];
  }
  return len;
}

// printf("%lu", v), returns the length (at most 20)
APB_HD int apbFormatU64(uint64_t v, char *out) {
  char tmp[20];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + v % 10u);
    v /= 10u;
  } while (v);
  for (int i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
  return n;
}

// ---- writeWithDynamicPrecision (ParallelVtkWriter.cpp:130-157) ------------------------------------------------------
// The reference decides with autopas::utils::Math::roundFloating (src/autopas/utils/Math.cpp:37-45):
//   factor = pow(10, precision - ceil(log10(|d|)));  rounded = round(d * factor) / factor
// and isNearAbs(rounded, border, pow(10, -precision)). The multiplication, round() and the division are IEEE operations
// (identical on the device); pow and log10 are libm calls whose last bit is implementation defined, so the library
// tabulates them once per process from the host's libm (vtk.cu: apbVtkTables): pow10[k + APB_VTK_POW_OFF] = pow(10, k)
// and log10Limit[n + APB_VTK_LOG_OFF] = the largest double whose ceil(log10()) is <= n.
#define APB_VTK_POW_MIN (-345)
#define APB_VTK_POW_MAX 345
#define APB_VTK_POW_OFF 345
#define APB_VTK_LOG_MIN (-324)
#define APB_VTK_LOG_MAX 309
#define APB_VTK_LOG_OFF 324
struct ApbVtkTables {
  double pow10[APB_VTK_POW_MAX - APB_VTK_POW_MIN + 1];
  double log10Limit[APB_VTK_LOG_MAX - APB_VTK_LOG_MIN + 1];
};

APB_HD double apbMulRn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
APB_HD double apbDivRn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
APB_HD double apbSubRn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, -b);
#else
  return a - b;
#endif
}
APB_HD double apbRoundHalfAway(double v) {  // std::round
#ifdef __CUDA_ARCH__
  return round(v);
#else
  return __builtin_round(v);
#endif
}
APB_HD double apbAbs(double v) {
#ifdef __CUDA_ARCH__
  return fabs(v);
#else
  return __builtin_fabs(v);
#endif
}

// ceil(log10(|d|)) as the host's libm evaluates it (d finite, non-zero)
APB_HD int apbCeilLog10(const ApbVtkTables &t, double ad) {
  int lo = APB_VTK_LOG_MIN, hi = APB_VTK_LOG_MAX;  // smallest n with ad <= log10Limit[n]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;  // (floor for negative sums: lo + hi is shifted arithmetically)
    if (ad <= t.log10Limit[mid + APB_VTK_LOG_OFF]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// precision the position is written with; -1 where the reference throws (indistinguishable from the border at 15 digits)
APB_HD int apbVtkPositionPrecision(const ApbVtkTables &t, double position, double border) {
  int precision = 6;
  if (apbSubRn(border, position) < 0.1) {
    for (;;) {
      double rounded = position;
      if (position != 0.0) {
        const double factor = t.pow10[precision - apbCeilLog10(t, apbAbs(position)) + APB_VTK_POW_OFF];
        rounded = apbDivRn(apbRoundHalfAway(apbMulRn(position, factor)), factor);
      }
      if (!(apbAbs(apbSubRn(rounded, border)) <= t.pow10[-precision + APB_VTK_POW_OFF])) break;
      if (++precision > 15) return -1;
    }
  }
  return precision;
}

// ---- the reading direction: decimal text -> double, as `stream >> value` (strtod) converts it -------------------------
// md-flexible's loader reads the checkpoint with operator>> (MDFlexConfig.cpp:31-46 readPayload), i.e. the correctly
// rounded double of the decimal string. value = D * 10^E with the digits D collected exactly (up to 40 significant
// digits); E >= 0: the product is an integer; E < 0: floor(D 2^s / 10^-E) with s chosen so that the quotient keeps more
// than 64 bits, remainders of the divisions are sticky. The top 53 bits (fewer for subnormal results) are rounded to
// nearest / even on what lies below. status: 0 fine, 1 not a decimal number this function supports (inf / nan / hex,
// more than 40 significant digits, stray characters).
APB_HD void apbBigAddSmall(ApbBig &b, uint32_t c) {
  uint64_t carry = c;
  for (int i = 0; i < b.n && carry; ++i) {
    const uint64_t t = static_cast<uint64_t>(b.w[i]) + carry;
    b.w[i] = static_cast<uint32_t>(t);
    carry = t >> 32;
  }
  if (carry) b.w[b.n++] = static_cast<uint32_t>(carry);
}
APB_HD int apbBigBitLength(const ApbBig &b) {
  const uint32_t top = b.w[b.n - 1];
  if (top == 0) return 0;  // (only the value zero has a zero top word)
  return 32 * (b.n - 1) + 64 - apbCountLeadingZeros64(static_cast<uint64_t>(top));
}
APB_HD double apbBitsToDouble(uint64_t bits) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(static_cast<long long>(bits));
#else
  double v;
  __builtin_memcpy(&v, &bits, 8);
  return v;
#endif
}

APB_HD double apbParseDouble(const char *s, int len, int &status) {
  status = 0;
  int i = 0;
  uint64_t sign = 0;
  if (i < len && (s[i] == '+' || s[i] == '-')) {
    if (s[i] == '-') sign = 1ull << 63;
    ++i;
  }
  ApbBig big;
  apbBigSet(big, 0);
  int nd = 0;          // significant digits collected
  long long E = 0;     // decimal exponent of the collected integer
  bool any = false, seenPoint = false;
  for (; i < len; ++i) {
    const char ch = s[i];
    if (ch >= '0' && ch <= '9') {
      any = true;
      if (seenPoint) --E;
      if (nd == 0 && ch == '0') continue;  // leading zeros
      if (nd >= 40) {
        status = 1;
        return 0.;
      }
      apbBigMulSmall(big, 10u);
      apbBigAddSmall(big, static_cast<uint32_t>(ch - '0'));
      ++nd;
    } else if (ch == '.' && !seenPoint) {
      seenPoint = true;
    } else {
      break;
    }
  }
  if (!any) {
    status = 1;
    return 0.;
  }
  if (i < len && (s[i] == 'e' || s[i] == 'E')) {
    ++i;
    bool negExp = false;
    if (i < len && (s[i] == '+' || s[i] == '-')) {
      negExp = s[i] == '-';
      ++i;
    }
    long long ex = 0;
    bool anyExp = false;
    for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) {
      anyExp = true;
      if (ex < 100000) ex = ex * 10 + (s[i] - '0');
    }
    if (!anyExp) {
      status = 1;
      return 0.;
    }
    E += negExp ? -ex : ex;
  }
  if (i != len) {
    status = 1;
    return 0.;
  }
  if (nd == 0) return apbBitsToDouble(sign);  // +-0
  // 10^(nd - 1 + E) <= value < 10^(nd + E)
  if (nd - 1 + E >= 309) return apbBitsToDouble(sign | 0x7ff0000000000000ull);
  if (nd + E < -324) return apbBitsToDouble(sign);  // below half of the smallest subnormal (2.47e-324)
  int s2 = 0;
  bool sticky = false;
  if (E >= 0) {
    long long kk = E;
    for (; kk >= 9; kk -= 9) apbBigMulSmall(big, 1000000000u);
    if (kk) apbBigMulSmall(big, static_cast<uint32_t>(apbPow10u64(static_cast<int>(kk))));
  } else {
    long long k = -E;
    s2 = 66 + static_cast<int>((k * 3402) >> 10) + 1 - apbBigBitLength(big);  // 3402 / 1024 > log2(10)
    if (s2 < 0) s2 = 0;
    apbBigShl(big, s2);
    for (; k > 9; k -= 9) sticky |= apbBigDivSmall(big, 1000000000u) != 0;
    sticky |= apbBigDivSmall(big, static_cast<uint32_t>(apbPow10u64(static_cast<int>(k)))) != 0;
  }
  const int nb = apbBigBitLength(big);
  int e2 = nb - 1 - s2;  // value in [2^e2, 2^(e2 + 1))
  if (e2 > 1023) return apbBitsToDouble(sign | 0x7ff0000000000000ull);
  int keep = 53;
  if (e2 < -1022) keep = 53 - (-1022 - e2);
  if (keep < 0) return apbBitsToDouble(sign);
  const int shift = nb - keep;
  uint64_t mant;
  if (shift <= 0) {
    mant = apbBigShr64(big, 0) << (-shift);  // exact (nb < 53 happens only for integers)
  } else {
    mant = keep ? apbBigShr64(big, shift) : 0ull;
    if (keep < 64 && keep > 0) mant &= (1ull << keep) - 1ull;  // (apbBigShr64 returns 64 bits)
    const bool rbit = apbBigBit(big, shift - 1) != 0;
    const bool below = sticky || apbBigAnyBelow(big, shift - 1);
    if (rbit && (below || (mant & 1ull))) ++mant;
  }
  if (keep < 53) return apbBitsToDouble(sign | mant);  // subnormal (a carry into bit 52 is the smallest normal)
  if (mant == (1ull << 53)) {
    mant >>= 1;
    if (++e2 > 1023) return apbBitsToDouble(sign | 0x7ff0000000000000ull);
  }
  return apbBitsToDouble(sign | (static_cast<uint64_t>(e2 + 1023) << 52) | (mant & ((1ull << 52) - 1ull)));
}

// decimal digits -> unsigned long (`stream >> size_t`); status 1: not a number of at most 19 digits
APB_HD uint64_t apbParseU64(const char *s, int len, int &status) {
  status = (len <= 0 || len > 19) ? 1 : 0;
  uint64_t v = 0;
  for (int i = 0; i < len && !status; ++i) {
    if (s[i] < '0' || s[i] > '9') status = 1;
    v = v * 10u + static_cast<uint64_t>(s[i] - '0');
  }
  return v;
}

// Host side: the two tables from this process's libm (the one the reference's writer would call on this machine).
#include <cmath>
#include <cstdio>
#include <cstdlib>
inline void apbVtkBuildTables(ApbVtkTables &t) {
  for (int k = APB_VTK_POW_MIN; k <= APB_VTK_POW_MAX; ++k) t.pow10[k + APB_VTK_POW_OFF] = std::pow(10, static_cast<double>(k));
  for (int n = APB_VTK_LOG_MIN; n <= APB_VTK_LOG_MAX; ++n) {
    char text[16];
    std::snprintf(text, sizeof text, "1e%d", n);
    double c = std::strtod(text, nullptr);  // nearest double to 10^n (0 / inf beyond the range)
    const double inf = std::numeric_limits<double>::infinity();
    for (int step = 0; step < 64; ++step) {  // libm may return exactly n a few ulp above 10^n
      const double up = std::nextafter(c, inf);
      if (up == c || !(std::ceil(std::log10(up)) <= n)) break;
      c = up;
    }
    for (int step = 0; step < 64 && c > 0. && !(std::ceil(std::log10(c)) <= n); ++step) c = std::nextafter(c, 0.);
    t.log10Limit[n + APB_VTK_LOG_OFF] = c;
  }
}
