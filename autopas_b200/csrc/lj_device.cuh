// Device-side pieces shared by the LJ kernels: parameters, the pair kernel and the statistics reduction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct LJParams {
  double cutoff2;
  double eps24, sigma2, shift6;  // non-mixing
  const double *mix;             // [T*T][3] = {eps24, sigma2, shift6} (ParticlePropertiesLibrary.h:324-328)
  int T;
  int applyShift;
  // derived constants of the regrouped pair kernel (pruned.cu: prPair): K1 = 2 eps24 sigma^12, K2 = -eps24 sigma^6
  double k1, k2;
  const double *mix4;  // [T*T][4] = {K1, K2, K1 / 2, shift6}
  int cutHiLo;         // high word of the bit pattern of cutoff^2, minus 1 (band in which dr2 is re-evaluated exactly)
};

struct LJStats {
  double upot;
  double vir[3];
  unsigned long long dist, kN3, kNoN3, gN3, gNoN3;
};

__device__ __forceinline__ void ljStatsZero(LJStats &s) {
  s.upot = 0.;
  s.vir[0] = s.vir[1] = s.vir[2] = 0.;
  s.dist = s.kN3 = s.kNoN3 = s.gN3 = s.gNoN3 = 0ULL;
}
__device__ __forceinline__ void ljStatsAdd(LJStats &s, const LJStats &o) {
  s.upot += o.upot;
  s.vir[0] += o.vir[0];
  s.vir[1] += o.vir[1];
  s.vir[2] += o.vir[2];
  s.dist += o.dist;
  s.kN3 += o.kN3;
  s.kNoN3 += o.kNoN3;
  s.gN3 += o.gN3;
  s.gNoN3 += o.gNoN3;
}
__device__ __forceinline__ void ljStatsWarpReduce(LJStats &s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s.upot += __shfl_xor_sync(0xffffffffu, s.upot, o);
    s.vir[0] += __shfl_xor_sync(0xffffffffu, s.vir[0], o);
    s.vir[1] += __shfl_xor_sync(0xffffffffu, s.vir[1], o);
    s.vir[2] += __shfl_xor_sync(0xffffffffu, s.vir[2], o);
    s.dist += __shfl_xor_sync(0xffffffffu, s.dist, o);
    s.kN3 += __shfl_xor_sync(0xffffffffu, s.kN3, o);
    s.kNoN3 += __shfl_xor_sync(0xffffffffu, s.kNoN3, o);
    s.gN3 += __shfl_xor_sync(0xffffffffu, s.gN3, o);
    s.gNoN3 += __shfl_xor_sync(0xffffffffu, s.gNoN3, o);
  }
}
// all threads of the block must call this; writes partials[index] (default: blockIdx.x). Butterfly order is fixed, so the sums are
// reproducible run to run.
__device__ __forceinline__ void ljStatsBlockReduce(LJStats &s, LJStats *partials, int index = -1) {
  __shared__ LJStats sh[32];
  ljStatsWarpReduce(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (warp == 0) {
    LJStats t;
    ljStatsZero(t);
    if (lane < ((blockDim.x + 31) >> 5)) t = sh[lane];
    ljStatsWarpReduce(t);
    if (lane == 0) partials[index >= 0 ? index : static_cast<int>(blockIdx.x)] = t;
  }
}

// dr2 = drx*drx + dry*dry + drz*drz with every product and sum rounded separately (LJFunctor.h:484-488). The cutoff
// mask `dr2 <= cutoff^2` must agree bit for bit with the oracle, so FMA contraction is ruled out here.
__device__ __forceinline__ double ljDist2(double drx, double dry, double drz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(drx, drx), __dmul_rn(dry, dry)), __dmul_rn(drz, drz));
}

// fac = epsilon24 * (lj12 + lj12m6) * invdr2 (LJFunctor.h:152-158) and
// potentialEnergy6 = epsilon24 * lj12m6 + shift6 (:174). upot6 is dead code when the caller does not use it.
template <bool MIX>
__device__ __forceinline__ double ljEval(const LJParams &p, double dr2, int ti, int tj, double &upot6) {
  double e24, s2, shift6;
  if (MIX) {
    const double *m = p.mix + 3 * (static_cast<size_t>(ti) * p.T + tj);
    e24 = __ldg(m);
    s2 = __ldg(m + 1);
    shift6 = p.applyShift ? __ldg(m + 2) : 0.;
  } else {
    e24 = p.eps24;
    s2 = p.sigma2;
    shift6 = p.shift6;
  }
  const double inv = 1. / dr2;
  double lj6 = s2 * inv;
  lj6 = lj6 * lj6 * lj6;
  const double lj12 = lj6 * lj6;
  const double lj12m6 = lj12 - lj6;
  upot6 = e24 * lj12m6 + shift6;
  return e24 * (lj12 + lj12m6) * inv;
}
