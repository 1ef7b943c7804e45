/* TEST INFRASTRUCTURE ONLY — never linked into or called from the product path (see oracle/Makefile).
 *
 * Plain-C restatement of the reference functors other than single-site LJ, as brute-force loops over particle pairs /
 * triplets. Expression order follows the reference's AoS functors; built with -ffp-contract=off.
 *   SPH kernels        applicationLibrary/sph/SPHLibrary/SPHKernels.h:37-87
 *   SPH density        applicationLibrary/sph/SPHLibrary/SPHCalcDensityFunctor.h:43-63
 *   SPH hydro force    applicationLibrary/sph/SPHLibrary/SPHCalcHydroForceFunctor.h:45-108
 *   Axilrod-Teller     applicationLibrary/molecularDynamics/molecularDynamicsLibrary/AxilrodTellerMutoFunctor.h:186-293
 *   LJ multi-site      applicationLibrary/molecularDynamics/molecularDynamicsLibrary/LJMultisiteFunctor.h:181-274,
 *                      src/autopas/utils/Quaternion.cpp:13-46
 * Pinned by tests/test_oracle_functors.py against the reference's own literals (SPHTest.cpp:16-50, ATMPotential.h) and
 * against fixtures produced by the unmodified reference (tests/golden/ npz files).
 * Every routine also returns, per particle, the sum of the magnitudes of the contributions ("scale"): the error norm
 * of the parity tests is |gpu - ref| <= 1e-12 * scale. */
#define _USE_MATH_DEFINES
#define _DEFAULT_SOURCE
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const double kSupport = 2.5;

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* SPHKernels::W(dr2, h) (:37-53) */
double orc_sph_W(double dr2, double h) {
  const double H = kSupport * h;
  if (dr2 < H * H) {
    const double s = sqrt(dr2) / H;
    const double s1 = 1.0 - s;
    const double s2 = fmax(0., 0.5 - s);
    double r = (s1 * s1 * s1) - 4.0 * (s2 * s2 * s2);
    r *= 16.0 / M_PI / (H * H * H);
    return r;
  }
  return 0.;
}

/* SPHKernels::gradW (:72-86) */
void orc_sph_gradW(const double *dr, double h, double *out) {
  const double H = kSupport * h;
  const double drabs = sqrt(dot3(dr, dr));
  const double s = drabs / H;
  const double s1 = (1.0 - s < 0) ? 0 : 1.0 - s;
  const double s2 = (0.5 - s < 0) ? 0 : 0.5 - s;
  double r = -3.0 * (s1 * s1) + 12.0 * (s2 * s2);
  r *= 16.0 / M_PI / (H * H * H);
  const double scale = r / (drabs * H + 1.0e-6 * h);
  for (int d = 0; d < 3; ++d) out[d] = dr[d] * scale;
}

/* SPHCalcDensityFunctor::AoSFunctor applied to every ordered pair (newton3 off semantics; newton3 on adds the same
 * terms). Only owned particles receive a result that matters; halos act as partners. */
void orc_sph_density(int64_t n, const double *x, const double *y, const double *z, const double *mass,
                     const double *smth, const int64_t *own, double *density, double *scale) {
  for (int64_t i = 0; i < n; ++i) {
    double rho = 0., sc = 0.;
    if (own[i] != 0) {
      for (int64_t j = 0; j < n; ++j) {
        if (j == i || own[j] == 0) continue;
        if (own[i] == 2 && own[j] == 2) continue; /* halo-halo pairs are never evaluated */
        const double dr[3] = {x[j] - x[i], y[j] - y[i], z[j] - z[i]};
        const double c = mass[j] * orc_sph_W(dot3(dr, dr), smth[i]);
        rho += c;
        sc += fabs(c);
      }
    }
    density[i] = rho;
    scale[i] = sc;
  }
}

/* SPHCalcHydroForceFunctor::AoSFunctor(i, j, newton3 = false) for every ordered pair */
void orc_sph_hydro(int64_t n, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                   const double *vz, const double *mass, const double *smth, const double *density,
                   const double *pressure, const double *snd, const int64_t *own, const double *vsigmaxIn, double *acc,
                   double *engDot, double *vsigmax, double *scale, double *scaleEng) {
  /* scale / scaleEng: sums of the magnitudes of the per-pair terms of the acceleration / of engDot (the rounding error
   * of a sum scales with those, not with the possibly cancelling result); scaleEng may be NULL */
  for (int64_t i = 0; i < n; ++i) {
    double a[3] = {0., 0., 0.}, e = 0., vm = vsigmaxIn ? vsigmaxIn[i] : 0., sc = 0., sce = 0.;
    if (own[i] != 0) {
      for (int64_t j = 0; j < n; ++j) {
        if (j == i || own[j] == 0) continue;
        if (own[i] == 2 && own[j] == 2) continue;
        const double dr[3] = {x[i] - x[j], y[i] - y[j], z[i] - z[j]};
        const double cutoff = smth[i] * kSupport;
        if (dot3(dr, dr) >= cutoff * cutoff) continue;
        const double dv[3] = {vx[i] - vx[j], vy[i] - vy[j], vz[i] - vz[j]};
        const double dvdr = dot3(dv, dr);
        const double wij = (dvdr < 0) ? dvdr / sqrt(dot3(dr, dr)) : 0;
        const double vsig = snd[i] + snd[j] - 3.0 * wij;
        if (vsig > vm) vm = vsig;
        const double AV = -0.5 * vsig * wij / (0.5 * (density[i] + density[j]));
        double gi[3], gj[3], g[3];
        orc_sph_gradW(dr, smth[i], gi);
        orc_sph_gradW(dr, smth[j], gj);
        for (int d = 0; d < 3; ++d) g[d] = (gi[d] + gj[d]) * 0.5;
        const double sc1 = pressure[i] / (density[i] * density[i]) + pressure[j] / (density[j] * density[j]) + AV;
        for (int d = 0; d < 3; ++d) {
          const double c = g[d] * (sc1 * mass[j]);
          a[d] -= c;
          sc += fabs(c);
        }
        const double scale2i = mass[j] * (pressure[i] / (density[i] * density[i]) + 0.5 * AV);
        e += dot3(g, dv) * scale2i;
        sce += (fabs(g[0] * dv[0]) + fabs(g[1] * dv[1]) + fabs(g[2] * dv[2])) * fabs(scale2i);
      }
    }
    acc[3 * i] = a[0];
    acc[3 * i + 1] = a[1];
    acc[3 * i + 2] = a[2];
    engDot[i] = e;
    vsigmax[i] = vm;
    scale[i] = sc;
    if (scaleEng) scaleEng[i] = sce;
  }
}

/* AxilrodTellerMutoFunctor::AoSFunctor(i, j, k, newton3 = false): force on i only. res: {sum of 3*Upot over owned i,
 * virial xyz over owned i, kernel calls with an owned first particle}. nuMix: [T*T*T] or NULL. */
void orc_atm(int64_t n, const double *x, const double *y, const double *z, const int64_t *type, const int64_t *own,
             double cutoff, double nuIn, int T, const double *nuMix, double *f, double *scale, double *res) {
  const double c2 = cutoff * cutoff;
  memset(res, 0, 5 * sizeof(double));
  int64_t *nb = malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    double F[3] = {0., 0., 0.}, sc = 0.;
    if (own[i] != 0) {
      int64_t cnt = 0;
      for (int64_t j = 0; j < n; ++j) {
        if (j == i || own[j] == 0) continue;
        const double d[3] = {x[j] - x[i], y[j] - y[i], z[j] - z[i]};
        if (dot3(d, d) <= c2) nb[cnt++] = j;
      }
      for (int64_t p = 0; p < cnt; ++p)
        for (int64_t q = p + 1; q < cnt; ++q) {
          const int64_t j = nb[p], k = nb[q];
          const double IJ[3] = {x[j] - x[i], y[j] - y[i], z[j] - z[i]};
          const double JK[3] = {x[k] - x[j], y[k] - y[j], z[k] - z[j]};
          const double KI[3] = {x[i] - x[k], y[i] - y[k], z[i] - z[k]};
          const double d2ij = dot3(IJ, IJ), d2jk = dot3(JK, JK), d2ki = dot3(KI, KI);
          if (d2ij > c2 || d2jk > c2 || d2ki > c2) continue;
          double nu = nuIn;
          if (nuMix) nu = nuMix[((size_t)type[i] * T + type[j]) * T + type[k]];
          const double all2 = d2ij * d2jk * d2ki;
          const double all5 = all2 * all2 * sqrt(all2);
          const double factor = 3.0 * nu / all5;
          const double IJdKI = dot3(IJ, KI), IJdJK = dot3(IJ, JK), JKdKI = dot3(JK, KI);
          const double allDots = IJdKI * IJdJK * JKdKI;
          double fi[3];
          for (int d = 0; d < 3; ++d) {
            const double dJK = JK[d] * IJdKI * (IJdJK - JKdKI);
            const double dIJ = IJ[d] * (IJdJK * JKdKI - d2jk * d2ki + 5.0 * allDots / d2ij);
            const double dKI = KI[d] * (-IJdJK * JKdKI + d2ij * d2jk - 5.0 * allDots / d2ki);
            fi[d] = (dJK + dIJ + dKI) * factor;
            F[d] += fi[d];
            sc += fabs(fi[d]);
          }
          if (own[i] == 1) {
            res[4] += 1.; /* lc_c01 newton3-off calls the functor for owned base particles only */
            res[0] += factor * (all2 - 3.0 * allDots);
            res[1] += fi[0] * x[i];
            res[2] += fi[1] * y[i];
            res[3] += fi[2] * z[i];
          }
        }
    }
    f[3 * i] = F[0];
    f[3 * i + 1] = F[1];
    f[3 * i + 2] = F[2];
    scale[i] = sc;
  }
  free(nb);
}

/* quaternion::rotateVectorOfPositions (Quaternion.cpp:13-46) for one position */
void orc_rotate(const double *q, const double *p, double *out) {
  const double ww = q[0] * q[0], wx = q[0] * q[1], wy = q[0] * q[2], wz = q[0] * q[3];
  const double xx = q[1] * q[1], xy = q[1] * q[2], xz = q[1] * q[3], yy = q[2] * q[2], yz = q[2] * q[3], zz = q[3] * q[3];
  const double r00 = ww + xx - yy - zz, r01 = 2. * (xy - wz), r02 = 2. * (xz + wy);
  const double r10 = 2. * (xy + wz), r11 = ww - xx + yy - zz, r12 = 2. * (yz - wx);
  const double r20 = 2. * (xz - wy), r21 = 2. * (yz + wx), r22 = ww - xx - yy + zz;
  out[0] = r00 * p[0] + r01 * p[1] + r02 * p[2];
  out[1] = r10 * p[0] + r11 * p[1] + r12 * p[2];
  out[2] = r20 * p[0] + r21 * p[1] + r22 * p[2];
}

/* LJMultisiteFunctor::AoSFunctor(A, B, newton3 = false) for every ordered pair. mix: [T*T][3] = {eps24, sigma2, shift6}.
 * res: {sum of Upot6 * [A owned], virial xyz}. */
void orc_multisite(int64_t n, const double *x, const double *y, const double *z, const double *q, const int64_t *molType,
                   const int64_t *own, double cutoff, int applyShift, int T, const double *mix, const int32_t *siteStart,
                   const double *sitePos, const int32_t *siteType, double *f, double *torque, double *scale,
                   double *res) {
  const double c2 = cutoff * cutoff;
  memset(res, 0, 4 * sizeof(double));
  for (int64_t a = 0; a < n; ++a) {
    double F[3] = {0., 0., 0.}, Tq[3] = {0., 0., 0.}, sc = 0.;
    if (own[a] != 0) {
      const int sA0 = siteStart[molType[a]], nA = siteStart[molType[a] + 1] - sA0;
      for (int64_t b = 0; b < n; ++b) {
        if (b == a || own[b] == 0) continue;
        if (own[a] == 2 && own[b] == 2) continue;
        const double dc[3] = {x[a] - x[b], y[a] - y[b], z[a] - z[b]};
        if (dot3(dc, dc) > c2) continue;
        const int sB0 = siteStart[molType[b]], nB = siteStart[molType[b] + 1] - sB0;
        for (int i = 0; i < nA; ++i) {
          double rA[3];
          orc_rotate(q + 4 * a, sitePos + 3 * (sA0 + i), rA);
          for (int j = 0; j < nB; ++j) {
            double rB[3];
            orc_rotate(q + 4 * b, sitePos + 3 * (sB0 + j), rB);
            double dr[3];
            for (int d = 0; d < 3; ++d) dr[d] = (dc[d] - rB[d]) + rA[d];
            const double dr2 = dot3(dr, dr);
            const double *m = mix + 3 * ((size_t)siteType[sA0 + i] * T + siteType[sB0 + j]);
            const double e24 = m[0], s2 = m[1], shift6 = applyShift ? m[2] : 0.;
            const double inv = 1. / dr2;
            const double lj2 = s2 * inv;
            const double lj6 = lj2 * lj2 * lj2;
            const double lj12 = lj6 * lj6;
            const double lj12m6 = lj12 - lj6;
            const double fac = e24 * (lj12 + lj12m6) * inv;
            const double fo[3] = {dr[0] * fac, dr[1] * fac, dr[2] * fac};
            for (int d = 0; d < 3; ++d) {
              F[d] += fo[d];
              sc += fabs(fo[d]);
            }
            Tq[0] += rA[1] * fo[2] - rA[2] * fo[1];
            Tq[1] += rA[2] * fo[0] - rA[0] * fo[2];
            Tq[2] += rA[0] * fo[1] - rA[1] * fo[0];
            if (own[a] == 1) {
              res[0] += e24 * lj12m6 + shift6;
              res[1] += dr[0] * fo[0];
              res[2] += dr[1] * fo[1];
              res[3] += dr[2] * fo[2];
            }
          }
        }
      }
    }
    for (int d = 0; d < 3; ++d) {
      f[3 * a + d] = F[d];
      torque[3 * a + d] = Tq[d];
    }
    scale[a] = sc;
  }
}
