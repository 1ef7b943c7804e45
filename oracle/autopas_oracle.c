/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (plain C) of the AutoPas reference algorithms on the short-range
 * interaction hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * (autopas_b200/) never does.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against the reference's own golden vectors
 * (LJFunctorTestNoGlobals.h:27-31, testingHelpers/LJPotential.h, CellBlock3DTest.cpp, the c08 offset tables of
 * LCC08CellHandlerUtilityTest.cpp, LJFunctorFlopCounterTest.cpp, VerletClusterListsTest.cpp: brute-force equivalence,
 * grid alignment, newton3 list relation) and, when oracle/_ref/libautopas_ref.so exists, against the unmodified reference compiled from
 * /root/reference (oracle/ref_driver.cpp).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference tree).
 * Compile with -ffp-contract=off: decisions (`<=`) must not depend on FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OWN_DUMMY 0
#define OWN_OWNED 1
#define OWN_HALO 2

#define F_SHIFT 1
#define F_MIXING 2
#define F_NEWTON3 4

typedef struct {
  double upot_sum, virial_sum[3];
  uint64_t num_dist_calls, num_kernel_calls_n3, num_kernel_calls_no_n3, num_global_calcs_n3, num_global_calcs_no_n3;
} orc_result;

typedef struct {
  double cutoff2, eps24, sigma2, shift6;
  const double *mix; /* [T*T][3] */
  int T, apply_shift, mixing;
} lj_params;

/* ParticlePropertiesLibrary::calcShift6 (ParticlePropertiesLibrary.h:576-582) */
double orc_calc_shift6(double eps24, double sigma2, double cutoff2) {
  const double s2 = sigma2 / cutoff2;
  const double s6 = s2 * s2 * s2;
  return eps24 * (s6 - s6 * s6);
}

/* ParticlePropertiesLibrary::calculateMixingCoefficients (ParticlePropertiesLibrary.h:444-473) */
void orc_mixing_table(int T, const double *eps, const double *sig, double cutoff, double *out) {
  const double rc2 = cutoff * cutoff;
  for (int i = 0; i < T; ++i)
    for (int j = 0; j < T; ++j) {
      const double e24 = 24 * sqrt(eps[i] * eps[j]);
      const double s = (sig[i] + sig[j]) / 2.0;
      const double s2 = s * s;
      double *o = out + 3 * ((size_t)i * T + j);
      o[0] = e24;
      o[1] = s2;
      o[2] = orc_calc_shift6(e24, s2, rc2);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * particle view used by the functor restatements
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  double *fscale; /* sum over pairs of |fx|+|fy|+|fz| per particle: the scale of the 1e-12 tolerance */
  const int64_t *type, *own;
} pview;

static void lj_coeffs(const lj_params *p, int64_t ti, int64_t tj, double *e24, double *s2, double *sh6) {
  if (p->mixing) {
    const double *m = p->mix + 3 * ((size_t)ti * p->T + tj);
    *e24 = m[0];
    *s2 = m[1];
    *sh6 = p->apply_shift ? m[2] : 0.;
  } else {
    *e24 = p->eps24;
    *s2 = p->sigma2;
    *sh6 = p->shift6;
  }
}

/* LJFunctor::SoAFunctorSingle (LJFunctor.h:204-364): pairs i<j of one index set, ALWAYS newton3 (:200-203, :345) */
static void soa_single(const lj_params *p, pview *v, const int *idx, int n, orc_result *r) {
  for (int a = 0; a < n; ++a) {
    const int i = idx[a];
    if (v->own[i] == OWN_DUMMY) continue;
    for (int b = a + 1; b < n; ++b) {
      const int j = idx[b];
      const int64_t ownJ = v->own[j];
      const double drx = v->x[i] - v->x[j], dry = v->y[i] - v->y[j], drz = v->z[i] - v->z[j];
      const double dr2 = drx * drx + dry * dry + drz * drz;
      const int mask = dr2 <= p->cutoff2 && ownJ != OWN_DUMMY;
      if (ownJ != OWN_DUMMY) r->num_dist_calls++;
      if (!mask) continue;
      double e24, s2, sh6;
      lj_coeffs(p, v->type[i], v->type[j], &e24, &s2, &sh6);
      const double inv = 1. / dr2;
      const double lj2 = s2 * inv;
      const double lj6 = lj2 * lj2 * lj2;
      const double lj12 = lj6 * lj6;
      const double lj12m6 = lj12 - lj6;
      const double fac = e24 * (lj12 + lj12m6) * inv;
      const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
      v->fx[i] += fx;
      v->fy[i] += fy;
      v->fz[i] += fz;
      v->fx[j] -= fx;
      v->fy[j] -= fy;
      v->fz[j] -= fz;
      const double mag = fabs(fx) + fabs(fy) + fabs(fz);
      v->fscale[i] += mag;
      v->fscale[j] += mag;
      r->num_kernel_calls_n3++;
      const double w = (v->own[i] == OWN_OWNED ? 1. : 0.) + (ownJ == OWN_OWNED ? 1. : 0.);
      r->upot_sum += (e24 * lj12m6 + sh6) * w;
      r->virial_sum[0] += drx * fx * w;
      r->virial_sum[1] += dry * fy * w;
      r->virial_sum[2] += drz * fz * w;
      r->num_global_calcs_n3++;
    }
  }
}

/* LJFunctor::SoAFunctorPairImpl<newton3> (LJFunctor.h:387-559) */
static void soa_pair(const lj_params *p, pview *v, const int *idx1, int n1, const int *idx2, int n2, int n3,
                     orc_result *r) {
  for (int a = 0; a < n1; ++a) {
    const int i = idx1[a];
    if (v->own[i] == OWN_DUMMY) continue;
    for (int b = 0; b < n2; ++b) {
      const int j = idx2[b];
      const int64_t ownJ = v->own[j];
      const double drx = v->x[i] - v->x[j], dry = v->y[i] - v->y[j], drz = v->z[i] - v->z[j];
      const double dr2 = drx * drx + dry * dry + drz * drz;
      const int mask = dr2 <= p->cutoff2 && ownJ != OWN_DUMMY;
      if (ownJ != OWN_DUMMY) r->num_dist_calls++;
      if (!mask) continue;
      double e24, s2, sh6;
      lj_coeffs(p, v->type[i], v->type[j], &e24, &s2, &sh6);
      const double inv = 1. / dr2;
      const double lj2 = s2 * inv;
      const double lj6 = lj2 * lj2 * lj2;
      const double lj12 = lj6 * lj6;
      const double lj12m6 = lj12 - lj6;
      const double fac = e24 * (lj12 + lj12m6) * inv;
      const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
      v->fx[i] += fx;
      v->fy[i] += fy;
      v->fz[i] += fz;
      const double mag = fabs(fx) + fabs(fy) + fabs(fz);
      v->fscale[i] += mag;
      if (n3) {
        v->fx[j] -= fx;
        v->fy[j] -= fy;
        v->fz[j] -= fz;
        v->fscale[j] += mag;
        r->num_kernel_calls_n3++;
      } else {
        r->num_kernel_calls_no_n3++;
      }
      const double w = (v->own[i] == OWN_OWNED ? 1. : 0.) + (n3 ? (ownJ == OWN_OWNED ? 1. : 0.) : 0.);
      r->upot_sum += (e24 * lj12m6 + sh6) * w;
      r->virial_sum[0] += drx * fx * w;
      r->virial_sum[1] += dry * fy * w;
      r->virial_sum[2] += drz * fz * w;
      if (n3)
        r->num_global_calcs_n3++;
      else
        r->num_global_calcs_no_n3++;
    }
  }
}

static void make_params(lj_params *p, double cutoff, int flags, int T, const double *eps, const double *sig,
                        double *mixbuf) {
  p->cutoff2 = cutoff * cutoff;
  p->apply_shift = (flags & F_SHIFT) != 0;
  p->mixing = (flags & F_MIXING) != 0;
  p->T = T;
  p->mix = mixbuf;
  if (p->mixing) {
    orc_mixing_table(T, eps, sig, cutoff, mixbuf);
  } else {
    /* LJFunctor::setParticleProperties(epsilon24, sigmaSquared) (LJFunctor.h:588-596) */
    p->eps24 = 24. * eps[0];
    p->sigma2 = sig[0] * sig[0];
    p->shift6 = p->apply_shift ? orc_calc_shift6(p->eps24, p->sigma2, p->cutoff2) : 0.;
  }
}

/* LJFunctor::endTraversal + getPotentialEnergy / getVirial (LJFunctor.h:661-720) */
void orc_lj_end_traversal(const orc_result *r, double *upot, double *virial) {
  double u = r->upot_sum;
  u *= 0.5;
  u /= 6.;
  *upot = u;
  *virial = r->virial_sum[0] * 0.5 + r->virial_sum[1] * 0.5 + r->virial_sum[2] * 0.5;
}

/* LJFunctor::getNumFLOPs (LJFunctor.h:776-789) */
uint64_t orc_lj_num_flops(const orc_result *r, int apply_shift) {
  const uint64_t gN3 = apply_shift ? 13 : 12, gNoN3 = apply_shift ? 9 : 8;
  return r->num_dist_calls * 8 + r->num_kernel_calls_n3 * 18 + r->num_kernel_calls_no_n3 * 15 +
         r->num_global_calcs_n3 * gN3 + r->num_global_calcs_no_n3 * gNoN3;
}

/* ------------------------------------------------------------------------------------------------------------------
 * LinkedCells
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  double box_min[3], box_max[3], halo_min[3], halo_max[3], cell_length[3], recip[3];
  int64_t cpd[3]; /* incl. halo */
  int64_t cpil, num_cells;
} lc_geom;

/* CellBlock3D::rebuild (containers/CellBlock3D.h:360-426) */
void orc_lc_geometry(const double *box_min, const double *box_max, double il, double csf, lc_geom *g) {
  g->cpil = csf >= 1.0 ? 1 : (int64_t)ceil(1.0 / csf);
  g->num_cells = 1;
  for (int d = 0; d < 3; ++d) {
    g->box_min[d] = box_min[d];
    g->box_max[d] = box_max[d];
    const double box_length = box_max[d] - box_min[d];
    uint64_t cells = (uint64_t)floor(box_length / (il * csf));
    if (cells < 1) cells = 1;
    g->cpd[d] = (int64_t)cells + 2 * g->cpil;
    g->cell_length[d] = box_length / (double)cells;
    g->recip[d] = (double)cells / box_length;
    g->halo_min[d] = box_min[d] - g->cpil * g->cell_length[d];
    g->halo_max[d] = box_max[d] + g->cpil * g->cell_length[d];
    g->num_cells *= g->cpd[d];
  }
}

/* CellBlock3D::get3DIndexOfPosition (:321-349) + threeToOneD (utils/ThreeDimensionalMapping.h:29-32) */
int64_t orc_lc_cell_index(const lc_geom *g, double px, double py, double pz) {
  const double pos[3] = {px, py, pz};
  int64_t idx[3];
  for (int d = 0; d < 3; ++d) {
    const long value = (long)floor((pos[d] - g->box_min[d]) * g->recip[d]) + g->cpil;
    int64_t v = value > 0 ? value : 0;
    if (v > g->cpd[d] - 1) v = g->cpd[d] - 1;
    if (pos[d] >= g->box_max[d]) {
      if (v < g->cpd[d] - g->cpil) v = g->cpd[d] - g->cpil;
    } else if (pos[d] < g->box_min[d] && v == g->cpil) {
      --v;
    } else if (pos[d] < g->box_max[d] && v == g->cpd[d] - g->cpil) {
      --v;
    }
    idx[d] = v;
  }
  return (idx[2] * g->cpd[1] + idx[1]) * g->cpd[0] + idx[0];
}

void orc_lc_cell_indices(const double *box_min, const double *box_max, double il, double csf, int64_t n,
                         const double *x, const double *y, const double *z, int64_t *out_cell, int64_t *out_cpd) {
  lc_geom g;
  orc_lc_geometry(box_min, box_max, il, csf, &g);
  for (int64_t i = 0; i < n; ++i) out_cell[i] = orc_lc_cell_index(&g, x[i], y[i], z[i]);
  for (int d = 0; d < 3; ++d) out_cpd[d] = g.cpd[d];
}

static int cell_can_own(const lc_geom *g, int64_t cx, int64_t cy, int64_t cz) {
  /* CellBlock3D::cellCanContainOwnedParticles: inside the non-halo block */
  return cx >= g->cpil && cx < g->cpd[0] - g->cpil && cy >= g->cpil && cy < g->cpd[1] - g->cpil && cz >= g->cpil &&
         cz < g->cpd[2] - g->cpil;
}

/*
 * The set of relative cell offsets the LinkedCells traversals pair a base cell with, as linear offset differences
 * (self = 0 included): every offset within the overlap whose cell-border distance is <= the interaction length
 * (LCC08CellHandlerUtility.cpp:67-159, filter :124), each unordered pair once. This is what orc_lj_linkedcells walks and
 * what the reference's c08 base step covers when its offset pairs are flattened to differences
 * (LCC08CellHandlerUtilityTest.cpp:27-233 holds the expected tables). Returns the count, fills out_lin (ascending).
 */
int64_t orc_lc_pair_offsets(const int64_t *cpd, const double *cell_length, double il, int64_t *out_lin, int64_t max_out) {
  int ov[3];
  for (int d = 0; d < 3; ++d) ov[d] = (int)ceil(il / cell_length[d]);
  const double il2 = il * il;
  int64_t count = 0;
  for (int oz = -ov[2]; oz <= ov[2]; ++oz)
    for (int oy = -ov[1]; oy <= ov[1]; ++oy)
      for (int ox = -ov[0]; ox <= ov[0]; ++ox) {
        const int64_t lin = ((int64_t)oz * cpd[1] + oy) * cpd[0] + ox;
        if (lin < 0) continue;
        const double dv[3] = {(abs(ox) > 1 ? abs(ox) - 1 : 0) * cell_length[0], (abs(oy) > 1 ? abs(oy) - 1 : 0) * cell_length[1],
                              (abs(oz) > 1 ? abs(oz) - 1 : 0) * cell_length[2]};
        if (!(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2] <= il2)) continue;
        if (count < max_out) out_lin[count] = lin;
        ++count;
      }
  /* ascending */
  for (int64_t i = 1; i < count && i < max_out; ++i) {
    const int64_t v = out_lin[i];
    int64_t j = i - 1;
    while (j >= 0 && out_lin[j] > v) {
      out_lin[j + 1] = out_lin[j];
      --j;
    }
    out_lin[j + 1] = v;
  }
  return count;
}

/*
 * LinkedCells + lc_c08 (== lc_c18 pair set) + LJFunctor SoA:
 * every cell: CellFunctor::processCell (baseFunctors/CellFunctor.h:141-160) -> SoAFunctorSingle;
 * every cell pair within the overlap whose border distance <= interaction length
 * (LCC08CellHandlerUtility.cpp:67-159, filter :124): CellFunctor::processCellPair (:162-191, :266-274) ->
 * SoAFunctorPair(c1,c2,n3) and, without newton3, also (c2,c1,false). Pairs of two halo-only cells are skipped (:173-184).
 */
int orc_lj_linkedcells(int64_t n, const double *x, const double *y, const double *z, const int64_t *type,
                       const int64_t *own, const double *box_min, const double *box_max, double cutoff, double skin,
                       double csf, int flags, int T, const double *eps, const double *sig, double *f /* 3n */,
                       double *fscale /* n */, orc_result *res, int64_t *out_cell /* n, may be NULL */) {
  lc_geom g;
  const double il = cutoff + skin;
  orc_lc_geometry(box_min, box_max, il, csf, &g);
  lj_params p;
  double *mixbuf = (double *)malloc(sizeof(double) * 3 * (size_t)(T > 0 ? T * T : 1));
  make_params(&p, cutoff, flags, T, eps, sig, mixbuf);
  const int n3 = (flags & F_NEWTON3) != 0;
  int64_t *zeros = NULL;
  if (!type) {
    zeros = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    type = zeros;
  }
  /* bin (stable in input order, like repeated addParticle) */
  int *cell = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *start = (int *)calloc((size_t)g.num_cells + 1, sizeof(int));
  for (int64_t i = 0; i < n; ++i) {
    if (own[i] == OWN_DUMMY) {
      cell[i] = -1;
      continue;
    }
    cell[i] = (int)orc_lc_cell_index(&g, x[i], y[i], z[i]);
    start[cell[i] + 1]++;
    if (out_cell) out_cell[i] = cell[i];
  }
  for (int64_t c = 0; c < g.num_cells; ++c) start[c + 1] += start[c];
  int *fill = (int *)malloc(sizeof(int) * ((size_t)g.num_cells + 1));
  memcpy(fill, start, sizeof(int) * ((size_t)g.num_cells + 1));
  int *order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i)
    if (cell[i] >= 0) order[fill[cell[i]]++] = (int)i;

  double *fx = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *fy = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *fz = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  memset(fscale, 0, sizeof(double) * (size_t)n);
  memset(res, 0, sizeof(*res));
  pview v = {x, y, z, fx, fy, fz, fscale, type, own};

  int ov[3];
  for (int d = 0; d < 3; ++d) ov[d] = (int)ceil(il / g.cell_length[d]);
  const double il2 = il * il;
  for (int64_t cz = 0; cz < g.cpd[2]; ++cz)
    for (int64_t cy = 0; cy < g.cpd[1]; ++cy)
      for (int64_t cx = 0; cx < g.cpd[0]; ++cx) {
        const int64_t c1 = (cz * g.cpd[1] + cy) * g.cpd[0] + cx;
        const int n1 = start[c1 + 1] - start[c1];
        if (n1 == 0) continue;
        const int own1 = cell_can_own(&g, cx, cy, cz);
        if (own1) soa_single(&p, &v, order + start[c1], n1, res);
        for (int oz = -ov[2]; oz <= ov[2]; ++oz)
          for (int oy = -ov[1]; oy <= ov[1]; ++oy)
            for (int ox = -ov[0]; ox <= ov[0]; ++ox) {
              const int64_t lin = ((int64_t)oz * g.cpd[1] + oy) * g.cpd[0] + ox;
              if (lin <= 0) continue; /* each unordered cell pair once */
              const int64_t nx = cx + ox, ny = cy + oy, nz = cz + oz;
              if (nx < 0 || ny < 0 || nz < 0 || nx >= g.cpd[0] || ny >= g.cpd[1] || nz >= g.cpd[2]) continue;
              const double dv[3] = {(abs(ox) > 1 ? abs(ox) - 1 : 0) * g.cell_length[0],
                                    (abs(oy) > 1 ? abs(oy) - 1 : 0) * g.cell_length[1],
                                    (abs(oz) > 1 ? abs(oz) - 1 : 0) * g.cell_length[2]};
              if (!(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2] <= il2)) continue;
              const int64_t c2 = c1 + lin;
              const int n2 = start[c2 + 1] - start[c2];
              if (n2 == 0) continue;
              const int own2 = cell_can_own(&g, nx, ny, nz);
              if (!own1 && !own2) continue;
              soa_pair(&p, &v, order + start[c1], n1, order + start[c2], n2, n3, res);
              if (!n3) soa_pair(&p, &v, order + start[c2], n2, order + start[c1], n1, 0, res);
            }
      }
  for (int64_t i = 0; i < n; ++i) {
    f[3 * i] = fx[i];
    f[3 * i + 1] = fy[i];
    f[3 * i + 2] = fz[i];
  }
  free(fx);
  free(fy);
  free(fz);
  free(cell);
  free(start);
  free(fill);
  free(order);
  free(mixbuf);
  free(zeros);
  return 0;
}

/* Brute force over all pairs with the canonical AoS kernel (LJFunctor::AoSFunctor, LJFunctor.h:123-198), newton3.
 * Used as an independent check of both traversals (like VerletClusterListsTest.cpp:128-198 does with N^2). */
int orc_lj_bruteforce(int64_t n, const double *x, const double *y, const double *z, const int64_t *type,
                      const int64_t *own, double cutoff, int flags, int T, const double *eps, const double *sig,
                      double *f, double *fscale, orc_result *res) {
  lj_params p;
  double *mixbuf = (double *)malloc(sizeof(double) * 3 * (size_t)(T > 0 ? T * T : 1));
  make_params(&p, cutoff, flags, T, eps, sig, mixbuf);
  memset(f, 0, sizeof(double) * 3 * (size_t)n);
  memset(fscale, 0, sizeof(double) * (size_t)n);
  memset(res, 0, sizeof(*res));
  for (int64_t i = 0; i < n; ++i) {
    if (own[i] == OWN_DUMMY) continue;
    for (int64_t j = i + 1; j < n; ++j) {
      if (own[j] == OWN_DUMMY) continue;
      if (own[i] == OWN_HALO && own[j] == OWN_HALO) continue; /* halo-halo never interacts in any container */
      res->num_dist_calls++;
      const double drx = x[i] - x[j], dry = y[i] - y[j], drz = z[i] - z[j];
      const double dr2 = drx * drx + dry * dry + drz * drz;
      if (dr2 > p.cutoff2) continue;
      double e24, s2, sh6;
      lj_coeffs(&p, type ? type[i] : 0, type ? type[j] : 0, &e24, &s2, &sh6);
      const double inv = 1. / dr2;
      double lj6 = s2 * inv;
      lj6 = lj6 * lj6 * lj6;
      const double lj12 = lj6 * lj6;
      const double lj12m6 = lj12 - lj6;
      const double fac = e24 * (lj12 + lj12m6) * inv;
      const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
      f[3 * i] += fx;
      f[3 * i + 1] += fy;
      f[3 * i + 2] += fz;
      f[3 * j] -= fx;
      f[3 * j + 1] -= fy;
      f[3 * j + 2] -= fz;
      const double mag = fabs(fx) + fabs(fy) + fabs(fz);
      fscale[i] += mag;
      fscale[j] += mag;
      res->num_kernel_calls_n3++;
      const double w = (own[i] == OWN_OWNED ? 1. : 0.) + (own[j] == OWN_OWNED ? 1. : 0.);
      res->upot_sum += (e24 * lj12m6 + sh6) * w;
      res->virial_sum[0] += drx * fx * w;
      res->virial_sum[1] += dry * fy * w;
      res->virial_sum[2] += drz * fz * w;
      res->num_global_calcs_n3++;
    }
  }
  free(mixbuf);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * VerletClusterLists
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  double box_min[3], box_max[3], halo_min[3], halo_max[3];
  double side[2], recip[2], il, il2;
  int64_t tpd[2], ntpil, num_towers, M;
} vcl_geom;

/* ClusterTowerBlock2D ctor (:36-42), estimateOptimalGridSideLength (:140-168), resize (:89-114) */
void orc_vcl_geometry(const double *box_min, const double *box_max, double il, int64_t M, int64_t num_particles,
                      vcl_geom *g) {
  g->il = il;
  g->il2 = il * il;
  g->M = M;
  for (int d = 0; d < 3; ++d) {
    g->box_min[d] = box_min[d];
    g->box_max[d] = box_max[d];
    g->halo_min[d] = box_min[d] - il;
    g->halo_max[d] = box_max[d] + il;
  }
  const double bs[3] = {box_max[0] - box_min[0], box_max[1] - box_min[1], box_max[2] - box_min[2]};
  if (num_particles == 0) {
    g->side[0] = bs[0];
    g->side[1] = bs[1];
    g->tpd[0] = g->tpd[1] = 3;
  } else {
    const double volume = bs[0] * bs[1] * bs[2];
    const double density = (double)num_particles / volume;
    const double optimal = cbrt((double)M / density);
    for (int d = 0; d < 2; ++d) {
      const double owned = ceil(bs[d] / optimal);
      const double side_new = bs[d] / owned;
      const double towers = owned + ceil(il / side_new) * 2.;
      g->side[d] = side_new;
      g->tpd[d] = (int64_t)(size_t)towers;
    }
  }
  g->ntpil = 0;
  for (int d = 0; d < 2; ++d) {
    g->recip[d] = 1. / g->side[d];
    const int64_t k = (int64_t)ceil(il / g->side[d]);
    if (k > g->ntpil) g->ntpil = k;
  }
  g->num_towers = g->tpd[0] * g->tpd[1];
}

/* ClusterTowerBlock2D::getTowerIndex2DAtPosition (:220-245), towerIndex2DTo1D = x + y*nx */
int64_t orc_vcl_tower_index(const vcl_geom *g, double px, double py) {
  const double pos[2] = {px, py};
  int64_t idx[2];
  for (int d = 0; d < 2; ++d) {
    const long value = (long)floor((pos[d] - g->box_min[d]) * g->recip[d]) + g->ntpil;
    int64_t v = value > 0 ? value : 0;
    if (v > g->tpd[d] - 1) v = g->tpd[d] - 1;
    if (pos[d] >= g->halo_max[d])
      v = g->tpd[d] - 1;
    else if (pos[d] < g->halo_min[d])
      v = 0;
    idx[d] = v;
  }
  return idx[0] + idx[1] * g->tpd[0];
}

typedef struct {
  double z;
  int64_t id;
  int idx;
} zsort_t;
static int zcmp(const void *a, const void *b) {
  const zsort_t *p = (const zsort_t *)a, *q = (const zsort_t *)b;
  if (p->z < q->z) return -1;
  if (p->z > q->z) return 1;
  if (p->id < q->id) return -1;
  if (p->id > q->id) return 1;
  return (p->idx > q->idx) - (p->idx < q->idx);
}

/* utils::ArrayMath::boxDistanceSquared (utils/ArrayMath.h:697-707) */
static double box_dist2(const double *amin, const double *amax, const double *bmin, const double *bmax) {
  double a2b[3], b2a[3];
  for (int d = 0; d < 3; ++d) {
    a2b[d] = fmax(0., amin[d] - bmax[d]);
    b2a[d] = fmax(0., bmin[d] - amax[d]);
  }
  return (a2b[0] * a2b[0] + a2b[1] * a2b[1] + a2b[2] * a2b[2]) + (b2a[0] * b2a[0] + b2a[1] * b2a[1] + b2a[2] * b2a[2]);
}

/* VerletClusterListsRebuilder::get1DInteractionCellIndexForTower / isForwardNeighbor (:313-354) */
static int64_t interaction_cell(const vcl_geom *g, int64_t tx, int64_t ty) {
  const int64_t numx = (int64_t)ceil(g->tpd[0] / (double)g->ntpil);
  return tx / g->ntpil + numx * (ty / g->ntpil);
}
static int is_forward(const vcl_geom *g, int64_t tx, int64_t ty, int64_t nx, int64_t ny) {
  const int64_t ca = interaction_cell(g, tx, ty), cb = interaction_cell(g, nx, ny);
  if (cb > ca) return 1;
  if (cb < ca) return 0;
  return nx + ny * g->tpd[0] >= tx + ty * g->tpd[0];
}

/* state kept between orc_lj_vcl and the dump getters */
static struct {
  int64_t num_slots, num_clusters, num_pairs, M, tpd[2];
  double side[2];
  int64_t *slot_particle; /* particle index per slot, -1 for padding */
  int64_t *slot_tower;
  int64_t *pairs;
} g_vcl;

void orc_vcl_dump_sizes(int64_t *sizes, double *side) {
  sizes[0] = g_vcl.num_clusters;
  sizes[1] = g_vcl.num_pairs;
  sizes[2] = g_vcl.tpd[0];
  sizes[3] = g_vcl.tpd[1];
  sizes[4] = g_vcl.M;
  side[0] = g_vcl.side[0];
  side[1] = g_vcl.side[1];
}
void orc_vcl_dump_copy(int64_t *slot_particle, int64_t *slot_tower, int64_t *pairs) {
  memcpy(slot_particle, g_vcl.slot_particle, sizeof(int64_t) * (size_t)g_vcl.num_slots);
  memcpy(slot_tower, g_vcl.slot_tower, sizeof(int64_t) * (size_t)g_vcl.num_slots);
  memcpy(pairs, g_vcl.pairs, sizeof(int64_t) * 2 * (size_t)g_vcl.num_pairs);
}

/*
 * VerletClusterLists: rebuildTowersAndClusters (VerletClusterListsRebuilder.h:67-143), ClusterTower::generateClusters
 * (ClusterTower.h:83-143) with the canonical (z, id) order, Cluster::getBoundingBox (Cluster.h:135-149),
 * updateNeighborLists / iterateNeighborTowers / calculateNeighborsBetweenTowers (:237-306, :366-409), then the
 * traversal: VCLClusterFunctor::processCluster (traversals/VCLClusterFunctor.h:38-96) over the owned clusters
 * (newton3 off: vcl_cluster_iteration / vcl_c01_balanced) or over all clusters (newton3 on: vcl_c06).
 */
int orc_lj_vcl(int64_t n, const double *x, const double *y, const double *z, const int64_t *type, const int64_t *own,
               const double *box_min, const double *box_max, double cutoff, double skin, int64_t M, int flags, int T,
               const double *eps, const double *sig, double *f, double *fscale, orc_result *res) {
  const double il = cutoff + skin;
  const int n3 = (flags & F_NEWTON3) != 0;
  int64_t *zeros = NULL;
  if (!type) {
    zeros = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    type = zeros;
  }
  int64_t num_actual = 0;
  for (int64_t i = 0; i < n; ++i) num_actual += own[i] != OWN_DUMMY;
  vcl_geom g;
  orc_vcl_geometry(box_min, box_max, il, M, num_actual, &g);
  /* sortParticlesIntoTowers (:217-232): inBox(haloBoxMin, haloBoxMax) half-open (utils/inBox.h:26-36) */
  const int64_t nt = g.num_towers;
  int64_t *tower = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int64_t *tcount = (int64_t *)calloc((size_t)nt + 1, sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) {
    tower[i] = -1;
    if (own[i] == OWN_DUMMY) continue;
    const int in = x[i] >= g.halo_min[0] && x[i] < g.halo_max[0] && y[i] >= g.halo_min[1] && y[i] < g.halo_max[1] &&
                   z[i] >= g.halo_min[2] && z[i] < g.halo_max[2];
    if (!in) continue;
    tower[i] = orc_vcl_tower_index(&g, x[i], y[i]);
    tcount[tower[i]]++;
  }
  int64_t *tstart = (int64_t *)calloc((size_t)nt + 1, sizeof(int64_t)); /* in slots, padded */
  for (int64_t t = 0; t < nt; ++t) tstart[t + 1] = tstart[t] + (tcount[t] + M - 1) / M * M;
  const int64_t num_slots = tstart[nt];
  const int64_t num_clusters = num_slots / M;
  /* slot arrays (own copies so that padding dummies exist) */
  const size_t ns = (size_t)(num_slots > 0 ? num_slots : 1);
  int64_t *slot_particle = (int64_t *)malloc(sizeof(int64_t) * ns);
  int64_t *slot_tower = (int64_t *)malloc(sizeof(int64_t) * ns);
  double *sx = (double *)malloc(sizeof(double) * ns), *sy = (double *)malloc(sizeof(double) * ns),
         *sz = (double *)malloc(sizeof(double) * ns);
  double *sfx = (double *)calloc(ns, sizeof(double)), *sfy = (double *)calloc(ns, sizeof(double)),
         *sfz = (double *)calloc(ns, sizeof(double)), *sfs = (double *)calloc(ns, sizeof(double));
  int64_t *stype = (int64_t *)calloc(ns, sizeof(int64_t)), *sown = (int64_t *)calloc(ns, sizeof(int64_t));
  for (int64_t s = 0; s < num_slots; ++s) slot_particle[s] = -1;
  {
    int64_t *fillc = (int64_t *)calloc((size_t)nt + 1, sizeof(int64_t));
    zsort_t *buf = (zsort_t *)malloc(sizeof(zsort_t) * (size_t)(n > 0 ? n : 1));
    /* gather per tower, sort by (z, id) */
    int64_t *offs = (int64_t *)calloc((size_t)nt + 1, sizeof(int64_t));
    for (int64_t t = 0; t < nt; ++t) offs[t + 1] = offs[t] + tcount[t];
    for (int64_t i = 0; i < n; ++i)
      if (tower[i] >= 0) {
        zsort_t e = {z[i], i, (int)i};
        buf[offs[tower[i]] + fillc[tower[i]]++] = e;
      }
    for (int64_t t = 0; t < nt; ++t) {
      qsort(buf + offs[t], (size_t)tcount[t], sizeof(zsort_t), zcmp);
      for (int64_t k = 0; k < tcount[t]; ++k) slot_particle[tstart[t] + k] = buf[offs[t] + k].idx;
      for (int64_t s = tstart[t]; s < tstart[t + 1]; ++s) slot_tower[s] = t;
    }
    free(fillc);
    free(buf);
    free(offs);
  }
  /* padding: copies of the last actual particle marked dummy (ClusterTower.h:96-104); later parked far away
   * (:152-161). A dummy never interacts, so only its position during the bounding-box computation matters. */
  for (int64_t t = 0; t < nt; ++t) {
    for (int64_t s = tstart[t]; s < tstart[t + 1]; ++s) {
      const int64_t pi = slot_particle[s];
      if (pi >= 0) {
        sx[s] = x[pi];
        sy[s] = y[pi];
        sz[s] = z[pi];
        stype[s] = type[pi];
        sown[s] = own[pi];
      } else {
        const int64_t last = tstart[t] + tcount[t] - 1;
        sx[s] = sx[last];
        sy[s] = sy[last];
        sz[s] = sz[last];
        stype[s] = stype[last];
        sown[s] = OWN_DUMMY;
      }
    }
  }
  /* bounding boxes + owned range per tower */
  const size_t ncl = (size_t)(num_clusters > 0 ? num_clusters : 1);
  double *bmin = (double *)malloc(sizeof(double) * 3 * ncl), *bmax = (double *)malloc(sizeof(double) * 3 * ncl);
  int *is_halo = (int *)calloc(ncl, sizeof(int));
  for (int64_t c = 0; c < num_clusters; ++c) {
    const int64_t s0 = c * M;
    double lo[3] = {sx[s0], sy[s0], sz[s0]}, hi[3] = {sx[s0 + M - 1], sy[s0 + M - 1], sz[s0 + M - 1]};
    for (int64_t k = 0; k < M; ++k) {
      lo[0] = fmin(lo[0], sx[s0 + k]);
      hi[0] = fmax(hi[0], sx[s0 + k]);
      lo[1] = fmin(lo[1], sy[s0 + k]);
      hi[1] = fmax(hi[1], sy[s0 + k]);
    }
    memcpy(bmin + 3 * c, lo, sizeof(lo));
    memcpy(bmax + 3 * c, hi, sizeof(hi));
  }
  for (int64_t t = 0; t < nt; ++t) {
    const int64_t c0 = tstart[t] / M, nc = (tstart[t + 1] - tstart[t]) / M;
    int64_t first_owned = nc, first_tail = nc;
    int found_owned = 0, found_tail = 0;
    for (int64_t c = 0; c < nc; ++c) {
      int any = 0;
      for (int64_t k = 0; k < M; ++k) any |= sown[(c0 + c) * M + k] == OWN_OWNED;
      const int contains = !found_tail && any;
      if (!found_owned && contains) {
        first_owned = c;
        found_owned = 1;
      }
      if (!found_tail && found_owned && !contains) {
        first_tail = c;
        found_tail = 1;
      }
    }
    for (int64_t c = 0; c < nc; ++c) is_halo[c0 + c] = c < first_owned || c >= first_tail;
  }
  /* neighbour lists */
  size_t cap = 1024, np = 0;
  int64_t *pairs = (int64_t *)malloc(sizeof(int64_t) * 2 * cap);
  int64_t *nstart = (int64_t *)calloc(ncl + 1, sizeof(int64_t));
  for (int64_t ty = 0; ty < g.tpd[1]; ++ty)
    for (int64_t tx = 0; tx < g.tpd[0]; ++tx) {
      /* our lists are stored per cluster in cluster order, so iterate clusters of this tower outermost */
      const int64_t t = tx + ty * g.tpd[0];
      const int64_t a0 = tstart[t] / M, na = (tstart[t + 1] - tstart[t]) / M;
      for (int64_t ca = 0; ca < na; ++ca) {
        const int64_t A = a0 + ca;
        nstart[A] = (int64_t)np;
        if (!n3 && is_halo[A]) continue;
        const int64_t minx = tx - g.ntpil > 0 ? tx - g.ntpil : 0, miny = ty - g.ntpil > 0 ? ty - g.ntpil : 0;
        const int64_t maxx = tx + g.ntpil < g.tpd[0] - 1 ? tx + g.ntpil : g.tpd[0] - 1;
        const int64_t maxy = ty + g.ntpil < g.tpd[1] - 1 ? ty + g.ntpil : g.tpd[1] - 1;
        for (int64_t ny = miny; ny <= maxy; ++ny) {
          const int64_t ady = llabs(ty - ny) - 1;
          const double disty = (double)(ady > 0 ? ady : 0) * g.side[1];
          for (int64_t nx = minx; nx <= maxx; ++nx) {
            if (n3 && !is_forward(&g, tx, ty, nx, ny)) continue;
            const int64_t adx = llabs(tx - nx) - 1;
            const double distx = (double)(adx > 0 ? adx : 0) * g.side[0];
            if (!(distx * distx + disty * disty <= g.il2)) continue;
            const int64_t tb = nx + ny * g.tpd[0];
            const int64_t b0 = tstart[tb] / M, nb = (tstart[tb + 1] - tstart[tb]) / M;
            for (int64_t cb = (tb == t && n3) ? ca + 1 : 0; cb < nb; ++cb) {
              const int64_t B = b0 + cb;
              if (B == A) continue;
              if (is_halo[A] && is_halo[B]) continue;
              if (box_dist2(bmin + 3 * A, bmax + 3 * A, bmin + 3 * B, bmax + 3 * B) <= g.il2) {
                if (np == cap) {
                  cap *= 2;
                  pairs = (int64_t *)realloc(pairs, sizeof(int64_t) * 2 * cap);
                }
                pairs[2 * np] = A;
                pairs[2 * np + 1] = B;
                ++np;
              }
            }
          }
        }
      }
    }
  /* nstart was filled in tower (y,x) order which is also ascending cluster order */
  nstart[num_clusters] = (int64_t)np;

  /* traversal */
  lj_params p;
  double *mixbuf = (double *)malloc(sizeof(double) * 3 * (size_t)(T > 0 ? T * T : 1));
  make_params(&p, cutoff, flags, T, eps, sig, mixbuf);
  memset(res, 0, sizeof(*res));
  pview v = {sx, sy, sz, sfx, sfy, sfz, sfs, stype, sown};
  int *idx1 = (int *)malloc(sizeof(int) * (size_t)M), *idx2 = (int *)malloc(sizeof(int) * (size_t)M);
  for (int64_t A = 0; A < num_clusters; ++A) {
    if (!n3 && is_halo[A]) continue;
    for (int64_t k = 0; k < M; ++k) idx1[k] = (int)(A * M + k);
    if (!is_halo[A]) soa_single(&p, &v, idx1, (int)M, res);
    for (int64_t e = nstart[A]; e < nstart[A + 1]; ++e) {
      const int64_t B = pairs[2 * e + 1];
      for (int64_t k = 0; k < M; ++k) idx2[k] = (int)(B * M + k);
      soa_pair(&p, &v, idx1, (int)M, idx2, (int)M, n3, res);
    }
  }
  memset(f, 0, sizeof(double) * 3 * (size_t)n);
  memset(fscale, 0, sizeof(double) * (size_t)n);
  for (int64_t s = 0; s < num_slots; ++s) {
    const int64_t pi = slot_particle[s];
    if (pi < 0) continue;
    f[3 * pi] = sfx[s];
    f[3 * pi + 1] = sfy[s];
    f[3 * pi + 2] = sfz[s];
    fscale[pi] = sfs[s];
  }
  /* keep the structure for the dump getters */
  free(g_vcl.slot_particle);
  free(g_vcl.slot_tower);
  free(g_vcl.pairs);
  g_vcl.num_slots = num_slots;
  g_vcl.num_clusters = num_clusters;
  g_vcl.num_pairs = (int64_t)np;
  g_vcl.M = M;
  g_vcl.tpd[0] = g.tpd[0];
  g_vcl.tpd[1] = g.tpd[1];
  g_vcl.side[0] = g.side[0];
  g_vcl.side[1] = g.side[1];
  g_vcl.slot_particle = slot_particle;
  g_vcl.slot_tower = slot_tower;
  g_vcl.pairs = pairs;
  free(tower);
  free(tcount);
  free(tstart);
  free(sx);
  free(sy);
  free(sz);
  free(sfx);
  free(sfy);
  free(sfz);
  free(sfs);
  free(stype);
  free(sown);
  free(bmin);
  free(bmax);
  free(is_halo);
  free(nstart);
  free(mixbuf);
  free(idx1);
  free(idx2);
  free(zeros);
  return 0;
}
