// Built-in functors other than single-site Lennard-Jones, on the gpuLinkedCells container (fp64):
//   * sphLib::SPHCalcDensityFunctor    (applicationLibrary/sph/SPHLibrary/SPHCalcDensityFunctor.h:43-63)
//   * sphLib::SPHCalcHydroForceFunctor (applicationLibrary/sph/SPHLibrary/SPHCalcHydroForceFunctor.h:45-108)
//     with the kernels of SPHKernels.h:37-87
//   * mdLib::AxilrodTellerMutoFunctor  (…/molecularDynamicsLibrary/AxilrodTellerMutoFunctor.h:186-293), triwise
//   * mdLib::LJMultisiteFunctor        (…/molecularDynamicsLibrary/LJMultisiteFunctor.h:181-274)
// Pair set = the reference's lc_c08 / lc_c18 cell pairs (all particle pairs of cell pairs whose border distance is at most
// the interaction length, halo-only cell pairs skipped, CellFunctor.h:173-184); one thread owns one particle i.
// Arithmetic follows the AoS functors expression by expression; every product and sum that feeds a comparison is rounded
// separately (__dmul_rn / __dadd_rn), like the oracle built with -ffp-contract=off.
#include "internal.cuh"
#include "lj_device.cuh"
#include "lc_warp.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// linked-cells pair walk shared by the kernels below
// ---------------------------------------------------------------------------------------------------------------------
struct LCWalk {
  LCGeom g;
  int64_t n;
  const int32_t *own;
  const int *slotCell, *cellStart;
  const int *stencil;  // 3 ints per entry, entry 0 = self
  int stencilN;
};

// Calls f(j) for every partner slot j of slot i. N3: each pair once (same cell: j > i; other cells: forward half of the
// stencil); otherwise every j != i of all stencil cells. Dummies are skipped, and so are cell pairs in which neither
// cell can hold owned particles.
template <bool N3, class F>
__device__ __forceinline__ void lcForEachPartner(const LCWalk &a, int64_t i, F &&f) {
  const LCGeom &g = a.g;
  const int c = a.slotCell[i];
  const int cx = c % g.cellsPerDim[0], cy = (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
            cz = c / (g.cellsPerDim[0] * g.cellsPerDim[1]);
  const bool canOwnI = apbCellCanOwn(g, cx, cy, cz);
  if (!canOwnI && !N3) return;  // forces on particles of halo cells are never used
  for (int s = 0; s < a.stencilN; ++s) {
    const int ox = a.stencil[3 * s], oy = a.stencil[3 * s + 1], oz = a.stencil[3 * s + 2];
    const int lin = (oz * g.cellsPerDim[1] + oy) * g.cellsPerDim[0] + ox;
    if (N3 && lin < 0) continue;
    const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
    if (nx < 0 || ny < 0 || nz < 0 || nx >= g.cellsPerDim[0] || ny >= g.cellsPerDim[1] || nz >= g.cellsPerDim[2]) continue;
    if (!canOwnI && !apbCellCanOwn(g, nx, ny, nz)) continue;
    const int c2 = c + lin;
    const int j0 = a.cellStart[c2], j1 = a.cellStart[c2 + 1];
    for (int j = j0; j < j1; ++j) {
      if (s == 0 && (N3 ? j <= i : j == i)) continue;
      if (a.own[j] == APB_OWN_DUMMY) continue;
      f(j);
    }
  }
}

// Every non-dummy j > i of all stencil cells, halo cells included (newton3 triplets are owned by their lowest slot,
// which may well be a halo copy).
template <class F>
__device__ __forceinline__ void lcForEachHigherSlot(const LCWalk &a, int64_t i, F &&f) {
  const LCGeom &g = a.g;
  const int c = a.slotCell[i];
  const int cx = c % g.cellsPerDim[0], cy = (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
            cz = c / (g.cellsPerDim[0] * g.cellsPerDim[1]);
  for (int s = 0; s < a.stencilN; ++s) {
    const int ox = a.stencil[3 * s], oy = a.stencil[3 * s + 1], oz = a.stencil[3 * s + 2];
    const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
    if (nx < 0 || ny < 0 || nz < 0 || nx >= g.cellsPerDim[0] || ny >= g.cellsPerDim[1] || nz >= g.cellsPerDim[2]) continue;
    const int c2 = c + (oz * g.cellsPerDim[1] + oy) * g.cellsPerDim[0] + ox;
    const int j1 = a.cellStart[c2 + 1];
    for (int j = max(a.cellStart[c2], static_cast<int>(i) + 1); j < j1; ++j)
      if (a.own[j] != APB_OWN_DUMMY) f(j);
  }
}

static LCWalk makeWalk(apb_handle h) {
  LCWalk w;
  w.g = h->lc;
  w.n = h->nslots;
  w.own = h->own;
  w.slotCell = static_cast<const int *>(h->slotCell.p);
  w.cellStart = static_cast<const int *>(h->start.p);
  w.stencil = static_cast<const int *>(h->stencilDev.p);
  w.stencilN = h->stencilN;
  return w;
}

__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
}

// ---------------------------------------------------------------------------------------------------------------------
// SPH kernels (SPHKernels.h:37-87)
// ---------------------------------------------------------------------------------------------------------------------
#define SPH_SUPPORT 2.5
#define SPH_PI 3.14159265358979323846

__device__ __forceinline__ double sphW(double dr2, double h) {
  const double H = SPH_SUPPORT * h;
  if (dr2 < __dmul_rn(H, H)) {
    const double s = sqrt(dr2) / H;
    const double s1 = 1.0 - s;
    const double s2 = fmax(0., 0.5 - s);
    double r = __dadd_rn(__dmul_rn(__dmul_rn(s1, s1), s1), -__dmul_rn(4.0, __dmul_rn(__dmul_rn(s2, s2), s2)));
    r = __dmul_rn(r, 16.0 / SPH_PI / __dmul_rn(__dmul_rn(H, H), H));
    return r;
  }
  return 0.;
}

// gradW(dr, h) = dr * scale; returns scale
__device__ __forceinline__ double sphGradWScale(double drabs, double h) {
  const double H = SPH_SUPPORT * h;
  const double s = drabs / H;
  const double s1 = (1.0 - s < 0) ? 0 : 1.0 - s;
  const double s2 = (0.5 - s < 0) ? 0 : 0.5 - s;
  double r = __dadd_rn(__dmul_rn(-3.0, __dmul_rn(s1, s1)), __dmul_rn(12.0, __dmul_rn(s2, s2)));
  r = __dmul_rn(r, 16.0 / SPH_PI / __dmul_rn(__dmul_rn(H, H), H));
  return r / __dadd_rn(__dmul_rn(drabs, H), __dmul_rn(1.0e-6, h));
}

struct SPHArgs {
  LCWalk w;
  const double *x, *y, *z, *vx, *vy, *vz, *mass, *smth, *pressure, *snd;
  double *density, *ax, *ay, *az, *engDot, *vsigmax;
};

// SPHCalcDensityFunctor::AoSFunctor (:43-63): rho_i += m_j W(dr, h_i); newton3: rho_j += m_i W(dr, h_j)
template <bool N3>
__global__ void __launch_bounds__(128) kSPHDensityLC(SPHArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.w.n || a.w.own[i] == APB_OWN_DUMMY) return;
  const double xi = a.x[i], yi = a.y[i], zi = a.z[i], hi = a.smth[i], mi = a.mass[i];
  double rho = 0.;
  lcForEachPartner<N3>(a.w, i, [&](int j) {
    const double drx = a.x[j] - xi, dry = a.y[j] - yi, drz = a.z[j] - zi;
    const double dr2 = dot3(drx, dry, drz, drx, dry, drz);
    rho += __dmul_rn(a.mass[j], sphW(dr2, hi));
    if (N3) {
      const double d2 = __dmul_rn(mi, sphW(dr2, a.smth[j]));
      if (d2 != 0.) atomicAdd(a.density + j, d2);
    }
  });
  if (N3)
    atomicAdd(a.density + i, rho);
  else
    a.density[i] += rho;
}

__device__ __forceinline__ void atomicMaxPositive(double *addr, double v) {
  // v_sig = c_i + c_j - 3 w with w <= 0 is positive: the bit patterns of non-negative doubles order like integers
  atomicMax(reinterpret_cast<long long *>(addr), __double_as_longlong(v));
}

// gradW(dr, h) scale with the normalisation 16 / pi / H^3 handed in (a per-particle factor; the list kernels precompute it,
// the others evaluate it here - the same operations on the same operands as sphGradWScale)
__device__ __forceinline__ double sphGradWNorm(double h) {
  const double H = SPH_SUPPORT * h;
  return 16.0 / SPH_PI / __dmul_rn(__dmul_rn(H, H), H);
}
__device__ __forceinline__ double sphGradWScaleNorm(double drabs, double h, double norm) {
  const double H = SPH_SUPPORT * h;
  const double s = drabs / H;
  const double s1 = (1.0 - s < 0) ? 0 : 1.0 - s;
  const double s2 = (0.5 - s < 0) ? 0 : 0.5 - s;
  double r = __dadd_rn(__dmul_rn(-3.0, __dmul_rn(s1, s1)), __dmul_rn(12.0, __dmul_rn(s2, s2)));
  r = __dmul_rn(r, norm);
  return r / __dadd_rn(__dmul_rn(drabs, H), __dmul_rn(1.0e-6, h));
}

// One pair of the hydro-force functor seen from particle i (SPHCalcHydroForceFunctor::AoSFunctor, :45-108), shared by
// the three kernel variants; the caller has already decided that the pair is inside i's support (dr2 < cut2).
struct SPHHydroI {
  double x, y, z, vx, vy, vz, h, m, rho, pOverRho2, c, norm;
};
struct SPHHydroSum {
  double ax = 0., ay = 0., az = 0., eng = 0., vmax = 0.;
};
template <bool N3>
__device__ __forceinline__ void sphHydroPair(const SPHArgs &a, const SPHHydroI &I, int j, double drx, double dry, double drz,
                                             double dr2, double vxj, double vyj, double vzj, double mj, double cj,
                                             double rhoj, double PjOverRho2, double hj, double normJ, SPHHydroSum &acc) {
  const double dvx = I.vx - vxj, dvy = I.vy - vyj, dvz = I.vz - vzj;
  const double dvdr = dot3(dvx, dvy, dvz, drx, dry, drz);
  const double drabs = sqrt(dr2);
  const double wij = (dvdr < 0) ? dvdr / drabs : 0;
  const double vsig = __dadd_rn(__dadd_rn(I.c, cj), -__dmul_rn(3.0, wij));
  acc.vmax = fmax(acc.vmax, vsig);
  const double AV = __dmul_rn(__dmul_rn(-0.5, vsig), wij) / __dmul_rn(0.5, __dadd_rn(I.rho, rhoj));
  // gradW_ij = (gradW(dr, h_i) + gradW(dr, h_j)) * 0.5, component by component
  const double gi = sphGradWScaleNorm(drabs, I.h, I.norm), gj = sphGradWScaleNorm(drabs, hj, normJ);
  const double gx = __dmul_rn(__dadd_rn(__dmul_rn(drx, gi), __dmul_rn(drx, gj)), 0.5);
  const double gy = __dmul_rn(__dadd_rn(__dmul_rn(dry, gi), __dmul_rn(dry, gj)), 0.5);
  const double gz = __dmul_rn(__dadd_rn(__dmul_rn(drz, gi), __dmul_rn(drz, gj)), 0.5);
  const double scale = __dadd_rn(__dadd_rn(I.pOverRho2, PjOverRho2), AV);
  const double si = __dmul_rn(scale, mj);
  acc.ax -= __dmul_rn(gx, si);
  acc.ay -= __dmul_rn(gy, si);
  acc.az -= __dmul_rn(gz, si);
  const double gdv = dot3(gx, gy, gz, dvx, dvy, dvz);
  const double scale2i = __dmul_rn(mj, __dadd_rn(I.pOverRho2, __dmul_rn(0.5, AV)));
  acc.eng += __dmul_rn(gdv, scale2i);
  if (N3) {
    atomicMaxPositive(a.vsigmax + j, vsig);
    const double sj = __dmul_rn(scale, I.m);
    atomicAdd(a.ax + j, __dmul_rn(gx, sj));
    atomicAdd(a.ay + j, __dmul_rn(gy, sj));
    atomicAdd(a.az + j, __dmul_rn(gz, sj));
    const double scale2j = __dmul_rn(I.m, __dadd_rn(PjOverRho2, __dmul_rn(0.5, AV)));
    atomicAdd(a.engDot + j, __dmul_rn(gdv, scale2j));
  }
}
__device__ __forceinline__ SPHHydroI sphHydroLoadI(const SPHArgs &a, int64_t i) {
  SPHHydroI I;
  I.x = a.x[i], I.y = a.y[i], I.z = a.z[i], I.vx = a.vx[i], I.vy = a.vy[i], I.vz = a.vz[i];
  I.h = a.smth[i], I.m = a.mass[i], I.rho = a.density[i], I.c = a.snd[i];
  I.pOverRho2 = a.pressure[i] / __dmul_rn(I.rho, I.rho);
  I.norm = sphGradWNorm(I.h);
  return I;
}
// partner j from the particle columns (cell-walk and warp kernels)
template <bool N3>
__device__ __forceinline__ void sphHydroPairFromColumns(const SPHArgs &a, const SPHHydroI &I, int j, double drx, double dry,
                                                        double drz, double dr2, SPHHydroSum &acc) {
  const double rhoj = a.density[j], hj = a.smth[j];
  sphHydroPair<N3>(a, I, j, drx, dry, drz, dr2, a.vx[j], a.vy[j], a.vz[j], a.mass[j], a.snd[j], rhoj,
                   a.pressure[j] / __dmul_rn(rhoj, rhoj), hj, sphGradWNorm(hj), acc);
}
template <bool N3>
__device__ __forceinline__ void sphHydroStore(const SPHArgs &a, int64_t i, const SPHHydroSum &acc) {
  if (N3) {
    atomicAdd(a.ax + i, acc.ax);
    atomicAdd(a.ay + i, acc.ay);
    atomicAdd(a.az + i, acc.az);
    atomicAdd(a.engDot + i, acc.eng);
    if (acc.vmax > 0.) atomicMaxPositive(a.vsigmax + i, acc.vmax);
  } else {
    a.ax[i] += acc.ax;
    a.ay[i] += acc.ay;
    a.az[i] += acc.az;
    a.engDot[i] += acc.eng;
    a.vsigmax[i] = fmax(a.vsigmax[i], acc.vmax);
  }
}

// SPHCalcHydroForceFunctor::AoSFunctor (:45-108), one thread per slot walking the stencil cells (round 1; APB_LC_KERNEL=thread)
template <bool N3>
__global__ void __launch_bounds__(128) kSPHHydroLC(SPHArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.w.n || a.w.own[i] == APB_OWN_DUMMY) return;
  const SPHHydroI I = sphHydroLoadI(a, i);
  const double cut = I.h * SPH_SUPPORT;
  const double cut2 = __dmul_rn(cut, cut);
  SPHHydroSum acc;
  lcForEachPartner<N3>(a.w, i, [&](int j) {
    const double drx = I.x - a.x[j], dry = I.y - a.y[j], drz = I.z - a.z[j];
    const double dr2 = dot3(drx, dry, drz, drx, dry, drz);
    if (dr2 >= cut2) return;
    sphHydroPairFromColumns<N3>(a, I, j, drx, dry, drz, dr2, acc);
  });
  sphHydroStore<N3>(a, i, acc);
}

// ---- one warp per particle slot (lc_warp.cuh) ---------------------------------------------------------------------------
// The SPH pair arithmetic is heavy (three divisions and a square root in the hydro force) and only 10-15 % of the
// candidates of the 27 stencil cells lie inside the kernel support: the lanes of a warp share the candidates of slot i,
// the hits are compacted and evaluated on full rows of 32, the lane partial sums are reduced with shuffles.
struct SPHWarpArgs {
  SPHArgs s;
  LCWarpGeom w;
  double interactionLength2;
};

template <bool N3>
__global__ void __launch_bounds__(LCW_WARPS * 32) kSPHDensityWarp(SPHWarpArgs wa) {
  __shared__ int queues[LCW_WARPS][64];
  const SPHArgs &a = wa.s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * LCW_WARPS + warp; i < a.w.n; i += static_cast<int64_t>(gridDim.x) * LCW_WARPS) {
    if (a.w.own[i] == APB_OWN_DUMMY) continue;
    const int c = a.w.slotCell[i];
    const LCGeom &g = a.w.g;
    const bool canOwnI = apbCellCanOwn(g, c % g.cellsPerDim[0], (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
                                       c / (g.cellsPerDim[0] * g.cellsPerDim[1]));
    if (!canOwnI && !N3) continue;
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i], hi = a.smth[i], mi = a.mass[i];
    // newton3: the pair also feeds rho_j with the partner's support, which may be the larger one
    const double H = SPH_SUPPORT * hi;
    const double reach2 = N3 ? wa.interactionLength2 : __dmul_rn(H, H);
    double rho = 0.;
    lcWarpWalk<N3>(
        wa.w, i, c, !canOwnI, queues[warp],
        [&](int j) {
          if (j == i || a.w.own[j] == APB_OWN_DUMMY) return false;
          const double drx = a.x[j] - xi, dry = a.y[j] - yi, drz = a.z[j] - zi;
          return dot3(drx, dry, drz, drx, dry, drz) < reach2;
        },
        [&](int j) {
          const double drx = a.x[j] - xi, dry = a.y[j] - yi, drz = a.z[j] - zi;
          const double dr2 = dot3(drx, dry, drz, drx, dry, drz);
          rho += __dmul_rn(a.mass[j], sphW(dr2, hi));
          if (N3) {
            const double d2 = __dmul_rn(mi, sphW(dr2, a.smth[j]));
            if (d2 != 0.) atomicAdd(a.density + j, d2);
          }
        });
    rho = lcWarpSum(rho);
    if (lane == 0) {
      if (N3)
        atomicAdd(a.density + i, rho);
      else
        a.density[i] += rho;
    }
  }
}

template <bool N3>
__global__ void __launch_bounds__(LCW_WARPS * 32) kSPHHydroWarp(SPHWarpArgs wa) {
  __shared__ int queues[LCW_WARPS][64];
  const SPHArgs &a = wa.s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * LCW_WARPS + warp; i < a.w.n; i += static_cast<int64_t>(gridDim.x) * LCW_WARPS) {
    if (a.w.own[i] == APB_OWN_DUMMY) continue;
    const int c = a.w.slotCell[i];
    const LCGeom &g = a.w.g;
    const bool canOwnI = apbCellCanOwn(g, c % g.cellsPerDim[0], (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
                                       c / (g.cellsPerDim[0] * g.cellsPerDim[1]));
    if (!canOwnI && !N3) continue;
    const SPHHydroI I = sphHydroLoadI(a, i);
    const double cut = I.h * SPH_SUPPORT;
    const double cut2 = __dmul_rn(cut, cut);
    SPHHydroSum acc;
    lcWarpWalk<N3>(
        wa.w, i, c, !canOwnI, queues[warp],
        [&](int j) {
          if (j == i || a.w.own[j] == APB_OWN_DUMMY) return false;
          const double drx = I.x - a.x[j], dry = I.y - a.y[j], drz = I.z - a.z[j];
          return dot3(drx, dry, drz, drx, dry, drz) < cut2;
        },
        [&](int j) {
          const double drx = I.x - a.x[j], dry = I.y - a.y[j], drz = I.z - a.z[j];
          sphHydroPairFromColumns<N3>(a, I, j, drx, dry, drz, dot3(drx, dry, drz, drx, dry, drz), acc);
        });
    acc.ax = lcWarpSum(acc.ax);
    acc.ay = lcWarpSum(acc.ay);
    acc.az = lcWarpSum(acc.az);
    acc.eng = lcWarpSum(acc.eng);
    acc.vmax = lcWarpMax(acc.vmax);
    if (lane == 0) sphHydroStore<N3>(a, i, acc);
  }
}

// ---- per-slot partner lists --------------------------------------------------------------------------------------------
// The 27 stencil cells hold ~1000 candidates per particle at SPH densities, of which ~110 lie inside the kernel support;
// walking them costs ~75 instructions per candidate whatever the kernel variant (measured, profiles/r02_lc_kernels.txt),
// and the density and the hydro-force pass walk the same cells. gpuLinkedCells therefore keeps, per slot, the list of
// partner slots within cutoff + skin (entry-major, coalesced across threads), built on first use after a rebuild and
// valid for as long as the cells are (particles move less than skin / 2 between rebuilds: every pair within the cutoff
// at traversal time is on the list - the guarantee a Verlet list gives, VerletListHelpers.h). The functor kernels test
// the current distance of every listed pair and do the pair arithmetic in place.
struct ATMArgs {
  LCWalk w;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type;
  double cutoff2, nu;
  const double *nuMix;  // [T*T*T] or null
  int T;
  int *nbrCount;  // per slot
  int *nbr;       // [cap][n] (entry-major: coalesced across threads)
  int cap;
  int *maxCount;
  LJStats *partials;
};

// HALF (newton3): only partners in higher slots, so that thread i owns the triplets in which it is the lowest slot
template <bool FILL, bool HALF>
__global__ void __launch_bounds__(128) kATMNeighbors(ATMArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (i < a.w.n && a.w.own[i] != APB_OWN_DUMMY) {
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    auto visit = [&](int j) {
      const double drx = a.x[j] - xi, dry = a.y[j] - yi, drz = a.z[j] - zi;
      if (dot3(drx, dry, drz, drx, dry, drz) <= a.cutoff2) {
        if (FILL) a.nbr[static_cast<size_t>(cnt) * a.w.n + i] = j;
        ++cnt;
      }
    };
    if (HALF)
      lcForEachHigherSlot(a.w, i, visit);
    else
      lcForEachPartner<false>(a.w, i, visit);
  }
  if (!FILL) {
    if (i < a.w.n) a.nbrCount[i] = cnt;
    int m = cnt;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(a.maxCount, m);
  }
}


static int ensureLCLists(apb_handle h, bool half) {
  if (h->lcListVersion == h->structureVersion && h->lcListHalf == (half ? 1 : 0)) return APB_OK;
  const int64_t n = h->nslots;
  ATMArgs a{};
  a.w = makeWalk(h);
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  const double il = h->cfg.cutoff + h->cfg.skin;
  a.cutoff2 = il * il;
  const int block = 128, grid = apbDivUp(n, block);
  APB_CHECK(apbEnsure(h, h->nbrCount, sizeof(int) * (n + 1)));
  a.nbrCount = static_cast<int *>(h->nbrCount.p);
  int *maxDev = reinterpret_cast<int *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 32);
  a.maxCount = maxDev;
  APB_CUDA(cudaMemsetAsync(maxDev, 0, 4, h->stream));
  ++h->launchCount;
  if (half)
    kATMNeighbors<false, true><<<grid, block, 0, h->stream>>>(a);
  else
    kATMNeighbors<false, false><<<grid, block, 0, h->stream>>>(a);
  APB_CUDA(cudaGetLastError());
  int cap = 0;
  APB_CUDA(cudaMemcpyAsync(&cap, maxDev, 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  APB_CHECK(apbEnsure(h, h->nbrList, sizeof(int) * static_cast<size_t>(std::max(cap, 1)) * n));
  a.nbr = static_cast<int *>(h->nbrList.p);
  a.cap = cap;
  if (cap > 0) {
    ++h->launchCount;
    if (half)
      kATMNeighbors<true, true><<<grid, block, 0, h->stream>>>(a);
    else
      kATMNeighbors<true, false><<<grid, block, 0, h->stream>>>(a);
    APB_CUDA(cudaGetLastError());
  }
  h->lcListVersion = h->structureVersion;
  h->lcListHalf = half ? 1 : 0;
  h->lcListCap = cap;
  return APB_OK;
}

struct SPHListArgs {
  SPHArgs s;
  const int *nbrCount, *nbr;
  // partner attributes packed per particle (kSPHPack): a pair reads 2 (density) or 6 (hydro force) 16-byte words from
  // one to three 32-byte sectors instead of 4 / 11 doubles from as many columns - the list kernels are bound by the
  // sectors their gathers touch, not by arithmetic
  const double2 *pack;
};

// density: {x, y, z, m}. hydro force: {x, y, z, m, vx, vy, vz, c, rho, P / rho^2, h, 16 / pi / H^3}; the last two
// per-particle factors are what the reference recomputes for every pair (SPHCalcHydroForceFunctor.h:86-96;
// SPHKernels.cpp gradW) - same operations on the same operands, so bit-identical.
template <bool HYDRO>
__global__ void kSPHPack(int64_t n, SPHArgs a, double2 *__restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 *o = out + static_cast<size_t>(i) * (HYDRO ? 6 : 2);
  // a slot that became a dummy after the lists were built (updateContainer with keepNeighborListsValid, deleteParticle)
  // is still listed: it is packed out of everybody's reach, like the reference functors skip dummies
  const double x = a.w.own[i] == APB_OWN_DUMMY ? 1e300 : a.x[i];
  o[0] = make_double2(x, a.y[i]);
  o[1] = make_double2(a.z[i], a.mass[i]);
  if (HYDRO) {
    const double rho = a.density[i], h = a.smth[i];
    o[2] = make_double2(a.vx[i], a.vy[i]);
    o[3] = make_double2(a.vz[i], a.snd[i]);
    o[4] = make_double2(rho, a.pressure[i] / __dmul_rn(rho, rho));
    o[5] = make_double2(h, sphGradWNorm(h));
  }
}

// The list kernels are bound by the latency of the partner gathers (ncu: long-scoreboard stalls dominate), so the
// entries are taken SPH_LIST_BATCH at a time: all indices, then all packed partner words, are in flight together before
// the first pair is evaluated.
#define SPH_LIST_BATCH 4
#ifndef SPH_HYDRO_BATCH
#define SPH_HYDRO_BATCH 1  // 12 registers per entry in flight; measured 6.51 / 6.54 / 10.4 ms for 1 / 2 / 4 (profiles/r02_lc_kernels.txt)
#endif
#ifndef SPH_HYDRO_MINBLOCKS
#define SPH_HYDRO_MINBLOCKS 8  // 64 registers (88 B spilled): 5.71 ms against 6.04 / 6.50 / 7.75 ms for 6 / 5 / 4 blocks per SM - resident warps hide the division chains
#endif

template <bool N3>
__global__ void __launch_bounds__(128) kSPHDensityList(SPHListArgs la) {
  const SPHArgs &a = la.s;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.w.n || a.w.own[i] == APB_OWN_DUMMY) return;
  const int cnt = la.nbrCount[i];
  if (cnt == 0) return;
  const double xi = a.x[i], yi = a.y[i], zi = a.z[i], hi = a.smth[i], mi = a.mass[i];
  double rho = 0.;
  for (int p0 = 0; p0 < cnt; p0 += SPH_LIST_BATCH) {
    int j[SPH_LIST_BATCH];
    double2 q0[SPH_LIST_BATCH], q1[SPH_LIST_BATCH];
#pragma unroll
    for (int b = 0; b < SPH_LIST_BATCH; ++b) j[b] = p0 + b < cnt ? la.nbr[static_cast<size_t>(p0 + b) * a.w.n + i] : -1;
#pragma unroll
    for (int b = 0; b < SPH_LIST_BATCH; ++b) {
      const size_t e = 2 * static_cast<size_t>(max(j[b], 0));
      q0[b] = __ldg(la.pack + e);
      q1[b] = __ldg(la.pack + e + 1);
    }
#pragma unroll
    for (int b = 0; b < SPH_LIST_BATCH; ++b) {
      if (j[b] < 0) continue;
      const double drx = q0[b].x - xi, dry = q0[b].y - yi, drz = q1[b].x - zi;
      const double dr2 = dot3(drx, dry, drz, drx, dry, drz);
      rho += __dmul_rn(q1[b].y, sphW(dr2, hi));
      if (N3) {
        const double d2 = __dmul_rn(mi, sphW(dr2, a.smth[j[b]]));
        if (d2 != 0.) atomicAdd(a.density + j[b], d2);
      }
    }
  }
  if (N3)
    atomicAdd(a.density + i, rho);
  else
    a.density[i] += rho;
}

template <bool N3>
__global__ void __launch_bounds__(128, SPH_HYDRO_MINBLOCKS) kSPHHydroList(SPHListArgs la) {
  const SPHArgs &a = la.s;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.w.n || a.w.own[i] == APB_OWN_DUMMY) return;
  const int cnt = la.nbrCount[i];
  if (cnt == 0) return;
  const SPHHydroI I = sphHydroLoadI(a, i);
  const double cut = I.h * SPH_SUPPORT;
  const double cut2 = __dmul_rn(cut, cut);
  SPHHydroSum acc;
  for (int p0 = 0; p0 < cnt; p0 += SPH_HYDRO_BATCH) {
    int jb[SPH_HYDRO_BATCH];
    double2 w[SPH_HYDRO_BATCH][6];
#pragma unroll
    for (int b = 0; b < SPH_HYDRO_BATCH; ++b) jb[b] = p0 + b < cnt ? la.nbr[static_cast<size_t>(p0 + b) * a.w.n + i] : -1;
#pragma unroll
    for (int b = 0; b < SPH_HYDRO_BATCH; ++b) {
      const double2 *pj = la.pack + 6 * static_cast<size_t>(max(jb[b], 0));
#pragma unroll
      for (int k = 0; k < 6; ++k) w[b][k] = __ldg(pj + k);
    }
#pragma unroll
    for (int b = 0; b < SPH_HYDRO_BATCH; ++b) {
      const int j = jb[b];
      if (j < 0) continue;
      // packed partner: {x, y} {z, m} {vx, vy} {vz, c} {rho, P / rho^2} {h, 16 / pi / H^3}
      const double drx = I.x - w[b][0].x, dry = I.y - w[b][0].y, drz = I.z - w[b][1].x;
      const double dr2 = dot3(drx, dry, drz, drx, dry, drz);
      if (dr2 >= cut2) continue;
      sphHydroPair<N3>(a, I, j, drx, dry, drz, dr2, w[b][2].x, w[b][2].y, w[b][3].x, w[b][1].y, w[b][3].y, w[b][4].x,
                       w[b][4].y, w[b][5].x, w[b][5].y, acc);
    }
  }
  sphHydroStore<N3>(a, i, acc);
}

static int computeSPH(apb_handle h, const apb_functor *f, int newton3, apb_traversal_result *out) {
  if (h->cfg.particle_kind != APB_PARTICLE_SPH) return h->fail(APB_ERR_NOT_APPLICABLE, "SPH functors need SPHParticle storage");
  if (h->cfg.container != APB_CONTAINER_LINKED_CELLS)
    return h->fail(APB_ERR_NOT_APPLICABLE, "SPH functors run on gpuLinkedCells (gpulc_c08 / gpulc_c18)");
  if (out) std::memset(out, 0, sizeof(*out));
  const int64_t n = h->nslots;
  if (n == 0) return APB_OK;
  SPHArgs a;
  a.w = makeWalk(h);
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.vx = h->col[APB_COL_VX];
  a.vy = h->col[APB_COL_VY];
  a.vz = h->col[APB_COL_VZ];
  a.mass = h->col[APB_COL_MASS];
  a.smth = h->col[APB_COL_SMTH];
  a.pressure = h->col[APB_COL_PRESSURE];
  a.snd = h->col[APB_COL_SNDSPEED];
  a.density = h->col[APB_COL_DENSITY];
  a.ax = h->col[APB_COL_FX];
  a.ay = h->col[APB_COL_FY];
  a.az = h->col[APB_COL_FZ];
  a.engDot = h->col[APB_COL_ENGDOT];
  a.vsigmax = h->col[APB_COL_VSIGMAX];
  const int lcKernel = apbLCKernelVariant(n, 3);  // 0 thread over the cells (round 1), 1 warp per slot, 3 partner lists
  if (lcKernel == 3) {
    APB_CHECK(ensureLCLists(h, newton3 != 0));
    SPHListArgs la;
    la.s = a;
    la.nbrCount = static_cast<const int *>(h->nbrCount.p);
    la.nbr = static_cast<const int *>(h->nbrList.p);
    const int lgrid = apbDivUp(n, 128);
    const bool hydro = f->kind == APB_FUNCTOR_SPH_HYDRO;
    APB_CHECK(apbEnsure(h, h->sortK2, sizeof(double2) * (hydro ? 6 : 2) * n));  // scratch of the rebuild, free between rebuilds
    la.pack = static_cast<const double2 *>(h->sortK2.p);
    ++h->launchCount;
    if (hydro)
      kSPHPack<true><<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, a, static_cast<double2 *>(h->sortK2.p));
    else
      kSPHPack<false><<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, a, static_cast<double2 *>(h->sortK2.p));
    ++h->launchCount;
    if (f->kind == APB_FUNCTOR_SPH_DENSITY) {
      if (newton3)
        kSPHDensityList<true><<<lgrid, 128, 0, h->stream>>>(la);
      else
        kSPHDensityList<false><<<lgrid, 128, 0, h->stream>>>(la);
    } else {
      if (newton3)
        kSPHHydroList<true><<<lgrid, 128, 0, h->stream>>>(la);
      else
        kSPHHydroList<false><<<lgrid, 128, 0, h->stream>>>(la);
    }
    APB_CUDA(cudaGetLastError());
    if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
    return APB_OK;
  }
  if (lcKernel != 0) {
    SPHWarpArgs wa;
    wa.s = a;
    wa.w.g = h->lc;
    wa.w.cellStart = a.w.cellStart;
    wa.w.stencilSorted = a.w.stencil + 3 * APB_MAX_STENCIL;
    wa.w.stencilN = h->stencilN;
    const double il = h->cfg.cutoff + h->cfg.skin;
    wa.interactionLength2 = il * il;
    const int grid = static_cast<int>(std::min<int64_t>(apbDivUp(n, LCW_WARPS), 148 * 16));
    ++h->launchCount;
    if (f->kind == APB_FUNCTOR_SPH_DENSITY) {
      if (newton3)
        kSPHDensityWarp<true><<<grid, LCW_WARPS * 32, 0, h->stream>>>(wa);
      else
        kSPHDensityWarp<false><<<grid, LCW_WARPS * 32, 0, h->stream>>>(wa);
    } else {
      if (newton3)
        kSPHHydroWarp<true><<<grid, LCW_WARPS * 32, 0, h->stream>>>(wa);
      else
        kSPHHydroWarp<false><<<grid, LCW_WARPS * 32, 0, h->stream>>>(wa);
    }
    APB_CUDA(cudaGetLastError());
    if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
    return APB_OK;
  }
  const int block = 128, grid = apbDivUp(n, block);
  ++h->launchCount;
  if (f->kind == APB_FUNCTOR_SPH_DENSITY) {
    if (newton3)
      kSPHDensityLC<true><<<grid, block, 0, h->stream>>>(a);
    else
      kSPHDensityLC<false><<<grid, block, 0, h->stream>>>(a);
  } else {
    if (newton3)
      kSPHHydroLC<true><<<grid, block, 0, h->stream>>>(a);
    else
      kSPHHydroLC<false><<<grid, block, 0, h->stream>>>(a);
  }
  APB_CUDA(cudaGetLastError());
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Axilrod-Teller-Muto (triwise). newton3 off: every particle i evaluates the triplets (i, j, k), j < k, of its own
// neighbourhood and receives its force (the reference's lc_c01 calls the functor three times per triplet with rotated
// roles, CellFunctor3B.h:267-273). newton3 on: every triplet once, by its lowest slot, forces on all three
// (CellFunctor3B.h:176-262 with newton3, AxilrodTellerMutoFunctor.h:240-290). Three kernels: neighbours of i within the
// cutoff (count, fill; newton3: higher slots only), then the triplets.
// ---------------------------------------------------------------------------------------------------------------------
template <bool MIX, bool STATS>
__global__ void __launch_bounds__(128) kATMTriplets(ATMArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats st;
  ljStatsZero(st);
  if (i < a.w.n && a.w.own[i] != APB_OWN_DUMMY) {
    const int cnt = a.nbrCount[i];
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    const bool owned = a.w.own[i] == APB_OWN_OWNED;
    double Fx = 0., Fy = 0., Fz = 0.;
    for (int p = 0; p < cnt; ++p) {
      const int j = a.nbr[static_cast<size_t>(p) * a.w.n + i];
      const double xj = a.x[j], yj = a.y[j], zj = a.z[j];
      const double ijx = xj - xi, ijy = yj - yi, ijz = zj - zi;
      const double d2ij = dot3(ijx, ijy, ijz, ijx, ijy, ijz);
      for (int q = p + 1; q < cnt; ++q) {
        const int k = a.nbr[static_cast<size_t>(q) * a.w.n + i];
        const double xk = a.x[k], yk = a.y[k], zk = a.z[k];
        const double jkx = xk - xj, jky = yk - yj, jkz = zk - zj;
        const double d2jk = dot3(jkx, jky, jkz, jkx, jky, jkz);
        if (STATS) ++st.dist;
        if (d2jk > a.cutoff2) continue;  // d2ij and d2ki are within the cutoff by construction of the list
        const double kix = xi - xk, kiy = yi - yk, kiz = zi - zk;
        const double d2ki = dot3(kix, kiy, kiz, kix, kiy, kiz);
        double nu = a.nu;
        if (MIX) nu = __ldg(a.nuMix + (static_cast<size_t>(ti) * a.T + a.type[j]) * a.T + a.type[k]);
        // AxilrodTellerMutoFunctor.h:217-251
        const double all2 = d2ij * d2jk * d2ki;
        const double all5 = all2 * all2 * sqrt(all2);
        const double factor = 3.0 * nu / all5;
        const double IJdKI = dot3(ijx, ijy, ijz, kix, kiy, kiz);
        const double IJdJK = dot3(ijx, ijy, ijz, jkx, jky, jkz);
        const double JKdKI = dot3(jkx, jky, jkz, kix, kiy, kiz);
        const double allDots = IJdKI * IJdJK * JKdKI;
        const double cJK = IJdKI * (IJdJK - JKdKI);
        const double cIJ = IJdJK * JKdKI - d2jk * d2ki + 5.0 * allDots / d2ij;
        const double cKI = -IJdJK * JKdKI + d2ij * d2jk - 5.0 * allDots / d2ki;
        const double fx = (jkx * cJK + ijx * cIJ + kix * cKI) * factor;
        const double fy = (jky * cJK + ijy * cIJ + kiy * cKI) * factor;
        const double fz = (jkz * cJK + ijz * cIJ + kiz * cKI) * factor;
        Fx += fx;
        Fy += fy;
        Fz += fz;
        if (STATS) {
          ++st.kNoN3;
          ++st.gNoN3;
          if (owned) {
            // potentialEnergy3 = factor (allDistsSquared - 3 allDotProducts); virial = f_i * r_i (:269-275)
            st.upot += factor * (all2 - 3.0 * allDots);
            st.vir[0] += fx * xi;
            st.vir[1] += fy * yi;
            st.vir[2] += fz * zi;
          }
        }
      }
    }
    a.fx[i] += Fx;
    a.fy[i] += Fy;
    a.fz[i] += Fz;
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

// newton3: every triplet once, by the thread of its lowest slot; forces on all three participants
// (AxilrodTellerMutoFunctor.h:240-251: F_j from its own three directions, F_k = -(F_i + F_j)), the two partners through
// RED.ADD.F64. Globals: 3 Upot and f_p * r_p for every owned participant (:269-290).
template <bool MIX, bool STATS>
__global__ void __launch_bounds__(128) kATMTripletsN3(ATMArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats st;
  ljStatsZero(st);
  if (i < a.w.n && a.w.own[i] != APB_OWN_DUMMY) {
    const int cnt = a.nbrCount[i];
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    const bool ownedI = a.w.own[i] == APB_OWN_OWNED;
    double Fx = 0., Fy = 0., Fz = 0.;
    for (int p = 0; p < cnt; ++p) {
      const int j = a.nbr[static_cast<size_t>(p) * a.w.n + i];
      const double xj = a.x[j], yj = a.y[j], zj = a.z[j];
      const double ijx = xj - xi, ijy = yj - yi, ijz = zj - zi;
      const double d2ij = dot3(ijx, ijy, ijz, ijx, ijy, ijz);
      const bool ownedJ = a.w.own[j] == APB_OWN_OWNED;
      double Fjx = 0., Fjy = 0., Fjz = 0.;
      for (int q = p + 1; q < cnt; ++q) {
        const int k = a.nbr[static_cast<size_t>(q) * a.w.n + i];
        const double xk = a.x[k], yk = a.y[k], zk = a.z[k];
        const double jkx = xk - xj, jky = yk - yj, jkz = zk - zj;
        const double d2jk = dot3(jkx, jky, jkz, jkx, jky, jkz);
        if (STATS) ++st.dist;
        if (d2jk > a.cutoff2) continue;
        const double kix = xi - xk, kiy = yi - yk, kiz = zi - zk;
        const double d2ki = dot3(kix, kiy, kiz, kix, kiy, kiz);
        double nu = a.nu;
        if (MIX) nu = __ldg(a.nuMix + (static_cast<size_t>(ti) * a.T + a.type[j]) * a.T + a.type[k]);
        const double all2 = d2ij * d2jk * d2ki;
        const double all5 = all2 * all2 * sqrt(all2);
        const double factor = 3.0 * nu / all5;
        const double IJdKI = dot3(ijx, ijy, ijz, kix, kiy, kiz);
        const double IJdJK = dot3(ijx, ijy, ijz, jkx, jky, jkz);
        const double JKdKI = dot3(jkx, jky, jkz, kix, kiy, kiz);
        const double allDots = IJdKI * IJdJK * JKdKI;
        const double cJK = IJdKI * (IJdJK - JKdKI);
        const double cIJ = IJdJK * JKdKI - d2jk * d2ki + 5.0 * allDots / d2ij;
        const double cKI = -IJdJK * JKdKI + d2ij * d2jk - 5.0 * allDots / d2ki;
        const double fix = (jkx * cJK + ijx * cIJ + kix * cKI) * factor;
        const double fiy = (jky * cJK + ijy * cIJ + kiy * cKI) * factor;
        const double fiz = (jkz * cJK + ijz * cIJ + kiz * cKI) * factor;
        // force on j (:241-247)
        const double jKI = IJdJK * (JKdKI - IJdKI);
        const double jIJ = -IJdKI * JKdKI + d2jk * d2ki - 5.0 * allDots / d2ij;
        const double jJK = IJdKI * JKdKI - d2ij * d2ki + 5.0 * allDots / d2jk;
        const double fjx = (kix * jKI + ijx * jIJ + jkx * jJK) * factor;
        const double fjy = (kiy * jKI + ijy * jIJ + jky * jJK) * factor;
        const double fjz = (kiz * jKI + ijz * jIJ + jkz * jJK) * factor;
        const double fkx = (fix + fjx) * (-1.0), fky = (fiy + fjy) * (-1.0), fkz = (fiz + fjz) * (-1.0);
        Fx += fix;
        Fy += fiy;
        Fz += fiz;
        Fjx += fjx;
        Fjy += fjy;
        Fjz += fjz;
        atomicAdd(a.fx + k, fkx);
        atomicAdd(a.fy + k, fky);
        atomicAdd(a.fz + k, fkz);
        if (STATS) {
          ++st.kN3;
          ++st.gN3;
          const double u3 = factor * (all2 - 3.0 * allDots);
          if (ownedI) {
            st.upot += u3;
            st.vir[0] += fix * xi;
            st.vir[1] += fiy * yi;
            st.vir[2] += fiz * zi;
          }
          if (ownedJ) {
            st.upot += u3;
            st.vir[0] += fjx * xj;
            st.vir[1] += fjy * yj;
            st.vir[2] += fjz * zj;
          }
          if (a.w.own[k] == APB_OWN_OWNED) {
            st.upot += u3;
            st.vir[0] += fkx * xk;
            st.vir[1] += fky * yk;
            st.vir[2] += fkz * zk;
          }
        }
      }
      atomicAdd(a.fx + j, Fjx);
      atomicAdd(a.fy + j, Fjy);
      atomicAdd(a.fz + j, Fjz);
    }
    atomicAdd(a.fx + i, Fx);
    atomicAdd(a.fy + i, Fy);
    atomicAdd(a.fz + i, Fz);
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

// One triplet (i, j, k) of the Axilrod-Teller-Muto functor from the thread that owns i
// (AxilrodTellerMutoFunctor.h:217-290): force on i into (Fx, Fy, Fz); newton3: forces on j and k through RED.ADD.F64
// (F_k = -(F_i + F_j)); globals 3 Upot and f_p * r_p for every owned participant (:269-290). Shared by the
// warp-per-slot and the two-pass kernel.
template <bool MIX, bool STATS, bool N3>
__device__ __forceinline__ void atmTriplet(const ATMArgs &a, double xi, double yi, double zi, int ti, bool ownedI, int j, int k,
                                           double xj, double yj, double zj, double xk, double yk, double zk, double &Fx,
                                           double &Fy, double &Fz, LJStats &st) {
  const double ijx = xj - xi, ijy = yj - yi, ijz = zj - zi;
  const double jkx = xk - xj, jky = yk - yj, jkz = zk - zj;
  const double kix = xi - xk, kiy = yi - yk, kiz = zi - zk;
  const double d2ij = dot3(ijx, ijy, ijz, ijx, ijy, ijz);
  const double d2jk = dot3(jkx, jky, jkz, jkx, jky, jkz);
  const double d2ki = dot3(kix, kiy, kiz, kix, kiy, kiz);
  double nu = a.nu;
  if (MIX) nu = __ldg(a.nuMix + (static_cast<size_t>(ti) * a.T + a.type[j]) * a.T + a.type[k]);
  // AxilrodTellerMutoFunctor.h:217-251
  const double all2 = d2ij * d2jk * d2ki;
  const double all5 = all2 * all2 * sqrt(all2);
  const double factor = 3.0 * nu / all5;
  const double IJdKI = dot3(ijx, ijy, ijz, kix, kiy, kiz);
  const double IJdJK = dot3(ijx, ijy, ijz, jkx, jky, jkz);
  const double JKdKI = dot3(jkx, jky, jkz, kix, kiy, kiz);
  const double allDots = IJdKI * IJdJK * JKdKI;
  const double cJK = IJdKI * (IJdJK - JKdKI);
  const double cIJ = IJdJK * JKdKI - d2jk * d2ki + 5.0 * allDots / d2ij;
  const double cKI = -IJdJK * JKdKI + d2ij * d2jk - 5.0 * allDots / d2ki;
  const double fix = (jkx * cJK + ijx * cIJ + kix * cKI) * factor;
  const double fiy = (jky * cJK + ijy * cIJ + kiy * cKI) * factor;
  const double fiz = (jkz * cJK + ijz * cIJ + kiz * cKI) * factor;
  Fx += fix;
  Fy += fiy;
  Fz += fiz;
  const double u3 = factor * (all2 - 3.0 * allDots);
  if (N3) {
    // force on j (:241-247), F_k = -(F_i + F_j)
    const double jKI = IJdJK * (JKdKI - IJdKI);
    const double jIJ = -IJdKI * JKdKI + d2jk * d2ki - 5.0 * allDots / d2ij;
    const double jJK = IJdKI * JKdKI - d2ij * d2ki + 5.0 * allDots / d2jk;
    const double fjx = (kix * jKI + ijx * jIJ + jkx * jJK) * factor;
    const double fjy = (kiy * jKI + ijy * jIJ + jky * jJK) * factor;
    const double fjz = (kiz * jKI + ijz * jIJ + jkz * jJK) * factor;
    const double fkx = (fix + fjx) * (-1.0), fky = (fiy + fjy) * (-1.0), fkz = (fiz + fjz) * (-1.0);
    atomicAdd(a.fx + j, fjx);
    atomicAdd(a.fy + j, fjy);
    atomicAdd(a.fz + j, fjz);
    atomicAdd(a.fx + k, fkx);
    atomicAdd(a.fy + k, fky);
    atomicAdd(a.fz + k, fkz);
    if (STATS) {
      ++st.kN3;
      ++st.gN3;
      if (ownedI) {
        st.upot += u3;
        st.vir[0] += fix * xi;
        st.vir[1] += fiy * yi;
        st.vir[2] += fiz * zi;
      }
      if (a.w.own[j] == APB_OWN_OWNED) {
        st.upot += u3;
        st.vir[0] += fjx * xj;
        st.vir[1] += fjy * yj;
        st.vir[2] += fjz * zj;
      }
      if (a.w.own[k] == APB_OWN_OWNED) {
        st.upot += u3;
        st.vir[0] += fkx * xk;
        st.vir[1] += fky * yk;
        st.vir[2] += fkz * zk;
      }
    }
  } else if (STATS) {
    ++st.kNoN3;
    ++st.gNoN3;
    if (ownedI) {
      // potentialEnergy3 = factor (allDistsSquared - 3 allDotProducts); virial = f_i * r_i (:269-275)
      st.upot += u3;
      st.vir[0] += fix * xi;
      st.vir[1] += fiy * yi;
      st.vir[2] += fiz * zi;
    }
  }
}

// ---- one warp per particle slot ----------------------------------------------------------------------------------------
// kATMCountWarp: neighbours of slot i within the cutoff (newton3: in higher slots), warp-cooperative walk; feeds the
// shared-memory capacity of kATMTripletsWarp. kATMTripletsWarp: the warp collects the neighbours of i into shared memory
// (positions, slot, type), its lanes share the (j, k) pairs of the neighbour list, test |r_jk| <= cutoff, compact the
// surviving triplets by ballot and evaluate them on full rows of 32 - a thread-per-particle loop executes the ~100 FP64
// instructions of a triplet for every (j, k) pair of which only ~17 % pass. No neighbour list in global memory.
// Triplet rules, globals and counters as in kATMTriplets / kATMTripletsN3.
struct ATMWarpArgs {
  ATMArgs a;
  LCWarpGeom w;
  int cap;  // neighbours per warp that fit in shared memory
};

template <bool HALF>
__global__ void __launch_bounds__(LCW_WARPS * 32) kATMCountWarp(ATMWarpArgs wa) {
  __shared__ int queues[LCW_WARPS][64];
  const ATMArgs &a = wa.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int most = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * LCW_WARPS + warp; i < a.w.n; i += static_cast<int64_t>(gridDim.x) * LCW_WARPS) {
    int cnt = 0;
    if (a.w.own[i] != APB_OWN_DUMMY) {
      const int c = a.w.slotCell[i];
      const LCGeom &g = a.w.g;
      const bool canOwnI = apbCellCanOwn(g, c % g.cellsPerDim[0], (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
                                         c / (g.cellsPerDim[0] * g.cellsPerDim[1]));
      if (HALF || canOwnI) {
        const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
        cnt = lcWarpWalk<HALF>(
            wa.w, i, c, false, queues[warp],
            [&](int j) {
              if (j == i || a.w.own[j] == APB_OWN_DUMMY) return false;
              const double drx = a.x[j] - xi, dry = a.y[j] - yi, drz = a.z[j] - zi;
              return dot3(drx, dry, drz, drx, dry, drz) <= a.cutoff2;
            },
            [](int) {});
      }
    }
    if (lane == 0) a.nbrCount[i] = cnt;
    most = max(most, cnt);
  }
  if (lane == 0 && most > 0) atomicMax(a.maxCount, most);
}

template <bool MIX, bool STATS, bool N3>
__global__ void __launch_bounds__(LCW_WARPS * 32) kATMTripletsWarp(ATMWarpArgs wa) {
  extern __shared__ __align__(16) unsigned char atmSmem[];
  __shared__ int queues[LCW_WARPS][64];
  const ATMArgs &a = wa.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cap = wa.cap;
  const unsigned below = (1u << lane) - 1u;
  double *nx = reinterpret_cast<double *>(atmSmem) + static_cast<size_t>(warp) * 3 * cap, *ny = nx + cap, *nz = ny + cap;
  int *nslot = reinterpret_cast<int *>(atmSmem + static_cast<size_t>(LCW_WARPS) * 24 * cap) + static_cast<size_t>(warp) * 2 * cap;
  int *queue = queues[warp];
  LJStats st;
  ljStatsZero(st);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * LCW_WARPS + warp; i < a.w.n; i += static_cast<int64_t>(gridDim.x) * LCW_WARPS) {
    const int ownI = a.w.own[i];
    if (ownI == APB_OWN_DUMMY) continue;
    const int cnt = a.nbrCount[i];
    if (cnt < 2) continue;
    const int c = a.w.slotCell[i];
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    const bool ownedI = ownI == APB_OWN_OWNED;
    // ---- neighbours of i -> shared memory, in walk order
    int filled = 0;
    lcWarpWalk<N3>(
        wa.w, i, c, false, queue,
        [&](int j) {
          if (j == i || a.w.own[j] == APB_OWN_DUMMY) return false;
          const double drx = a.x[j] - xi, dry = a.y[j] - yi, drz = a.z[j] - zi;
          return dot3(drx, dry, drz, drx, dry, drz) <= a.cutoff2;
        },
        [&](int j) {
          const int e = filled + lane;
          nx[e] = a.x[j];
          ny[e] = a.y[j];
          nz[e] = a.z[j];
          nslot[e] = j;
          filled += 32;
        });
    __syncwarp();
    double Fx = 0., Fy = 0., Fz = 0.;
    auto triplet = [&](int pq) {
      const int p = pq >> 16, q = pq & 0xFFFF;
      // (mixing reads the partner types through their slots)
      atmTriplet<MIX, STATS, N3>(a, xi, yi, zi, ti, ownedI, nslot[p], nslot[q], nx[p], ny[p], nz[p], nx[q], ny[q], nz[q], Fx, Fy,
                                 Fz, st);
    };
    // ---- (j, k) pairs of the list: lanes over k for every j; survivors compacted, evaluated on full rows
    int qn = 0;
    for (int p = 0; p + 1 < cnt; ++p) {
      const double xj = nx[p], yj = ny[p], zj = nz[p];
      for (int base = p + 1; base < cnt; base += 32) {
        const int q = base + lane;
        bool hit = false;
        if (q < cnt) {
          const double jkx = nx[q] - xj, jky = ny[q] - yj, jkz = nz[q] - zj;
          hit = dot3(jkx, jky, jkz, jkx, jky, jkz) <= a.cutoff2;  // d2ij and d2ki are within the cutoff by construction
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) queue[qn + __popc(m & below)] = (p << 16) | q;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
          const int pq = queue[lane];
          const int carry = lane < qn - 32 ? queue[32 + lane] : 0;
          __syncwarp();
          if (lane < qn - 32) queue[lane] = carry;
          qn -= 32;
          triplet(pq);
          __syncwarp();
        }
      }
    }
    if (lane < qn) triplet(queue[lane]);
    __syncwarp();
    if (STATS && lane == 0) st.dist += static_cast<unsigned long long>(cnt) * (cnt - 1) / 2;
    Fx = lcWarpSum(Fx);
    Fy = lcWarpSum(Fy);
    Fz = lcWarpSum(Fz);
    if (lane == 0) {
      if (N3) {
        atomicAdd(a.fx + i, Fx);
        atomicAdd(a.fy + i, Fy);
        atomicAdd(a.fz + i, Fz);
      } else {
        a.fx[i] += Fx;
        a.fy[i] += Fy;
        a.fz[i] += Fz;
      }
    }
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

// Two passes per thread over the neighbour list of its slot (kATMNeighbors). Pass 1 tests every (j, k) pair of the list
// for |r_jk| <= cutoff and records the survivors as one 64-bit mask per j in shared memory; pass 2 walks the set bits.
// kATMTriplets / kATMTripletsN3 evaluate the ~100 FP64 instructions of the triplet inside the (j, k) loop, where the
// whole warp pays for them whenever one lane's pair survives (28 % of the pairs do, so practically always); here a lane
// only runs the triplet arithmetic for its own survivors and the warp is done when its busiest lane is (lanes of a warp
// are slots of the same cell: their counts differ by ~20 %). Pass 2 is software-pipelined: the partner slots and positions
// of the next surviving triplet are loaded before the current one is evaluated (measured without it: long-scoreboard
// stalls of 7.6 per issue at 30 % resident warps, profiles/r02_lc_kernels.txt). Needs at most 64 neighbours per slot.
template <bool MIX, bool STATS, bool N3>
__global__ void __launch_bounds__(128) kATMTripletsMasked(ATMArgs a) {
  extern __shared__ unsigned long long atmMasks[];  // [cap][128]
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats st;
  ljStatsZero(st);
  const int ownI = i < a.w.n ? a.w.own[i] : APB_OWN_DUMMY;
  const int cnt = ownI != APB_OWN_DUMMY ? a.nbrCount[i] : 0;
  if (cnt >= 2) {
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    const bool ownedI = ownI == APB_OWN_OWNED;
    unsigned long long *mine = atmMasks + threadIdx.x;
    // ---- pass 1: which (j, k) pairs are within the cutoff of each other (d2ij and d2ki are, by construction of the list)
    for (int p = 0; p + 1 < cnt; ++p) {
      const int j = a.nbr[static_cast<size_t>(p) * a.w.n + i];
      const double xj = a.x[j], yj = a.y[j], zj = a.z[j];
      unsigned long long m = 0ULL;
#pragma unroll 4
      for (int q = p + 1; q < cnt; ++q) {
        const int k = a.nbr[static_cast<size_t>(q) * a.w.n + i];
        const double jkx = a.x[k] - xj, jky = a.y[k] - yj, jkz = a.z[k] - zj;
        if (dot3(jkx, jky, jkz, jkx, jky, jkz) <= a.cutoff2) m |= 1ULL << q;
      }
      mine[p * 128] = m;
    }
    if (STATS) st.dist += static_cast<unsigned long long>(cnt) * (cnt - 1) / 2;
    double Fx = 0., Fy = 0., Fz = 0.;
    auto triplet = [&](int j, int k, double xj, double yj, double zj, double xk, double yk, double zk) {
      atmTriplet<MIX, STATS, N3>(a, xi, yi, zi, ti, ownedI, j, k, xj, yj, zj, xk, yk, zk, Fx, Fy, Fz, st);
    };
    // ---- pass 2: the surviving triplets of this slot, one after the other; the loads of triplet t + 1 are in flight
    // while triplet t is evaluated
    int p = 0;
    unsigned long long m = mine[0];
    auto nextPair = [&](int &pp, int &qq) {
      while (m == 0ULL && ++p + 1 < cnt) m = mine[p * 128];
      if (m == 0ULL) return false;
      pp = p;
      qq = __ffsll(static_cast<long long>(m)) - 1;
      m &= m - 1ULL;
      return true;
    };
    int pc = 0, qc = 0;
    bool have = nextPair(pc, qc);
    int j = 0, k = 0;
    double xj = 0., yj = 0., zj = 0., xk = 0., yk = 0., zk = 0.;
    if (have) {
      j = a.nbr[static_cast<size_t>(pc) * a.w.n + i];
      k = a.nbr[static_cast<size_t>(qc) * a.w.n + i];
      xj = a.x[j], yj = a.y[j], zj = a.z[j], xk = a.x[k], yk = a.y[k], zk = a.z[k];
    }
    while (have) {
      int pn = 0, qn = 0, jn = 0, kn = 0;
      double xjn = 0., yjn = 0., zjn = 0., xkn = 0., ykn = 0., zkn = 0.;
      const bool haveNext = nextPair(pn, qn);
      if (haveNext) {
        jn = a.nbr[static_cast<size_t>(pn) * a.w.n + i];
        kn = a.nbr[static_cast<size_t>(qn) * a.w.n + i];
        xjn = a.x[jn], yjn = a.y[jn], zjn = a.z[jn], xkn = a.x[kn], ykn = a.y[kn], zkn = a.z[kn];
      }
      triplet(j, k, xj, yj, zj, xk, yk, zk);
      have = haveNext;
      j = jn, k = kn;
      xj = xjn, yj = yjn, zj = zjn, xk = xkn, yk = ykn, zk = zkn;
    }
    if (N3) {
      atomicAdd(a.fx + i, Fx);
      atomicAdd(a.fy + i, Fy);
      atomicAdd(a.fz + i, Fz);
    } else {
      a.fx[i] += Fx;
      a.fy[i] += Fy;
      a.fz[i] += Fz;
    }
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

int apbFinishStats(apb_handle h, int numBlocks, bool stats, const apb_functor *f, apb_traversal_result *out);

static int computeATM(apb_handle h, const apb_functor *f, int newton3, apb_traversal_result *out) {
  if (h->cfg.particle_kind != APB_PARTICLE_LJ) return h->fail(APB_ERR_NOT_APPLICABLE, "AxilrodTellerMutoFunctor needs MoleculeLJ particles");
  // the reference offers triwise traversals for LinkedCells / DirectSum only (CompatibleTraversals.h:219-231)
  if (h->cfg.container != APB_CONTAINER_LINKED_CELLS)
    return h->fail(APB_ERR_NOT_APPLICABLE, "triwise functors run on gpuLinkedCells only");
  if (!(f->cutoff > 0.) || f->cutoff > h->cfg.cutoff)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "functor cutoff must be in (0, container cutoff]");
  const bool mix = f->flags & APB_FUNCTOR_USE_MIXING;
  const bool stats = f->flags & (APB_FUNCTOR_CALC_GLOBALS | APB_FUNCTOR_COUNT_FLOPS);
  const int64_t n = h->nslots;
  if (n == 0) return apbFinishStats(h, 0, stats, f, out);
  ATMArgs a;
  a.w = makeWalk(h);
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.fx = h->col[APB_COL_FX];
  a.fy = h->col[APB_COL_FY];
  a.fz = h->col[APB_COL_FZ];
  a.type = h->type;
  a.cutoff2 = f->cutoff * f->cutoff;
  a.nu = f->nu;
  a.nuMix = nullptr;
  a.T = 0;
  if (mix) {
    // mixing: mixing_table holds nu_ijk = cbrt(nu_i nu_j nu_k) row-major [T*T*T] (ParticlePropertiesLibrary.h:460-470)
    if (f->num_types <= 0 || !f->mixing_table) return h->fail(APB_ERR_INVALID_ARGUMENT, "ATM mixing needs num_types and a nu table");
    const size_t cnt = static_cast<size_t>(f->num_types) * f->num_types * f->num_types;
    APB_CHECK(apbEnsure(h, h->mixDev, cnt * 8));
    APB_CUDA(cudaMemcpyAsync(h->mixDev.p, f->mixing_table, cnt * 8, cudaMemcpyHostToDevice, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    h->mixHostCache.clear();
    a.nuMix = static_cast<const double *>(h->mixDev.p);
    a.T = f->num_types;
  }
  const int block = 128, grid = apbDivUp(n, block);
  h->lcListVersion = -1;  // the buffers of the cached partner lists (ensureLCLists) are reused below
  APB_CHECK(apbEnsure(h, h->nbrCount, sizeof(int) * (n + 1)));
  a.nbrCount = static_cast<int *>(h->nbrCount.p);
  a.nbr = nullptr;
  a.cap = 0;
  int *maxDev = reinterpret_cast<int *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 32);
  a.maxCount = maxDev;
  APB_CUDA(cudaMemsetAsync(maxDev, 0, 4, h->stream));
  const int lcKernel = apbLCKernelVariant(n, 0);  // 0 thread per slot over neighbour lists, 1 warp per slot
  const bool lcThreadKernel = lcKernel != 1;
  ATMWarpArgs wa;
  wa.w.g = h->lc;
  wa.w.cellStart = a.w.cellStart;
  wa.w.stencilSorted = a.w.stencil + 3 * APB_MAX_STENCIL;
  wa.w.stencilN = h->stencilN;
  const int wgrid = static_cast<int>(std::min<int64_t>(apbDivUp(n, LCW_WARPS), 148 * 16));
  ++h->launchCount;
  if (!lcThreadKernel) {
    wa.a = a;
    wa.cap = 0;
    if (newton3)
      kATMCountWarp<true><<<wgrid, LCW_WARPS * 32, 0, h->stream>>>(wa);
    else
      kATMCountWarp<false><<<wgrid, LCW_WARPS * 32, 0, h->stream>>>(wa);
  } else if (newton3) {
    kATMNeighbors<false, true><<<grid, block, 0, h->stream>>>(a);
  } else {
    kATMNeighbors<false, false><<<grid, block, 0, h->stream>>>(a);
  }
  APB_CUDA(cudaGetLastError());
  int cap = 0;
  APB_CUDA(cudaMemcpyAsync(&cap, maxDev, 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  // the warp kernel keeps the neighbours of its 8 slots in shared memory (32 B each, rounded up to full rows of 32);
  // beyond 640 neighbours per particle the list-based thread kernels below take over
  const int capRows = (cap + 31) / 32 * 32;
  if (!lcThreadKernel && capRows <= 640 && capRows < 65536) {
    if (cap < 2) return apbFinishStats(h, 0, stats, f, out);
    wa.a = a;
    wa.cap = capRows;
    APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * wgrid));
    wa.a.partials = static_cast<LJStats *>(h->partials.p);
    const size_t smem = static_cast<size_t>(LCW_WARPS) * 32 * capRows;
    ++h->launchCount;
    const int wsel = (newton3 ? 4 : 0) | (mix ? 2 : 0) | (stats ? 1 : 0);
#define ATM_WARP_LAUNCH(MIXV, STATSV, N3V)                                                                           \
  do {                                                                                                               \
    if (smem > 48 * 1024)                                                                                            \
      APB_CUDA(cudaFuncSetAttribute(kATMTripletsWarp<MIXV, STATSV, N3V>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    static_cast<int>(smem)));                                                        \
    kATMTripletsWarp<MIXV, STATSV, N3V><<<wgrid, LCW_WARPS * 32, smem, h->stream>>>(wa);                             \
  } while (0)
    switch (wsel) {
      case 0: ATM_WARP_LAUNCH(false, false, false); break;
      case 1: ATM_WARP_LAUNCH(false, true, false); break;
      case 2: ATM_WARP_LAUNCH(true, false, false); break;
      case 3: ATM_WARP_LAUNCH(true, true, false); break;
      case 4: ATM_WARP_LAUNCH(false, false, true); break;
      case 5: ATM_WARP_LAUNCH(false, true, true); break;
      case 6: ATM_WARP_LAUNCH(true, false, true); break;
      default: ATM_WARP_LAUNCH(true, true, true); break;
    }
    APB_CUDA(cudaGetLastError());
    return apbFinishStats(h, wgrid, stats, f, out);
  }
  // (with counts from the warp kernel the fill pass below recounts: same candidates, same result)
  a.cap = cap;
  APB_CHECK(apbEnsure(h, h->nbrList, sizeof(int) * static_cast<size_t>(std::max(cap, 1)) * n));
  a.nbr = static_cast<int *>(h->nbrList.p);
  if (cap > 0) {
    ++h->launchCount;
    if (newton3)
      kATMNeighbors<true, true><<<grid, block, 0, h->stream>>>(a);
    else
      kATMNeighbors<true, false><<<grid, block, 0, h->stream>>>(a);
    APB_CUDA(cudaGetLastError());
  }
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * grid));
  a.partials = static_cast<LJStats *>(h->partials.p);
  ++h->launchCount;
  const int sel = (newton3 ? 4 : 0) | (mix ? 2 : 0) | (stats ? 1 : 0);
  static const bool atmInline = getenv("APB_ATM_INLINE") != nullptr;  // A/B switch: triplet arithmetic inside the pair loop
  if (cap <= 64 && !atmInline) {
    const size_t smem = sizeof(unsigned long long) * 128 * static_cast<size_t>(std::max(cap, 1));
#define ATM_MASKED_LAUNCH(MIXV, STATSV, N3V)                                                                             \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      APB_CUDA(cudaFuncSetAttribute(kATMTripletsMasked<MIXV, STATSV, N3V>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    static_cast<int>(smem)));                                                          \
    kATMTripletsMasked<MIXV, STATSV, N3V><<<grid, block, smem, h->stream>>>(a);                                        \
  } while (0)
    switch (sel) {
      case 0: ATM_MASKED_LAUNCH(false, false, false); break;
      case 1: ATM_MASKED_LAUNCH(false, true, false); break;
      case 2: ATM_MASKED_LAUNCH(true, false, false); break;
      case 3: ATM_MASKED_LAUNCH(true, true, false); break;
      case 4: ATM_MASKED_LAUNCH(false, false, true); break;
      case 5: ATM_MASKED_LAUNCH(false, true, true); break;
      case 6: ATM_MASKED_LAUNCH(true, false, true); break;
      default: ATM_MASKED_LAUNCH(true, true, true); break;
    }
    APB_CUDA(cudaGetLastError());
    return apbFinishStats(h, grid, stats, f, out);
  }
  switch (sel) {
    case 0: kATMTriplets<false, false><<<grid, block, 0, h->stream>>>(a); break;
    case 1: kATMTriplets<false, true><<<grid, block, 0, h->stream>>>(a); break;
    case 2: kATMTriplets<true, false><<<grid, block, 0, h->stream>>>(a); break;
    case 3: kATMTriplets<true, true><<<grid, block, 0, h->stream>>>(a); break;
    case 4: kATMTripletsN3<false, false><<<grid, block, 0, h->stream>>>(a); break;
    case 5: kATMTripletsN3<false, true><<<grid, block, 0, h->stream>>>(a); break;
    case 6: kATMTripletsN3<true, false><<<grid, block, 0, h->stream>>>(a); break;
    default: kATMTripletsN3<true, true><<<grid, block, 0, h->stream>>>(a); break;
  }
  APB_CUDA(cudaGetLastError());
  return apbFinishStats(h, grid, stats, f, out);
}

// ---------------------------------------------------------------------------------------------------------------------
// LJMultisiteFunctor::AoSFunctor (LJMultisiteFunctor.h:181-274): centre-of-mass cutoff, sites rotated by the molecule's
// quaternion (utils/Quaternion.cpp:13-46), site-site LJ without a per-site cutoff, force on the centre of mass and
// torque r_site x f. One thread per molecule i; newton3 off evaluates both directions (like the LJ kernels).
// ---------------------------------------------------------------------------------------------------------------------
#define MS_MAX_SITES 16
struct MSArgs {
  LCWalk w;
  const double *x, *y, *z, *q0, *q1, *q2, *q3;
  double *fx, *fy, *fz, *tx, *ty, *tz;
  const int32_t *type;
  double cutoff2;
  const int *siteStart;
  const double *sitePos;
  const int *siteType;
  const double *mix;  // [T*T][3] = {eps24, sigma2, shift6}
  int T;
  int applyShift;
  LJStats *partials;
};

// autopas::utils::quaternion::rotateVectorOfPositions (Quaternion.cpp:13-46): rotation matrix from (q0..q3) applied to p
__device__ __forceinline__ void msRotate(double q0, double q1, double q2, double q3, double px, double py, double pz,
                                         double &rx, double &ry, double &rz) {
  const double q00 = q0 * q0, q01 = q0 * q1, q02 = q0 * q2, q03 = q0 * q3;
  const double q11 = q1 * q1, q12 = q1 * q2, q13 = q1 * q3, q22 = q2 * q2, q23 = q2 * q3, q33 = q3 * q3;
  const double r00 = q00 + q11 - q22 - q33, r01 = 2. * (q12 - q03), r02 = 2. * (q13 + q02);
  const double r10 = 2. * (q12 + q03), r11 = q00 - q11 + q22 - q33, r12 = 2. * (q23 - q01);
  const double r20 = 2. * (q13 - q02), r21 = 2. * (q23 + q01), r22 = q00 - q11 - q22 + q33;
  rx = r00 * px + r01 * py + r02 * pz;
  ry = r10 * px + r11 * py + r12 * pz;
  rz = r20 * px + r21 * py + r22 * pz;
}

template <bool STATS, bool N3>
__global__ void __launch_bounds__(128) kLJMultisiteLC(MSArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats st;
  ljStatsZero(st);
  if (i < a.w.n && a.w.own[i] != APB_OWN_DUMMY) {
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int mi = a.type[i];
    const int sA0 = a.siteStart[mi], nA = a.siteStart[mi + 1] - sA0;
    double rAx[MS_MAX_SITES], rAy[MS_MAX_SITES], rAz[MS_MAX_SITES];
    for (int s = 0; s < nA; ++s)
      msRotate(a.q0[i], a.q1[i], a.q2[i], a.q3[i], a.sitePos[3 * (sA0 + s)], a.sitePos[3 * (sA0 + s) + 1],
               a.sitePos[3 * (sA0 + s) + 2], rAx[s], rAy[s], rAz[s]);
    const double wI = a.w.own[i] == APB_OWN_OWNED ? 1. : 0.;
    double Fx = 0., Fy = 0., Fz = 0., Tx = 0., Ty = 0., Tz = 0.;
    lcForEachPartner<N3>(a.w, i, [&](int j) {
      const double dcx = xi - a.x[j], dcy = yi - a.y[j], dcz = zi - a.z[j];
      if (dot3(dcx, dcy, dcz, dcx, dcy, dcz) > a.cutoff2) return;  // centre-of-mass cutoff (:191)
      const int mj = a.type[j];
      const int sB0 = a.siteStart[mj], nB = a.siteStart[mj + 1] - sB0;
      const double wJ = a.w.own[j] == APB_OWN_OWNED ? 1. : 0.;
      double FBx = 0., FBy = 0., FBz = 0., TBx = 0., TBy = 0., TBz = 0.;
      for (int sb = 0; sb < nB; ++sb) {
        double rBx, rBy, rBz;
        msRotate(a.q0[j], a.q1[j], a.q2[j], a.q3[j], a.sitePos[3 * (sB0 + sb)], a.sitePos[3 * (sB0 + sb) + 1],
                 a.sitePos[3 * (sB0 + sb) + 2], rBx, rBy, rBz);
        const int tb = a.siteType[sB0 + sb];
        for (int sa = 0; sa < nA; ++sa) {
          // displacement between the sites: dr_CoM + r_siteA - r_siteB (:218-220)
          const double drx = (dcx - rBx) + rAx[sa], dry = (dcy - rBy) + rAy[sa], drz = (dcz - rBz) + rAz[sa];
          const double dr2 = dot3(drx, dry, drz, drx, dry, drz);
          const double *m = a.mix + 3 * (static_cast<size_t>(a.siteType[sA0 + sa]) * a.T + tb);
          const double e24 = __ldg(m), s2 = __ldg(m + 1), shift6 = a.applyShift ? __ldg(m + 2) : 0.;
          const double inv = 1. / dr2;
          const double lj2 = s2 * inv;
          const double lj6 = lj2 * lj2 * lj2;
          const double lj12 = lj6 * lj6;
          const double lj12m6 = lj12 - lj6;
          const double fac = e24 * (lj12 + lj12m6) * inv;
          const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
          Fx += fx;
          Fy += fy;
          Fz += fz;
          // torque on A: r_siteA x f (:236-238)
          Tx += rAy[sa] * fz - rAz[sa] * fy;
          Ty += rAz[sa] * fx - rAx[sa] * fz;
          Tz += rAx[sa] * fy - rAy[sa] * fx;
          if (N3) {
            FBx -= fx;
            FBy -= fy;
            FBz -= fz;
            TBx -= rBy * fz - rBz * fy;
            TBy -= rBz * fx - rBx * fz;
            TBz -= rBx * fy - rBy * fx;
          }
          if (STATS) {
            const double w = N3 ? wI + wJ : wI;
            st.upot += (e24 * lj12m6 + shift6) * w;
            st.vir[0] += drx * fx * w;
            st.vir[1] += dry * fy * w;
            st.vir[2] += drz * fz * w;
          }
        }
      }
      if (N3) {
        atomicAdd(a.fx + j, FBx);
        atomicAdd(a.fy + j, FBy);
        atomicAdd(a.fz + j, FBz);
        atomicAdd(a.tx + j, TBx);
        atomicAdd(a.ty + j, TBy);
        atomicAdd(a.tz + j, TBz);
      }
    });
    if (N3) {
      atomicAdd(a.fx + i, Fx);
      atomicAdd(a.fy + i, Fy);
      atomicAdd(a.fz + i, Fz);
      atomicAdd(a.tx + i, Tx);
      atomicAdd(a.ty + i, Ty);
      atomicAdd(a.tz + i, Tz);
    } else {
      a.fx[i] += Fx;
      a.fy[i] += Fy;
      a.fz[i] += Fz;
      a.tx[i] += Tx;
      a.ty[i] += Ty;
      a.tz[i] += Tz;
    }
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

static int computeMultisite(apb_handle h, const apb_functor *f, int newton3, apb_traversal_result *out) {
  if (h->cfg.particle_kind != APB_PARTICLE_MULTISITE)
    return h->fail(APB_ERR_NOT_APPLICABLE, "LJMultisiteFunctor needs MultisiteMoleculeLJ storage");
  if (h->cfg.container != APB_CONTAINER_LINKED_CELLS)
    return h->fail(APB_ERR_NOT_APPLICABLE, "LJMultisiteFunctor runs on gpuLinkedCells (gpulc_c08 / gpulc_c18)");
  if (!(f->cutoff > 0.) || f->cutoff > h->cfg.cutoff)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "functor cutoff must be in (0, container cutoff]");
  if (f->num_mol_types <= 0 || !f->site_start || !f->site_positions || !f->site_types || f->num_types <= 0 || !f->mixing_table)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "LJMultisiteFunctor needs the site tables and the site-type mixing table");
  const bool stats = f->flags & APB_FUNCTOR_CALC_GLOBALS;
  const int64_t n = h->nslots;
  if (n == 0) return apbFinishStats(h, 0, stats, f, out);
  const int nMol = f->num_mol_types;
  const int nSites = f->site_start[nMol];
  for (int m = 0; m < nMol; ++m)
    if (f->site_start[m + 1] - f->site_start[m] > MS_MAX_SITES || f->site_start[m + 1] < f->site_start[m])
      return h->fail(APB_ERR_NOT_APPLICABLE, "molecules with more than 16 sites are not supported");
  // one device buffer: mixing table | site positions | site start | site types
  const size_t mixBytes = sizeof(double) * 3 * f->num_types * f->num_types;
  const size_t posBytes = sizeof(double) * 3 * nSites;
  const size_t startBytes = sizeof(int) * (nMol + 1), typeBytes = sizeof(int) * nSites;
  APB_CHECK(apbEnsure(h, h->mixDev, mixBytes + posBytes + startBytes + typeBytes + 64));
  char *base = static_cast<char *>(h->mixDev.p);
  APB_CUDA(cudaMemcpyAsync(base, f->mixing_table, mixBytes, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(base + mixBytes, f->site_positions, posBytes, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(base + mixBytes + posBytes, f->site_start, startBytes, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(base + mixBytes + posBytes + startBytes, f->site_types, typeBytes, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  h->mixHostCache.clear();
  MSArgs a;
  a.w = makeWalk(h);
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.q0 = h->col[APB_COL_Q0];
  a.q1 = h->col[APB_COL_Q1];
  a.q2 = h->col[APB_COL_Q2];
  a.q3 = h->col[APB_COL_Q3];
  a.fx = h->col[APB_COL_FX];
  a.fy = h->col[APB_COL_FY];
  a.fz = h->col[APB_COL_FZ];
  a.tx = h->col[APB_COL_TX];
  a.ty = h->col[APB_COL_TY];
  a.tz = h->col[APB_COL_TZ];
  a.type = h->type;
  a.cutoff2 = f->cutoff * f->cutoff;
  a.mix = reinterpret_cast<const double *>(base);
  a.sitePos = reinterpret_cast<const double *>(base + mixBytes);
  a.siteStart = reinterpret_cast<const int *>(base + mixBytes + posBytes);
  a.siteType = reinterpret_cast<const int *>(base + mixBytes + posBytes + startBytes);
  a.T = f->num_types;
  a.applyShift = (f->flags & APB_FUNCTOR_APPLY_SHIFT) ? 1 : 0;
  const int block = 128, grid = apbDivUp(n, block);
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * grid));
  a.partials = static_cast<LJStats *>(h->partials.p);
  ++h->launchCount;
  const int sel = (stats ? 2 : 0) | (newton3 ? 1 : 0);
  switch (sel) {
    case 0: kLJMultisiteLC<false, false><<<grid, block, 0, h->stream>>>(a); break;
    case 1: kLJMultisiteLC<false, true><<<grid, block, 0, h->stream>>>(a); break;
    case 2: kLJMultisiteLC<true, false><<<grid, block, 0, h->stream>>>(a); break;
    default: kLJMultisiteLC<true, true><<<grid, block, 0, h->stream>>>(a); break;
  }
  APB_CUDA(cudaGetLastError());
  return apbFinishStats(h, grid, stats, f, out);
}

int apbComputeOtherFunctor(apb_handle h, const apb_functor *f, int newton3, apb_traversal_result *out) {
  switch (f->kind) {
    case APB_FUNCTOR_SPH_DENSITY:
    case APB_FUNCTOR_SPH_HYDRO: return computeSPH(h, f, newton3, out);
    case APB_FUNCTOR_ATM: return computeATM(h, f, newton3, out);
    case APB_FUNCTOR_LJ_MULTISITE: return computeMultisite(h, f, newton3, out);
    default: return h->fail(APB_ERR_NOT_APPLICABLE, "functor kind not implemented on the GPU path");
  }
}

// AxilrodTellerMutoFunctor::endTraversal (:360-386) + getPotentialEnergy / getVirial (:392-420)
extern "C" void apb_atm_end_traversal(const apb_traversal_result *raw, double *upot, double *virial) {
  double u = raw->upot_sum;
  u /= 3.;
  u /= 3.;
  if (upot) *upot = u;
  if (virial) *virial = raw->virial_sum[0] + raw->virial_sum[1] + raw->virial_sum[2];
}

// AxilrodTellerMutoFunctor::getNumFLOPs (:476-486)
extern "C" uint64_t apb_atm_num_flops(const apb_traversal_result *r) {
  return r->num_dist_calls * 24 + r->num_kernel_calls_n3 * 100 + r->num_kernel_calls_no_n3 * 59 +
         r->num_global_calcs_n3 * 24 + r->num_global_calcs_no_n3 * 10;
}
