// Lennard-Jones 12-6 force kernels (fp64) for the gpuLinkedCells and gpuVerletClusterLists containers.
// Arithmetic restated from mdLib::LJFunctor (applicationLibrary/molecularDynamics/molecularDynamicsLibrary/LJFunctor.h):
//   pair kernel :146-159 (AoS, canonical) / :480-499 (SoA), globals :518-531, counters :510-516, 533-539,
//   SoAFunctorSingle is always newton3 (:200-203, :345), N3-off cell pairs are evaluated in both directions
//   (baseFunctors/CellFunctor.h:266-274), halo-only cell pairs are skipped (:173-184),
//   VCL cluster traversal: traversals/VCLClusterFunctor.h:38-96.
// The kernels are FP64-pipe bound (no tensor cores: not a contraction). One thread owns one particle i and keeps its
// force in registers; with newton3 the reaction on j goes through fp64 atomics (RED.ADD.F64).
#include "internal.cuh"
#include "lj_device.cuh"
#include "lc_warp.cuh"

// ------------------------------------------------------------------------------------------------------------------
// block reduction of per-thread statistics into one partial per block; final pass sums partials in fixed order
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) kReducePartials(const LJStats *__restrict__ partials, int numBlocks, apb_traversal_result *out) {
  __shared__ LJStats sh[32];
  LJStats s;
  ljStatsZero(s);
  // a thread's partials are loaded four at a time (independent loads in flight), summed in index order
  int b = threadIdx.x;
  for (; b + 3 * static_cast<int>(blockDim.x) < numBlocks; b += 4 * blockDim.x) {
    const LJStats p0 = partials[b], p1 = partials[b + blockDim.x], p2 = partials[b + 2 * blockDim.x],
                  p3 = partials[b + 3 * blockDim.x];
    ljStatsAdd(s, p0);
    ljStatsAdd(s, p1);
    ljStatsAdd(s, p2);
    ljStatsAdd(s, p3);
  }
  for (; b < numBlocks; b += blockDim.x) ljStatsAdd(s, partials[b]);
  ljStatsWarpReduce(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (warp == 0) {
    LJStats t;
    ljStatsZero(t);
    if (lane < (blockDim.x >> 5)) t = sh[lane];
    ljStatsWarpReduce(t);
    if (lane == 0) {
      out->upot_sum = t.upot;
      out->virial_sum[0] = t.vir[0];
      out->virial_sum[1] = t.vir[1];
      out->virial_sum[2] = t.vir[2];
      out->num_dist_calls = t.dist;
      out->num_kernel_calls_n3 = t.kN3;
      out->num_kernel_calls_no_n3 = t.kNoN3;
      out->num_global_calcs_n3 = t.gN3;
      out->num_global_calcs_no_n3 = t.gNoN3;
    }
  }
}

// first stage for many partials (16 M-particle containers have ~10^5 tiles): block b sums the contiguous slice b and
// writes one LJStats behind the partials; the single-block kernel above then finishes. Slices and orders are fixed.
__global__ void __launch_bounds__(256) kReducePartialsStage1(const LJStats *__restrict__ partials, int numBlocks, int slice,
                                                             LJStats *__restrict__ out) {
  LJStats s;
  ljStatsZero(s);
  const int b0 = blockIdx.x * slice, b1 = min(b0 + slice, numBlocks);
  for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x) ljStatsAdd(s, partials[b]);
  ljStatsBlockReduce(s, out);
}

int apbReducePartials(apb_handle h, int numBlocks, apb_traversal_result *dst) {
  const LJStats *partials = static_cast<const LJStats *>(h->partials.p);
  // (one block of 1024 threads needs 22 us for the 8 k partials of a 2 M-particle container, the two stages 12 us)
  if (numBlocks > 2048) {
    const int stage1Blocks = 128, slice = (numBlocks + stage1Blocks - 1) / stage1Blocks;
    // room behind the partials: apbEnsure may move the buffer, so the caller's partials are re-read from the handle
    APB_CHECK(apbEnsure(h, h->partials2, sizeof(LJStats) * stage1Blocks));
    LJStats *mid = static_cast<LJStats *>(h->partials2.p);
    ++h->launchCount, kReducePartialsStage1<<<stage1Blocks, 256, 0, h->stream>>>(partials, numBlocks, slice, mid);
    ++h->launchCount, kReducePartials<<<1, 1024, 0, h->stream>>>(mid, stage1Blocks, dst);
  } else {
    ++h->launchCount, kReducePartials<<<1, 1024, 0, h->stream>>>(partials, numBlocks, dst);
  }
  APB_CUDA(cudaGetLastError());
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// LinkedCells: one thread per particle slot (slots are sorted by cell), neighbour cells from the stencil
// ------------------------------------------------------------------------------------------------------------------
struct LCArgs {
  LCGeom g;
  int64_t n;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type, *own;
  const int *slotCell, *cellStart;
  const int *stencil;  // 3 ints per entry, entry 0 = self
  int stencilN;
  int processHaloCells;  // also compute (discarded) forces on particles of halo cells, like the reference does
  LJParams p;
  LJStats *partials;
};

template <bool MIX, bool STATS, bool N3>
__global__ void __launch_bounds__(128) kLJLinkedCells(LCArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats st;
  ljStatsZero(st);
  const bool inRange = i < a.n;
  const int ownI = inRange ? a.own[i] : APB_OWN_DUMMY;
  if (ownI != APB_OWN_DUMMY) {
    const LCGeom &g = a.g;
    const int c = a.slotCell[i];
    const int cx = c % g.cellsPerDim[0], cy = (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
              cz = c / (g.cellsPerDim[0] * g.cellsPerDim[1]);
    const bool canOwnI = apbCellCanOwn(g, cx, cy, cz);
    if (canOwnI || N3 || a.processHaloCells) {
      const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
      const int ti = MIX ? a.type[i] : 0;
      const double wI = ownI == APB_OWN_OWNED ? 1. : 0.;
      double fxa = 0., fya = 0., fza = 0.;
      for (int s = 0; s < a.stencilN; ++s) {
        const int ox = a.stencil[3 * s], oy = a.stencil[3 * s + 1], oz = a.stencil[3 * s + 2];
        const int lin = (oz * g.cellsPerDim[1] + oy) * g.cellsPerDim[0] + ox;
        if (N3 && lin < 0) continue;  // forward neighbours only: each cell pair once
        const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
        if (nx < 0 || ny < 0 || nz < 0 || nx >= g.cellsPerDim[0] || ny >= g.cellsPerDim[1] || nz >= g.cellsPerDim[2])
          continue;
        const bool canOwnJ = apbCellCanOwn(g, nx, ny, nz);
        if (!canOwnI && !canOwnJ) continue;  // CellFunctor.h:173-184
        const int c2 = c + lin;
        const int j0 = a.cellStart[c2], j1 = a.cellStart[c2 + 1];
        const bool self = s == 0;
        for (int j = j0; j < j1; ++j) {
          if (self && (N3 ? j <= i : j == i)) continue;
          const int ownJ = a.own[j];
          if (ownJ == APB_OWN_DUMMY) continue;
          const double drx = xi - a.x[j], dry = yi - a.y[j], drz = zi - a.z[j];
          const double dr2 = ljDist2(drx, dry, drz);
          const bool hit = dr2 <= a.p.cutoff2;
          // counters: a same-cell pair is one newton3 evaluation in the reference, whatever the newton3 flag
          const bool countThis = STATS && (!self || N3 || j > i);
          if (countThis) ++st.dist;
          if (hit) {
            double upot6;
            const double fac = ljEval<MIX>(a.p, dr2, ti, MIX ? a.type[j] : 0, upot6);
            const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
            fxa += fx;
            fya += fy;
            fza += fz;
            if (N3) {
              atomicAdd(&a.fx[j], -fx);
              atomicAdd(&a.fy[j], -fy);
              atomicAdd(&a.fz[j], -fz);
            }
            if (STATS) {
              const double wJ = ownJ == APB_OWN_OWNED ? 1. : 0.;
              // globals: every evaluated direction adds upot6 * [i owned] (+ [j owned] with newton3)
              const double w = N3 ? wI + wJ : wI;
              st.upot += upot6 * w;
              st.vir[0] += drx * fx * w;
              st.vir[1] += dry * fy * w;
              st.vir[2] += drz * fz * w;
              if (countThis) {
                if (N3 || self) {
                  ++st.kN3;
                  ++st.gN3;
                } else {
                  ++st.kNoN3;
                  ++st.gNoN3;
                }
              }
            }
          }
        }
      }
      if (N3) {
        atomicAdd(&a.fx[i], fxa);
        atomicAdd(&a.fy[i], fya);
        atomicAdd(&a.fz[i], fza);
      } else {
        a.fx[i] += fxa;
        a.fy[i] += fya;
        a.fz[i] += fza;
      }
    }
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

// The same traversal with one warp per particle slot (lc_warp.cuh): the lanes share the candidates of the stencil cells,
// hits are compacted and evaluated on full rows. Counters, ownership weights and the halo-cell rules are those of
// kLJLinkedCells above. Forces of slot i are reduced over the lanes with shuffles; newton3 scatters -f to the partner
// with one RED per component. The warps of a block take consecutive slots (same cell: the candidate loads hit in L1)
// and stride over the slots.
template <bool MIX, bool STATS, bool N3>
__global__ void __launch_bounds__(LCW_WARPS * 32) kLJLinkedCellsWarp(LCArgs a, LCWarpGeom w) {
  __shared__ int queues[LCW_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LJStats st;
  ljStatsZero(st);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * LCW_WARPS + warp; i < a.n; i += static_cast<int64_t>(gridDim.x) * LCW_WARPS) {
    const int ownI = a.own[i];
    if (ownI == APB_OWN_DUMMY) continue;
    const LCGeom &g = a.g;
    const int c = a.slotCell[i];
    const int cx = c % g.cellsPerDim[0], cy = (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
              cz = c / (g.cellsPerDim[0] * g.cellsPerDim[1]);
    const bool canOwnI = apbCellCanOwn(g, cx, cy, cz);
    if (!(canOwnI || N3 || a.processHaloCells)) continue;
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    const double wI = ownI == APB_OWN_OWNED ? 1. : 0.;
    const int self0 = a.cellStart[c], self1 = a.cellStart[c + 1];
    double fxa = 0., fya = 0., fza = 0.;
    lcWarpWalk<N3>(
        w, i, c, !canOwnI, queues[warp],
        [&](int j) {
          if (!N3 && j == i) return false;
          if (a.own[j] == APB_OWN_DUMMY) return false;
          const double drx = xi - a.x[j], dry = yi - a.y[j], drz = zi - a.z[j];
          const bool self = j >= self0 && j < self1;
          // counters: a same-cell pair is one newton3 evaluation in the reference, whatever the newton3 flag
          if (STATS && (!self || N3 || j > i)) ++st.dist;
          return ljDist2(drx, dry, drz) <= a.p.cutoff2;
        },
        [&](int j) {
          const double drx = xi - a.x[j], dry = yi - a.y[j], drz = zi - a.z[j];
          const double dr2 = ljDist2(drx, dry, drz);
          double upot6;
          const double fac = ljEval<MIX>(a.p, dr2, ti, MIX ? a.type[j] : 0, upot6);
          const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
          fxa += fx;
          fya += fy;
          fza += fz;
          if (N3) {
            atomicAdd(&a.fx[j], -fx);
            atomicAdd(&a.fy[j], -fy);
            atomicAdd(&a.fz[j], -fz);
          }
          if (STATS) {
            const double wJ = a.own[j] == APB_OWN_OWNED ? 1. : 0.;
            const double wgt = N3 ? wI + wJ : wI;
            st.upot += upot6 * wgt;
            st.vir[0] += drx * fx * wgt;
            st.vir[1] += dry * fy * wgt;
            st.vir[2] += drz * fz * wgt;
            const bool self = j >= self0 && j < self1;
            if (!self || N3 || j > i) {
              if (N3 || self) {
                ++st.kN3;
                ++st.gN3;
              } else {
                ++st.kNoN3;
                ++st.gNoN3;
              }
            }
          }
        });
    fxa = lcWarpSum(fxa);
    fya = lcWarpSum(fya);
    fza = lcWarpSum(fza);
    if (lane == 0) {
      if (N3) {
        atomicAdd(&a.fx[i], fxa);
        atomicAdd(&a.fy[i], fya);
        atomicAdd(&a.fz[i], fza);
      } else {
        a.fx[i] += fxa;
        a.fy[i] += fya;
        a.fz[i] += fza;
      }
    }
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

// ------------------------------------------------------------------------------------------------------------------
// VerletClusterLists, list-faithful traversal: every listed cluster pair costs M x M distance evaluations
// (VCLClusterFunctor.h:38-96). One thread per slot; the M lanes of a cluster walk the cluster's list together and read
// the neighbour cluster's particles through broadcast loads.
// ------------------------------------------------------------------------------------------------------------------
struct VCLArgs {
  int64_t n;
  int M;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type, *own;
  const int *clIsHalo, *nbrStart, *nbrList;
  LJParams p;
  LJStats *partials;
};

template <bool MIX, bool STATS, bool N3>
__global__ void __launch_bounds__(128) kLJClusterPairs(VCLArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats st;
  ljStatsZero(st);
  const bool inRange = i < a.n;
  const int ownI = inRange ? a.own[i] : APB_OWN_DUMMY;
  if (ownI != APB_OWN_DUMMY) {
    const int M = a.M;
    const int64_t A = i / M;
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    const double wI = ownI == APB_OWN_OWNED ? 1. : 0.;
    double fxa = 0., fya = 0., fza = 0.;
    // (1) inside the cluster: SoAFunctorSingle, skipped for halo clusters (VCLClusterFunctor.h:39-41).
    //     Both directions are evaluated (f_ji = -f_ij bit for bit), counted once as the reference's newton3 pair.
    if (!a.clIsHalo[A]) {
      const int64_t s0 = A * M;
      for (int k = 0; k < M; ++k) {
        const int64_t j = s0 + k;
        if (j == i) continue;
        const int ownJ = a.own[j];
        if (ownJ == APB_OWN_DUMMY) continue;
        const double drx = xi - a.x[j], dry = yi - a.y[j], drz = zi - a.z[j];
        const double dr2 = ljDist2(drx, dry, drz);
        const bool hit = dr2 <= a.p.cutoff2;
        const bool countThis = STATS && j > i;
        if (countThis) ++st.dist;
        if (hit) {
          double upot6;
          const double fac = ljEval<MIX>(a.p, dr2, ti, MIX ? a.type[j] : 0, upot6);
          const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
          fxa += fx;
          fya += fy;
          fza += fz;
          if (STATS) {
            st.upot += upot6 * wI;
            st.vir[0] += drx * fx * wI;
            st.vir[1] += dry * fy * wI;
            st.vir[2] += drz * fz * wI;
            if (countThis) {
              ++st.kN3;
              ++st.gN3;
            }
          }
        }
      }
    }
    // (2) listed neighbour clusters: SoAFunctorPair(A, B, newton3)
    const int e0 = a.nbrStart[A], e1 = a.nbrStart[A + 1];
    for (int e = e0; e < e1; ++e) {
      const int64_t s0 = static_cast<int64_t>(a.nbrList[e]) * M;
      for (int k = 0; k < M; ++k) {
        const int64_t j = s0 + k;
        const int ownJ = a.own[j];
        if (ownJ == APB_OWN_DUMMY) continue;
        const double drx = xi - a.x[j], dry = yi - a.y[j], drz = zi - a.z[j];
        const double dr2 = ljDist2(drx, dry, drz);
        const bool hit = dr2 <= a.p.cutoff2;
        if (STATS) ++st.dist;
        if (hit) {
          double upot6;
          const double fac = ljEval<MIX>(a.p, dr2, ti, MIX ? a.type[j] : 0, upot6);
          const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
          fxa += fx;
          fya += fy;
          fza += fz;
          if (N3) {
            atomicAdd(&a.fx[j], -fx);
            atomicAdd(&a.fy[j], -fy);
            atomicAdd(&a.fz[j], -fz);
          }
          if (STATS) {
            const double wJ = ownJ == APB_OWN_OWNED ? 1. : 0.;
            const double w = N3 ? wI + wJ : wI;
            st.upot += upot6 * w;
            st.vir[0] += drx * fx * w;
            st.vir[1] += dry * fy * w;
            st.vir[2] += drz * fz * w;
            if (N3) {
              ++st.kN3;
              ++st.gN3;
            } else {
              ++st.kNoN3;
              ++st.gNoN3;
            }
          }
        }
      }
    }
    if (N3) {
      atomicAdd(&a.fx[i], fxa);
      atomicAdd(&a.fy[i], fya);
      atomicAdd(&a.fz[i], fza);
    } else {
      a.fx[i] += fxa;
      a.fy[i] += fya;
      a.fz[i] += fza;
    }
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

// ------------------------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------------------------
int apbPrepareLJParams(apb_handle h, const apb_functor *f, LJParams &p) {
  if (!(f->cutoff > 0.)) return h->fail(APB_ERR_INVALID_ARGUMENT, "functor cutoff must be > 0");
  // AutoPas::computeInteractions checks functor cutoff <= container cutoff (AutoPasImpl.h:118-130)
  if (f->cutoff > h->cfg.cutoff)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "functor cutoff exceeds the container's interaction cutoff");
  p.cutoff2 = f->cutoff * f->cutoff;
  p.eps24 = f->epsilon24;
  p.sigma2 = f->sigma_squared;
  p.shift6 = (f->flags & APB_FUNCTOR_APPLY_SHIFT) ? apb_lj_calc_shift6(f->epsilon24, f->sigma_squared, p.cutoff2) : 0.;
  p.applyShift = (f->flags & APB_FUNCTOR_APPLY_SHIFT) ? 1 : 0;
  p.T = 0;
  p.mix = nullptr;
  p.mix4 = nullptr;
  {
    const double s6 = p.sigma2 * p.sigma2 * p.sigma2;
    p.k1 = 2. * p.eps24 * s6 * s6;
    p.k2 = -p.eps24 * s6;
    long long bits;
    std::memcpy(&bits, &p.cutoff2, 8);
    p.cutHiLo = static_cast<int>(bits >> 32) - 1;
  }
  if (f->flags & APB_FUNCTOR_USE_MIXING) {
    if (f->num_types <= 0 || !f->mixing_table)
      return h->fail(APB_ERR_INVALID_ARGUMENT, "mixing functor needs num_types > 0 and a mixing table");
    const size_t cnt = static_cast<size_t>(f->num_types) * f->num_types * 3;
    if (h->mixHostCache.size() != cnt || h->mixHostShift != (f->flags & APB_FUNCTOR_APPLY_SHIFT) ||
        std::memcmp(h->mixHostCache.data(), f->mixing_table, cnt * 8) != 0) {
      // device buffer: the caller's table | the derived {K1, K2, K1 / 2, shift6} table (32-byte entries)
      const size_t pairs = cnt / 3, off4 = (cnt * 8 + 31) & ~size_t(31);
      std::vector<double> derived(pairs * 4);
      for (size_t k = 0; k < pairs; ++k) {
        const double e24 = f->mixing_table[3 * k], s2 = f->mixing_table[3 * k + 1];
        const double s6 = s2 * s2 * s2;
        derived[4 * k] = 2. * e24 * s6 * s6;
        derived[4 * k + 1] = -e24 * s6;
        derived[4 * k + 2] = e24 * s6 * s6;
        derived[4 * k + 3] = (f->flags & APB_FUNCTOR_APPLY_SHIFT) ? f->mixing_table[3 * k + 2] : 0.;
      }
      APB_CHECK(apbEnsure(h, h->mixDev, off4 + pairs * 32));
      APB_CUDA(cudaMemcpyAsync(h->mixDev.p, f->mixing_table, cnt * 8, cudaMemcpyHostToDevice, h->stream));
      APB_CUDA(cudaMemcpyAsync(static_cast<char *>(h->mixDev.p) + off4, derived.data(), pairs * 32, cudaMemcpyHostToDevice, h->stream));
      APB_CUDA(cudaStreamSynchronize(h->stream));
      h->mixHostCache.assign(f->mixing_table, f->mixing_table + cnt);
      h->mixHostShift = f->flags & APB_FUNCTOR_APPLY_SHIFT;
    }
    p.T = f->num_types;
    p.mix = static_cast<const double *>(h->mixDev.p);
    p.mix4 = reinterpret_cast<const double *>(static_cast<const char *>(h->mixDev.p) + ((cnt * 8 + 31) & ~size_t(31)));
  }
  return APB_OK;
}

int apbFinishStats(apb_handle h, int numBlocks, bool stats, const apb_functor *f, apb_traversal_result *out) {
  apb_traversal_result host;
  std::memset(&host, 0, sizeof(host));
  if (h->asyncResultDev) {
    // device-resident loop (apb_run_steps): the reduced accumulators stay on the device, no host sync per step
    if (stats && numBlocks > 0) {
      APB_CHECK(apbReducePartials(h, numBlocks, h->asyncResultDev));
    } else {
      APB_CUDA(cudaMemsetAsync(h->asyncResultDev, 0, sizeof(apb_traversal_result), h->stream));
    }
    return APB_OK;
  }
  if (stats && numBlocks > 0) {
    APB_CHECK(apbReducePartials(h, numBlocks, static_cast<apb_traversal_result *>(h->result.p)));
    APB_CUDA(cudaMemcpyAsync(&host, h->result.p, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
  }
  APB_CUDA(cudaStreamSynchronize(h->stream));
  apbMaskResultByFlags(host, f->flags);
  if (out) *out = host;
  return APB_OK;
}

#define LJ_DISPATCH(KERNEL, grid, block, args)                                                     \
  do {                                                                                             \
    const int sel = (mix ? 4 : 0) | (stats ? 2 : 0) | (n3 ? 1 : 0);                                \
    switch (sel) {                                                                                 \
      case 0: ++h->launchCount, KERNEL<false, false, false><<<grid, block, 0, h->stream>>>(args); break;             \
      case 1: ++h->launchCount, KERNEL<false, false, true><<<grid, block, 0, h->stream>>>(args); break;              \
      case 2: ++h->launchCount, KERNEL<false, true, false><<<grid, block, 0, h->stream>>>(args); break;              \
      case 3: ++h->launchCount, KERNEL<false, true, true><<<grid, block, 0, h->stream>>>(args); break;               \
      case 4: ++h->launchCount, KERNEL<true, false, false><<<grid, block, 0, h->stream>>>(args); break;              \
      case 5: ++h->launchCount, KERNEL<true, false, true><<<grid, block, 0, h->stream>>>(args); break;               \
      case 6: ++h->launchCount, KERNEL<true, true, false><<<grid, block, 0, h->stream>>>(args); break;               \
      default: ++h->launchCount, KERNEL<true, true, true><<<grid, block, 0, h->stream>>>(args); break;               \
    }                                                                                              \
  } while (0)

int apbComputeLJPruned(apb_handle h, const apb_functor *f, const LJParams &p, bool mix, bool stats, bool n3,
                       apb_traversal_result *out);

int apbComputeLJ(apb_handle h, int traversal, const apb_functor *f, int newton3, apb_traversal_result *out) {
  LJParams p;
  APB_CHECK(apbPrepareLJParams(h, f, p));
  const bool mix = f->flags & APB_FUNCTOR_USE_MIXING;
  const bool stats = f->flags & (APB_FUNCTOR_CALC_GLOBALS | APB_FUNCTOR_COUNT_FLOPS);
  const bool n3 = newton3 != 0;
  const int64_t n = h->nslots;
  if (traversal == APB_TRAVERSAL_GPUVCL_PRUNED) return apbComputeLJPruned(h, f, p, mix, stats, n3, out);
  const int block = 128;
  int grid = apbDivUp(n, block);
  if (n == 0) return apbFinishStats(h, 0, stats, f, out);
  // gpuLinkedCells kernel variant: one warp per slot while the system is too small to fill the GPU with one thread per
  // slot (3x faster at 10 k slots), else one thread per slot; APB_LC_KERNEL=thread|warp overrides (A/B runs,
  // profiles/r02_lc_kernels.txt)
  const int lcKernel = apbLCKernelVariant(n, 0);
  if (h->cfg.container == APB_CONTAINER_LINKED_CELLS && lcKernel == 1)
    grid = static_cast<int>(std::min<int64_t>(apbDivUp(n, LCW_WARPS), 148 * 16));
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * grid));
  if (h->cfg.container == APB_CONTAINER_LINKED_CELLS) {
    LCArgs a;
    a.g = h->lc;
    a.n = n;
    a.x = h->col[APB_COL_X];
    a.y = h->col[APB_COL_Y];
    a.z = h->col[APB_COL_Z];
    a.fx = h->col[APB_COL_FX];
    a.fy = h->col[APB_COL_FY];
    a.fz = h->col[APB_COL_FZ];
    a.type = h->type;
    a.own = h->own;
    a.slotCell = static_cast<const int *>(h->slotCell.p);
    a.cellStart = static_cast<const int *>(h->start.p);
    a.stencil = static_cast<const int *>(h->stencilDev.p);
    a.stencilN = h->stencilN;
    // with FLOP counting requested reproduce the reference's full evaluation set (incl. forces on halo-cell
    // particles, which are discarded); otherwise skip that wasted work
    a.processHaloCells = (f->flags & APB_FUNCTOR_COUNT_FLOPS) ? 1 : 0;
    a.p = p;
    a.partials = static_cast<LJStats *>(h->partials.p);
    if (lcKernel != 1) {
      LJ_DISPATCH(kLJLinkedCells, grid, block, a);
    } else {
      LCWarpGeom w;
      w.g = h->lc;
      w.cellStart = a.cellStart;
      w.stencilSorted = a.stencil + 3 * APB_MAX_STENCIL;
      w.stencilN = h->stencilN;
#define LJ_LCW_ARGS a, w
      LJ_DISPATCH(kLJLinkedCellsWarp, grid, LCW_WARPS * 32, LJ_LCW_ARGS);
    }
  } else {
    VCLArgs a;
    a.n = n;
    a.M = h->cfg.cluster_size;
    a.x = h->col[APB_COL_X];
    a.y = h->col[APB_COL_Y];
    a.z = h->col[APB_COL_Z];
    a.fx = h->col[APB_COL_FX];
    a.fy = h->col[APB_COL_FY];
    a.fz = h->col[APB_COL_FZ];
    a.type = h->type;
    a.own = h->own;
    a.clIsHalo = static_cast<const int *>(h->clIsHalo.p);
    a.nbrStart = static_cast<const int *>(h->nbrStart.p);
    a.nbrList = static_cast<const int *>(h->nbrList.p);
    a.p = p;
    a.partials = static_cast<LJStats *>(h->partials.p);
    LJ_DISPATCH(kLJClusterPairs, grid, block, a);
  }
  APB_CUDA(cudaGetLastError());
  return apbFinishStats(h, grid, stats, f, out);
}

int apbCheckTraversal(apb_handle h, int traversal, int newton3);
int apbComputeOtherFunctor(apb_handle h, const apb_functor *f, int newton3, apb_traversal_result *out);

extern "C" int apb_compute_interactions(apb_handle h, int32_t traversal, const apb_functor *functor, int32_t newton3,
                                        apb_traversal_result *out) {
  APB_ENTRY(h);
  if (!functor) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_compute_interactions: null functor");
  APB_CHECK(apbCheckTraversal(h, traversal, newton3));
  if (!h->structureValid)
    return h->fail(APB_ERR_STATE, "apb_compute_interactions: particles were added / removed since the last "
                                  "apb_rebuild_neighbor_lists");
  // (gpuvcl_pruned refines the newton3-off cluster-pair list in both modes and keeps track of its own lists)
  if (h->cfg.container == APB_CONTAINER_VERLET_CLUSTER_LISTS &&
      h->builtNewton3 != (traversal == APB_TRAVERSAL_GPUVCL_PRUNED ? 0 : (newton3 ? 1 : 0)))
    return h->fail(APB_ERR_STATE, "apb_compute_interactions: cluster-pair lists were built for the other newton3 mode "
                                  "(VerletClusterListsRebuilder.h:153-163: list contents depend on newton3)");
  switch (functor->kind) {
    case APB_FUNCTOR_LJ:
      if (h->cfg.particle_kind != APB_PARTICLE_LJ)
        return h->fail(APB_ERR_NOT_APPLICABLE, "LJFunctor needs MoleculeLJ particles");
      return apbComputeLJ(h, traversal, functor, newton3, out);
    default:
      return apbComputeOtherFunctor(h, functor, newton3, out);
  }
}
