// Structure (re)builds on the device:
//   * LinkedCells: cell binning as a stable counting sort (replaces LinkedCells::updateContainer re-binning,
//     containers/linkedCells/LinkedCells.h:152-202, and CellBlock3D::get1DIndexOfPosition, CellBlock3D.h:321-349)
//   * VerletClusterLists: towers, z-sorted clusters, cluster AABBs and the cluster-pair neighbour list
//     (containers/verletClusterLists/VerletClusterListsRebuilder.h:67-143, 237-256, 366-409; ClusterTower.h:83-143;
//     Cluster.h:135-149; utils/ArrayMath.h:697-707)
// All kernels here are HBM-bound integer / compare work; decisions that must be bit-exact use explicit
// round-to-nearest intrinsics so that no FMA contraction can change a `<=` outcome.
#include <algorithm>
#include <array>
#include <vector>
#include <cmath>

#include "internal.cuh"

// ------------------------------------------------------------------------------------------------------------------
// host geometry
// ------------------------------------------------------------------------------------------------------------------
void apbComputeLCGeom(const apb_config &cfg, LCGeom &g) {
  // CellBlock3D::rebuild (CellBlock3D.h:360-426), same expressions in the same order
  const double il = cfg.cutoff + cfg.skin;
  const double csf = cfg.cell_size_factor;
  g.cellsPerInteractionLength = csf >= 1.0 ? 1 : static_cast<int>(std::ceil(1.0 / csf));
  g.numCells = 1;
  for (int d = 0; d < 3; ++d) {
    g.boxMin[d] = cfg.box_min[d];
    g.boxMax[d] = cfg.box_max[d];
    const double boxLength = g.boxMax[d] - g.boxMin[d];
    const unsigned long cellsPerDim =
        std::max(static_cast<unsigned long>(std::floor(boxLength / (il * csf))), 1ul);
    g.cellsPerDim[d] = static_cast<int>(cellsPerDim + 2 * g.cellsPerInteractionLength);
    g.cellLength[d] = boxLength / static_cast<double>(cellsPerDim);
    g.cellLengthReciprocal[d] = static_cast<double>(cellsPerDim) / boxLength;
    g.haloBoxMin[d] = g.boxMin[d] - g.cellsPerInteractionLength * g.cellLength[d];
    g.haloBoxMax[d] = g.boxMax[d] + g.cellsPerInteractionLength * g.cellLength[d];
    g.numCells *= g.cellsPerDim[d];
  }
}

int apbComputeStencil(apb_handle h) {
  // Pair set of lc_c08 / lc_c18: cells whose index offset lies within the overlap and whose border distance is at most
  // the interaction length (LCC08CellHandlerUtility.cpp:67-159, filter at :124; LCC18Traversal.h:115-211).
  const LCGeom &g = h->lc;
  const double il = h->cfg.cutoff + h->cfg.skin;
  const double il2 = il * il;
  int ov[3];
  for (int d = 0; d < 3; ++d) ov[d] = static_cast<int>(std::ceil(il / g.cellLength[d]));
  h->stencilN = 0;
  auto push = [&](int x, int y, int z) {
    if (h->stencilN >= APB_MAX_STENCIL) return false;
    h->stencil[h->stencilN][0] = x;
    h->stencil[h->stencilN][1] = y;
    h->stencil[h->stencilN][2] = z;
    ++h->stencilN;
    return true;
  };
  push(0, 0, 0);
  for (int z = -ov[2]; z <= ov[2]; ++z)
    for (int y = -ov[1]; y <= ov[1]; ++y)
      for (int x = -ov[0]; x <= ov[0]; ++x) {
        if (x == 0 && y == 0 && z == 0) continue;
        const double dv[3] = {std::max(0, std::abs(x) - 1) * g.cellLength[0], std::max(0, std::abs(y) - 1) * g.cellLength[1],
                              std::max(0, std::abs(z) - 1) * g.cellLength[2]};
        const double d2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
        if (d2 <= il2) {
          if (!push(x, y, z))
            return h->fail(APB_ERR_NOT_APPLICABLE, "cell size factor too small: neighbour stencil exceeds " +
                                                       std::to_string(APB_MAX_STENCIL) + " cells");
        }
      }
  // behind the list (self first): the same cells in (z, y, x) order, for the warp-cooperative walk (lc_warp.cuh), which
  // merges cells adjacent in x into one slot range
  std::vector<std::array<int, 3>> sorted(h->stencilN);
  for (int s = 0; s < h->stencilN; ++s) sorted[s] = {h->stencil[s][2], h->stencil[s][1], h->stencil[s][0]};
  std::sort(sorted.begin(), sorted.end());
  std::vector<int> flat(3 * APB_MAX_STENCIL, 0);
  for (int s = 0; s < h->stencilN; ++s) {
    flat[3 * s] = sorted[s][2];
    flat[3 * s + 1] = sorted[s][1];
    flat[3 * s + 2] = sorted[s][0];
  }
  APB_CHECK(apbEnsure(h, h->stencilDev, sizeof(int) * 6 * APB_MAX_STENCIL));
  APB_CUDA(cudaMemcpy(h->stencilDev.p, h->stencil, sizeof(int) * 3 * h->stencilN, cudaMemcpyHostToDevice));
  APB_CUDA(cudaMemcpy(static_cast<int *>(h->stencilDev.p) + 3 * APB_MAX_STENCIL, flat.data(), sizeof(int) * 3 * APB_MAX_STENCIL,
                      cudaMemcpyHostToDevice));
  return APB_OK;
}

static void computeVCLGeom(const apb_config &cfg, int64_t numParticles, VCLGeom &g) {
  // ClusterTowerBlock2D ctor (:36-42), estimateOptimalGridSideLength (:140-168), resize (:89-114)
  const double il = cfg.cutoff + cfg.skin;
  g.interactionLength = il;
  g.interactionLengthSqr = il * il;
  g.clusterSize = cfg.cluster_size;
  for (int d = 0; d < 3; ++d) {
    g.boxMin[d] = cfg.box_min[d];
    g.boxMax[d] = cfg.box_max[d];
    g.haloBoxMin[d] = cfg.box_min[d] - il;
    g.haloBoxMax[d] = cfg.box_max[d] + il;
  }
  const double boxSize[3] = {g.boxMax[0] - g.boxMin[0], g.boxMax[1] - g.boxMin[1], g.boxMax[2] - g.boxMin[2]};
  if (numParticles == 0) {
    g.side[0] = boxSize[0];
    g.side[1] = boxSize[1];
    g.towersPerDim[0] = g.towersPerDim[1] = 3;
  } else {
    const double volume = boxSize[0] * boxSize[1] * boxSize[2];
    const double density = static_cast<double>(numParticles) / volume;
    const double optimalSideLength = std::cbrt(static_cast<double>(cfg.cluster_size) / density);
    for (int d = 0; d < 2; ++d) {
      const double numTowersOwned = std::ceil(boxSize[d] / optimalSideLength);
      const double sideNew = boxSize[d] / numTowersOwned;
      const double numTowers = numTowersOwned + std::ceil(il / sideNew) * 2.;
      g.side[d] = sideNew;
      g.towersPerDim[d] = static_cast<int>(static_cast<size_t>(numTowers));
    }
  }
  int ntpil = 0;
  for (int d = 0; d < 2; ++d) {
    g.sideReciprocal[d] = 1. / g.side[d];
    ntpil = std::max(ntpil, static_cast<int>(std::ceil(il / g.side[d])));
  }
  g.numTowersPerInteractionLength = ntpil;
  g.numTowers = g.towersPerDim[0] * g.towersPerDim[1];
}

// ------------------------------------------------------------------------------------------------------------------
// counting sort pieces
// ------------------------------------------------------------------------------------------------------------------
__global__ void kKeysLC(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                        const double *__restrict__ z, const int32_t *__restrict__ own, LCGeom g, int *__restrict__ key,
                        int *__restrict__ rank, int *__restrict__ count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = -1;
  if (own[i] != APB_OWN_DUMMY) {
    k = apbCellIndexLC(g, x[i], y[i], z[i]);
    rank[i] = atomicAdd(&count[k], 1);
  }
  key[i] = k;
}


__global__ void kPadCounts(int64_t n, const int *__restrict__ count, int *__restrict__ padded, int M, int *maxCount) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int c = 0;
  if (i < n) {
    c = count[i];
    padded[i] = (c + M - 1) / M * M;
  }
  // block max -> one atomic
  for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
  if ((threadIdx.x & 31) == 0 && c > 0) atomicMax(maxCount, c);
}

__global__ void kScatterPerm(int64_t n, const int *__restrict__ key, const int *__restrict__ rank,
                             const int *__restrict__ start, int *__restrict__ perm) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = key[i];
  if (k >= 0) perm[start[k] + rank[i]] = static_cast<int>(i);
}

// Block-per-segment bitonic sort of perm[start[s] .. start[s]+count[s]).
// MODE 0: ascending source slot (makes the counting sort stable, i.e. deterministic and history preserving)
// MODE 1: ascending (z, id) — the canonical order used for clusters; the reference sorts by z only with an unstable
//         std::sort (cells/FullParticleCell.h:260-263), so ties there are implementation-defined.
template <int MODE>
__global__ void kSegSort(const int *__restrict__ start, const int *__restrict__ count, int *__restrict__ perm,
                         const double *__restrict__ z, const int64_t *__restrict__ id, int useGlobal, double *gK1,
                         long long *gK2, int *gV) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int seg = blockIdx.x;
  const int n = count[seg];
  if (n <= 1) return;
  const int s0 = start[seg];
  int P = 1;
  while (P < n) P <<= 1;
  double *k1;
  long long *k2;
  int *v;
  if (useGlobal) {
    k1 = gK1 + 2 * static_cast<size_t>(s0);
    k2 = gK2 + 2 * static_cast<size_t>(s0);
    v = gV + 2 * static_cast<size_t>(s0);
  } else {
    if (MODE == 1) {
      k1 = reinterpret_cast<double *>(smemRaw);
      k2 = reinterpret_cast<long long *>(k1 + P);
      v = reinterpret_cast<int *>(k2 + P);
    } else {
      k1 = nullptr;
      k2 = nullptr;
      v = reinterpret_cast<int *>(smemRaw);
    }
  }
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    if (t < n) {
      const int p = perm[s0 + t];
      v[t] = p;
      if (MODE == 1) {
        k1[t] = z[p];
        k2[t] = id[p];
      }
    } else {
      v[t] = 0x7fffffff;
      if (MODE == 1) {
        k1[t] = INFINITY;
        k2[t] = 0x7fffffffffffffffLL;
      }
    }
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += blockDim.x) {
        const int u = t ^ j;
        if (u > t) {
          const bool asc = (t & k) == 0;
          bool gt;  // element t > element u
          if (MODE == 1) {
            const double a = k1[t], b = k1[u];
            gt = a > b || (a == b && (k2[t] > k2[u] || (k2[t] == k2[u] && v[t] > v[u])));
          } else {
            gt = v[t] > v[u];
          }
          if (gt == asc) {
            const int tv = v[t];
            v[t] = v[u];
            v[u] = tv;
            if (MODE == 1) {
              const double ta = k1[t];
              k1[t] = k1[u];
              k1[u] = ta;
              const long long tb = k2[t];
              k2[t] = k2[u];
              k2[u] = tb;
            }
          }
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < n; t += blockDim.x) perm[s0 + t] = v[t];
}

__global__ void kSlotCell(int64_t m, const int *__restrict__ perm, const int *__restrict__ key, int *__restrict__ slotCell) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int p = perm[q];
  slotCell[q] = p >= 0 ? key[p] : -1;
}

#define SEGSORT_SMEM_LIMIT (200 * 1024)
int apbInitBuildAttributes(apb_handle h) {
  APB_CUDA(cudaFuncSetAttribute(kSegSort<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEGSORT_SMEM_LIMIT));
  APB_CUDA(cudaFuncSetAttribute(kSegSort<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEGSORT_SMEM_LIMIT));
  return APB_OK;
}

static int segSort(apb_handle h, int mode, int64_t numSeg, int maxCount, const int *start, const int *count, int *perm) {
  if (maxCount <= 1 || numSeg == 0) return APB_OK;
  int P = 1;
  while (P < maxCount) P <<= 1;
  const size_t elem = mode == 1 ? 20 : 4;
  const size_t smem = static_cast<size_t>(P) * elem;
  const size_t smemLimit = SEGSORT_SMEM_LIMIT;
  int useGlobal = smem > smemLimit;
  int threads = P / 2;
  threads = std::max(32, std::min(1024, threads));
  double *gK1 = nullptr;
  long long *gK2 = nullptr;
  int *gV = nullptr;
  if (useGlobal) {
    const size_t total = 2 * static_cast<size_t>(h->nslots > 0 ? h->nslots : 1) + 2 * static_cast<size_t>(P);
    APB_CHECK(apbEnsure(h, h->sortK1, total * 8));
    APB_CHECK(apbEnsure(h, h->sortK2, total * 8));
    APB_CHECK(apbEnsure(h, h->sortV, total * 4));
    gK1 = static_cast<double *>(h->sortK1.p);
    gK2 = static_cast<long long *>(h->sortK2.p);
    gV = static_cast<int *>(h->sortV.p);
  }
  const size_t dyn = useGlobal ? 0 : smem;
  if (mode == 1) {
    ++h->launchCount, kSegSort<1><<<static_cast<unsigned>(numSeg), threads, dyn, h->stream>>>(start, count, perm, h->col[APB_COL_Z], h->id,
                                                                           useGlobal, gK1, gK2, gV);
  } else {
    ++h->launchCount, kSegSort<0><<<static_cast<unsigned>(numSeg), threads, dyn, h->stream>>>(start, count, perm, nullptr, nullptr,
                                                                           useGlobal, gK1, gK2, gV);
  }
  APB_CUDA(cudaGetLastError());
  return APB_OK;
}

// ---- tower sort with z bins (VerletClusterLists) --------------------------------------------------------------------
// Sorting every tower by (z, id) is done as a counting sort on the finer key (tower, z bin) followed by an all-pairs
// rank sort inside each bin (a dozen particles): O(N * bin size) comparisons instead of a bitonic network over the whole
// tower. z bins are monotone in z, so bin order followed by in-bin order is the total (z, id) order the reference's
// sorted towers have (ClusterTower.h:83-100) with the canonical tie-break.
__global__ void kKeysVCLBins(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                             const double *__restrict__ z, const int32_t *__restrict__ own, VCLGeom g, int ZB,
                             double zScale, int *__restrict__ key, int *__restrict__ rank, int *__restrict__ binCount) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int k = -1;
  if (i < n && own[i] != APB_OWN_DUMMY) {
    const double px = x[i], py = y[i], pz = z[i];
    // VerletClusterListsRebuilder::sortParticlesIntoTowers (:217-232): only particles inside the halo box are kept
    const bool in = px >= g.haloBoxMin[0] && px < g.haloBoxMax[0] && py >= g.haloBoxMin[1] && py < g.haloBoxMax[1] &&
                    pz >= g.haloBoxMin[2] && pz < g.haloBoxMax[2];
    if (in) {
      int zb = static_cast<int>((pz - g.haloBoxMin[2]) * zScale);
      zb = zb < 0 ? 0 : (zb > ZB - 1 ? ZB - 1 : zb);
      k = apbTowerIndex(g, px, py) * ZB + zb;
    }
  }
  // warp-aggregated arrival rank: storage is tower-major from the previous build, so most lanes of a warp share a key
  // and one atomic per group replaces up to 32 colliding ones
  const unsigned grp = __match_any_sync(0xffffffffu, k);
  const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
  int base = 0;
  if (lane == leader && k >= 0) base = atomicAdd(&binCount[k], __popc(grp));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (i < n) {
    key[i] = k;
    if (k >= 0) rank[i] = base + __popc(grp & ((1u << lane) - 1u));
  }
}

// one warp per tower: exclusive prefix of its bins (in place), particle count, M-padded count, global maximum
__global__ void kTowerBins(int numTowers, int ZB, int M, int *__restrict__ binCount, int *__restrict__ binOff,
                           int *__restrict__ count, int *__restrict__ padded, int *maxCount) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= numTowers) return;
  int run = 0;
  for (int b0 = 0; b0 < ZB; b0 += 32) {
    const int b = b0 + lane;
    const int c = b < ZB ? binCount[t * ZB + b] : 0;
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (b < ZB) binOff[t * ZB + b] = run + incl - c;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    count[t] = run;
    padded[t] = (run + M - 1) / M * M;
    if (run > 0) atomicMax(maxCount, run);
  }
}

__global__ void kScatterPermBins(int64_t n, int ZB, const int *__restrict__ key, const int *__restrict__ rank,
                                 const int *__restrict__ start, const int *__restrict__ binOff, int *__restrict__ perm) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = key[i];
  if (k >= 0) perm[start[k / ZB] + binOff[k] + rank[i]] = static_cast<int>(i);
}

// BIN_G lanes per (tower, z bin) - a bin holds about a dozen particles, so two bins share a warp: rank of every member
// among the members by (z, id, slot); permOut receives the sorted bin. The loops run for the larger of the two bins (the
// shuffles need the whole warp); a lane without a member contributes nothing.
#define BIN_G 16
__global__ void kBinSort(int numBins, int ZB, const int *__restrict__ start, const int *__restrict__ binOff,
                         const int *__restrict__ binCount, const int *__restrict__ permIn, int *__restrict__ permOut,
                         const double *__restrict__ z, const int64_t *__restrict__ id) {
  const int bin = (blockIdx.x * blockDim.x + threadIdx.x) / BIN_G, lane = threadIdx.x & (BIN_G - 1);
  const int cnt = bin < numBins ? binCount[bin] : 0;
  int cntMax = cnt;
  for (int w = BIN_G; w < 32; w <<= 1) cntMax = max(cntMax, __shfl_xor_sync(0xffffffffu, cntMax, w));
  if (cntMax == 0) return;  // warp-uniform
  const int b0 = cnt > 0 ? start[bin / ZB] + binOff[bin] : 0;
  for (int ci = 0; ci < cntMax; ci += BIN_G) {
    const bool mine = ci + lane < cnt;
    const int p = mine ? permIn[b0 + ci + lane] : 0;
    const double zi = mine ? z[p] : 0.;
    // Fast pass: rank by z alone (two shuffles and one comparison per member) and note whether any two members of the
    // bin share their z. Only then - a lattice at rest; practically never once the particles move - the full
    // (z, id, slot) comparison below decides, as ClusterTower's sort with the canonical tie-break does.
    int r = 0;
    bool tie = false;
    for (int cj = 0; cj < cntMax; cj += BIN_G) {
      const bool have = cj + lane < cnt;
      const double zj = have ? (cj == ci ? zi : z[permIn[b0 + cj + lane]]) : 0.;
      const int m = min(BIN_G, cnt - cj);
#pragma unroll 4
      for (int k = 0; k < BIN_G; ++k) {
        const double zk = __shfl_sync(0xffffffffu, zj, k, BIN_G);
        r += k < m && zk < zi;
        tie |= mine && k < m && zk == zi && !(cj == ci && k == lane);
      }
    }
    if (__any_sync(0xffffffffu, tie)) {
      const long long idi = mine ? id[p] : 0;
      r = 0;
      for (int cj = 0; cj < cntMax; cj += BIN_G) {
        const bool have = cj + lane < cnt;
        const int pj = have ? (cj == ci ? p : permIn[b0 + cj + lane]) : 0;
        const double zj = have ? (cj == ci ? zi : z[pj]) : 0.;
        const long long idj = have ? (cj == ci ? idi : id[pj]) : 0;
        const int m = min(BIN_G, cnt - cj);  // members of this group's bin in the chunk (<= 0: none)
#pragma unroll 4
        for (int k = 0; k < BIN_G; ++k) {
          const double zk = __shfl_sync(0xffffffffu, zj, k, BIN_G);
          const long long idk = __shfl_sync(0xffffffffu, idj, k, BIN_G);
          const int pk = __shfl_sync(0xffffffffu, pj, k, BIN_G);
          r += k < m && (zk < zi || (zk == zi && (idk < idi || (idk == idi && pk < p))));
        }
      }
    }
    if (mine) permOut[b0 + r] = p;
  }
}

// scratch words behind the result struct: [0..1] scan totals (int64), [8] max count (int32)
static long long *scratchTotals(apb_handle h) {
  return reinterpret_cast<long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result));
}
static int *scratchMax(apb_handle h) {
  return reinterpret_cast<int *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 32);
}

// ------------------------------------------------------------------------------------------------------------------
// LinkedCells rebuild
// ------------------------------------------------------------------------------------------------------------------
int apbRebuildLinkedCells(apb_handle h) {
  if (!h->ownedInsideBox) h->ownedKnown = false;  // (see apbRebuildVCL)
  h->countsTrusted = false;
  const int64_t n = h->nslots;
  const int64_t nc = h->lc.numCells;
  APB_CHECK(apbEnsure(h, h->key, sizeof(int) * std::max<int64_t>(n, 1)));
  APB_CHECK(apbEnsure(h, h->rank, sizeof(int) * std::max<int64_t>(n, 1)));
  APB_CHECK(apbEnsure(h, h->perm, sizeof(int) * std::max<int64_t>(n, 1)));
  APB_CHECK(apbEnsure(h, h->slotCell, sizeof(int) * std::max<int64_t>(n, 1)));
  APB_CHECK(apbEnsure(h, h->count, sizeof(int) * (nc + 1)));
  APB_CHECK(apbEnsure(h, h->start, sizeof(int) * (nc + 1)));
  int *key = static_cast<int *>(h->key.p), *rank = static_cast<int *>(h->rank.p), *perm = static_cast<int *>(h->perm.p);
  int *count = static_cast<int *>(h->count.p), *start = static_cast<int *>(h->start.p);
  APB_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (nc + 1), h->stream));
  APB_CUDA(cudaMemsetAsync(scratchMax(h), 0, sizeof(int), h->stream));
  if (n > 0) {
    ++h->launchCount, kKeysLC<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->own,
                                                     h->lc, key, rank, count);
    APB_CUDA(cudaGetLastError());
  }
  // reuse kPadCounts with M = 1 only for the max
  APB_CHECK(apbEnsure(h, h->nbrCount, sizeof(int) * (nc + 1)));
  ++h->launchCount, kPadCounts<<<apbDivUp(nc + 1, 256), 256, 0, h->stream>>>(nc + 1, count, static_cast<int *>(h->nbrCount.p), 1, scratchMax(h));
  APB_CHECK(apbExclusiveScan(h, count, start, nc + 1, scratchTotals(h)));
  long long total = 0;
  int maxCount = 0;
  APB_CUDA(cudaMemcpyAsync(&total, scratchTotals(h), 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(&maxCount, scratchMax(h), 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (n > 0) {
    ++h->launchCount, kScatterPerm<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, key, rank, start, perm);
    APB_CUDA(cudaGetLastError());
    APB_CHECK(segSort(h, 0, nc, maxCount, start, count, perm));
    if (total > 0) {
      ++h->launchCount, kSlotCell<<<apbDivUp(total, 256), 256, 0, h->stream>>>(total, perm, key, static_cast<int *>(h->slotCell.p));
      APB_CUDA(cudaGetLastError());
    }
  }
  APB_CHECK(apbRemapHaloLinks(h, perm, n, total));
  APB_CHECK(apbPermuteStorage(h, perm, total));
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  h->numCells = nc;
  h->structureValid = true;
  ++h->structureVersion;
  h->countsValid = false;
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// VerletClusterLists rebuild
// ------------------------------------------------------------------------------------------------------------------
// Padding dummies are parked like ClusterTower::setDummyValues (ClusterTower.h:152-161) with the arguments of
// VerletClusterListsRebuilder::rebuildNeighborListsAndFillClusters (:157-162).
__global__ void kParkDummies(int64_t m, const int32_t *__restrict__ own, const int *__restrict__ slotTower,
                             const int *__restrict__ start, const int *__restrict__ padded, double *x, double *y,
                             double *z, double startX, double dist) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m || own[q] != APB_OWN_DUMMY) return;
  const int t = slotTower[q];
  const int index = start[t] + padded[t] - static_cast<int>(q);  // 1 for the last slot of the tower
  x[q] = startX + static_cast<double>(t) * dist;
  y[q] = 0.;
  z[q] = dist * static_cast<double>(index);
}

// per slot tower for padded storage: towers are contiguous slot ranges
__global__ void kSlotTower(int numTowers, const int *__restrict__ start, const int *__restrict__ padded,
                           int *__restrict__ slotTower) {
  const int t = blockIdx.x;
  if (t >= numTowers) return;
  const int s0 = start[t], np = padded[t];
  for (int k = threadIdx.x; k < np; k += blockDim.x) slotTower[s0 + k] = t;
}

// M lanes per cluster (32 / M clusters per warp): bounding box over the ACTUAL members. The reference computes the box
// while the padding dummies sit on the last actual particle (ClusterTower.h:170-180), which gives the same box: z from
// first / last member (sorted), x and y as min / max (Cluster.h:135-149). Every lane reads one slot (coalesced) and the
// group reduces with shuffles; one thread per cluster walked its 32 slots serially (0.64 ms at 16 M particles).
__global__ void kClusterBoxes(int64_t numClusters, int M, int logM, const double *__restrict__ x,
                              const double *__restrict__ y, const double *__restrict__ z, const int32_t *__restrict__ own,
                              const int *__restrict__ slotTower, double *__restrict__ bmin, double *__restrict__ bmax,
                              int *__restrict__ hasOwned, int *__restrict__ clTower) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // slot
  const int64_t c = s >> logM;
  const int k = static_cast<int>(s) & (M - 1);
  const bool in = c < numClusters;
  const int o = in ? own[s] : APB_OWN_DUMMY;
  const bool live = o != APB_OWN_DUMMY;  // dummies only trail; slot 0 of a cluster is always a particle
  double lx = live ? x[s] : 1e308, hx = live ? x[s] : -1e308;
  double ly = live ? y[s] : 1e308, hy = live ? y[s] : -1e308;
  double lz = live ? z[s] : 1e308, hz = live ? z[s] : -1e308;  // z-sorted: min = first, max = last live member
  int ownedAny = o == APB_OWN_OWNED;
  for (int w = 1; w < M; w <<= 1) {
    lx = fmin(lx, __shfl_xor_sync(0xffffffffu, lx, w));
    hx = fmax(hx, __shfl_xor_sync(0xffffffffu, hx, w));
    ly = fmin(ly, __shfl_xor_sync(0xffffffffu, ly, w));
    hy = fmax(hy, __shfl_xor_sync(0xffffffffu, hy, w));
    lz = fmin(lz, __shfl_xor_sync(0xffffffffu, lz, w));
    hz = fmax(hz, __shfl_xor_sync(0xffffffffu, hz, w));
    ownedAny |= __shfl_xor_sync(0xffffffffu, ownedAny, w);
  }
  if (in && k == 0) {
    bmin[c] = lx;
    bmin[numClusters + c] = ly;
    bmin[2 * numClusters + c] = lz;
    bmax[c] = hx;
    bmax[numClusters + c] = hy;
    bmax[2 * numClusters + c] = hz;
    hasOwned[c] = ownedAny;
    clTower[c] = slotTower[s];
  }
}

// ClusterTower::generateClusters (ClusterTower.h:109-141): [firstOwnedCluster, firstTailHaloCluster)
__global__ void kTowerRanges(int numTowers, int M, const int *__restrict__ start, const int *__restrict__ padded,
                             const int *__restrict__ hasOwned, int *__restrict__ twFirstCluster,
                             int *__restrict__ twNumClusters, int *__restrict__ twFirstOwned,
                             int *__restrict__ twFirstTail, int *__restrict__ clIsHalo) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= numTowers) return;
  const int c0 = start[t] / M, nc = padded[t] / M;
  int firstOwned = nc, firstTail = nc;
  bool foundOwned = false, foundTail = false;
  for (int c = 0; c < nc; ++c) {
    const bool containsOwned = !foundTail && hasOwned[c0 + c];
    if (!foundOwned && containsOwned) {
      firstOwned = c;
      foundOwned = true;
    }
    if (!foundTail && foundOwned && !containsOwned) {
      firstTail = c;
      foundTail = true;
    }
  }
  twFirstCluster[t] = c0;
  twNumClusters[t] = nc;
  twFirstOwned[t] = c0 + firstOwned;
  twFirstTail[t] = c0 + firstTail;
  for (int c = 0; c < nc; ++c) clIsHalo[c0 + c] = (c < firstOwned || c >= firstTail);
}

struct NbrArgs {
  VCLGeom g;
  int64_t numClusters;
  int newton3;
  const double *bmin, *bmax;
  const int *clTower, *clIsHalo, *twFirstCluster, *twNumClusters;
};

// utils::ArrayMath::boxDistanceSquared (ArrayMath.h:697-707): dot(aToB,aToB) + dot(bToA,bToA), each dot = (x*x+y*y)+z*z,
// every product and sum rounded separately (the oracle is compiled with -ffp-contract=off)
__device__ __forceinline__ double boxDist2(const NbrArgs &a, int64_t A, int64_t B) {
  const int64_t n = a.numClusters;
  double s1 = 0., s2 = 0.;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double aMin = a.bmin[d * n + A], aMax = a.bmax[d * n + A];
    const double bMin = a.bmin[d * n + B], bMax = a.bmax[d * n + B];
    const double aToB = fmax(0., __dsub_rn(aMin, bMax));
    const double bToA = fmax(0., __dsub_rn(bMin, aMax));
    s1 = d == 0 ? __dmul_rn(aToB, aToB) : __dadd_rn(s1, __dmul_rn(aToB, aToB));
    s2 = d == 0 ? __dmul_rn(bToA, bToA) : __dadd_rn(s2, __dmul_rn(bToA, bToA));
  }
  return __dadd_rn(s1, s2);
}

// VerletClusterListsRebuilder::get1DInteractionCellIndexForTower / isForwardNeighbor (:313-354)
__device__ __forceinline__ int interactionCell(const VCLGeom &g, int tx, int ty) {
  const int n = g.numTowersPerInteractionLength;
  const int numX = static_cast<int>(ceil(g.towersPerDim[0] / static_cast<double>(n)));
  return tx / n + numX * (ty / n);
}
__device__ __forceinline__ bool isForwardNeighbor(const VCLGeom &g, int tx, int ty, int nx, int ny) {
  const int ca = interactionCell(g, tx, ty), cb = interactionCell(g, nx, ny);
  if (cb > ca) return true;
  if (cb < ca) return false;
  return nx + ny * g.towersPerDim[0] >= tx + ty * g.towersPerDim[0];
}

// One thread per cluster A. FILL = false counts, FILL = true writes the list in the reference's iteration order
// (neighbour towers y-major then x, clusters ascending): updateNeighborLists (:237-256), iterateNeighborTowers
// (:279-306), calculateNeighborsBetweenTowers (:366-409).
template <bool FILL>
__global__ void kNeighborLists(NbrArgs a, int *__restrict__ nbrCount, const int *__restrict__ nbrStart,
                               int *__restrict__ nbrList) {
  const int64_t A = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (A >= a.numClusters) return;
  const VCLGeom &g = a.g;
  const bool haloA = a.clIsHalo[A];
  int cnt = 0;
  int *out = FILL ? nbrList + nbrStart[A] : nullptr;
  if (a.newton3 || !haloA) {
    const int t = a.clTower[A];
    const int tx = t % g.towersPerDim[0], ty = t / g.towersPerDim[0];
    const int n = g.numTowersPerInteractionLength;
    const int minX = max(tx - n, 0), minY = max(ty - n, 0);
    const int maxX = min(tx + n, g.towersPerDim[0] - 1), maxY = min(ty + n, g.towersPerDim[1] - 1);
    const int64_t nc = a.numClusters;
    const double aMinZ = a.bmin[2 * nc + A], aMaxZ = a.bmax[2 * nc + A];
    for (int ny = minY; ny <= maxY; ++ny) {
      const double distY = max(0, abs(ty - ny) - 1) * g.side[1];
      for (int nx = minX; nx <= maxX; ++nx) {
        if (a.newton3 && !isForwardNeighbor(g, tx, ty, nx, ny)) continue;
        const double distX = max(0, abs(tx - nx) - 1) * g.side[0];
        const double d2 = __dadd_rn(__dmul_rn(distX, distX), __dmul_rn(distY, distY));
        if (!(d2 <= g.interactionLengthSqr)) continue;
        const int tb = nx + ny * g.towersPerDim[0];
        const int b0 = a.twFirstCluster[tb], nb = a.twNumClusters[tb];
        if (nb == 0) continue;
        const bool sameTower = tb == t;
        int lo = (sameTower && a.newton3) ? static_cast<int>(A - b0) + 1 : 0;
        int hi = nb;
        // z window. Clusters of a tower are z-sorted, so their zmin and zmax are non-decreasing. A cluster whose z gap
        // alone exceeds the interaction length can never pass the box test (all other terms are >= 0 and rounding is
        // monotone), so narrowing [lo, hi) by binary search does not change the resulting set.
        {
          int l = lo, r = hi;  // first B with NOT (aMinZ - bMaxZ > 0 && (gap)^2 > il^2)
          while (l < r) {
            const int mid = (l + r) >> 1;
            const double gap = __dsub_rn(aMinZ, a.bmax[2 * nc + b0 + mid]);
            const bool tooLow = gap > 0. && __dmul_rn(gap, gap) > g.interactionLengthSqr;
            if (tooLow) l = mid + 1; else r = mid;
          }
          lo = l;
          l = lo;
          r = hi;  // first B with (bMinZ - aMaxZ > 0 && gap^2 > il^2)
          while (l < r) {
            const int mid = (l + r) >> 1;
            const double gap = __dsub_rn(a.bmin[2 * nc + b0 + mid], aMaxZ);
            const bool tooHigh = gap > 0. && __dmul_rn(gap, gap) > g.interactionLengthSqr;
            if (tooHigh) r = mid; else l = mid + 1;
          }
          hi = l;
        }
        for (int k = lo; k < hi; ++k) {
          const int64_t B = b0 + k;
          if (B == A) continue;
          if (haloA && a.clIsHalo[B]) continue;
          if (boxDist2(a, A, B) <= g.interactionLengthSqr) {
            if (FILL) out[cnt] = static_cast<int>(B);
            ++cnt;
          }
        }
      }
    }
  }
  if (!FILL) nbrCount[A] = cnt;
}

int apbRebuildVCL(apb_handle h, int newton3) {
  const int M = h->cfg.cluster_size;
  int64_t owned = 0, halo = 0;
  if (!h->countsTrusted) h->countsValid = false;  // trusted: set by the one-pass halo generation that just ran
  h->countsTrusted = false;
  if (!h->ownedInsideBox) h->ownedKnown = false;  // the sort drops particles outside the halo box, owned ones included
  APB_CHECK(apb_get_num_particles(h, &owned, &halo));
  computeVCLGeom(h->cfg, owned + halo, h->vcl);
  const VCLGeom &g = h->vcl;
  const int64_t n = h->nslots;
  const int64_t nt = g.numTowers;
  // z bins per tower: about a dozen particles per bin on average
  const double avgPerTower = static_cast<double>(owned + halo) / static_cast<double>(std::max<int64_t>(nt, 1));
  const int ZB = static_cast<int>(std::min(512.0, std::max(1.0, std::ceil(avgPerTower / 12.0))));
  const double zScale = static_cast<double>(ZB) / (g.haloBoxMax[2] - g.haloBoxMin[2]);
  const int64_t numBins = nt * ZB;
  if (numBins + 1 > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "too many towers");
  APB_CHECK(apbEnsure(h, h->key, sizeof(int) * std::max<int64_t>(n, 1)));
  APB_CHECK(apbEnsure(h, h->rank, sizeof(int) * std::max<int64_t>(std::max<int64_t>(n, nt + 1), 1)));
  APB_CHECK(apbEnsure(h, h->count, sizeof(int) * (nt + 1)));
  APB_CHECK(apbEnsure(h, h->start, sizeof(int) * (nt + 1)));
  APB_CHECK(apbEnsure(h, h->nbrCount, sizeof(int) * (nt + 1)));  // padded counts (temporarily)
  APB_CHECK(apbEnsure(h, h->sortK1, sizeof(int) * (numBins + 1)));  // bin counts
  APB_CHECK(apbEnsure(h, h->sortK2, sizeof(int) * (numBins + 1)));  // bin offsets inside the tower
  int *key = static_cast<int *>(h->key.p), *rank = static_cast<int *>(h->rank.p);
  int *count = static_cast<int *>(h->count.p), *start = static_cast<int *>(h->start.p);
  int *padded = static_cast<int *>(h->nbrCount.p);
  int *binCount = static_cast<int *>(h->sortK1.p), *binOff = static_cast<int *>(h->sortK2.p);
  APB_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (nt + 1), h->stream));
  APB_CUDA(cudaMemsetAsync(padded, 0, sizeof(int) * (nt + 1), h->stream));
  APB_CUDA(cudaMemsetAsync(binCount, 0, sizeof(int) * (numBins + 1), h->stream));
  APB_CUDA(cudaMemsetAsync(scratchMax(h), 0, sizeof(int), h->stream));
  if (n > 0) {
    ++h->launchCount, kKeysVCLBins<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z],
                                                          h->own, g, ZB, zScale, key, rank, binCount);
    APB_CUDA(cudaGetLastError());
  }
  ++h->launchCount, kTowerBins<<<apbDivUp(nt * 32, 256), 256, 0, h->stream>>>(static_cast<int>(nt), ZB, M, binCount, binOff, count,
                                                                  padded, scratchMax(h));
  APB_CHECK(apbExclusiveScan(h, padded, start, nt + 1, scratchTotals(h)));
  long long total = 0;
  int maxCount = 0;
  APB_CUDA(cudaMemcpyAsync(&total, scratchTotals(h), 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(&maxCount, scratchMax(h), 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  h->vclMaxTowerCount = maxCount;
  APB_CHECK(apbEnsure(h, h->perm, sizeof(int) * std::max<int64_t>(total, 1)));
  APB_CHECK(apbEnsure(h, h->sortV, sizeof(int) * std::max<int64_t>(total, 1)));
  APB_CHECK(apbEnsure(h, h->slotCell, sizeof(int) * std::max<int64_t>(total, 1)));
  int *perm = static_cast<int *>(h->perm.p), *permUnsorted = static_cast<int *>(h->sortV.p);
  int *slotTower = static_cast<int *>(h->slotCell.p);
  if (total > 0) APB_CUDA(cudaMemsetAsync(perm, 0xFF, sizeof(int) * total, h->stream));
  if (n > 0 && total > 0) {
    ++h->launchCount, kScatterPermBins<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, ZB, key, rank, start, binOff, permUnsorted);
    ++h->launchCount, kBinSort<<<apbDivUp(numBins * BIN_G, 256), 256, 0, h->stream>>>(static_cast<int>(numBins), ZB, start, binOff, binCount,
                                                                      permUnsorted, perm, h->col[APB_COL_Z], h->id);
    APB_CUDA(cudaGetLastError());
  }
  APB_CHECK(apbRemapHaloLinks(h, perm, n, total));
  APB_CHECK(apbPermuteStorage(h, perm, total));
  const int64_t numClusters = total / M;
  h->numClusters = numClusters;
  h->numPairs = 0;
  // keep padded counts: move them to their own buffer (nbrCount is reused below)
  APB_CHECK(apbEnsure(h, h->twNumClusters, sizeof(int) * (nt + 1)));
  APB_CHECK(apbEnsure(h, h->twFirstCluster, sizeof(int) * (nt + 1)));
  APB_CHECK(apbEnsure(h, h->twFirstOwned, sizeof(int) * (nt + 1)));
  APB_CHECK(apbEnsure(h, h->twFirstTailHalo, sizeof(int) * (nt + 1)));
  APB_CHECK(apbEnsure(h, h->rank, sizeof(int) * (nt + 1)));  // rank no longer needed: holds padded counts
  int *paddedKeep = static_cast<int *>(h->rank.p);
  APB_CUDA(cudaMemcpyAsync(paddedKeep, padded, sizeof(int) * (nt + 1), cudaMemcpyDeviceToDevice, h->stream));
  if (total > 0) {
    ++h->launchCount, kSlotTower<<<static_cast<unsigned>(nt), 128, 0, h->stream>>>(static_cast<int>(nt), start, paddedKeep, slotTower);
    const double dist = g.interactionLength * 2;
    const double startX = 1000 * g.haloBoxMax[0];
    ++h->launchCount, kParkDummies<<<apbDivUp(total, 256), 256, 0, h->stream>>>(total, h->own, slotTower, start, paddedKeep,
                                                              h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z],
                                                              startX, dist);
    APB_CUDA(cudaGetLastError());
  }
  const int64_t ncAlloc = std::max<int64_t>(numClusters, 1);
  APB_CHECK(apbEnsure(h, h->clBoxMin, sizeof(double) * 3 * ncAlloc));
  APB_CHECK(apbEnsure(h, h->clBoxMax, sizeof(double) * 3 * ncAlloc));
  APB_CHECK(apbEnsure(h, h->clHasOwned, sizeof(int) * ncAlloc));
  APB_CHECK(apbEnsure(h, h->clIsHalo, sizeof(int) * ncAlloc));
  APB_CHECK(apbEnsure(h, h->clTower, sizeof(int) * ncAlloc));
  APB_CHECK(apbEnsure(h, h->nbrCount, sizeof(int) * (ncAlloc + 1)));
  APB_CHECK(apbEnsure(h, h->nbrStart, sizeof(int) * (ncAlloc + 1)));
  if (numClusters > 0) {
    int logM = 0;
    while ((1 << logM) < M) ++logM;
    ++h->launchCount, kClusterBoxes<<<apbDivUp(numClusters * M, 256), 256, 0, h->stream>>>(
        numClusters, M, logM, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->own, slotTower,
        static_cast<double *>(h->clBoxMin.p), static_cast<double *>(h->clBoxMax.p), static_cast<int *>(h->clHasOwned.p),
        static_cast<int *>(h->clTower.p));
    APB_CUDA(cudaGetLastError());
  }
  ++h->launchCount, kTowerRanges<<<apbDivUp(nt, 128), 128, 0, h->stream>>>(
      static_cast<int>(nt), M, start, paddedKeep, static_cast<int *>(h->clHasOwned.p),
      static_cast<int *>(h->twFirstCluster.p), static_cast<int *>(h->twNumClusters.p),
      static_cast<int *>(h->twFirstOwned.p), static_cast<int *>(h->twFirstTailHalo.p), static_cast<int *>(h->clIsHalo.p));
  APB_CUDA(cudaGetLastError());
  // neighbour lists: count -> scan -> fill
  NbrArgs a;
  a.g = g;
  a.numClusters = numClusters;
  a.newton3 = newton3 ? 1 : 0;
  a.bmin = static_cast<double *>(h->clBoxMin.p);
  a.bmax = static_cast<double *>(h->clBoxMax.p);
  a.clTower = static_cast<int *>(h->clTower.p);
  a.clIsHalo = static_cast<int *>(h->clIsHalo.p);
  a.twFirstCluster = static_cast<int *>(h->twFirstCluster.p);
  a.twNumClusters = static_cast<int *>(h->twNumClusters.p);
  int *nbrCount = static_cast<int *>(h->nbrCount.p), *nbrStart = static_cast<int *>(h->nbrStart.p);
  long long numPairs = 0;
  APB_CUDA(cudaMemsetAsync(nbrCount, 0, sizeof(int) * (numClusters + 1), h->stream));
  if (numClusters > 0) {
    ++h->launchCount, kNeighborLists<false><<<apbDivUp(numClusters, 64), 64, 0, h->stream>>>(a, nbrCount, nullptr, nullptr);
    APB_CUDA(cudaGetLastError());
  }
  APB_CHECK(apbExclusiveScan(h, nbrCount, nbrStart, numClusters + 1, scratchTotals(h)));
  APB_CUDA(cudaMemcpyAsync(&numPairs, scratchTotals(h), 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (numPairs > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "cluster-pair list exceeds 2^31 entries");
  APB_CHECK(apbEnsure(h, h->nbrList, sizeof(int) * std::max<long long>(numPairs, 1)));
  if (numPairs > 0) {
    ++h->launchCount, kNeighborLists<true><<<apbDivUp(numClusters, 64), 64, 0, h->stream>>>(a, nbrCount, nbrStart,
                                                                          static_cast<int *>(h->nbrList.p));
    APB_CUDA(cudaGetLastError());
  }
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  h->numPairs = numPairs;
  h->numCells = nt;
  h->structureValid = true;
  ++h->structureVersion;
  h->builtNewton3 = newton3 ? 1 : 0;
  h->ownDirty = false;
  h->prunedValid = false;
  h->countsValid = false;
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
static bool traversalMatchesContainer(int container, int traversal) {
  if (container == APB_CONTAINER_LINKED_CELLS)
    return traversal == APB_TRAVERSAL_GPULC_C08 || traversal == APB_TRAVERSAL_GPULC_C18;
  return traversal >= APB_TRAVERSAL_GPUVCL_CLUSTER_ITERATION && traversal <= APB_TRAVERSAL_GPUVCL_PRUNED;
}

int apbCheckTraversal(apb_handle h, int traversal, int newton3) {
  if (!traversalMatchesContainer(h->cfg.container, traversal))
    return h->fail(APB_ERR_NOT_APPLICABLE, "traversal option is not compatible with this container "
                                           "(CompatibleTraversals.h: gpulc_* <-> gpuLinkedCells, gpuvcl_* <-> gpuVerletClusterLists)");
  // CompatibleTraversals.h:142-151: cluster_iteration and c01_balanced support newton3 off only
  if (newton3 && (traversal == APB_TRAVERSAL_GPUVCL_CLUSTER_ITERATION ||
                  traversal == APB_TRAVERSAL_GPUVCL_C01_BALANCED))
    return h->fail(APB_ERR_NOT_APPLICABLE, "this traversal supports newton3 = off only");
  return APB_OK;
}

extern "C" int apb_rebuild_neighbor_lists(apb_handle h, int32_t traversal, int32_t newton3) {
  APB_ENTRY(h);
  APB_CHECK(apbCheckTraversal(h, traversal, newton3));
  if (h->cfg.container == APB_CONTAINER_LINKED_CELLS) {
    APB_CHECK(apbRebuildLinkedCells(h));
    return apbSnapshotRebuildPositions(h);
  }
  // gpuvcl_pruned refines the full (newton3 off) cluster-pair list in both newton3 modes
  APB_CHECK(apbRebuildVCL(h, traversal == APB_TRAVERSAL_GPUVCL_PRUNED ? 0 : newton3));
  if (traversal == APB_TRAVERSAL_GPUVCL_PRUNED) APB_CHECK(apbBuildPruned(h, newton3));
  return apbSnapshotRebuildPositions(h);
}

extern "C" int apb_get_geometry(apb_handle h, apb_geometry *out) {
  APB_ENTRY(h);
  if (!out) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_geometry: null argument");
  std::memset(out, 0, sizeof(*out));
  out->interaction_length = h->cfg.cutoff + h->cfg.skin;
  out->num_slots = h->nslots;
  if (h->cfg.container == APB_CONTAINER_LINKED_CELLS) {
    for (int d = 0; d < 3; ++d) {
      out->cells_per_dim[d] = h->lc.cellsPerDim[d];
      out->cell_length[d] = h->lc.cellLength[d];
    }
    out->num_cells = h->lc.numCells;
  } else {
    out->cells_per_dim[0] = h->vcl.towersPerDim[0];
    out->cells_per_dim[1] = h->vcl.towersPerDim[1];
    out->cells_per_dim[2] = 1;
    out->cell_length[0] = h->vcl.side[0];
    out->cell_length[1] = h->vcl.side[1];
    out->num_cells = h->vcl.numTowers;
    out->cluster_size = h->cfg.cluster_size;
    out->num_clusters = h->numClusters;
    out->num_cluster_pairs = h->numPairs;
    out->towers_per_interaction_length = h->vcl.numTowersPerInteractionLength;
  }
  return APB_OK;
}


extern "C" int apb_debug_cell_of_slot(apb_handle h, int64_t *out) {
  APB_ENTRY(h);
  if (!h->structureValid) return h->fail(APB_ERR_STATE, "apb_debug_cell_of_slot: structure is not built");
  if (h->nslots == 0) return APB_OK;
  std::vector<int> tmp(h->nslots);
  APB_CUDA(cudaMemcpy(tmp.data(), h->slotCell.p, sizeof(int) * h->nslots, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < h->nslots; ++i) out[i] = tmp[i];
  return APB_OK;
}

extern "C" int apb_debug_cluster_pairs(apb_handle h, int64_t *out) {
  APB_ENTRY(h);
  if (!h->structureValid || h->cfg.container != APB_CONTAINER_VERLET_CLUSTER_LISTS)
    return h->fail(APB_ERR_STATE, "apb_debug_cluster_pairs: cluster lists are not built");
  if (h->numPairs == 0) return APB_OK;
  std::vector<int> starts(h->numClusters + 1), list(h->numPairs);
  APB_CUDA(cudaMemcpy(starts.data(), h->nbrStart.p, sizeof(int) * (h->numClusters + 1), cudaMemcpyDeviceToHost));
  APB_CUDA(cudaMemcpy(list.data(), h->nbrList.p, sizeof(int) * h->numPairs, cudaMemcpyDeviceToHost));
  for (int64_t A = 0; A < h->numClusters; ++A)
    for (int e = starts[A]; e < starts[A + 1]; ++e) {
      out[2 * static_cast<int64_t>(e)] = A;
      out[2 * static_cast<int64_t>(e) + 1] = list[e];
    }
  return APB_OK;
}
