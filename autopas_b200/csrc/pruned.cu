// gpuvcl_pruned — the B200-native traversal of the gpuVerletClusterLists container (newton3 off).
//
// The reference's cluster traversal evaluates M x M distances for every listed cluster pair
// (traversals/VCLClusterFunctor.h:80-95); at liquid density only ~10-20 % of those lie inside the cutoff, and on a GPU
// the FP64 pipe pays for all of them. Here the cluster-pair list (built exactly like the reference's) is refined once
// per rebuild into per-particle lists: for every particle i of an owned cluster, the partners j (from its own cluster
// and from the listed neighbour clusters) with |r_i - r_j|^2 <= (cutoff + skin)^2. That is a superset of every pair
// that can come inside the cutoff before the next rebuild (same skin argument as the cluster list itself), so forces
// and globals are identical to the list-faithful traversal; only the number of distance evaluations drops.
//
// Layout for the force kernel: a CTA owns a tile of 128 consecutive slots (4 warps). The union of clusters its
// particles interact with is staged once in shared memory (coalesced), and each lane walks its private list of 16-bit
// indices into that staged tile (lists are stored transposed per warp: entry k of lane l at base + 32 k + l, so a warp
// reads 64 contiguous bytes per step). The inner loop is: 1 coalesced LDG.U16, 3 LDS.64 gathers, the LJ kernel in
// fp64, no atomics.
#include <algorithm>

#include "internal.cuh"
#include "lj_device.cuh"

#define PR_TILE 128
#define PR_CAND_MAX 8192  // candidate cluster ids gathered per tile before sort/unique

struct PrunedArgs {
  int64_t nslots;
  int64_t numClusters;
  int M;
  int numTiles;
  const double *x, *y, *z;
  const int32_t *own;
  const int *clIsHalo, *nbrStart, *nbrList;
  double il2;
};

// ---- stage sets ----------------------------------------------------------------------------------------------------
// One CTA per tile: gather {own clusters} U {listed neighbours of its non-halo clusters}, sort, unique.
template <bool FILL>
__global__ void __launch_bounds__(PR_TILE) kPrunedStage(PrunedArgs a, int *__restrict__ numStaged,
                                                        const int *__restrict__ stagedStart, int *__restrict__ staged,
                                                        int *__restrict__ maxStaged, int *__restrict__ overflow) {
  __shared__ int cand[PR_CAND_MAX];
  __shared__ int nCand;
  __shared__ int chunkCount[PR_TILE];
  const int tile = blockIdx.x;
  const int64_t s0 = static_cast<int64_t>(tile) * PR_TILE;
  const int64_t s1 = min(s0 + PR_TILE, a.nslots);
  const int c0 = static_cast<int>(s0 / a.M), c1 = static_cast<int>((s1 + a.M - 1) / a.M);
  if (threadIdx.x == 0) nCand = 0;
  __syncthreads();
  // eligible clusters of this tile: non-halo (newton3 off: halo clusters own no list and no self interaction)
  bool any = false;
  for (int c = c0 + threadIdx.x; c < c1; c += PR_TILE) {
    if (a.clIsHalo[c]) continue;
    any = true;
    const int e0 = a.nbrStart[c], e1 = a.nbrStart[c + 1];
    const int base = atomicAdd(&nCand, e1 - e0 + 1);
    if (base + (e1 - e0 + 1) <= PR_CAND_MAX) {
      cand[base] = c;
      for (int e = e0; e < e1; ++e) cand[base + 1 + e - e0] = a.nbrList[e];
    }
  }
  const int anyBlock = __syncthreads_or(any);
  if (!anyBlock) {
    if (!FILL && threadIdx.x == 0) numStaged[tile] = 0;
    return;
  }
  int n = nCand;
  if (n > PR_CAND_MAX) {
    if (threadIdx.x == 0) {
      atomicExch(overflow, 1);
      if (!FILL) numStaged[tile] = 0;
    }
    return;
  }
  int P = 1;
  while (P < n) P <<= 1;
  for (int t = n + threadIdx.x; t < P; t += PR_TILE) cand[t] = 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += PR_TILE) {
        const int u = t ^ j;
        if (u > t) {
          const bool asc = (t & k) == 0;
          const int va = cand[t], vb = cand[u];
          if ((va > vb) == asc) {
            cand[t] = vb;
            cand[u] = va;
          }
        }
      }
      __syncthreads();
    }
  }
  // unique with a block scan over per-thread chunks
  const int chunk = (n + PR_TILE - 1) / PR_TILE;
  const int b = threadIdx.x * chunk, e = min(b + chunk, n);
  int cnt = 0;
  for (int t = b; t < e; ++t) cnt += (t == 0 || cand[t] != cand[t - 1]);
  chunkCount[threadIdx.x] = cnt;
  __syncthreads();
  // simple Hillis-Steele inclusive scan over 128 counts
  for (int o = 1; o < PR_TILE; o <<= 1) {
    const int v = threadIdx.x >= o ? chunkCount[threadIdx.x - o] : 0;
    __syncthreads();
    chunkCount[threadIdx.x] += v;
    __syncthreads();
  }
  const int total = chunkCount[PR_TILE - 1];
  if (!FILL) {
    if (threadIdx.x == 0) {
      numStaged[tile] = total;
      atomicMax(maxStaged, total);
    }
  } else {
    int pos = stagedStart[tile] + chunkCount[threadIdx.x] - cnt;
    for (int t = b; t < e; ++t)
      if (t == 0 || cand[t] != cand[t - 1]) staged[pos++] = cand[t];
  }
}

// ---- per-particle lists ----------------------------------------------------------------------------------------------
// One CTA per tile, thread t <-> slot. FILL = false: per-warp maximum list length. FILL = true: write the lists.
template <bool FILL>
__global__ void __launch_bounds__(PR_TILE) kPrunedLists(PrunedArgs a, const int *__restrict__ stagedStart,
                                                        const int *__restrict__ staged, int *__restrict__ warpLen,
                                                        const int *__restrict__ warpListStart,
                                                        unsigned short *__restrict__ lists) {
  extern __shared__ int stg[];  // staged cluster ids of this tile (sorted)
  const int tile = blockIdx.x;
  const int g0 = stagedStart[tile], nS = stagedStart[tile + 1] - g0;
  for (int t = threadIdx.x; t < nS; t += PR_TILE) stg[t] = staged[g0 + t];
  __syncthreads();
  const int64_t i = static_cast<int64_t>(tile) * PR_TILE + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int warpGlobal = tile * (PR_TILE / 32) + (threadIdx.x >> 5);
  int cnt = 0;
  unsigned short *out = nullptr;
  int len = 0;
  if (FILL) {
    len = warpLen[warpGlobal];
    out = lists + static_cast<size_t>(warpListStart[warpGlobal]) * 32 + lane;
  }
  if (i < a.nslots && nS > 0) {
    const int ownI = a.own[i];
    const int A = static_cast<int>(i / a.M);
    // forces on halo particles are never used and carry no weight in the globals: they get no list
    if (ownI == APB_OWN_OWNED && !a.clIsHalo[A]) {
      const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
      const int e0 = a.nbrStart[A], e1 = a.nbrStart[A + 1];
      for (int e = e0 - 1; e < e1; ++e) {
        const int B = e < e0 ? A : a.nbrList[e];
        int lo = 0, hi = nS;  // lower_bound(B)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (stg[mid] < B) lo = mid + 1; else hi = mid;
        }
        const int64_t sB = static_cast<int64_t>(B) * a.M;
        for (int k = 0; k < a.M; ++k) {
          const int64_t j = sB + k;
          if (j == i || a.own[j] == APB_OWN_DUMMY) continue;
          const double dr2 = ljDist2(xi - a.x[j], yi - a.y[j], zi - a.z[j]);
          if (dr2 <= a.il2) {
            if (FILL) out[static_cast<size_t>(cnt) * 32] = static_cast<unsigned short>(lo * a.M + k);
            ++cnt;
          }
        }
      }
    }
  }
  if (!FILL) {
    int m = cnt;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) warpLen[warpGlobal] = m;
  } else {
    for (int k = cnt; k < len; ++k) out[static_cast<size_t>(k) * 32] = 0xFFFF;
  }
}

int apbBuildPruned(apb_handle h) {
  if (!h->structureValid || h->builtNewton3 != 0)
    return h->fail(APB_ERR_STATE, "gpuvcl_pruned needs cluster lists built with newton3 off");
  const int M = h->cfg.cluster_size;
  const int64_t n = h->nslots;
  const int numTiles = apbDivUp(n, PR_TILE);
  h->prunedTiles = numTiles;
  h->prunedMaxStaged = 0;
  if (numTiles == 0) {
    h->prunedValid = true;
    return APB_OK;
  }
  PrunedArgs a;
  a.nslots = n;
  a.numClusters = h->numClusters;
  a.M = M;
  a.numTiles = numTiles;
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.own = h->own;
  a.clIsHalo = static_cast<const int *>(h->clIsHalo.p);
  a.nbrStart = static_cast<const int *>(h->nbrStart.p);
  a.nbrList = static_cast<const int *>(h->nbrList.p);
  a.il2 = h->vcl.interactionLengthSqr;
  const int numWarps = numTiles * (PR_TILE / 32);
  APB_CHECK(apbEnsure(h, h->prNumStaged, sizeof(int) * (numTiles + 1)));
  APB_CHECK(apbEnsure(h, h->prStagedStart, sizeof(int) * (numTiles + 1)));
  APB_CHECK(apbEnsure(h, h->prWarpLen, sizeof(int) * (numWarps + 1)));
  APB_CHECK(apbEnsure(h, h->prWarpStart, sizeof(int) * (numWarps + 1)));
  int *numStaged = static_cast<int *>(h->prNumStaged.p), *stagedStart = static_cast<int *>(h->prStagedStart.p);
  int *warpLen = static_cast<int *>(h->prWarpLen.p), *warpStart = static_cast<int *>(h->prWarpStart.p);
  char *scratch = static_cast<char *>(h->result.p) + sizeof(apb_traversal_result);
  long long *totals = reinterpret_cast<long long *>(scratch);
  int *maxStagedDev = reinterpret_cast<int *>(scratch + 32), *overflowDev = reinterpret_cast<int *>(scratch + 36);
  APB_CUDA(cudaMemsetAsync(scratch + 32, 0, 8, h->stream));
  APB_CUDA(cudaMemsetAsync(numStaged, 0, sizeof(int) * (numTiles + 1), h->stream));
  ++h->launchCount, kPrunedStage<false><<<numTiles, PR_TILE, 0, h->stream>>>(a, numStaged, nullptr, nullptr, maxStagedDev, overflowDev);
  APB_CUDA(cudaGetLastError());
  APB_CHECK(apbExclusiveScan(h, numStaged, stagedStart, numTiles + 1, totals));
  long long totalStaged = 0;
  int hostMisc[2] = {0, 0};
  APB_CUDA(cudaMemcpyAsync(&totalStaged, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(hostMisc, scratch + 32, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (hostMisc[1]) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: a tile interacts with more than " +
                                                              std::to_string(PR_CAND_MAX) + " cluster-list entries");
  const int maxStaged = hostMisc[0];
  if (static_cast<int64_t>(maxStaged) * M > 65534)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged tile exceeds 16-bit indices");
  const size_t smemForce = static_cast<size_t>(maxStaged) * M * 28;
  if (smemForce > 200 * 1024)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged tile (" + std::to_string(maxStaged * M) +
                                               " particles) does not fit shared memory; use a larger cluster size");
  h->prunedMaxStaged = maxStaged;
  APB_CHECK(apbEnsure(h, h->prStaged, sizeof(int) * std::max<long long>(totalStaged, 1)));
  int *staged = static_cast<int *>(h->prStaged.p);
  ++h->launchCount, kPrunedStage<true><<<numTiles, PR_TILE, 0, h->stream>>>(a, numStaged, stagedStart, staged, maxStagedDev, overflowDev);
  APB_CUDA(cudaGetLastError());
  const size_t smemLists = sizeof(int) * std::max(maxStaged, 1);
  if (smemLists > 48 * 1024) {
    APB_CUDA(cudaFuncSetAttribute(kPrunedLists<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemLists)));
    APB_CUDA(cudaFuncSetAttribute(kPrunedLists<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemLists)));
  }
  APB_CUDA(cudaMemsetAsync(warpLen, 0, sizeof(int) * (numWarps + 1), h->stream));
  ++h->launchCount, kPrunedLists<false><<<numTiles, PR_TILE, smemLists, h->stream>>>(a, stagedStart, staged, warpLen, nullptr, nullptr);
  APB_CUDA(cudaGetLastError());
  APB_CHECK(apbExclusiveScan(h, warpLen, warpStart, numWarps + 1, totals));
  long long totalRows = 0;
  APB_CUDA(cudaMemcpyAsync(&totalRows, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (totalRows > 0x7fffffffLL / 32) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: lists exceed 2^31 entries");
  h->prunedRows = totalRows;
  APB_CHECK(apbEnsure(h, h->prLists, sizeof(unsigned short) * 32 * std::max<long long>(totalRows, 1)));
  ++h->launchCount, kPrunedLists<true><<<numTiles, PR_TILE, smemLists, h->stream>>>(a, stagedStart, staged, warpLen, warpStart,
                                                                  static_cast<unsigned short *>(h->prLists.p));
  APB_CUDA(cudaGetLastError());
  APB_CUDA(cudaStreamSynchronize(h->stream));
  h->prunedValid = true;
  return APB_OK;
}

// ---- force kernel ------------------------------------------------------------------------------------------------
struct PrunedForceArgs {
  int64_t nslots;
  int M;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type, *own;
  const int *stagedStart, *staged, *warpLen, *warpStart;
  const unsigned short *lists;
  int maxStagedParticles;
  LJParams p;
  LJStats *partials;
};

template <bool MIX, bool STATS>
__global__ void __launch_bounds__(PR_TILE) kLJPruned(PrunedForceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double *sx = reinterpret_cast<double *>(smemRaw);
  double *sy = sx + a.maxStagedParticles;
  double *sz = sy + a.maxStagedParticles;
  int *stype = reinterpret_cast<int *>(sz + a.maxStagedParticles);
  const int tile = blockIdx.x;
  const int g0 = a.stagedStart[tile], nS = a.stagedStart[tile + 1] - g0;
  const int M = a.M;
  const int nP = nS * M;
  for (int e = threadIdx.x; e < nP; e += PR_TILE) {
    const int64_t slot = static_cast<int64_t>(a.staged[g0 + e / M]) * M + (e % M);
    // particles deleted since the list build (ownership dummy) are moved out of reach: dr2 = inf fails the cutoff test
    const bool dead = a.own[slot] == APB_OWN_DUMMY;
    sx[e] = dead ? 1e300 : a.x[slot];
    sy[e] = a.y[slot];
    sz[e] = a.z[slot];
    if (MIX) stype[e] = a.type[slot];
  }
  __syncthreads();
  LJStats st;
  ljStatsZero(st);
  const int64_t i = static_cast<int64_t>(tile) * PR_TILE + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int warpGlobal = tile * (PR_TILE / 32) + (threadIdx.x >> 5);
  const int len = a.warpLen[warpGlobal];
  if (len > 0 && i < a.nslots && a.own[i] == APB_OWN_OWNED) {
    const unsigned short *list = a.lists + static_cast<size_t>(a.warpStart[warpGlobal]) * 32 + lane;
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int ti = MIX ? a.type[i] : 0;
    double fxa = 0., fya = 0., fza = 0.;
    for (int k = 0; k < len; ++k) {
      const unsigned idx = list[static_cast<size_t>(k) * 32];
      if (idx == 0xFFFFu) continue;
      const double drx = xi - sx[idx], dry = yi - sy[idx], drz = zi - sz[idx];
      const double dr2 = ljDist2(drx, dry, drz);
      if (STATS) ++st.dist;
      if (dr2 <= a.p.cutoff2) {
        double upot6;
        const double fac = ljEval<MIX>(a.p, dr2, ti, MIX ? stype[idx] : 0, upot6);
        const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
        fxa += fx;
        fya += fy;
        fza += fz;
        if (STATS) {
          // only owned particles carry lists, so the weight [i owned] is 1
          st.upot += upot6;
          st.vir[0] += drx * fx;
          st.vir[1] += dry * fy;
          st.vir[2] += drz * fz;
          ++st.kNoN3;
          ++st.gNoN3;
        }
      }
    }
    a.fx[i] += fxa;
    a.fy[i] += fya;
    a.fz[i] += fza;
  }
  if (STATS) ljStatsBlockReduce(st, a.partials);
}

int apbFinishStats(apb_handle h, int numBlocks, bool stats, const apb_functor *f, apb_traversal_result *out);

int apbComputeLJPruned(apb_handle h, const apb_functor *f, const LJParams &p, bool mix, bool stats,
                       apb_traversal_result *out) {
  if (!h->prunedValid) APB_CHECK(apbBuildPruned(h));
  const int numTiles = h->prunedTiles;
  if (numTiles == 0) return apbFinishStats(h, 0, stats, f, out);
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * numTiles));
  PrunedForceArgs a;
  a.nslots = h->nslots;
  a.M = h->cfg.cluster_size;
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.fx = h->col[APB_COL_FX];
  a.fy = h->col[APB_COL_FY];
  a.fz = h->col[APB_COL_FZ];
  a.type = h->type;
  a.own = h->own;
  a.stagedStart = static_cast<const int *>(h->prStagedStart.p);
  a.staged = static_cast<const int *>(h->prStaged.p);
  a.warpLen = static_cast<const int *>(h->prWarpLen.p);
  a.warpStart = static_cast<const int *>(h->prWarpStart.p);
  a.lists = static_cast<const unsigned short *>(h->prLists.p);
  a.maxStagedParticles = std::max(h->prunedMaxStaged * a.M, 1);
  a.p = p;
  a.partials = static_cast<LJStats *>(h->partials.p);
  const size_t smem = static_cast<size_t>(a.maxStagedParticles) * (mix ? 28 : 24);
  const int sel = (mix ? 2 : 0) | (stats ? 1 : 0);
#define PR_LAUNCH(MIXV, STATSV)                                                                                      \
  do {                                                                                                               \
    if (smem > 48 * 1024)                                                                                            \
      APB_CUDA(cudaFuncSetAttribute(kLJPruned<MIXV, STATSV>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                    static_cast<int>(smem)));                                                       \
    ++h->launchCount, kLJPruned<MIXV, STATSV><<<numTiles, PR_TILE, smem, h->stream>>>(a);                                              \
  } while (0)
  switch (sel) {
    case 0: PR_LAUNCH(false, false); break;
    case 1: PR_LAUNCH(false, true); break;
    case 2: PR_LAUNCH(true, false); break;
    default: PR_LAUNCH(true, true); break;
  }
  APB_CUDA(cudaGetLastError());
  return apbFinishStats(h, numTiles, stats, f, out);
}
