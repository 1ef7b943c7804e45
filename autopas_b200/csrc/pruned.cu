// gpuvcl_pruned — the B200-native traversal of the gpuVerletClusterLists container (newton3 off).
//
// The reference's cluster traversal evaluates M x M distances for every listed cluster pair
// (traversals/VCLClusterFunctor.h:80-95); at liquid density only ~10-20 % of those lie inside the cutoff, and on a GPU
// the FP64 pipe pays for all of them. Here the cluster-pair list (built exactly like the reference's) is refined once
// per rebuild into per-particle lists: for every owned particle i of an owned cluster, the partners j (from its own
// cluster and from the listed neighbour clusters) with |r_i - r_j|^2 <= (cutoff + skin)^2. That is a superset of every
// pair that can come inside the cutoff before the next rebuild (same skin argument as the cluster list itself), so
// forces and globals equal the list-faithful traversal; only the number of distance evaluations drops (hit rate
// ~0.7 instead of ~0.1).
//
// Layout for the force kernel: a CTA owns a tile of up to 8 warp chunks (32 consecutive slots of one tower each) taken
// from a 2 x 2 block of towers and two consecutive z chunks, i.e. a compact brick, so that the union of the clusters its
// particles interact with is small. That union is staged once in shared memory (coalesced loads), and each lane walks
// its private list of 16-bit indices into the staged tile. Lists are stored per warp in rows of 32 lanes x 4 entries
// (one 8-byte load per lane per 4 pairs, 256 contiguous bytes per warp); rows are prefetched two ahead. The pair kernel
// is branch-free fp64; there are no atomics. Padding entries point at a sentinel slot parked at 1e300, which fails the
// cutoff test.
#include <algorithm>

#include "internal.cuh"
#include "lj_device.cuh"

#define PR_WARPS 8
#define PR_TILE (PR_WARPS * 32)
#define PR_CAND_MAX 8192  // candidate cluster ids gathered per tile before sort/unique

struct PrunedArgs {
  int M, logM;
  int numTiles;
  const int *chunkFirst, *chunkNum;  // [numTiles * PR_WARPS]: first slot and slot count of each warp chunk (-1 / 0)
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
  if (threadIdx.x == 0) nCand = 0;
  __syncthreads();
  // eligible clusters of this tile: non-halo (newton3 off: halo clusters own no list and no self interaction).
  // thread (warp w, lane l) looks at cluster l of chunk w (a chunk holds 32 / M <= 32 clusters)
  bool any = false;
  {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int first = a.chunkFirst[tile * PR_WARPS + w], num = a.chunkNum[tile * PR_WARPS + w];
    const int nc = (num + a.M - 1) >> a.logM;
    if (first >= 0 && l < nc) {
      const int c = (first >> a.logM) + l;
      if (!a.clIsHalo[c]) {
        any = true;
        const int e0 = a.nbrStart[c], e1 = a.nbrStart[c + 1];
        const int base = atomicAdd(&nCand, e1 - e0 + 1);
        if (base + (e1 - e0 + 1) <= PR_CAND_MAX) {
          cand[base] = c;
          for (int e = e0; e < e1; ++e) cand[base + 1 + e - e0] = a.nbrList[e];
        }
      }
    }
  }
  const int anyBlock = __syncthreads_or(any);
  if (!anyBlock) {
    if (!FILL && threadIdx.x == 0) numStaged[tile] = 0;
    return;
  }
  const int n = nCand;
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
  const int b = min(static_cast<int>(threadIdx.x) * chunk, n), e = min(b + chunk, n);
  int cnt = 0;
  for (int t = b; t < e; ++t) cnt += (t == 0 || cand[t] != cand[t - 1]);
  chunkCount[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 1; o < PR_TILE; o <<= 1) {  // Hillis-Steele inclusive scan
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
// One CTA per tile, thread (warp w, lane l) <-> slot l of chunk w. FILL = false: per-warp maximum list length in rows
// of 4 entries. FILL = true: write the lists; entry k of lane l lives at row (k / 4): rowBase + l * 4 + (k % 4).
template <bool FILL>
__global__ void __launch_bounds__(PR_TILE) kPrunedLists(PrunedArgs a, const int *__restrict__ stagedStart,
                                                        const int *__restrict__ staged, int *__restrict__ warpRows,
                                                        const int *__restrict__ warpRowStart,
                                                        unsigned short *__restrict__ lists) {
  extern __shared__ int stg[];  // staged cluster ids of this tile (sorted)
  const int tile = blockIdx.x;
  const int g0 = stagedStart[tile], nS = stagedStart[tile + 1] - g0;
  for (int t = threadIdx.x; t < nS; t += PR_TILE) stg[t] = staged[g0 + t];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpGlobal = tile * PR_WARPS + warp;
  const int first = a.chunkFirst[warpGlobal], num = a.chunkNum[warpGlobal];
  if (first < 0) {  // warp without a chunk (whole warp)
    if (!FILL && lane == 0) warpRows[warpGlobal] = 0;
    return;
  }
  const int64_t i = static_cast<int64_t>(first) + lane;
  const unsigned short sentinel = static_cast<unsigned short>(nS * a.M);
  int cnt = 0;
  unsigned short *out = nullptr;
  int rows = 0;
  if (FILL) {
    rows = warpRows[warpGlobal];
    out = lists + static_cast<size_t>(warpRowStart[warpGlobal]) * 128 + lane * 4;
  }
  if (lane < num && nS > 0) {
    const int ownI = a.own[i];
    const int A = static_cast<int>(i >> a.logM);
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
        const int64_t sB = static_cast<int64_t>(B) << a.logM;
        for (int k = 0; k < a.M; ++k) {
          const int64_t j = sB + k;
          if (j == i || a.own[j] == APB_OWN_DUMMY) continue;
          const double dr2 = ljDist2(xi - a.x[j], yi - a.y[j], zi - a.z[j]);
          if (dr2 <= a.il2) {
            if (FILL) out[static_cast<size_t>(cnt >> 2) * 128 + (cnt & 3)] = static_cast<unsigned short>(lo * a.M + k);
            ++cnt;
          }
        }
      }
    }
  }
  if (!FILL) {
    int m = (cnt + 3) >> 2;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) warpRows[warpGlobal] = m;
  } else {
    for (int k = cnt; k < rows * 4; ++k) out[static_cast<size_t>(k >> 2) * 128 + (k & 3)] = sentinel;
  }
}

int apbBuildPruned(apb_handle h) {
  if (!h->structureValid || h->builtNewton3 != 0)
    return h->fail(APB_ERR_STATE, "gpuvcl_pruned needs cluster lists built with newton3 off");
  const int M = h->cfg.cluster_size;
  int logM = 0;
  while ((1 << logM) < M) ++logM;
  const int64_t n = h->nslots;
  h->prunedTiles = 0;
  h->prunedMaxStaged = 0;
  if (n == 0) {
    h->prunedValid = true;
    return APB_OK;
  }
  // tiles: bricks of 2 x 2 towers x 2 consecutive 32-slot chunks (towers are contiguous slot ranges, z sorted)
  const int64_t nt = h->vcl.numTowers;
  const int nx = h->vcl.towersPerDim[0], ny = h->vcl.towersPerDim[1];
  std::vector<int> towerStart(nt + 1);
  APB_CUDA(cudaMemcpy(towerStart.data(), h->start.p, sizeof(int) * (nt + 1), cudaMemcpyDeviceToHost));
  std::vector<int> cFirst, cNum;
  for (int by = 0; by < ny; by += 2) {
    for (int bx = 0; bx < nx; bx += 2) {
      int towers[4], ntw = 0, maxChunks = 0;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          if (bx + dx >= nx || by + dy >= ny) continue;
          const int t = (bx + dx) + (by + dy) * nx;
          towers[ntw++] = t;
          maxChunks = std::max(maxChunks, (towerStart[t + 1] - towerStart[t] + 31) / 32);
        }
      for (int zc = 0; zc < maxChunks; zc += 2) {
        int used = 0;
        int first[PR_WARPS], num[PR_WARPS];
        for (int q = 0; q < ntw; ++q) {
          const int t = towers[q];
          const int slots = towerStart[t + 1] - towerStart[t];
          for (int k = zc; k < zc + 2; ++k) {
            if (k * 32 >= slots) continue;
            first[used] = towerStart[t] + k * 32;
            num[used] = std::min(32, slots - k * 32);
            ++used;
          }
        }
        if (used == 0) continue;
        for (int w = 0; w < PR_WARPS; ++w) {
          cFirst.push_back(w < used ? first[w] : -1);
          cNum.push_back(w < used ? num[w] : 0);
        }
      }
    }
  }
  const int numTiles = static_cast<int>(cFirst.size() / PR_WARPS);
  const int numWarps = numTiles * PR_WARPS;
  h->prunedTiles = numTiles;
  h->prunedWarps = numWarps;
  if (numTiles == 0) {
    h->prunedValid = true;
    return APB_OK;
  }
  APB_CHECK(apbEnsure(h, h->prTileFirst, sizeof(int) * numWarps));
  APB_CHECK(apbEnsure(h, h->prTileNum, sizeof(int) * numWarps));
  APB_CUDA(cudaMemcpyAsync(h->prTileFirst.p, cFirst.data(), sizeof(int) * numWarps, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(h->prTileNum.p, cNum.data(), sizeof(int) * numWarps, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));  // host vectors go out of scope
  PrunedArgs a;
  a.M = M;
  a.logM = logM;
  a.numTiles = numTiles;
  a.chunkFirst = static_cast<const int *>(h->prTileFirst.p);
  a.chunkNum = static_cast<const int *>(h->prTileNum.p);
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.own = h->own;
  a.clIsHalo = static_cast<const int *>(h->clIsHalo.p);
  a.nbrStart = static_cast<const int *>(h->nbrStart.p);
  a.nbrList = static_cast<const int *>(h->nbrList.p);
  a.il2 = h->vcl.interactionLengthSqr;
  APB_CHECK(apbEnsure(h, h->prNumStaged, sizeof(int) * (numTiles + 1)));
  APB_CHECK(apbEnsure(h, h->prStagedStart, sizeof(int) * (numTiles + 1)));
  APB_CHECK(apbEnsure(h, h->prWarpLen, sizeof(int) * (numWarps + 1)));
  APB_CHECK(apbEnsure(h, h->prWarpStart, sizeof(int) * (numWarps + 1)));
  int *numStaged = static_cast<int *>(h->prNumStaged.p), *stagedStart = static_cast<int *>(h->prStagedStart.p);
  int *warpRows = static_cast<int *>(h->prWarpLen.p), *warpStart = static_cast<int *>(h->prWarpStart.p);
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
                                                              std::to_string(PR_CAND_MAX) +
                                                              " cluster-list entries; use a larger cluster size");
  const int maxStaged = hostMisc[0];
  if (static_cast<int64_t>(maxStaged) * M > 65534)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged tile exceeds 16-bit indices");
  const size_t smemForce = (static_cast<size_t>(maxStaged) * M + 2) * 28;
  if (smemForce > 200 * 1024)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged tile (" + std::to_string(maxStaged * M) +
                                               " particles) does not fit shared memory; use a larger cluster size");
  h->prunedMaxStaged = maxStaged;
  APB_CHECK(apbEnsure(h, h->prStaged, sizeof(int) * std::max<long long>(totalStaged, 1)));
  int *staged = static_cast<int *>(h->prStaged.p);
  ++h->launchCount, kPrunedStage<true><<<numTiles, PR_TILE, 0, h->stream>>>(a, numStaged, stagedStart, staged, maxStagedDev, overflowDev);
  APB_CUDA(cudaGetLastError());
  const size_t smemLists = sizeof(int) * std::max(maxStaged, 1);
  if (smemLists > 40 * 1024) {
    APB_CUDA(cudaFuncSetAttribute(kPrunedLists<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemLists)));
    APB_CUDA(cudaFuncSetAttribute(kPrunedLists<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemLists)));
  }
  APB_CUDA(cudaMemsetAsync(warpRows, 0, sizeof(int) * (numWarps + 1), h->stream));
  ++h->launchCount, kPrunedLists<false><<<numTiles, PR_TILE, smemLists, h->stream>>>(a, stagedStart, staged, warpRows, nullptr, nullptr);
  APB_CUDA(cudaGetLastError());
  APB_CHECK(apbExclusiveScan(h, warpRows, warpStart, numWarps + 1, totals));
  long long totalRows = 0;
  APB_CUDA(cudaMemcpyAsync(&totalRows, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (totalRows > 0x7fffffffLL / 128) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: lists exceed 2^31 entries");
  h->prunedRows = totalRows;
  APB_CHECK(apbEnsure(h, h->prLists, sizeof(unsigned short) * 128 * std::max<long long>(totalRows, 1)));
  ++h->launchCount, kPrunedLists<true><<<numTiles, PR_TILE, smemLists, h->stream>>>(a, stagedStart, staged, warpRows, warpStart,
                                                                  static_cast<unsigned short *>(h->prLists.p));
  APB_CUDA(cudaGetLastError());
  APB_CUDA(cudaStreamSynchronize(h->stream));
  h->prunedValid = true;
  if (getenv("APB_DEBUG"))
    fprintf(stderr, "[apb] pruned build: slots %lld tiles %d maxStagedClusters %d totalStaged %lld rows %lld (entries %lld)\n",
            static_cast<long long>(n), numTiles, maxStaged, totalStaged, totalRows, totalRows * 128);
  return APB_OK;
}

// ---- force kernel ------------------------------------------------------------------------------------------------
struct PrunedForceArgs {
  int M, logM;
  const int *chunkFirst, *chunkNum;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type, *own;
  const int *stagedStart, *staged, *warpRows, *warpRowStart;
  const unsigned short *lists;
  int stagedCapacity;  // particles incl. the sentinel slot, rounded up to even
  LJParams p;
  LJStats *partials;
};

// reciprocal by Newton-Raphson on the hardware seed: MUFU.RCP64H + 4 DFMA, relative error ~1 ulp, no slow path.
// inf / NaN inputs (sentinel distances) give garbage that the caller discards with a select.
__device__ __forceinline__ double prRcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <bool MIX, bool STATS>
struct PairAcc {
  double fx = 0., fy = 0., fz = 0.;
  double upot = 0., vx = 0., vy = 0., vz = 0.;
  unsigned dist = 0, hits = 0;
};

// one pair, branch-free. Same formula as LJFunctor.h:146-159 with fac regrouped as
// (lj6 * invdr2) * (48 eps * lj6 - 24 eps) = eps24 * (lj12 + lj12m6) * invdr2; differences are at the 1e-16 level.
template <bool MIX, bool STATS>
__device__ __forceinline__ void prPair(const LJParams &p, double xi, double yi, double zi, int ti, const double *sx,
                                       const double *sy, const double *sz, const int *stype, unsigned idx,
                                       unsigned sentinel, PairAcc<MIX, STATS> &acc) {
  const double drx = xi - sx[idx], dry = yi - sy[idx], drz = zi - sz[idx];
  const double dr2 = fma(drz, drz, fma(dry, dry, drx * drx));
  const bool hit = dr2 <= p.cutoff2;
  double e24, s2, shift6;
  if (MIX) {
    const double *m = p.mix + 3 * (static_cast<size_t>(ti) * p.T + stype[idx]);
    e24 = __ldg(m);
    s2 = __ldg(m + 1);
    shift6 = p.applyShift ? __ldg(m + 2) : 0.;
  } else {
    e24 = p.eps24;
    s2 = p.sigma2;
    shift6 = p.shift6;
  }
  const double inv = prRcp(dr2);
  const double lj2 = s2 * inv;
  const double lj6 = lj2 * lj2 * lj2;
  const double t = fma(e24 + e24, lj6, -e24);
  double fac = (lj6 * inv) * t;
  fac = hit ? fac : 0.;
  const double fx = drx * fac, fy = dry * fac, fz = drz * fac;
  acc.fx += fx;
  acc.fy += fy;
  acc.fz += fz;
  if (STATS) {
    // potentialEnergy6 = eps24 * (lj12 - lj6) + shift6 (LJFunctor.h:174); only owned particles carry lists: weight 1
    const double upot6 = fma(e24 * lj6, lj6 - 1.0, shift6);
    acc.upot += hit ? upot6 : 0.;
    acc.vx = fma(drx, fx, acc.vx);
    acc.vy = fma(dry, fy, acc.vy);
    acc.vz = fma(drz, fz, acc.vz);
    acc.dist += idx != sentinel;
    acc.hits += hit;
  }
}

template <bool MIX, bool STATS>
__global__ void __launch_bounds__(PR_TILE) kLJPruned(PrunedForceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double *sx = reinterpret_cast<double *>(smemRaw);
  double *sy = sx + a.stagedCapacity;
  double *sz = sy + a.stagedCapacity;
  int *stype = reinterpret_cast<int *>(sz + a.stagedCapacity);
  const int tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpGlobal = tile * PR_WARPS + warp;
  const int g0 = a.stagedStart[tile], nS = a.stagedStart[tile + 1] - g0;
  const int nP = nS << a.logM;
  const int mask = a.M - 1;
  // issue the first list rows before staging so that their latency overlaps the staging loads
  const int first = a.chunkFirst[warpGlobal];
  const int rows = first >= 0 ? a.warpRows[warpGlobal] : 0;
  const unsigned sentinel = static_cast<unsigned>(nP);
  const unsigned sent2 = sentinel | (sentinel << 16);
  const uint2 *list = reinterpret_cast<const uint2 *>(a.lists) +
                      (rows > 0 ? static_cast<size_t>(a.warpRowStart[warpGlobal]) * 32 + lane : 0);
  uint2 cur = make_uint2(sent2, sent2), nxt = cur;
  if (rows > 0) cur = __ldg(list);
  if (rows > 1) nxt = __ldg(list + 32);
  for (int e = threadIdx.x; e < nP; e += PR_TILE) {
    const int64_t slot = (static_cast<int64_t>(a.staged[g0 + (e >> a.logM)]) << a.logM) + (e & mask);
    // particles deleted since the list build (ownership dummy) are moved out of reach
    const bool dead = a.own[slot] == APB_OWN_DUMMY;
    sx[e] = dead ? 1e300 : a.x[slot];
    sy[e] = a.y[slot];
    sz[e] = a.z[slot];
    if (MIX) stype[e] = a.type[slot];
  }
  if (threadIdx.x == 0) {  // sentinel slot for padding entries
    sx[nP] = 1e300;
    sy[nP] = 0.;
    sz[nP] = 0.;
    if (MIX) stype[nP] = 0;
  }
  __syncthreads();
  PairAcc<MIX, STATS> acc;
  if (rows > 0) {
    const int64_t i = static_cast<int64_t>(first) + lane;
    const bool active = lane < a.chunkNum[warpGlobal] && a.own[i] == APB_OWN_OWNED;
    const double xi = active ? a.x[i] : 0., yi = active ? a.y[i] : 0., zi = active ? a.z[i] : 0.;
    const int ti = (MIX && active) ? a.type[i] : 0;
    for (int r = 0; r < rows; ++r) {
      uint2 nn = make_uint2(sent2, sent2);
      if (r + 2 < rows) nn = __ldg(list + static_cast<size_t>(r + 2) * 32);
      if (!active) cur = make_uint2(sent2, sent2);  // particle deleted after the list build: no interactions
      prPair<MIX, STATS>(a.p, xi, yi, zi, ti, sx, sy, sz, stype, cur.x & 0xFFFFu, sentinel, acc);
      prPair<MIX, STATS>(a.p, xi, yi, zi, ti, sx, sy, sz, stype, cur.x >> 16, sentinel, acc);
      prPair<MIX, STATS>(a.p, xi, yi, zi, ti, sx, sy, sz, stype, cur.y & 0xFFFFu, sentinel, acc);
      prPair<MIX, STATS>(a.p, xi, yi, zi, ti, sx, sy, sz, stype, cur.y >> 16, sentinel, acc);
      cur = nxt;
      nxt = nn;
    }
    if (active) {
      a.fx[i] += acc.fx;
      a.fy[i] += acc.fy;
      a.fz[i] += acc.fz;
    }
  }
  if (STATS) {
    LJStats st;
    ljStatsZero(st);
    st.upot = acc.upot;
    st.vir[0] = acc.vx;
    st.vir[1] = acc.vy;
    st.vir[2] = acc.vz;
    st.dist = acc.dist;
    st.kNoN3 = acc.hits;
    st.gNoN3 = acc.hits;
    ljStatsBlockReduce(st, a.partials);
  }
}

int apbFinishStats(apb_handle h, int numBlocks, bool stats, const apb_functor *f, apb_traversal_result *out);

int apbComputeLJPruned(apb_handle h, const apb_functor *f, const LJParams &p, bool mix, bool stats,
                       apb_traversal_result *out) {
  if (!h->prunedValid) APB_CHECK(apbBuildPruned(h));
  const int numTiles = h->prunedTiles;
  if (numTiles == 0) return apbFinishStats(h, 0, stats, f, out);
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * numTiles));
  PrunedForceArgs a;
  a.M = h->cfg.cluster_size;
  a.logM = 0;
  while ((1 << a.logM) < a.M) ++a.logM;
  a.chunkFirst = static_cast<const int *>(h->prTileFirst.p);
  a.chunkNum = static_cast<const int *>(h->prTileNum.p);
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
  a.warpRows = static_cast<const int *>(h->prWarpLen.p);
  a.warpRowStart = static_cast<const int *>(h->prWarpStart.p);
  a.lists = static_cast<const unsigned short *>(h->prLists.p);
  a.stagedCapacity = (h->prunedMaxStaged * a.M + 2) & ~1;
  a.p = p;
  a.partials = static_cast<LJStats *>(h->partials.p);
  const size_t smem = static_cast<size_t>(a.stagedCapacity) * (mix ? 28 : 24);
  const int sel = (mix ? 2 : 0) | (stats ? 1 : 0);
#define PR_LAUNCH(MIXV, STATSV)                                                                                      \
  do {                                                                                                               \
    if (smem > 40 * 1024)                                                                                            \
      APB_CUDA(cudaFuncSetAttribute(kLJPruned<MIXV, STATSV>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                    static_cast<int>(smem)));                                                       \
    ++h->launchCount, kLJPruned<MIXV, STATSV><<<numTiles, PR_TILE, smem, h->stream>>>(a);                            \
  } while (0)
  switch (sel) {
    case 0: PR_LAUNCH(false, false); break;
    case 1: PR_LAUNCH(false, true); break;
    case 2: PR_LAUNCH(true, false); break;
    default: PR_LAUNCH(true, true); break;
  }
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      h->poisoned = true;
      return h->fail(APB_ERR_CUDA, std::string("kLJPruned launch failed: ") + cudaGetErrorString(e) + " (tiles " +
                                       std::to_string(numTiles) + ", dynamic smem " + std::to_string(smem) +
                                       " B, staged clusters max " + std::to_string(h->prunedMaxStaged) + ")");
    }
  }
  return apbFinishStats(h, numTiles, stats, f, out);
}
