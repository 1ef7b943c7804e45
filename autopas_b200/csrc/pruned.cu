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
// particles interact with is small. The particles of that union that some list of the tile refers to are staged once in
// shared memory (asynchronous 8-byte copies; (x, y) pairs and a z array at compile-time offsets), and each lane walks
// its private list of 16-bit entries (staged index * 16). Lists are stored per warp in rows of 32 lanes x 4 entries
// (one 8-byte load per lane per 4 pairs, 256 contiguous bytes per warp); rows are prefetched four ahead. The pair math
// is unconditional fp64, only the accumulation is predicated; forces leave through one RED.ADD per component and
// particle (single writer). Padding entries point at sentinel slots parked at 1e100, which fail the cutoff test.
// Tiles that stage no halo copy are "interior": apb_run_steps evaluates them while the halo refresh is in flight.
#include <algorithm>

#include "internal.cuh"
#include "lj_device.cuh"

#ifndef PR_WARPS
#define PR_WARPS 8  // warp chunks per tile (8: 2 x 2 towers x 2 chunks, 16: 2 x 2 towers x 4 chunks, 4: 2 x 1 towers x 2 chunks)
#endif
#define PR_ZC (PR_WARPS >= 8 ? PR_WARPS / 4 : 2)  // consecutive 32-slot chunks of one tower per tile (power of two)
#define PR_TILE (PR_WARPS * 32)
#ifndef PR_MINBLOCKS
#define PR_MINBLOCKS (32 / PR_WARPS)  // 32 resident warps per SM: 64 registers per thread
#endif
#ifdef PR_ZDUP
#define PR_MINBLOCKS_CAP2048 3
#define PR_MINBLOCKS_CAP4096 1
#define PR_BYTES_XYZ 32  // (x, y) + two z copies per staged particle
#else
#define PR_MINBLOCKS_CAP2048 PR_MINBLOCKS
#define PR_MINBLOCKS_CAP4096 2
#define PR_BYTES_XYZ 24
#endif
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
// The counting pass (FILL = false) also keeps its result in a fixed-stride scratch row of the tile when it fits
// (`early`, `earlyStride`): the sets then only have to be compacted into their final place (kPrunedStageCompact) instead
// of being gathered, sorted and made unique a second time. A tile that does not fit raises overflow[7] and the FILL
// pass runs as before.
template <bool FILL>
__global__ void __launch_bounds__(PR_TILE) kPrunedStage(PrunedArgs a, int *__restrict__ numStaged,
                                                        const int *__restrict__ stagedStart, int *__restrict__ staged,
                                                        int *__restrict__ maxStaged, int *__restrict__ overflow,
                                                        int *__restrict__ early, int earlyStride) {
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
      if (early && total > earlyStride) atomicExch(overflow + 7, 1);  // (flag word at scratch + 64)
    }
    if (early && total <= earlyStride) {
      int pos = tile * earlyStride + chunkCount[threadIdx.x] - cnt;
      for (int t = b; t < e; ++t)
        if (t == 0 || cand[t] != cand[t - 1]) early[pos++] = cand[t];
    }
  } else {
    int pos = stagedStart[tile] + chunkCount[threadIdx.x] - cnt;
    for (int t = b; t < e; ++t)
      if (t == 0 || cand[t] != cand[t - 1]) staged[pos++] = cand[t];
  }
}

// fixed-stride rows of the counting pass -> the compact sets; one warp per tile
__global__ void kPrunedStageCompact(int numTiles, const int *__restrict__ numStaged, const int *__restrict__ stagedStart,
                                    const int *__restrict__ early, int earlyStride, int *__restrict__ staged) {
  const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= numTiles) return;
  const int n = numStaged[tile], dst = stagedStart[tile];
  for (int t = lane; t < n; t += 32) staged[dst + t] = early[static_cast<size_t>(tile) * earlyStride + t];
}

// ---- tile table ----------------------------------------------------------------------------------------------------
// Tile (bx, by, zc) = towers (2bx..2bx+1, 2by..2by+1) x 32-slot chunks (2zc, 2zc+1) of each: a compact brick of up to
// 8 warp chunks. Towers are contiguous, z-sorted slot ranges, so chunk k of tower t is slots [start[t] + 32k, +32).
__global__ void kPrunedTiles(int numTiles, int nbx, int nzc, int nx, int ny, int shift, const int *__restrict__ towerStart,
                             int *__restrict__ chunkFirst, int *__restrict__ chunkNum) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= numTiles * PR_WARPS) return;
  const int tile = g / PR_WARPS, w = g % PR_WARPS;
  const int zc = tile % nzc, b = tile / nzc;
  const int bx = b % nbx, by = b / nbx;
  // `shift` aligns the 2 x 2 tower blocks with the first owned tower, so that halo towers (which carry no lists) do not
  // share a tile - and a CTA's lifetime - with owned ones
  const int wz = w % PR_ZC, wt = w / PR_ZC;  // chunk within the tower, tower within the 2 x 2 (2 x 1) block
  const int tx = 2 * bx - shift + (wt & 1), ty = PR_WARPS >= 8 ? 2 * by - shift + (wt >> 1) : by, k = PR_ZC * zc + wz;
  int first = -1, num = 0;
  if (tx >= 0 && ty >= 0 && tx < nx && ty < ny) {
    const int t = tx + ty * nx;
    const int s0 = towerStart[t], slots = towerStart[t + 1] - s0;
    if (k * 32 < slots) {
      first = s0 + k * 32;
      num = min(32, slots - k * 32);
    }
  }
  chunkFirst[g] = first;
  chunkNum[g] = num;
}

// ---- partner masks -------------------------------------------------------------------------------------------------
// One CTA per tile, thread (warp w, lane l) <-> slot l of chunk w. For every owned particle i of a non-halo cluster A and
// every entry B of {A} U list(A): the M-bit mask of partners j in B with |ri - rj|^2 <= (cutoff + skin)^2, evaluated
// in fp32 on tile-relative coordinates with a threshold widened by the worst-case rounding error, i.e. a (very slightly)
// conservative superset of the fp64 decision: FP32 issues at twice the FP64 rate and the test is the hot loop of the
// rebuild. The force kernel re-tests every pair exactly in fp64 against the cutoff, so forces do not depend on this.
// Masks go to masks[(nbrStart[A] + A + e) * M + (i % M)], e = 0 for A itself. Also produced: rows per warp (list length
// in rows of 4 entries), and per staged cluster the mask of particles referenced by any list of the tile together with
// its exclusive prefix (compact index base): the force kernel stages only referenced particles.
struct MaskOut {
  unsigned *masks;
  int *entryLo;  // per mask row: index of the partner cluster in the tile's staged set (saves kPrunedFill the search)
  int *warpRows;
  unsigned *used;   // [totalStaged]
  int *cbase;       // [totalStaged]
  int *numCompact;  // [numTiles]
  int *maxCompact;  // [0] largest compact tile, [1] longest list in rows
  unsigned long long *totalEntries;
};

__device__ __forceinline__ int prLowerBound(const int *stg, int nS, int B) {
  int lo = 0, hi = nS;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (stg[mid] < B) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// N3 (newton3 lists): every owned-owned pair is listed once, by the particle in the lower slot; pairs with a halo partner
// stay with the owned particle (halo particles own no list and receive no force). The force kernel then applies the
// reaction to the listed partner (LJFunctor::SoAFunctorPairImpl<true>, LJFunctor.h:387-559).
template <bool UNIFORM, bool N3>
__global__ void __launch_bounds__(PR_TILE) kPrunedMasks(PrunedArgs a, const int *__restrict__ stagedStart,
                                                        const int *__restrict__ staged, MaskOut o) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ float sRed[PR_WARPS];
  __shared__ int sScan[PR_TILE];
  const int tile = blockIdx.x;
  const int g0 = stagedStart[tile], nS = stagedStart[tile + 1] - g0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpGlobal = tile * PR_WARPS + warp;
  if (nS == 0) {
    if (lane == 0) o.warpRows[warpGlobal] = 0;
    if (threadIdx.x == 0) o.numCompact[tile] = 0;
    return;
  }
  const int nP = nS << a.logM;
  float4 *rel = reinterpret_cast<float4 *>(smemRaw);
  int *stg = reinterpret_cast<int *>(rel + nP);
  unsigned *used = reinterpret_cast<unsigned *>(stg + nS);
  unsigned *haloBits = used + nS;  // UNIFORM: bit k = particle k of the staged cluster is a halo copy
  for (int t = threadIdx.x; t < nS; t += PR_TILE) {
    stg[t] = staged[g0 + t];
    used[t] = 0u;
  }
  __syncthreads();
  // tile-relative fp32 coordinates; origin = first staged particle (same address for all threads: broadcast)
  const int64_t s00 = static_cast<int64_t>(stg[0]) << a.logM;
  const double ox = a.x[s00], oy = a.y[s00], oz = a.z[s00];
  const int mask = a.M - 1;
  float ext = 0.f;
  for (int e = threadIdx.x; e < nP; e += PR_TILE) {
    const int64_t slot = (static_cast<int64_t>(stg[e >> a.logM]) << a.logM) + (e & mask);
    float4 r;
    const int ownE = a.own[slot];
    if (ownE == APB_OWN_DUMMY) {
      // padding dummies are nobody's partner: dr2 overflows to +inf (UNIFORM: |r|^2 = +inf)
      r = make_float4(1e30f, 0.f, 0.f, UNIFORM ? __int_as_float(0x7f800000) : 0.f);
    } else {
      r = make_float4(static_cast<float>(a.x[slot] - ox), static_cast<float>(a.y[slot] - oy),
                      static_cast<float>(a.z[slot] - oz), (N3 && ownE == APB_OWN_HALO) ? 1.f : 0.f);
      ext = fmaxf(ext, fmaxf(fabsf(r.x), fmaxf(fabsf(r.y), fabsf(r.z))));
      if (UNIFORM) r.w = fmaf(r.z, r.z, fmaf(r.y, r.y, r.x * r.x));
    }
    rel[e] = r;
    if (UNIFORM && N3) {  // M == 32: a warp loads one staged cluster per pass (nP is a multiple of 32)
      const unsigned hb = __ballot_sync(0xffffffffu, ownE == APB_OWN_HALO);
      if (lane == 0) haloBits[e >> 5] = hb;
    }
  }
  for (int s = 16; s > 0; s >>= 1) ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, s));
  if (lane == 0) sRed[warp] = ext;
  __syncthreads();
  ext = sRed[0];
#pragma unroll
  for (int w = 1; w < PR_WARPS; ++w) ext = fmaxf(ext, sRed[w]);
  // error budget of the fp32 test: coordinates are rounded to fp32 (<= ext 2^-24 each), the differences and the sum of
  // squares add a few ulp: |dr2_f32 - dr2| <= 2 sqrt(3) il (3 ext 2^-24) + il^2 2^-22  ~  6.2e-7 il ext + 2.4e-7 il^2.
  // The threshold below carries more than twice that.
  const double il = sqrt(a.il2);
  // UNIFORM evaluates dr2 = |ri|^2 + |rj|^2 - 2 ri.rj with |rj|^2 precomputed (3 FFMA per test instead of 3 FADD + FMUL +
  // 2 FFMA): the squares (<= 3 ext^2 each), the three FMA roundings (partial sums <= 9 ext^2) and thr - |ri|^2 add at
  // most 48 ulp(ext^2) = 2.9e-6 ext^2 to the error of dr2; the threshold carries twice that on top.
  const float thr = __double2float_ru(a.il2 * (1.0 + 1e-6) + 1.5e-6 * il * static_cast<double>(ext) +
                                      (UNIFORM ? 6e-6 * static_cast<double>(ext) * static_cast<double>(ext) : 0.0));
  const float thrBox = thr * 1.00001f;

  const int first = a.chunkFirst[warpGlobal], num = a.chunkNum[warpGlobal];
  int cnt = 0;
  if (first >= 0) {
    const int64_t i = static_cast<int64_t>(first) + lane;
    const bool laneIn = lane < num;
    const int A = static_cast<int>((laneIn ? i : static_cast<int64_t>(first)) >> a.logM);
    // forces on halo particles are never used and carry no weight in the globals: they get no list
    const bool active = laneIn && a.own[i] == APB_OWN_OWNED && !a.clIsHalo[A];
    if (UNIFORM) {
      // M == 32: the warp is one cluster, all lanes walk the same entries and test the same partner (smem broadcast)
      if (!a.clIsHalo[A]) {
        const int loA = prLowerBound(stg, nS, A);
        const float4 ri = rel[(loA << 5) + lane];
        const float nxi = -2.f * ri.x, nyi = -2.f * ri.y, nzi = -2.f * ri.z, thri = thr - ri.w;
        // fp32 bounding box of the active particles of A
        float bx0 = active ? ri.x : 3e38f, bx1 = active ? ri.x : -3e38f;
        float by0 = active ? ri.y : 3e38f, by1 = active ? ri.y : -3e38f;
        float bz0 = active ? ri.z : 3e38f, bz1 = active ? ri.z : -3e38f;
        for (int s = 16; s > 0; s >>= 1) {
          bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, s));
          bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, s));
          by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, s));
          by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, s));
          bz0 = fminf(bz0, __shfl_xor_sync(0xffffffffu, bz0, s));
          bz1 = fmaxf(bz1, __shfl_xor_sync(0xffffffffu, bz1, s));
        }
        const int e0 = a.nbrStart[A], e1 = a.nbrStart[A + 1];
        unsigned *mrow = o.masks + (static_cast<size_t>(e0) + A) * 32 + lane;
        // The partner clusters of A and their positions in the staged set are fetched 32 entries at a time, one per lane
        // (one coalesced load and 32 binary searches side by side instead of a dependent global load and a serial search
        // per entry), and handed out by shuffles.
        const int nE = e1 - e0 + 1;  // entry 0 is A itself
        int Bl = A, lol = 0;
        for (int e = e0 - 1; e < e1; ++e, mrow += 32) {
          const int q = (e - (e0 - 1)) & 31;
          if (q == 0) {
            const int idx = e - (e0 - 1) + lane;
            if (idx < nE) {
              Bl = idx == 0 ? A : __ldg(a.nbrList + e0 + idx - 1);
              lol = prLowerBound(stg, nS, Bl);
              o.entryLo[static_cast<size_t>(e0) + A + idx] = lol;
            }
          }
          const int B = __shfl_sync(0xffffffffu, Bl, q);
          const int lo = __shfl_sync(0xffffffffu, lol, q);
          const float4 *rb = rel + (lo << 5);
          // candidate partners: lane k tests particle k of B against the box of A
          const float4 pk = rb[lane];
          const float gx = fmaxf(0.f, fmaxf(bx0 - pk.x, pk.x - bx1)), gy = fmaxf(0.f, fmaxf(by0 - pk.y, pk.y - by1)),
                      gz = fmaxf(0.f, fmaxf(bz0 - pk.z, pk.z - bz1));
          unsigned cand = __ballot_sync(0xffffffffu, fmaf(gz, gz, fmaf(gy, gy, gx * gx)) <= thrBox);
          // Groups of eight partners (B is z-sorted, candidates come in runs) are skipped when the box test left none
          // in them; inside a group every partner is tested with compile-time offsets and bit masks: 9 instructions
          // per test instead of 19 for the bit-serial walk over the candidate mask. A partner that failed the box
          // test is farther than the threshold from every active particle of A, so the masks are the same.
          unsigned m = 0u;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            if ((cand >> (8 * g8)) & 0xFFu) {
#pragma unroll
              for (int k = 8 * g8; k < 8 * g8 + 8; ++k) {
                const float4 pj = rb[k];
                if (fmaf(nxi, pj.x, fmaf(nyi, pj.y, fmaf(nzi, pj.z, pj.w))) <= thri) m |= 1u << k;
              }
            }
          }
          if (e < e0) m &= ~(1u << lane);
          if (N3) {
            const unsigned halos = haloBits[lo];  // bit k: particle k of B is a halo copy
            m &= halos | (B > A ? 0xffffffffu : (B == A ? (0xfffffffeu << lane) : 0u));
          }
          if (!active) m = 0u;
          *mrow = m;
          cnt += __popc(m);
          const unsigned wm = __reduce_or_sync(0xffffffffu, m);
          if (lane == 0 && wm) atomicOr(&used[lo], wm);
        }
      }
    } else if (active) {
      const int loA = prLowerBound(stg, nS, A);
      const int li = static_cast<int>(i) & mask;
      const float4 ri = rel[(loA << a.logM) + li];
      const int e0 = a.nbrStart[A], e1 = a.nbrStart[A + 1];
      for (int e = e0 - 1; e < e1; ++e) {
        const int B = e < e0 ? A : a.nbrList[e];
        const int lo = prLowerBound(stg, nS, B);
        const float4 *rb = rel + (lo << a.logM);
        unsigned m = 0u;
        for (int k = 0; k < a.M; ++k) {
          const float4 pj = rb[k];
          const float dx = ri.x - pj.x, dy = ri.y - pj.y, dz = ri.z - pj.z;
          const bool mine = !N3 || pj.w != 0.f || B > A || (B == A && k > li);
          if (mine && fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= thr) m |= 1u << k;
        }
        if (e < e0) m &= ~(1u << li);
        o.masks[(static_cast<size_t>(e) + 1 + A) * a.M + li] = m;
        o.entryLo[static_cast<size_t>(e) + 1 + A] = lo;  // same value from every lane of the cluster
        cnt += __popc(m);
        if (m) atomicOr(&used[lo], m);
      }
    }
  }
  {
    int m = (cnt + 3) >> 2;
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    if (lane == 0) {
      o.warpRows[warpGlobal] = m;
      if (m) atomicMax(o.maxCompact + 1, m);
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, static_cast<unsigned>(cnt));
    if (lane == 0 && tot) atomicAdd(o.totalEntries, static_cast<unsigned long long>(tot));
  }
  __syncthreads();
  // compact index base per staged cluster = exclusive prefix of popc(used) (block scan over per-thread chunks)
  const int chunk = (nS + PR_TILE - 1) / PR_TILE;
  const int b = min(static_cast<int>(threadIdx.x) * chunk, nS), e = min(b + chunk, nS);
  int c = 0;
  for (int t = b; t < e; ++t) c += __popc(used[t]);
  sScan[threadIdx.x] = c;
  __syncthreads();
  for (int s = 1; s < PR_TILE; s <<= 1) {
    const int v = threadIdx.x >= s ? sScan[threadIdx.x - s] : 0;
    __syncthreads();
    sScan[threadIdx.x] += v;
    __syncthreads();
  }
  int run = sScan[threadIdx.x] - c;
  for (int t = b; t < e; ++t) {
    o.used[g0 + t] = used[t];
    o.cbase[g0 + t] = run;
    run += __popc(used[t]);
  }
  if (threadIdx.x == PR_TILE - 1) {
    o.numCompact[tile] = sScan[PR_TILE - 1];
    atomicMax(o.maxCompact, sScan[PR_TILE - 1]);
  }
}

// ---- lists ---------------------------------------------------------------------------------------------------------
// Expands the masks into per-lane lists of 16-bit entries (compact index * 16 = byte offset of the partner's (x, y)): entry k of lane l lives in row k / 4 at
// rowBase + l * 4 + (k % 4). Also writes the tile's compact slot table (which slot each compact index stages).
template <bool UNIFORM>
__global__ void __launch_bounds__(PR_TILE) kPrunedFill(PrunedArgs a, const int *__restrict__ stagedStart,
                                                       const int *__restrict__ staged, MaskOut o,
                                                       const int *__restrict__ warpRowStart,
                                                       unsigned short *__restrict__ lists, int *__restrict__ compactSlot,
                                                       int smemRowsPerWarp, int sched, int *__restrict__ tileHalo,
                                                       int sentBase) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int tile = blockIdx.x;
  const int g0 = stagedStart[tile], nS = stagedStart[tile + 1] - g0;
  if (nS == 0) {
    if (threadIdx.x == 0) tileHalo[tile] = 0;
    return;
  }
  int *stg = reinterpret_cast<int *>(smemRaw);
  unsigned *used = reinterpret_cast<unsigned *>(stg + nS);
  int *cbase = reinterpret_cast<int *>(used + nS);
  for (int t = threadIdx.x; t < nS; t += PR_TILE) {
    stg[t] = staged[g0 + t];
    used[t] = o.used[g0 + t];
    cbase[t] = o.cbase[g0 + t];
  }
  __syncthreads();
  const int mask = a.M - 1;
  int *cs = compactSlot + (static_cast<size_t>(g0) << a.logM);
  // a tile is a boundary tile if any particle it stages is a halo copy: its forces have to wait for the halo refresh
  int stagesHalo = 0;
  for (int e = threadIdx.x; e < (nS << a.logM); e += PR_TILE) {
    const int s = e >> a.logM, k = e & mask;
    const unsigned u = used[s];
    if ((u >> k) & 1u) {
      const int slot = (stg[s] << a.logM) + k;
      cs[cbase[s] + __popc(u & ((1u << k) - 1u))] = slot;
      stagesHalo |= a.own[slot] == APB_OWN_HALO;
    }
  }
  stagesHalo = __syncthreads_or(stagesHalo);
  if (threadIdx.x == 0) tileHalo[tile] = stagesHalo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpGlobal = tile * PR_WARPS + warp;
  const int first = a.chunkFirst[warpGlobal], num = a.chunkNum[warpGlobal];
  const int rows = o.warpRows[warpGlobal];
  if (first < 0 || rows == 0) return;
  // The warp's rows are assembled in shared memory (2-byte scattered writes) and copied out coalesced; directly in
  // global memory only if a warp's rows exceed the shared-memory budget. Padding entries point at one of the 16 sentinel
  // slots that follow the staged particles (one per bank class).
  //
  // Order of a lane's entries (sched != 0). The force kernel gathers the partner of every lane from shared memory: an
  // LDS.128 for (x, y) - a quarter-warp per wavefront, bank group = index mod 8 - and an LDS.64 for z - a half-warp per
  // wavefront, bank pair = index mod 16. In list order the indices of the 32 lanes are unrelated and a row costs ~14
  // wavefronts instead of 6 (tools/sim/bank_conflicts.py; ncu: LSU data pipe 86 % busy, the first limiter of
  // kLJPruned). So each lane's list is laid out on a Latin square: at slot t lane l reads a partner of class
  // (l + t) mod 16, i.e. the j-th entry of class k sits at slot ((k - l) mod 16) + 16 j. At every slot the 16 lanes of a
  // half-warp then hit 16 different bank pairs, and the 8 lanes of a quarter 8 different bank groups: conflict-free by
  // construction, with no communication between lanes while the list is built. An entry whose slot is beyond the warp's
  // rows or already taken (a class with more entries than visits: about one entry in eight) goes to the lane's last free
  // slot instead (open addressing from the end, where most lanes idle) and may collide there; the slots left over take
  // the sentinel of the slot's own class. ~7.6 wavefronts per row in the model. Only the order inside a lane changes:
  // the same pairs are evaluated.
  const int R = rows * 4;
  unsigned short *gout = lists + static_cast<size_t>(warpRowStart[warpGlobal]) * 128;
  const bool viaSmem = rows <= smemRowsPerWarp;
  const bool latin = viaSmem && sched;
  // Shared-memory assembly is lane-major: lane l owns the R consecutive 2-byte slots at l * stride (stride = an odd
  // number of 8-byte rows, so that the transposing copy-out - lane l reads its row r, the warp writes 256 contiguous
  // bytes - is free of bank conflicts). Appending an entry is then one STS and one pointer increment.
  unsigned char *blocks = smemRaw + ((static_cast<size_t>(nS) * 12 + 15) & ~size_t(15));
  const int strideRows = smemRowsPerWarp | 1;
  unsigned short *mine = reinterpret_cast<unsigned short *>(blocks) +
                         (static_cast<size_t>(warp) * 32 + lane) * strideRows * 4;  // viaSmem only
  const int64_t i = static_cast<int64_t>(first) + lane;
  const bool laneIn = lane < num;
  const int A = static_cast<int>((laneIn ? i : static_cast<int64_t>(first)) >> a.logM);
  const bool active = laneIn && a.own[i] == APB_OWN_OWNED && !a.clIsHalo[A];
  const int li = static_cast<int>(i) & mask;
  const int e0 = active ? a.nbrStart[A] : 0, e1 = active ? a.nbrStart[A + 1] : -1;
  const unsigned sentMine = static_cast<unsigned>(sentBase + (lane & 15)) << 4;
  int cnt = 0, spill = R - 1;
  unsigned long long C0 = 0ULL, C1 = 0ULL;  // Latin: entries placed so far per class, 8 bits each (classes 0-7, 8-15)
  if (latin)
    for (int r = 0; r < rows; ++r) *reinterpret_cast<uint2 *>(mine + r * 4) = make_uint2(~0u, ~0u);
  // masks and partner positions are fetched eight entries at a time (independent loads in flight together: the loop is
  // bound by global-memory latency otherwise), then expanded one after the other
  constexpr int PR_FILL_BATCH = 8;
  const size_t rowBase = static_cast<size_t>(e0) + A;  // mask row of entry 0 (A itself)
  const int nE = e1 - e0 + 1;                          // 0 for an inactive lane
  unsigned short *wp = mine;
  for (int eb = 0; eb < nE; eb += PR_FILL_BATCH) {
    unsigned mm[PR_FILL_BATCH];
    int ll[PR_FILL_BATCH];
#pragma unroll
    for (int q = 0; q < PR_FILL_BATCH; ++q) {
      const bool v = eb + q < nE;
      mm[q] = v ? __ldg(o.masks + (rowBase + eb + q) * a.M + li) : 0u;
      ll[q] = v ? __ldg(o.entryLo + rowBase + eb + q) : 0;
    }
#pragma unroll
    for (int q = 0; q < PR_FILL_BATCH; ++q) {
      unsigned m = mm[q];
      const unsigned u = used[ll[q]];
      const unsigned cb16 = static_cast<unsigned>(cbase[ll[q]]) << 4;
      if (viaSmem && !latin) {
        // lowest set bit k of m: its rank among the referenced particles of the cluster is popc(u & bits below k), and
        // the bits below k are ~m & (m - 1) - no bit index, no shift
        while (m) {
          const unsigned t = m - 1u;
          *wp++ = static_cast<unsigned short>(cb16 + (static_cast<unsigned>(__popc(u & ~m & t)) << 4));
          m &= t;
        }
      } else {
        while (m) {
          const unsigned t = m - 1u;
          const unsigned ci16 = cb16 + (static_cast<unsigned>(__popc(u & ~m & t)) << 4);
          m &= t;
          if (latin) {
            const unsigned cls = (ci16 >> 4) & 15u, sh = (cls & 7u) * 8u;
            unsigned j;
            if (cls & 8u) {
              j = static_cast<unsigned>(C1 >> sh) & 0xFFu;
              C1 += 1ULL << sh;
            } else {
              j = static_cast<unsigned>(C0 >> sh) & 0xFFu;
              C0 += 1ULL << sh;
            }
            int t2 = static_cast<int>((cls - lane) & 15u) + 16 * static_cast<int>(j);
            if (t2 >= R || mine[t2] != 0xFFFFu) {
              while (mine[spill] != 0xFFFFu) --spill;  // cnt <= R: one is free
              t2 = spill--;
            }
            mine[t2] = static_cast<unsigned short>(ci16);
          } else {
            gout[static_cast<size_t>(cnt >> 2) * 128 + lane * 4 + (cnt & 3)] = static_cast<unsigned short>(ci16);
          }
          ++cnt;
        }
      }
    }
  }
  if (viaSmem && !latin) cnt = static_cast<int>(wp - mine);
  if (latin) {
    // empty slots: the sentinel of the slot's own class (sentBase is a multiple of 16: sentinel sentBase + c has class c)
    for (int t = 0; t < R; ++t)
      if (mine[t] == 0xFFFFu) mine[t] = static_cast<unsigned short>((sentBase + ((lane + t) & 15)) << 4);
  } else if (viaSmem) {
    for (int t = cnt; t < R; ++t) mine[t] = static_cast<unsigned short>(sentMine);
  } else {
    for (int t = cnt; t < R; ++t) gout[static_cast<size_t>(t >> 2) * 128 + lane * 4 + (t & 3)] = static_cast<unsigned short>(sentMine);
  }
  if (viaSmem) {
    uint2 *dst = reinterpret_cast<uint2 *>(gout) + lane;
    for (int r = 0; r < rows; ++r) dst[r * 32] = *reinterpret_cast<const uint2 *>(mine + r * 4);
  }
}

// Tile order for the split force step: interior tiles (stage no halo copy) first, boundary tiles after them, each group
// in ascending tile order (deterministic). order[pos] = tile, *numInterior = size of the first group. One block.
__global__ void __launch_bounds__(1024) kPrunedTileOrder(int numTiles, const int *__restrict__ tileHalo,
                                                        int *__restrict__ order, int *__restrict__ numInterior) {
  __shared__ int sScan[1024];
  const int chunk = (numTiles + 1023) / 1024;
  const int b = min(static_cast<int>(threadIdx.x) * chunk, numTiles), e = min(b + chunk, numTiles);
  int c = 0;
  for (int t = b; t < e; ++t) c += tileHalo[t] == 0;
  sScan[threadIdx.x] = c;
  __syncthreads();
  for (int s = 1; s < 1024; s <<= 1) {
    const int v = threadIdx.x >= s ? sScan[threadIdx.x - s] : 0;
    __syncthreads();
    sScan[threadIdx.x] += v;
    __syncthreads();
  }
  const int total = sScan[1023];
  int posI = sScan[threadIdx.x] - c, posB = total + (b - posI);
  for (int t = b; t < e; ++t) {
    if (tileHalo[t] == 0) order[posI++] = t; else order[posB++] = t;
  }
  if (threadIdx.x == 0) *numInterior = total;
}

int apbBuildPruned(apb_handle h, int newton3) {
  // The per-particle lists are refined from the full (newton3 off) cluster-pair list in both modes; with newton3 the
  // half that the lower slot owns is kept (kPrunedMasks<*, true>).
  if (!h->structureValid || h->builtNewton3 != 0)
    return h->fail(APB_ERR_STATE, "gpuvcl_pruned needs the cluster-pair list of the newton3-off mode");
  const bool n3 = newton3 != 0;
  h->prunedNewton3 = n3 ? 1 : 0;
  const int M = h->cfg.cluster_size;
  int logM = 0;
  while ((1 << logM) < M) ++logM;
  if ((1 << logM) != M || M > 32)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned needs a power-of-two cluster size <= 32");
  const int64_t n = h->nslots;
  h->prunedTiles = 0;
  h->prunedMaxStaged = 0;
  h->prunedMaxCompact = 0;
  if (n == 0) {
    h->prunedValid = true;
    return APB_OK;
  }
  // tiles: bricks of 2 x 2 towers x 2 consecutive 32-slot chunks
  const int nx = h->vcl.towersPerDim[0], ny = h->vcl.towersPerDim[1];
  const int shift = h->vcl.numTowersPerInteractionLength & 1;
  const int nbx = (nx + shift + 1) / 2, nby = PR_WARPS >= 8 ? (ny + shift + 1) / 2 : ny;
  const int maxSlots = (h->vclMaxTowerCount + M - 1) / M * M;
  const int nzc = std::max(1, ((maxSlots + 31) / 32 + PR_ZC - 1) / PR_ZC);
  const int64_t numTiles64 = static_cast<int64_t>(nbx) * nby * nzc;
  if (numTiles64 * PR_WARPS > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: too many tiles");
  const int numTiles = static_cast<int>(numTiles64);
  const int numWarps = numTiles * PR_WARPS;
  h->prunedTiles = numTiles;
  h->prunedWarps = numWarps;
  APB_CHECK(apbEnsure(h, h->prTileFirst, sizeof(int) * numWarps));
  APB_CHECK(apbEnsure(h, h->prTileNum, sizeof(int) * numWarps));
  ++h->launchCount, kPrunedTiles<<<apbDivUp(numWarps, 256), 256, 0, h->stream>>>(
      numTiles, nbx, nzc, nx, ny, shift, static_cast<const int *>(h->start.p), static_cast<int *>(h->prTileFirst.p),
      static_cast<int *>(h->prTileNum.p));
  APB_CUDA(cudaGetLastError());
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
  APB_CHECK(apbEnsure(h, h->prNumCompact, sizeof(int) * (numTiles + 1)));
  APB_CHECK(apbEnsure(h, h->prTileHalo, sizeof(int) * (numTiles + 1)));
  APB_CHECK(apbEnsure(h, h->prTileOrder, sizeof(int) * (numTiles + 2)));
  int *numStaged = static_cast<int *>(h->prNumStaged.p), *stagedStart = static_cast<int *>(h->prStagedStart.p);
  int *warpRows = static_cast<int *>(h->prWarpLen.p), *warpStart = static_cast<int *>(h->prWarpStart.p);
  char *scratch = static_cast<char *>(h->result.p) + sizeof(apb_traversal_result);
  long long *totals = reinterpret_cast<long long *>(scratch);
  int *maxStagedDev = reinterpret_cast<int *>(scratch + 32), *overflowDev = reinterpret_cast<int *>(scratch + 36);
  int *maxCompactDev = reinterpret_cast<int *>(scratch + 40);
  APB_CUDA(cudaMemsetAsync(scratch + 32, 0, 36, h->stream));  // ... + the early-overflow word at scratch + 64
  APB_CUDA(cudaMemsetAsync(numStaged, 0, sizeof(int) * (numTiles + 1), h->stream));
  // scratch rows for the sets found by the counting pass: as many clusters as kPrunedMasks can stage in shared memory
  // (a larger set fails below anyway)
  const int earlyStride = static_cast<int>(200 * 1024 / (static_cast<size_t>(M) * 16 + 12)) + 1;
  int *early = nullptr;
  if (static_cast<size_t>(numTiles) * earlyStride * sizeof(int) <= (size_t(1) << 30)) {
    APB_CHECK(apbEnsure(h, h->prStageEarly, sizeof(int) * static_cast<size_t>(numTiles) * earlyStride));
    early = static_cast<int *>(h->prStageEarly.p);
  }
  ++h->launchCount, kPrunedStage<false><<<numTiles, PR_TILE, 0, h->stream>>>(a, numStaged, nullptr, nullptr, maxStagedDev, overflowDev, early, earlyStride);
  APB_CUDA(cudaGetLastError());
  APB_CHECK(apbExclusiveScan(h, numStaged, stagedStart, numTiles + 1, totals));
  long long totalStaged = 0;
  int hostMisc[4] = {0, 0, 0, 0};
  APB_CUDA(cudaMemcpyAsync(&totalStaged, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(hostMisc, scratch + 32, 8, cudaMemcpyDeviceToHost, h->stream));  // {largest set, overflow}
  int earlyOverflow = 0;
  APB_CUDA(cudaMemcpyAsync(&earlyOverflow, scratch + 64, 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (hostMisc[1]) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: a tile interacts with more than " +
                                                              std::to_string(PR_CAND_MAX) +
                                                              " cluster-list entries; use a larger cluster size");
  const int maxStaged = hostMisc[0];
  const size_t smemMasks = static_cast<size_t>(maxStaged) * M * 16 + static_cast<size_t>(maxStaged) * 12 + 16;
  if (smemMasks > 200 * 1024)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged tile (" + std::to_string(maxStaged * M) +
                                               " particles) does not fit shared memory; use a larger cluster size");
  if (totalStaged * M > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged sets exceed 2^31 slots");
  h->prunedMaxStaged = maxStaged;
  const long long stagedAlloc = std::max<long long>(totalStaged, 1);
  APB_CHECK(apbEnsure(h, h->prStaged, sizeof(int) * stagedAlloc));
  APB_CHECK(apbEnsure(h, h->prUsed, sizeof(int) * stagedAlloc));
  APB_CHECK(apbEnsure(h, h->prCbase, sizeof(int) * stagedAlloc));
  APB_CHECK(apbEnsure(h, h->prCompactSlot, sizeof(int) * stagedAlloc * M));
  APB_CHECK(apbEnsure(h, h->prMasks, sizeof(unsigned) * static_cast<size_t>(h->numPairs + h->numClusters + 1) * M));
  int *staged = static_cast<int *>(h->prStaged.p);
  if (early && !earlyOverflow)
    ++h->launchCount, kPrunedStageCompact<<<apbDivUp(static_cast<int64_t>(numTiles) * 32, 256), 256, 0, h->stream>>>(
        numTiles, numStaged, stagedStart, early, earlyStride, staged);
  else
    ++h->launchCount, kPrunedStage<true><<<numTiles, PR_TILE, 0, h->stream>>>(a, numStaged, stagedStart, staged, maxStagedDev, overflowDev, nullptr, 0);
  APB_CUDA(cudaGetLastError());
  MaskOut o;
  o.masks = static_cast<unsigned *>(h->prMasks.p);
  APB_CHECK(apbEnsure(h, h->prEntryLo, sizeof(int) * static_cast<size_t>(h->numPairs + h->numClusters + 2)));
  o.entryLo = static_cast<int *>(h->prEntryLo.p);
  o.warpRows = warpRows;
  o.used = static_cast<unsigned *>(h->prUsed.p);
  o.cbase = static_cast<int *>(h->prCbase.p);
  o.numCompact = static_cast<int *>(h->prNumCompact.p);
  o.maxCompact = maxCompactDev;
  o.totalEntries = reinterpret_cast<unsigned long long *>(scratch + 48);
  const bool uniform = M == 32;
  APB_CUDA(cudaMemsetAsync(warpRows, 0, sizeof(int) * (numWarps + 1), h->stream));
  if (uniform && n3)
    ++h->launchCount, kPrunedMasks<true, true><<<numTiles, PR_TILE, smemMasks, h->stream>>>(a, stagedStart, staged, o);
  else if (uniform)
    ++h->launchCount, kPrunedMasks<true, false><<<numTiles, PR_TILE, smemMasks, h->stream>>>(a, stagedStart, staged, o);
  else if (n3)
    ++h->launchCount, kPrunedMasks<false, true><<<numTiles, PR_TILE, smemMasks, h->stream>>>(a, stagedStart, staged, o);
  else
    ++h->launchCount, kPrunedMasks<false, false><<<numTiles, PR_TILE, smemMasks, h->stream>>>(a, stagedStart, staged, o);
  APB_CUDA(cudaGetLastError());
  APB_CHECK(apbExclusiveScan(h, warpRows, warpStart, numWarps + 1, totals));
  long long totalRows = 0;
  APB_CUDA(cudaMemcpyAsync(&totalRows, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(hostMisc + 2, maxCompactDev, 8, cudaMemcpyDeviceToHost, h->stream));  // + max rows
  unsigned long long totalEntries = 0;
  APB_CUDA(cudaMemcpyAsync(&totalEntries, scratch + 48, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (totalRows > 0x7fffffffLL / 128) return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: lists exceed 2^31 entries");
  const int maxCompact = hostMisc[2];
  if (maxCompact + 16 > 4096)  // list entries are 16-bit byte offsets (index * 16)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: a tile stages more than 4080 particles; use a smaller cluster size");
  if ((static_cast<size_t>(maxCompact) + 18) * 28 > 200 * 1024)
    return h->fail(APB_ERR_NOT_APPLICABLE, "gpuvcl_pruned: staged tile (" + std::to_string(maxCompact) +
                                               " particles) does not fit shared memory");
  h->prunedMaxCompact = maxCompact;
  // shared-memory layout of the force kernels: staged particles + 16 sentinel slots at the end (padding list entries
  // point at them, so the lists are built for one layout)
  h->prunedCap = maxCompact + 16 <= 1280 ? 1280 : (maxCompact + 16 <= 2048 ? 2048 : 4096);
  const int sentBase = h->prunedCap - 16;
  h->prunedRows = totalRows;
  h->prunedEntries = totalEntries;
  APB_CHECK(apbEnsure(h, h->prLists, sizeof(unsigned short) * 128 * (totalRows + 8)));  // 8 rows of slack: unchecked prefetch
  // per-warp list blocks in shared memory if the longest list fits (256 B per row)
  const int maxRows = hostMisc[3];
  const size_t smemFillBase = (static_cast<size_t>(maxStaged) * 12 + 15) & ~size_t(15);
  const int smemRowsPerWarp = smemFillBase + static_cast<size_t>(PR_WARPS) * (maxRows | 1) * 256 <= 150 * 1024 ? maxRows : 0;
  const size_t smemFill = smemFillBase + static_cast<size_t>(PR_WARPS) * (smemRowsPerWarp | 1) * 256 + 16;
  // Latin layout of the lists (kPrunedFill; lists assembled in shared memory only). Measured on B200 (C2, 1 M particles,
  // profiles/r02_latin_lists.txt): shared-memory wavefronts of kLJPruned 40.3 M -> 28.1 M, LSU data pipe 86 % -> 69 %,
  // FP64 pipe 53 % -> 60 %, kernel 0.197 -> 0.177 ms per step; but kPrunedFill pays for the placement (rebuild 1.44 ->
  // 2.0 ms). At the reference's rebuild frequency of 10 that is a net loss (0.341 -> 0.378 ms per step), so the layout
  // is opt-in: APB_LIST_SCHEDULE=1 (worth it from about 25 steps per rebuild).
  static const bool schedEnv = getenv("APB_LIST_SCHEDULE") != nullptr && atoi(getenv("APB_LIST_SCHEDULE")) != 0;
  const int sched = schedEnv;
  if (uniform)
    ++h->launchCount, kPrunedFill<true><<<numTiles, PR_TILE, smemFill, h->stream>>>(
        a, stagedStart, staged, o, warpStart, static_cast<unsigned short *>(h->prLists.p), static_cast<int *>(h->prCompactSlot.p), smemRowsPerWarp, sched,
        static_cast<int *>(h->prTileHalo.p), sentBase);
  else
    ++h->launchCount, kPrunedFill<false><<<numTiles, PR_TILE, smemFill, h->stream>>>(
        a, stagedStart, staged, o, warpStart, static_cast<unsigned short *>(h->prLists.p), static_cast<int *>(h->prCompactSlot.p), smemRowsPerWarp, sched,
        static_cast<int *>(h->prTileHalo.p), sentBase);
  APB_CUDA(cudaGetLastError());
  // order[0 .. numTiles) and, behind it, the number of interior tiles
  ++h->launchCount, kPrunedTileOrder<<<1, 1024, 0, h->stream>>>(numTiles, static_cast<const int *>(h->prTileHalo.p),
                                                               static_cast<int *>(h->prTileOrder.p),
                                                               static_cast<int *>(h->prTileOrder.p) + numTiles);
  APB_CUDA(cudaGetLastError());
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  h->prunedValid = true;
  if (getenv("APB_DEBUG"))
    fprintf(stderr, "[apb] pruned build: slots %lld tiles %d maxStagedClusters %d maxCompact %d totalStaged %lld rows %lld (entries %lld)\n",
            static_cast<long long>(n), numTiles, maxStaged, maxCompact, totalStaged, totalRows, totalRows * 128);
  return APB_OK;
}

// ---- force kernel ------------------------------------------------------------------------------------------------
struct PrunedForceArgs {
  int M, logM;
  const int *chunkFirst, *chunkNum;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type, *own;
  const int *stagedStart, *numCompact, *compactSlot, *warpRows, *warpRowStart;
  const unsigned short *lists;
  unsigned long long totalEntries;  // list entries written at build time = distance evaluations per call
  LJParams p;
  LJStats *partials;
  // split step: part 0 = all tiles in natural order; 1 = interior tiles order[0 .. *numInterior); 2 = boundary tiles
  // order[*numInterior .. numTiles). A CTA beyond its part's range exits at once. Partials are indexed by position.
  const int *tileOrder, *numInterior;
  int part, numTiles;
  int sentBase;  // first of the 16 sentinel slots of the staged layout (prunedCap - 16)
  // apb_run_steps: the force of an owned particle is stored as global force + pair sum instead of being added to a
  // column the integrator reset beforehand (one column pass less on each side; same value bit for bit)
  int overwrite;
  double gx, gy, gz;
};

// reciprocal from the hardware seed: MUFU.RCP64H (relative error ~2^-20, it reads the high word only) followed by one
// cubic step  r' = r + r (e + e^2),  e = 1 - x r  (3 DFMA, error ~e^3 = 2^-60, i.e. rounding level), no slow path.
// inf / NaN never occur for the sentinel distance (1e200), see below.
__device__ __forceinline__ double prRcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
}

#define PR_FAR 1e100  // x coordinate of the sentinels / of deleted partners: dr2 ~ 1e200 is finite and fails every cutoff

template <bool MIX, bool STATS, bool VIR3>
struct PairAcc {
  double fx = 0., fy = 0., fz = 0.;
  double sb2 = 0., sb = 0.;  // non-mixing: sums of b^2 and b over hits, b = 1 / dr2^3 (see prPair)
  double upot = 0., vt = 0.;  // mixing: sum of potentialEnergy6 and of the virial trace
  double vx = 0., vy = 0., vz = 0.;  // VIR3: virial per component
  unsigned dist = 0, hits = 0;
};

// One pair: 17 FP64-pipe instructions without globals, 19 with (non-mixing), +5 with per-component virial.
// Same quantities as LJFunctor.h:146-159 / :174-190, regrouped around b = invdr2^3 so that sigma and epsilon fold into
// two constants K1 = 2 eps24 sigma^12 and K2 = -eps24 sigma^6 (per type pair when mixing):
//   fac           = eps24 (lj12 + lj12m6) invdr2 = invdr2^4 (K1 b + K2)
//   Upot6         = eps24 (lj12 - lj6) + shift6  = b (K1/2 b + K2) + shift6
//   virial trace  = dr . f = dr2 fac             = b (K1 b + K2)              (dr2 invdr2^4 = b)
// so Upot and the virial trace (what LJFunctor::getVirial returns, LJFunctor.h:719) need only sum b^2 and sum b.
// Differences to the reference's operation order are at the 1e-16 level per pair.
// Cutoff decision: dr2 is evaluated with two FMAs (3 instead of 5 FP64 instructions) and compared through the high
// word of its bit pattern (non-negative doubles order like integers; off the FP64 pipe). The reference rounds every
// product and sum separately (LJFunctor.h:484-488) and the two values can differ by a few ulp, so the fast path only
// accepts pairs whose high word is below the one of cutoff^2 minus 1; a pair whose high word is within +-1 of it
// (relative distance to the cutoff < 3e-6: about one row in 10^4) raises `near`, and the caller re-evaluates that row
// the reference's way (prPairExact) - the decision stays bit-identical.
// A miss skips the (predicated) accumulation instead of multiplying by a mask.
// `e16` is the partner's index in the staged tile times 16 = the byte offset of its (x, y) pair; z lives in a second
// array behind the CAP pairs. One LDS.128 + one LDS.64 with immediate offsets instead of three LDS.64 at a 24-byte
// stride: one instruction less per pair and about 10 % fewer shared-memory wavefronts for the random gather.
template <bool MIX, bool STATS, bool VIR3>
__device__ __forceinline__ void prAccumulate(const LJParams &p, int ti, const int *stype, unsigned e16, double drx,
                                             double dry, double drz, double dx2, double b, double t, double fac,
                                             double k2, PairAcc<MIX, STATS, VIR3> &acc) {
  acc.fx = fma(drx, fac, acc.fx);
  acc.fy = fma(dry, fac, acc.fy);
  acc.fz = fma(drz, fac, acc.fz);
  if (STATS) {
    if (MIX) {
      const double2 *m = reinterpret_cast<const double2 *>(p.mix4) + 2 * (static_cast<size_t>(ti) * p.T + stype[e16 >> 4]);
      const double2 as = __ldg(m + 1);  // {K1 / 2, shift6}
      acc.upot = fma(b, fma(as.x, b, k2), acc.upot);
      acc.upot += as.y;
      if (!VIR3) acc.vt = fma(b, t, acc.vt);
    } else {
      acc.sb2 = fma(b, b, acc.sb2);
      acc.sb += b;
    }
    if (VIR3) {
      acc.vx = fma(dx2, fac, acc.vx);
      acc.vy = fma(dry * dry, fac, acc.vy);
      acc.vz = fma(drz * drz, fac, acc.vz);
    }
    ++acc.hits;
  }
}

// z of the staged particle behind list entry e16 (= index * 16).
// Default layout: one z array behind the CAP (x, y) pairs, element index * 8.
// PR_ZDUP (experiment): two interleaved copies, index c at ((c >> 3) * 16 + (c & 7)) * 8 and 64 bytes behind it; lanes
// 0-7 of each half-warp read the first copy (banks 0-15), lanes 8-15 the second (banks 16-31), so the LDS.64 of a
// half-warp is conflict-free whenever each quarter reads 8 different classes mod 8.
template <int CAP>
__device__ __forceinline__ double prLoadZ(const unsigned char *sxyz, unsigned e16) {
#ifdef PR_ZDUP
  return *reinterpret_cast<const double *>(sxyz + CAP * 16 + (threadIdx.x & 8u) * 8u + (e16 - ((e16 & 0x70u) >> 1)));
#else
  return *reinterpret_cast<const double *>(sxyz + CAP * 16 + (e16 >> 1));
#endif
}
__device__ __forceinline__ int prZIndex(int e) {  // element index (in doubles, first copy) of staged particle e
#ifdef PR_ZDUP
  return ((e >> 3) << 4) + (e & 7);
#else
  return e;
#endif
}

template <bool MIX, bool STATS, bool DEAD, bool VIR3, int CAP>
__device__ __forceinline__ void prPair(const LJParams &p, double xi, double yi, double zi, int ti,
                                       const unsigned char *sxyz, const int *stype, unsigned e16, unsigned sentinel16,
                                       PairAcc<MIX, STATS, VIR3> &acc, bool &near) {
#ifdef PR_EXP_NOCONFLICT
  e16 = (((e16 >> 4) & ~15u) | (threadIdx.x & 15u)) << 4;  // timing experiment only: conflict-free gather
#endif
  const double2 pxy = *reinterpret_cast<const double2 *>(sxyz + e16);
  const double pz = prLoadZ<CAP>(sxyz, e16);
  const double drx = xi - pxy.x, dry = yi - pxy.y, drz = zi - pz;
  const double dx2 = drx * drx;
  const double dr2 = fma(drz, drz, fma(dry, dry, dx2));
  const int band = __double2hiint(dr2) - p.cutHiLo;  // cutHiLo = high word of cutoff^2 minus 1
  const bool hit = band < 0;
  near |= static_cast<unsigned>(band) <= 2u;
  double k1, k2;
  if (MIX) {
    const double2 *m = reinterpret_cast<const double2 *>(p.mix4) + 2 * (static_cast<size_t>(ti) * p.T + stype[e16 >> 4]);
    const double2 k = __ldg(m);
    k1 = k.x;
    k2 = k.y;
  } else {
    k1 = p.k1;
    k2 = p.k2;
  }
  const double inv = prRcp(dr2);
  const double a2 = inv * inv;
  double b = a2 * inv;
  const double c = a2 * a2;
  const double t = fma(k1, b, k2);
  double fac = c * t;
  // keep the pair math unconditional (the four pairs of a row interleave and hide each other's FP64 latency); only the
  // accumulation below is predicated
  asm volatile("" : "+d"(fac), "+d"(b));
  if (STATS && DEAD) acc.dist += e16 < sentinel16;
  if (hit) prAccumulate<MIX, STATS, VIR3>(p, ti, stype, e16, drx, dry, drz, dx2, b, t, fac, k2, acc);
}

// the rare path: a pair inside the band around the cutoff that prPair left out, decided like the reference does
template <bool MIX, bool STATS, bool VIR3, int CAP>
__device__ __forceinline__ void prPairExact(const LJParams &p, double xi, double yi, double zi, int ti,
                                            const unsigned char *sxyz, const int *stype, unsigned e16,
                                            PairAcc<MIX, STATS, VIR3> &acc) {
  const double2 pxy = *reinterpret_cast<const double2 *>(sxyz + e16);
  const double pz = prLoadZ<CAP>(sxyz, e16);
  const double drx = xi - pxy.x, dry = yi - pxy.y, drz = zi - pz;
  const double dx2 = drx * drx;
  const double dr2 = fma(drz, drz, fma(dry, dry, dx2));
  if (static_cast<unsigned>(__double2hiint(dr2) - p.cutHiLo) > 2u) return;  // prPair has dealt with it
  const double dr2x = ljDist2(drx, dry, drz);
  if (__double_as_longlong(dr2x) > __double_as_longlong(p.cutoff2)) return;
  double k1 = p.k1, k2 = p.k2;
  if (MIX) {
    const double2 *m = reinterpret_cast<const double2 *>(p.mix4) + 2 * (static_cast<size_t>(ti) * p.T + stype[e16 >> 4]);
    const double2 k = __ldg(m);
    k1 = k.x;
    k2 = k.y;
  }
  const double inv = prRcp(dr2x);
  const double a2 = inv * inv;
  const double b = a2 * inv;
  const double t = fma(k1, b, k2);
  prAccumulate<MIX, STATS, VIR3>(p, ti, stype, e16, drx, dry, drz, dx2, b, t, (a2 * a2) * t, k2, acc);
}

__device__ __forceinline__ void prCpAsync8(void *smemDst, const void *gmemSrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smemDst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmemSrc) : "memory");
}

// DEAD: the ownership column changed since the list build (particles deleted / marked dummy). Those partners are moved
// out of reach while staging and distance evaluations are counted per pair; otherwise their number is the build-time
// entry count.
// CAP: staged particles (incl. the 16 sentinels) the shared-memory layout is compiled for: (x, y) pairs at 16 CAP bytes,
// z behind them, types behind those - compile-time offsets keep the gather free of address arithmetic.
template <bool MIX, bool STATS, bool DEAD, bool VIR3, int CAP>
__global__ void __launch_bounds__(PR_TILE, CAP <= 2048 ? PR_MINBLOCKS_CAP2048 : PR_MINBLOCKS_CAP4096) kLJPruned(PrunedForceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  unsigned char *sxyz = smemRaw;
  int *stype = reinterpret_cast<int *>(smemRaw + static_cast<size_t>(CAP) * PR_BYTES_XYZ);
  int pos = blockIdx.x;
  if (a.part != 0) {
    const int nI = *a.numInterior;
    if (a.part == 1 ? pos >= nI : (pos += nI) >= a.numTiles) return;
  }
  const int tile = a.part != 0 ? a.tileOrder[pos] : pos;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpGlobal = tile * PR_WARPS + warp;
  // Prologue, ordered so that independent global loads are in flight together: (0) the per-tile and per-warp table
  // entries, loaded unconditionally and ahead of the empty-tile test (all valid for every warp of the grid) so that they
  // travel together instead of one dependent round trip after the other, (1) this warp's list rows (prefetched four
  // ahead in a register ring; the list buffer has slack behind its end, so the prefetch needs no bounds check), (2) the
  // tile's slot table, (3) the staged positions as asynchronous 8-byte copies (LDGSTS, no register round trip), (4)
  // this lane's own particle.
  const int nP = __ldg(a.numCompact + tile);
  const int first = __ldg(a.chunkFirst + warpGlobal), rowsRaw = __ldg(a.warpRows + warpGlobal);
  const int rowStart = __ldg(a.warpRowStart + warpGlobal), chunkNum = __ldg(a.chunkNum + warpGlobal);
  const int stagedStart = __ldg(a.stagedStart + tile);
  const bool addEntries = STATS && !DEAD && pos == 0 && threadIdx.x == 0;
  if (nP == 0) {  // empty tile (block-uniform)
    if (STATS && threadIdx.x == 0) {
      LJStats st;
      ljStatsZero(st);
      if (addEntries) st.dist = a.totalEntries;
      a.partials[pos] = st;
    }
    return;
  }
#ifdef PR_EXP_NOLOOP
  const int rows = first >= 0 ? min(rowsRaw, PR_EXP_NOLOOP) : 0;  // timing experiment only
#else
  const int rows = first >= 0 ? rowsRaw : 0;
#endif
  const uint2 *list = reinterpret_cast<const uint2 *>(a.lists) + (rows > 0 ? static_cast<size_t>(rowStart) * 32 + lane : lane);
  uint2 q0 = __ldg(list), q1 = __ldg(list + 32), q2 = __ldg(list + 64), q3 = __ldg(list + 96);
  const int *cs = a.compactSlot + (static_cast<size_t>(stagedStart) << a.logM);
  double *sxy = reinterpret_cast<double *>(sxyz), *sz = reinterpret_cast<double *>(sxyz + CAP * 16);
  constexpr int PR_STAGE_UNROLL = 8;
  int slots[PR_STAGE_UNROLL];
#pragma unroll
  for (int k = 0; k < PR_STAGE_UNROLL; ++k) {
    const int e = threadIdx.x + k * PR_TILE;
    slots[k] = e < nP ? __ldg(cs + e) : -1;
  }
  const int64_t i = static_cast<int64_t>(first >= 0 ? first : 0) + lane;
  const bool active = rows > 0 && lane < chunkNum && a.own[i] == APB_OWN_OWNED;
  // a slot without an owned particle (its rows hold padding only, unless it was deleted after the build) sits far away
  const double xi = active ? a.x[i] : 0.5 * PR_FAR, yi = active ? a.y[i] : 0., zi = active ? a.z[i] : 0.;
  const int ti = (MIX && active) ? a.type[i] : 0;
#pragma unroll
  for (int k = 0; k < PR_STAGE_UNROLL; ++k) {
    const int e = threadIdx.x + k * PR_TILE;
    if (slots[k] >= 0) {
      prCpAsync8(sxy + 2 * e, a.x + slots[k]);
      prCpAsync8(sxy + 2 * e + 1, a.y + slots[k]);
      prCpAsync8(sz + prZIndex(e), a.z + slots[k]);
#ifdef PR_ZDUP
      prCpAsync8(sz + prZIndex(e) + 8, a.z + slots[k]);
#endif
      if (MIX) stype[e] = a.type[slots[k]];
    }
  }
  for (int e = threadIdx.x + PR_STAGE_UNROLL * PR_TILE; e < nP; e += PR_TILE) {
    const int slot = __ldg(cs + e);
    prCpAsync8(sxy + 2 * e, a.x + slot);
    prCpAsync8(sxy + 2 * e + 1, a.y + slot);
    prCpAsync8(sz + prZIndex(e), a.z + slot);
#ifdef PR_ZDUP
    prCpAsync8(sz + prZIndex(e) + 8, a.z + slot);
#endif
    if (MIX) stype[e] = a.type[slot];
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const unsigned sentinel16 = static_cast<unsigned>(a.sentBase) << 4;
  if (threadIdx.x < 16) {  // sentinel slots for padding entries, one per bank class
    sxy[2 * (a.sentBase + threadIdx.x)] = PR_FAR;
    sxy[2 * (a.sentBase + threadIdx.x) + 1] = 0.;
    sz[prZIndex(a.sentBase + threadIdx.x)] = 0.;
#ifdef PR_ZDUP
    sz[prZIndex(a.sentBase + threadIdx.x) + 8] = 0.;
#endif
    if (MIX) stype[a.sentBase + threadIdx.x] = 0;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (DEAD) {
    for (int e = threadIdx.x; e < nP; e += PR_TILE)
      if (a.own[cs[e]] == APB_OWN_DUMMY) sxy[2 * e] = PR_FAR;
    __syncthreads();
  }
  PairAcc<MIX, STATS, VIR3> acc;
#define PR_ROW(Q)                                                                                              \
  do {                                                                                                         \
    bool near = false;                                                                                         \
    prPair<MIX, STATS, DEAD, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).x & 0xFFFFu), sentinel16, acc, near); \
    prPair<MIX, STATS, DEAD, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).x >> 16), sentinel16, acc, near);     \
    prPair<MIX, STATS, DEAD, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).y & 0xFFFFu), sentinel16, acc, near); \
    prPair<MIX, STATS, DEAD, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).y >> 16), sentinel16, acc, near);     \
    if (near) {                                                                                                \
      prPairExact<MIX, STATS, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).x & 0xFFFFu), acc);                 \
      prPairExact<MIX, STATS, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).x >> 16), acc);                     \
      prPairExact<MIX, STATS, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).y & 0xFFFFu), acc);                 \
      prPairExact<MIX, STATS, VIR3, CAP>(a.p, xi, yi, zi, ti, sxyz, stype, ((Q).y >> 16), acc);                     \
    }                                                                                                          \
  } while (0)
  int r = 0;
  list += 128;
  for (; r + 4 <= rows; r += 4, list += 128) {
    PR_ROW(q0);
    q0 = __ldg(list);
    PR_ROW(q1);
    q1 = __ldg(list + 32);
    PR_ROW(q2);
    q2 = __ldg(list + 64);
    PR_ROW(q3);
    q3 = __ldg(list + 96);
  }
  if (r < rows) {
    PR_ROW(q0);
    if (r + 1 < rows) {
      PR_ROW(q1);
      if (r + 2 < rows) PR_ROW(q2);
    }
  }
  if (active) {
    if (a.overwrite) {
      a.fx[i] = a.gx + acc.fx;
      a.fy[i] = a.gy + acc.fy;
      a.fz[i] = a.gz + acc.fz;
    } else {
      // single writer per slot: fire-and-forget RED.ADD.F64 instead of a load / add / store round trip
      atomicAdd(a.fx + i, acc.fx);
      atomicAdd(a.fy + i, acc.fy);
      atomicAdd(a.fz + i, acc.fz);
    }
  }
  if (STATS) {
    // Block sums of the raw accumulators only (2 doubles + the hit count; 5 with the per-component virial) instead of the
    // nine fields of LJStats: fixed butterfly order inside a warp, fixed order over the warps - reproducible run to run.
    // potentialEnergy6 = eps24 * (lj12 - lj6) + shift6 (LJFunctor.h:174); only owned particles carry lists: weight 1.
    // Non-mixing: Upot = K1/2 sum b^2 + K2 sum b + hits shift6; the trace of the virial is K1 sum b^2 + K2 sum b.
    constexpr int NV = VIR3 ? 5 : 2;
    __shared__ double sRedD[PR_WARPS][NV];
    __shared__ unsigned sRedU[PR_WARPS][2];
    double v[NV];
    v[0] = MIX ? acc.upot : acc.sb2;
    v[1] = MIX ? acc.vt : acc.sb;
    if (VIR3) v[2] = acc.vx, v[3] = acc.vy, v[4] = acc.vz;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    const unsigned hitsW = __reduce_add_sync(0xffffffffu, acc.hits);
    const unsigned distW = DEAD ? __reduce_add_sync(0xffffffffu, acc.dist) : 0u;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) sRedD[warp][k] = v[k];
      sRedU[warp][0] = hitsW;
      sRedU[warp][1] = distW;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t[NV];
      unsigned long long hits = 0ULL, dist = 0ULL;
#pragma unroll
      for (int k = 0; k < NV; ++k) t[k] = 0.;
#pragma unroll
      for (int w = 0; w < PR_WARPS; ++w) {
#pragma unroll
        for (int k = 0; k < NV; ++k) t[k] += sRedD[w][k];
        hits += sRedU[w][0];
        dist += sRedU[w][1];
      }
      LJStats st;
      ljStatsZero(st);
      st.upot = MIX ? t[0] : fma(0.5 * a.p.k1, t[0], fma(a.p.k2, t[1], static_cast<double>(hits) * a.p.shift6));
      st.vir[0] = VIR3 ? t[2] : (MIX ? t[1] : fma(a.p.k1, t[0], a.p.k2 * t[1]));
      st.vir[1] = VIR3 ? t[3] : 0.;
      st.vir[2] = VIR3 ? t[4] : 0.;
      st.dist = DEAD ? dist : (addEntries ? a.totalEntries : 0ULL);
      st.kNoN3 = hits;
      st.gNoN3 = hits;
      a.partials[pos] = st;
    }
  }
}


#undef PR_ROW

// ---- newton3 variant ------------------------------------------------------------------------------------------------
// Same tiles, staging and list format; the lists hold every owned-owned pair once (kPrunedMasks<*, true>). A hit adds
// f to the lane's own particle in registers and -f to the partner through three RED.ADD.F64 on its force columns
// (LJFunctor::SoAFunctorPairImpl<true>: LJFunctor.h:499-516); halo partners receive nothing. Globals carry the
// reference's weight [i owned] + [j owned] (LJFunctor.h:525-527), i.e. 2 for an owned partner and 1 for a halo copy, so
// they equal the newton3-off sums; every hit counts as a newton3 kernel call. Half the pair evaluations of kLJPruned,
// but the scatter costs more than the arithmetic saves on this machine (REDs: ~1.3 cycles per lane and SM against 0.3
// cycles per pair of FP64 work; measured in profiles/r02_newton3.txt) - the option exists so that the tuner can see it.
template <bool MIX, bool STATS, int CAP>
__global__ void __launch_bounds__(PR_TILE, CAP <= 2048 ? 3 : 1) kLJPrunedN3(PrunedForceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  unsigned char *sxyz = smemRaw;
  int *sslot = reinterpret_cast<int *>(smemRaw + static_cast<size_t>(CAP) * PR_BYTES_XYZ);  // owned partner: slot, else -1
  int *stype = sslot + CAP;
  int pos = blockIdx.x;
  if (a.part != 0) {
    const int nI = *a.numInterior;
    if (a.part == 1 ? pos >= nI : (pos += nI) >= a.numTiles) return;
  }
  const int tile = a.part != 0 ? a.tileOrder[pos] : pos;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpGlobal = tile * PR_WARPS + warp;
  const int nP = a.numCompact[tile];
  if (nP == 0) {
    if (STATS && threadIdx.x == 0) {
      LJStats st;
      ljStatsZero(st);
      a.partials[pos] = st;
    }
    return;
  }
  const int first = a.chunkFirst[warpGlobal];
  const int rows = first >= 0 ? a.warpRows[warpGlobal] : 0;
  const uint2 *list = reinterpret_cast<const uint2 *>(a.lists) +
                      (rows > 0 ? static_cast<size_t>(a.warpRowStart[warpGlobal]) * 32 + lane : lane);
  const int *cs = a.compactSlot + (static_cast<size_t>(a.stagedStart[tile]) << a.logM);
  double *sxy = reinterpret_cast<double *>(sxyz), *sz = reinterpret_cast<double *>(sxyz + CAP * 16);
  for (int e = threadIdx.x; e < nP; e += PR_TILE) {
    const int slot = __ldg(cs + e);
    const int o = a.own[slot];
    sxy[2 * e] = o == APB_OWN_DUMMY ? PR_FAR : a.x[slot];  // deleted since the list build: out of reach
    sxy[2 * e + 1] = a.y[slot];
    sz[prZIndex(e)] = a.z[slot];
#ifdef PR_ZDUP
    sz[prZIndex(e) + 8] = a.z[slot];
#endif
    sslot[e] = o == APB_OWN_OWNED ? slot : -1;
    if (MIX) stype[e] = a.type[slot];
  }
  if (threadIdx.x < 16) {
    sxy[2 * (a.sentBase + threadIdx.x)] = PR_FAR;
    sxy[2 * (a.sentBase + threadIdx.x) + 1] = 0.;
    sz[prZIndex(a.sentBase + threadIdx.x)] = 0.;
#ifdef PR_ZDUP
    sz[prZIndex(a.sentBase + threadIdx.x) + 8] = 0.;
#endif
    sslot[a.sentBase + threadIdx.x] = -1;
    if (MIX) stype[a.sentBase + threadIdx.x] = 0;
  }
  const int64_t i = static_cast<int64_t>(first >= 0 ? first : 0) + lane;
  const bool active = rows > 0 && lane < a.chunkNum[warpGlobal] && a.own[i] == APB_OWN_OWNED;
  const double xi = active ? a.x[i] : 0.5 * PR_FAR, yi = active ? a.y[i] : 0., zi = active ? a.z[i] : 0.;
  const int ti = (MIX && active) ? a.type[i] : 0;
  __syncthreads();
  const unsigned sentinel16 = static_cast<unsigned>(a.sentBase) << 4;
  double fx = 0., fy = 0., fz = 0.;
  double upot = 0., vx = 0., vy = 0., vz = 0.;
  unsigned dist = 0, hits = 0;
  for (int r = 0; r < rows; ++r, list += 32) {
    const uint2 q = __ldg(list);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const unsigned e16 = s == 0 ? (q.x & 0xFFFFu) : (s == 1 ? (q.x >> 16) : (s == 2 ? (q.y & 0xFFFFu) : (q.y >> 16)));
      const double2 pxy = *reinterpret_cast<const double2 *>(sxyz + e16);
      const double pz = prLoadZ<CAP>(sxyz, e16);
      const double drx = xi - pxy.x, dry = yi - pxy.y, drz = zi - pz;
      double dr2 = fma(drz, drz, fma(dry, dry, drx * drx));
      const int band = __double2hiint(dr2) - a.p.cutHiLo;
      bool hit = band < 0;
      if (static_cast<unsigned>(band) <= 2u) {  // within a few ulp of the cutoff: decide like the reference (prPairExact)
        dr2 = ljDist2(drx, dry, drz);
        hit = __double_as_longlong(dr2) <= __double_as_longlong(a.p.cutoff2);
      }
      if (STATS) dist += e16 < sentinel16;
      if (hit) {
        double k1 = a.p.k1, k2 = a.p.k2, khalf = 0.5 * a.p.k1, shift6 = a.p.shift6;
        if (MIX) {
          const double2 *m = reinterpret_cast<const double2 *>(a.p.mix4) + 2 * (static_cast<size_t>(ti) * a.p.T + stype[e16 >> 4]);
          const double2 k = __ldg(m), as = __ldg(m + 1);
          k1 = k.x, k2 = k.y, khalf = as.x, shift6 = as.y;
        }
        const double inv = prRcp(dr2);
        const double a2 = inv * inv;
        const double b = a2 * inv;
        const double fac = (a2 * a2) * fma(k1, b, k2);
        const double px = drx * fac, py = dry * fac, pzf = drz * fac;
        fx += px, fy += py, fz += pzf;
        const int sj = sslot[e16 >> 4];
        if (sj >= 0) {
          atomicAdd(a.fx + sj, -px);
          atomicAdd(a.fy + sj, -py);
          atomicAdd(a.fz + sj, -pzf);
        }
        if (STATS) {
          const double w = sj >= 0 ? 2. : 1.;
          upot = fma(w, fma(b, fma(khalf, b, k2), shift6), upot);
          vx = fma(w * drx, px, vx);
          vy = fma(w * dry, py, vy);
          vz = fma(w * drz, pzf, vz);
          ++hits;
        }
      }
    }
  }
  if (active) {
    atomicAdd(a.fx + i, fx);
    atomicAdd(a.fy + i, fy);
    atomicAdd(a.fz + i, fz);
  }
  if (STATS) {
    LJStats st;
    ljStatsZero(st);
    st.upot = upot;
    st.vir[0] = vx;
    st.vir[1] = vy;
    st.vir[2] = vz;
    st.dist = dist;
    st.kN3 = hits;
    st.gN3 = hits;
    ljStatsBlockReduce(st, a.partials, pos);
  }
}

int apbFinishStats(apb_handle h, int numBlocks, bool stats, const apb_functor *f, apb_traversal_result *out);

int apbComputeLJPruned(apb_handle h, const apb_functor *f, const LJParams &p, bool mix, bool stats, bool n3,
                       apb_traversal_result *out) {
  if (!h->prunedValid || h->prunedNewton3 != (n3 ? 1 : 0)) APB_CHECK(apbBuildPruned(h, n3 ? 1 : 0));
  const int numTiles = h->prunedTiles;
  if (numTiles == 0) return apbFinishStats(h, 0, stats, f, out);
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
  a.overwrite = (h->forceOverwrite && !n3) ? 1 : 0;
  a.gx = h->forceG[0];
  a.gy = h->forceG[1];
  a.gz = h->forceG[2];
  a.type = h->type;
  a.own = h->own;
  a.stagedStart = static_cast<const int *>(h->prStagedStart.p);
  a.numCompact = static_cast<const int *>(h->prNumCompact.p);
  a.compactSlot = static_cast<const int *>(h->prCompactSlot.p);
  a.warpRows = static_cast<const int *>(h->prWarpLen.p);
  a.warpRowStart = static_cast<const int *>(h->prWarpStart.p);
  a.lists = static_cast<const unsigned short *>(h->prLists.p);
  a.totalEntries = h->prunedEntries;
  a.p = p;
  // shared-memory layout compiled for 2048 or 4096 staged particles (4 or 2 CTAs per SM)
  const int cap = h->prunedCap <= 2048 ? 2048 : 4096;
  a.sentBase = h->prunedCap - 16;
  const size_t smem = static_cast<size_t>(cap) * (mix ? PR_BYTES_XYZ + 4 : PR_BYTES_XYZ);
  // per-component virial only on request: LJFunctor exposes the sum alone (getVirial, LJFunctor.h:719)
  const bool vir3 = stats && !(f->flags & APB_FUNCTOR_VIRIAL_TRACE);
  const int sel = (mix ? 8 : 0) | (stats ? 4 : 0) | (h->ownDirty ? 2 : 0) | (vir3 ? 1 : 0);
  const int part = h->prunedPart;  // set by apb_run_steps around the two halves of a split step
  a.part = part;
  a.numTiles = numTiles;
  a.tileOrder = static_cast<const int *>(h->prTileOrder.p);
  a.numInterior = a.tileOrder + numTiles;
  const int numBlocks = numTiles;
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * numBlocks));
  a.partials = static_cast<LJStats *>(h->partials.p);
  if (n3) {
    const size_t smemN3 = static_cast<size_t>(cap) * (PR_BYTES_XYZ + (mix ? 8 : 4));
#define PR_LAUNCH_N3(MIXV, STATSV)                                                                                      \
  do {                                                                                                                  \
    if (cap == 2048)                                                                                                    \
      ++h->launchCount, kLJPrunedN3<MIXV, STATSV, 2048><<<numTiles, PR_TILE, smemN3, h->stream>>>(a);                   \
    else                                                                                                                \
      ++h->launchCount, kLJPrunedN3<MIXV, STATSV, 4096><<<numTiles, PR_TILE, smemN3, h->stream>>>(a);                   \
  } while (0)
    if (mix && stats) PR_LAUNCH_N3(true, true);
    else if (mix) PR_LAUNCH_N3(true, false);
    else if (stats) PR_LAUNCH_N3(false, true);
    else PR_LAUNCH_N3(false, false);
#undef PR_LAUNCH_N3
    APB_CUDA(cudaGetLastError());
    if (part == 1) return APB_OK;
    return apbFinishStats(h, numBlocks, stats, f, out);
  }
#define PR_LAUNCH_CAP(MIXV, STATSV, DEADV, VIRV, CAPV)                                                               \
  do {                                                                                                               \
    ++h->launchCount, kLJPruned<MIXV, STATSV, DEADV, VIRV, CAPV><<<numTiles, PR_TILE, smem, h->stream>>>(a);         \
  } while (0)
#define PR_LAUNCH(MIXV, STATSV, DEADV, VIRV)                                                                         \
  do {                                                                                                               \
    if (cap == 2048)                                                                                                 \
      PR_LAUNCH_CAP(MIXV, STATSV, DEADV, VIRV, 2048);                                                                \
    else                                                                                                             \
      PR_LAUNCH_CAP(MIXV, STATSV, DEADV, VIRV, 4096);                                                                \
  } while (0)
  switch (sel) {
    case 0: PR_LAUNCH(false, false, false, false); break;
    case 2: PR_LAUNCH(false, false, true, false); break;
    case 4: PR_LAUNCH(false, true, false, false); break;
    case 5: PR_LAUNCH(false, true, false, true); break;
    case 6: PR_LAUNCH(false, true, true, false); break;
    case 7: PR_LAUNCH(false, true, true, true); break;
    case 8: PR_LAUNCH(true, false, false, false); break;
    case 10: PR_LAUNCH(true, false, true, false); break;
    case 12: PR_LAUNCH(true, true, false, false); break;
    case 13: PR_LAUNCH(true, true, false, true); break;
    case 14: PR_LAUNCH(true, true, true, false); break;
    default: PR_LAUNCH(true, true, true, true); break;
  }
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      h->poisoned = true;
      return h->fail(APB_ERR_CUDA, std::string("kLJPruned launch failed: ") + cudaGetErrorString(e) + " (tiles " +
                                       std::to_string(numTiles) + ", dynamic smem " + std::to_string(smem) +
                                       " B, staged particles max " + std::to_string(h->prunedMaxCompact) + ")");
    }
  }
  if (part == 1) return APB_OK;  // the boundary half follows and finishes the statistics
  return apbFinishStats(h, numBlocks, stats, f, out);
}

// Opt-in shared-memory sizes, set once per handle (apb_create) instead of before every launch.
int apbInitPrunedAttributes(apb_handle h) {
  const int big = 200 * 1024 + 1024;
  APB_CUDA(cudaFuncSetAttribute(kPrunedMasks<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  APB_CUDA(cudaFuncSetAttribute(kPrunedMasks<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  APB_CUDA(cudaFuncSetAttribute(kPrunedMasks<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  APB_CUDA(cudaFuncSetAttribute(kPrunedMasks<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  APB_CUDA(cudaFuncSetAttribute(kPrunedFill<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  APB_CUDA(cudaFuncSetAttribute(kPrunedFill<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
#define PR_ATTR(MIXV, STATSV, DEADV, VIRV)                                                                              \
  APB_CUDA(cudaFuncSetAttribute(kLJPruned<MIXV, STATSV, DEADV, VIRV, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                2048 * (PR_BYTES_XYZ + 4)));                                                                            \
  APB_CUDA(cudaFuncSetAttribute(kLJPruned<MIXV, STATSV, DEADV, VIRV, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                4096 * (PR_BYTES_XYZ + 4)))
  PR_ATTR(false, false, false, false);
  PR_ATTR(false, false, true, false);
  PR_ATTR(false, true, false, false);
  PR_ATTR(false, true, false, true);
  PR_ATTR(false, true, true, false);
  PR_ATTR(false, true, true, true);
  PR_ATTR(true, false, false, false);
  PR_ATTR(true, false, true, false);
  PR_ATTR(true, true, false, false);
  PR_ATTR(true, true, false, true);
  PR_ATTR(true, true, true, false);
  PR_ATTR(true, true, true, true);
#undef PR_ATTR
#define PR_ATTR_N3(MIXV, STATSV)                                                                                       \
  APB_CUDA(cudaFuncSetAttribute(kLJPrunedN3<MIXV, STATSV, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                2048 * (PR_BYTES_XYZ + 8)));                                                           \
  APB_CUDA(cudaFuncSetAttribute(kLJPrunedN3<MIXV, STATSV, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                4096 * (PR_BYTES_XYZ + 8)))
  PR_ATTR_N3(false, false);
  PR_ATTR_N3(false, true);
  PR_ATTR_N3(true, false);
  PR_ATTR_N3(true, true);
#undef PR_ATTR_N3
  return APB_OK;
}
