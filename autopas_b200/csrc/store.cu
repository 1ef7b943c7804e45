// Device-resident SoA particle storage, host-mirror transfers and container maintenance behind the C ABI.
// Reference semantics restated here (not code): ParticleContainerInterface storage virtuals
// (containers/ParticleContainerInterface.h:112-352), LeavingParticleCollector.h:85-118,
// LinkedCells.h:152-202, VerletClusterLists.h:362-397.
#include <algorithm>
#include <cmath>
#include <new>

#include "internal.cuh"

static std::string g_createError;

// ------------------------------------------------------------------------------------------------------------------
// memory helpers
// ------------------------------------------------------------------------------------------------------------------
// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold lifted in
// apb_create, so that growing a buffer never costs a device-wide synchronisation (cudaFree) or a trip to the driver's
// page allocator once the pool is warm: the AutoTuner times every rebuild, the first ones included
// (LogicHandler.h:1066-1141). A first allocation carries 25 % headroom: sizes that follow the particle configuration
// (list rows, halo copies, staged sets) drift by a few per cent from rebuild to rebuild.
int apbEnsure(apb_handle h, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return APB_OK;
  size_t want = std::max(bytes + bytes / 4, b.cap + b.cap / 2);
  want = (want + 255) & ~size_t(255);
  void *q = nullptr;
  cudaError_t e = cudaMallocAsync(&q, want, h->stream);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    return h->fail(APB_ERR_OUT_OF_MEMORY, "device allocation of " + std::to_string(want) + " bytes failed");
  }
  APB_CUDA(e);
  if (b.p) APB_CUDA(cudaFreeAsync(b.p, h->stream));  // stream-ordered: earlier kernels may still read the old block
  b.p = q;
  b.cap = want;
  ++h->allocCount;
  return APB_OK;
}

int apbEnsurePinned(apb_handle h, size_t bytes) {
  if (bytes <= h->pinnedCap) return APB_OK;
  if (h->pinned) APB_CUDA(cudaFreeHost(h->pinned));
  h->pinned = nullptr;
  h->pinnedCap = 0;
  size_t want = std::max(bytes, h->pinnedCap + h->pinnedCap / 2);
  APB_CUDA(cudaMallocHost(&h->pinned, want));
  h->pinnedCap = want;
  return APB_OK;
}

template <class T>
static int growArray(apb_handle h, T *&p, int64_t oldN, int64_t newCap) {
  T *q = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&q), sizeof(T) * newCap, h->stream);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    return h->fail(APB_ERR_OUT_OF_MEMORY, "device allocation of particle column failed");
  }
  APB_CUDA(e);
  if (p && oldN > 0) APB_CUDA(cudaMemcpyAsync(q, p, sizeof(T) * oldN, cudaMemcpyDeviceToDevice, h->stream));
  if (p) APB_CUDA(cudaFreeAsync(p, h->stream));
  p = q;
  ++h->allocCount;
  return APB_OK;
}

int apbReserveSlots(apb_handle h, int64_t slots) {
  if (slots <= h->cap) return APB_OK;
  int64_t newCap = std::max<int64_t>(slots + slots / 4, h->cap + h->cap / 2);
  newCap = (newCap + 1023) & ~int64_t(1023);
  for (int c = 0; c < APB_NUM_COLUMNS; ++c) {
    if (!h->active[c]) continue;
    APB_CHECK(growArray(h, h->col[c], h->nslots, newCap));
    APB_CHECK(growArray(h, h->colTmp[c], 0, newCap));
  }
  APB_CHECK(growArray(h, h->id, h->nslots, newCap));
  APB_CHECK(growArray(h, h->idTmp, 0, newCap));
  APB_CHECK(growArray(h, h->type, h->nslots, newCap));
  APB_CHECK(growArray(h, h->typeTmp, 0, newCap));
  APB_CHECK(growArray(h, h->own, h->nslots, newCap));
  APB_CHECK(growArray(h, h->ownTmp, 0, newCap));
  h->cap = newCap;
  return APB_OK;
}

int apbInitKernelAttributes(apb_handle h) {
  APB_CHECK(apbInitBuildAttributes(h));
  APB_CHECK(apbInitPrunedAttributes(h));
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// exclusive scan (int32), three-phase, recursive over block sums. HBM-bound: reads n, writes n (+ n/1024 sums).
// ------------------------------------------------------------------------------------------------------------------
#define SCAN_BLOCK 1024
// Exclusive scan in ONE launch for any length: chained tiles with decoupled look-back. A block takes the next tile by
// ticket (atomicInc wraps to zero after the last tile, so the counter needs no reset), scans its 4096 elements with
// coalesced 16-byte loads, publishes {epoch, flag, value} for its tile and walks back over its predecessors' entries
// until it meets an inclusive prefix. Entries of earlier scans carry an older epoch and read as "not ready", so the
// status array is never cleared. (Round 1 scanned up to 2^18 elements with one block - strided, uncoalesced, 43 us for
// the 60 k clusters of a 2 M-particle container, nine times per rebuild - and longer arrays with four launches.)
#define SCAN_TILE 4096
__device__ __forceinline__ unsigned long long scanPack(unsigned long long epoch, unsigned flag, int value) {
  return (epoch << 34) | (static_cast<unsigned long long>(flag) << 32) | static_cast<unsigned>(value);
}
__global__ void __launch_bounds__(1024) kScanChained(const int *__restrict__ in, int *__restrict__ out, int64_t n,
                                                     long long *total, unsigned long long *status, unsigned *ticket,
                                                     unsigned numTiles, unsigned long long epoch) {
  __shared__ int warpSums[32];
  __shared__ unsigned sTile;
  __shared__ int sPrefix;
  if (threadIdx.x == 0) sTile = atomicInc(ticket, numTiles - 1);
  __syncthreads();
  const unsigned tile = sTile;
  const int64_t base = static_cast<int64_t>(tile) * SCAN_TILE + 4 * threadIdx.x;
  int v[4] = {0, 0, 0, 0};
  if (base + 3 < n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    const int4 q = *reinterpret_cast<const int4 *>(in + base);
    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
  } else {
    for (int k = 0; k < 4; ++k)
      if (base + k < n) v[k] = in[base + k];
  }
  const int sum = v[0] + v[1] + v[2] + v[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warpSums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warpSums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warpSums[lane] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tileTotal = warpSums[31];
    int prefix = 0;
    volatile unsigned long long *st = status;
    if (tile == 0) {
      st[0] = scanPack(epoch, 2u, tileTotal);
    } else {
      st[tile] = scanPack(epoch, 1u, tileTotal);
      __threadfence();
      for (int t = static_cast<int>(tile) - 1;; --t) {
        unsigned long long w;
        do {
          w = st[t];
        } while ((w >> 34) != epoch || ((w >> 32) & 3u) == 0u);
        prefix += static_cast<int>(static_cast<unsigned>(w));
        if (((w >> 32) & 3u) == 2u) break;
      }
      st[tile] = scanPack(epoch, 2u, prefix + tileTotal);
    }
    sPrefix = prefix;
    if (total && tile == numTiles - 1) *total = static_cast<long long>(prefix) + tileTotal;
  }
  __syncthreads();
  int run = sPrefix + (warp == 0 ? 0 : warpSums[warp - 1]) + incl - sum;
  if (base + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    int4 q;
    q.x = run;
    q.y = run + v[0];
    q.z = q.y + v[1];
    q.w = q.z + v[2];
    *reinterpret_cast<int4 *>(out + base) = q;
  } else {
    for (int k = 0; k < 4; ++k)
      if (base + k < n) {
        out[base + k] = run;
        run += v[k];
      }
  }
}

int apbExclusiveScan(apb_handle h, const int *in, int *out, int64_t n, long long *totalDev) {
  if (n <= 0) {
    if (totalDev) APB_CUDA(cudaMemsetAsync(totalDev, 0, sizeof(long long), h->stream));
    return APB_OK;
  }
  const int64_t numTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  const size_t need = sizeof(unsigned long long) * (numTiles + 2);
  if (h->scanTmp.cap < need) {  // status array + ticket counter behind it; zero = an epoch no scan ever uses
    const size_t before = h->scanTmp.cap;
    APB_CHECK(apbEnsure(h, h->scanTmp, need + 4096));
    (void)before;
    APB_CUDA(cudaMemsetAsync(h->scanTmp.p, 0, h->scanTmp.cap, h->stream));
    h->scanTiles = static_cast<int64_t>(h->scanTmp.cap / sizeof(unsigned long long)) - 1;
  }
  unsigned long long *status = static_cast<unsigned long long *>(h->scanTmp.p);
  unsigned *ticket = reinterpret_cast<unsigned *>(status + h->scanTiles);
  if (((++h->scanEpoch) & ((1ULL << 30) - 1)) == 0) ++h->scanEpoch;  // zero is the cleared state
  ++h->launchCount, kScanChained<<<static_cast<unsigned>(numTiles), 1024, 0, h->stream>>>(
      in, out, n, totalDev, status, ticket, static_cast<unsigned>(numTiles), h->scanEpoch & ((1ULL << 30) - 1));
  APB_CUDA(cudaGetLastError());
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// life cycle
// ------------------------------------------------------------------------------------------------------------------
extern "C" const char *apb_last_error(apb_handle h) { return h ? h->err.c_str() : g_createError.c_str(); }

extern "C" int apb_create(const apb_config *config, apb_handle *out) {
  if (!config || !out) {
    g_createError = "apb_create: null argument";
    return APB_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  const apb_config &c = *config;
  for (int d = 0; d < 3; ++d) {
    if (!(c.box_max[d] > c.box_min[d])) {
      g_createError = "apb_create: box_max must be larger than box_min";
      return APB_ERR_INVALID_ARGUMENT;
    }
    // LogicHandler::checkMinimalSize (LogicHandler.h:968-979): box must be at least one interaction length
    if (c.box_max[d] - c.box_min[d] < c.cutoff + c.skin) {
      g_createError = "apb_create: box is smaller than cutoff + skin in dimension " + std::to_string(d);
      return APB_ERR_INVALID_ARGUMENT;
    }
  }
  if (!(c.cutoff > 0.) || c.skin < 0.) {
    g_createError = "apb_create: cutoff must be > 0 and skin >= 0";
    return APB_ERR_INVALID_ARGUMENT;
  }
  if (c.container != APB_CONTAINER_LINKED_CELLS && c.container != APB_CONTAINER_VERLET_CLUSTER_LISTS) {
    g_createError = "apb_create: unknown container option";
    return APB_ERR_INVALID_ARGUMENT;
  }
  if (c.container == APB_CONTAINER_VERLET_CLUSTER_LISTS) {
    const int m = c.cluster_size;
    if (!(m == 1 || m == 2 || m == 4 || m == 8 || m == 16 || m == 32)) {
      g_createError = "apb_create: gpuVerletClusterLists supports cluster sizes 1,2,4,8,16,32 (a warp tile)";
      return APB_ERR_NOT_APPLICABLE;
    }
  } else if (!(c.cell_size_factor > 0.)) {
    g_createError = "apb_create: cell_size_factor must be > 0";
    return APB_ERR_INVALID_ARGUMENT;
  }
  if (c.particle_kind < APB_PARTICLE_LJ || c.particle_kind > APB_PARTICLE_SPH) {
    g_createError = "apb_create: unknown particle kind";
    return APB_ERR_INVALID_ARGUMENT;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_createError = std::string("apb_create: no CUDA device available (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback";
    return APB_ERR_CUDA;
  }
  if (c.device < 0 || c.device >= ndev) {
    g_createError = "apb_create: device ordinal out of range";
    return APB_ERR_INVALID_ARGUMENT;
  }
  e = cudaSetDevice(c.device);
  if (e != cudaSuccess) {
    g_createError = std::string("apb_create: cudaSetDevice: ") + cudaGetErrorString(e);
    return APB_ERR_CUDA;
  }
  apb_handle h = new (std::nothrow) apb_handle_s();
  if (!h) {
    g_createError = "apb_create: out of host memory";
    return APB_ERR_OUT_OF_MEMORY;
  }
  h->cfg = c;
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    // the exchange stream outranks the compute stream: its small pack / NCCL / unpack kernels must get SM slots while
    // the interior force kernel still has blocks waiting (split step of apb_run_steps)
    int leastPriority = 0, greatestPriority = 0;
    cudaDeviceGetStreamPriorityRange(&leastPriority, &greatestPriority);
    e = cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, greatestPriority);
  }
  if (e != cudaSuccess) {
    g_createError = std::string("apb_create: cudaStreamCreate: ") + cudaGetErrorString(e);
    delete h;
    return APB_ERR_CUDA;
  }
  {
    // keep freed blocks in the pool instead of returning them to the driver at the next synchronisation
    cudaMemPool_t pool = nullptr;
    unsigned long long threshold = ~0ULL;
    if (cudaDeviceGetDefaultMemPool(&pool, c.device) != cudaSuccess ||
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold) != cudaSuccess) {
      g_createError = "apb_create: the device has no stream-ordered memory pool";
      delete h;
      return APB_ERR_CUDA;
    }
  }
  {
    // per-function attributes (opt-in shared memory sizes) are set once here, not per launch
    int rc = apbInitKernelAttributes(h);
    if (rc != APB_OK) {
      g_createError = h->err;
      apb_destroy(h);
      return rc;
    }
  }
  for (int k = APB_COL_X; k <= APB_COL_FZ; ++k) h->active[k] = true;
  if (c.particle_kind == APB_PARTICLE_LJ || c.particle_kind == APB_PARTICLE_MULTISITE) {
    for (int k = APB_COL_OLDFX; k <= APB_COL_OLDFZ; ++k) h->active[k] = true;
  }
  if (c.particle_kind == APB_PARTICLE_MULTISITE) {
    for (int k = APB_COL_Q0; k <= APB_COL_TZ; ++k) h->active[k] = true;
  }
  if (c.particle_kind == APB_PARTICLE_SPH) {
    for (int k = APB_COL_MASS; k <= APB_COL_VSIGMAX; ++k) h->active[k] = true;
  }
  if (c.container == APB_CONTAINER_LINKED_CELLS) {
    apbComputeLCGeom(c, h->lc);
    h->numCells = h->lc.numCells;
    int rc = apbComputeStencil(h);
    if (rc != APB_OK) {
      g_createError = h->err;
      apb_destroy(h);
      return rc;
    }
  }
  int rc = apbEnsure(h, h->result, sizeof(apb_traversal_result) + 256);
  if (rc != APB_OK) {
    g_createError = h->err;
    apb_destroy(h);
    return rc;
  }
  *out = h;
  return APB_OK;
}

extern "C" int apb_destroy(apb_handle h) {
  if (!h) return APB_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int c = 0; c < APB_NUM_COLUMNS; ++c) {
    if (h->col[c]) cudaFree(h->col[c]);
    if (h->colTmp[c]) cudaFree(h->colTmp[c]);
  }
  void *arrs[] = {h->id, h->idTmp, h->type, h->typeTmp, h->own, h->ownTmp};
  for (void *p : arrs)
    if (p) cudaFree(p);
  DevBuf *bufs[] = {&h->key,          &h->rank,         &h->count,        &h->start,       &h->perm,
                    &h->slotCell,     &h->scanTmp,      &h->sortK1,       &h->sortK2,      &h->sortV,
                    &h->stencilDev,   &h->clBoxMin,     &h->clBoxMax,     &h->clHasOwned,  &h->clIsHalo,
                    &h->clTower,      &h->twFirstCluster, &h->twNumClusters, &h->twFirstOwned, &h->twFirstTailHalo,
                    &h->nbrCount,     &h->nbrStart,     &h->nbrList,      &h->prNumStaged, &h->prStagedStart, &h->prStaged, &h->prWarpLen, &h->prWarpStart, &h->prLists,
                    &h->partials,     &h->result,       &h->mixDev,       &h->leaverIdx, &h->idStage, &h->prTileHalo, &h->prTileOrder, &h->prMasks, &h->prUsed, &h->prCbase,
                    &h->prNumCompact, &h->prCompactSlot, &h->haloAllSrc, &h->haloAllDst, &h->haloAllCode, &h->prEntryLo, &h->partials2};
  for (DevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  DevBuf *ctrl[] = {&h->thermoDev, &h->rAtRebuild, &h->remBuf, &h->vtkCtl, &h->vtkTables, &h->vtkFlag, &h->vtkLen, &h->vtkOut};
  for (DevBuf *b : ctrl)
    if (b->p) cudaFree(b->p);
  DevBuf *more[] = {&h->prStageEarly, &h->prTileFirst, &h->prTileNum, &h->prTileWarp, &h->loopResults, &h->invPerm, &h->xbuf[0], &h->xbuf[1], &h->xbuf[2], &h->xbuf[3], &h->massDev};
  for (DevBuf *b : more)
    if (b->p) cudaFree(b->p);
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s) {
      if (h->link[d][s].sendIdx.p) cudaFree(h->link[d][s].sendIdx.p);
      if (h->link[d][s].recvSlot.p) cudaFree(h->link[d][s].recvSlot.p);
    }
  apbCommDestroy(h);
  for (cudaEvent_t e : h->evSplit)
    if (e) cudaEventDestroy(e);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  cudaGetLastError();
  delete h;
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// adding / deleting particles
// ------------------------------------------------------------------------------------------------------------------
__global__ void kFillAppended(int64_t first, int64_t n, int64_t *id, int32_t *type, int32_t *own, int32_t ownership,
                              int hasIds, int hasTypes) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!hasIds) id[first + i] = i;
  if (!hasTypes) type[first + i] = 0;
  own[first + i] = ownership;
}

extern "C" int apb_add_particles(apb_handle h, int64_t n, const double *x, const double *y, const double *z,
                                 const int64_t *ids, const int32_t *types, int32_t ownership, int32_t check_box) {
  APB_ENTRY(h);
  h->ownedKnown = false;  // the number of owned particles may change
  h->ownedInsideBox = false;  // positions / ownership may change: the one-pass halo images need apb_migrate first
  h->noHalos = false;
  if (n < 0 || (n > 0 && (!x || !y || !z))) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_add_particles: null positions");
  if (ownership != APB_OWN_OWNED && ownership != APB_OWN_HALO)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_add_particles: ownership must be owned (1) or halo (2)");
  if (n == 0) return APB_OK;
  if (check_box && ownership == APB_OWN_OWNED) {
    // utils::inBox is half-open [lo, hi) (utils/inBox.h:26-36); LogicHandler::addParticle throws outside
    const double *p[3] = {x, y, z};
    for (int d = 0; d < 3; ++d) {
      const double lo = h->cfg.box_min[d], hi = h->cfg.box_max[d];
      for (int64_t i = 0; i < n; ++i) {
        if (!(p[d][i] >= lo && p[d][i] < hi)) {
          return h->fail(APB_ERR_PARTICLE_OUTSIDE, "apb_add_particles: owned particle " + std::to_string(i) +
                                                       " is outside the container box");
        }
      }
    }
  }
  const int64_t first = h->nslots;
  APB_CHECK(apbReserveSlots(h, first + n));
  const double *src[3] = {x, y, z};
  for (int d = 0; d < 3; ++d)
    APB_CUDA(cudaMemcpyAsync(h->col[APB_COL_X + d] + first, src[d], sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  for (int c = APB_COL_VX; c < APB_NUM_COLUMNS; ++c) {
    if (h->active[c]) APB_CUDA(cudaMemsetAsync(h->col[c] + first, 0, sizeof(double) * n, h->stream));
  }
  if (ids) APB_CUDA(cudaMemcpyAsync(h->id + first, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, h->stream));
  if (types) APB_CUDA(cudaMemcpyAsync(h->type + first, types, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
  ++h->launchCount, kFillAppended<<<apbDivUp(n, 256), 256, 0, h->stream>>>(first, n, h->id, h->type, h->own, ownership, ids != nullptr,
                                                         types != nullptr);
  APB_CUDA(cudaGetLastError());
  APB_CUDA(cudaStreamSynchronize(h->stream));
  h->nslots = first + n;
  h->structureValid = false;
  h->prunedValid = false;
  h->countsValid = false;
  apbForgetHaloLinks(h);
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// md-flexible's MPI wire format (examples/md-flexible/src/ParticleSerializationTools.cpp:43-146): one record of
// AttributesSize = 120 bytes per MoleculeLJ, the attributes memcpy'd in this order (:43-58), each 8 bytes wide -
// id (size_t), posX posY posZ, velocityX..Z, forceX..Z, oldForceX..Z (double), typeId (size_t), ownershipState (int64_t,
// OwnershipState.h:20-29). Lets a hybrid run hand device-resident particles to md-flexible's MPI exchange between nodes
// (RegularGridDecomposition.cpp: sendParticles / receiveParticles) and take its messages back.
// One thread per (record, word): the 15-word records are written / read with fully coalesced 8-byte accesses.
// ------------------------------------------------------------------------------------------------------------------
#define APB_WIRE_WORDS 15
struct WireColumns {
  const double *c[12];  // x y z vx vy vz fx fy fz oldFx oldFy oldFz
  double *w[12];
};
__global__ void kWireSelect(int64_t n, const int32_t *__restrict__ own, int mask, int *__restrict__ flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = own[i];
  flag[i] = (o == APB_OWN_OWNED && (mask & 1)) || (o == APB_OWN_HALO && (mask & 2));
}
__global__ void kWirePack(int64_t n, const int *__restrict__ flag, const int *__restrict__ pos, WireColumns cols,
                          const int64_t *__restrict__ id, const int32_t *__restrict__ type, const int32_t *__restrict__ own,
                          unsigned long long *__restrict__ out) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t i = g / APB_WIRE_WORDS;
  const int w = static_cast<int>(g % APB_WIRE_WORDS);
  if (i >= n || !flag[i]) return;
  unsigned long long v;
  if (w == 0)
    v = static_cast<unsigned long long>(id[i]);
  else if (w <= 12)
    v = static_cast<unsigned long long>(__double_as_longlong(cols.c[w - 1][i]));
  else if (w == 13)
    v = static_cast<unsigned long long>(type[i]);
  else
    v = static_cast<unsigned long long>(static_cast<long long>(own[i]));
  out[static_cast<size_t>(pos[i]) * APB_WIRE_WORDS + w] = v;
}
__global__ void kWireUnpack(int64_t m, int64_t first, const unsigned long long *__restrict__ in, WireColumns cols,
                            int64_t *__restrict__ id, int32_t *__restrict__ type, int32_t *__restrict__ own, int *bad) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t r = g / APB_WIRE_WORDS;
  const int w = static_cast<int>(g % APB_WIRE_WORDS);
  if (r >= m) return;
  const unsigned long long v = in[g];
  const int64_t t = first + r;
  if (w == 0)
    id[t] = static_cast<int64_t>(v);
  else if (w <= 12)
    cols.w[w - 1][t] = __longlong_as_double(static_cast<long long>(v));
  else if (w == 13)
    type[t] = static_cast<int32_t>(v);
  else {
    const long long o = static_cast<long long>(v);
    if (o != APB_OWN_OWNED && o != APB_OWN_HALO) atomicExch(bad, 1);  // a dummy has no business on the wire
    own[t] = o == APB_OWN_HALO ? APB_OWN_HALO : APB_OWN_OWNED;
  }
}

static int wireColumns(apb_handle h, WireColumns &wc) {
  if (h->cfg.particle_kind != APB_PARTICLE_LJ)
    return h->fail(APB_ERR_NOT_APPLICABLE, "the 120-byte wire record is the one of MoleculeLJ (single-site mode)");
  for (int k = 0; k < 12; ++k) wc.c[k] = wc.w[k] = h->col[APB_COL_X + k];
  return APB_OK;
}

extern "C" int apb_serialize_particles(apb_handle h, int32_t ownershipMask, void *dst, int64_t capacityRecords,
                                       int64_t *outNum) {
  APB_ENTRY(h);
  if (outNum) *outNum = 0;
  if (!(ownershipMask & 3) || capacityRecords < 0 || (capacityRecords > 0 && !dst))
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_serialize_particles: bad argument");
  WireColumns wc;
  APB_CHECK(wireColumns(h, wc));
  const int64_t n = h->nslots;
  if (n == 0) return APB_OK;
  APB_CHECK(apbEnsure(h, h->key, sizeof(int) * n));
  APB_CHECK(apbEnsure(h, h->rank, sizeof(int) * (n + 1)));
  int *flag = static_cast<int *>(h->key.p), *pos = static_cast<int *>(h->rank.p);
  long long *totals = reinterpret_cast<long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result));
  ++h->launchCount, kWireSelect<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, h->own, ownershipMask, flag);
  APB_CHECK(apbExclusiveScan(h, flag, pos, n, totals));
  long long m = 0;
  APB_CUDA(cudaMemcpyAsync(&m, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (outNum) *outNum = m;
  if (m > capacityRecords)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_serialize_particles: " + std::to_string(m) + " records do not fit the buffer");
  if (m == 0) return APB_OK;
  APB_CHECK(apbEnsure(h, h->sortK1, static_cast<size_t>(m) * APB_WIRE_WORDS * 8));
  unsigned long long *out = static_cast<unsigned long long *>(h->sortK1.p);
  // (kWirePack reads `flag` after the scan wrote `pos`: the scan is out of place)
  ++h->launchCount, kWirePack<<<apbDivUp(n * APB_WIRE_WORDS, 256), 256, 0, h->stream>>>(n, flag, pos, wc, h->id, h->type, h->own, out);
  APB_CUDA(cudaGetLastError());
  APB_CUDA(cudaMemcpyAsync(dst, out, static_cast<size_t>(m) * APB_WIRE_WORDS * 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_deserialize_particles(apb_handle h, const void *src, int64_t numRecords) {
  APB_ENTRY(h);
  if (numRecords < 0 || (numRecords > 0 && !src)) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_deserialize_particles: bad argument");
  WireColumns wc;
  APB_CHECK(wireColumns(h, wc));
  if (numRecords == 0) return APB_OK;
  h->ownedKnown = false;
  h->ownedInsideBox = false;
  h->noHalos = false;
  const int64_t first = h->nslots;
  APB_CHECK(apbReserveSlots(h, first + numRecords));
  APB_CHECK(wireColumns(h, wc));  // the columns may have moved
  APB_CHECK(apbEnsure(h, h->sortK1, static_cast<size_t>(numRecords) * APB_WIRE_WORDS * 8));
  int *bad = reinterpret_cast<int *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 32);
  APB_CUDA(cudaMemsetAsync(bad, 0, 4, h->stream));
  APB_CUDA(cudaMemcpyAsync(h->sortK1.p, src, static_cast<size_t>(numRecords) * APB_WIRE_WORDS * 8, cudaMemcpyHostToDevice, h->stream));
  ++h->launchCount, kWireUnpack<<<apbDivUp(numRecords * APB_WIRE_WORDS, 256), 256, 0, h->stream>>>(
      numRecords, first, static_cast<const unsigned long long *>(h->sortK1.p), wc, h->id, h->type, h->own, bad);
  APB_CUDA(cudaGetLastError());
  int hostBad = 0;
  APB_CUDA(cudaMemcpyAsync(&hostBad, bad, 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (hostBad) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_deserialize_particles: a record is neither owned nor halo");
  h->nslots = first + numRecords;
  h->structureValid = false;
  h->prunedValid = false;
  h->countsValid = false;
  apbForgetHaloLinks(h);
  return APB_OK;
}

extern "C" int apb_delete_all_particles(apb_handle h) {
  APB_ENTRY(h);
  h->ownedKnown = false;  // the number of owned particles may change
  h->nslots = 0;
  h->structureValid = false;
  h->prunedValid = false;
  h->countsValid = false;
  h->numClusters = h->numPairs = 0;
  apbForgetHaloLinks(h);
  return APB_OK;
}

// The recorded halo links (send / receive / image slot lists) describe copies made by the last generating
// apb_exchange_halos. Whatever replaces the particle set behind their back - deleteAllParticles, particles or ownership
// states supplied by the host - makes them meaningless: the next apb_exchange_halos must not take the refresh branch.
void apbForgetHaloLinks(apb_handle h) {
  h->haloLinksValid = false;
  h->haloAllMode = false;
  h->haloAllN = 0;
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s) h->link[d][s].nSend = h->link[d][s].nRecv = 0;
}

// ParticleContainerInterface::reserve(numParticles, numParticlesHaloEstimate) (containers/ParticleContainerInterface.h:95)
extern "C" int apb_reserve(apb_handle h, int64_t numParticles, int64_t numParticlesHaloEstimate) {
  APB_ENTRY(h);
  if (numParticles < 0 || numParticlesHaloEstimate < 0) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_reserve: negative count");
  int64_t slots = numParticles + numParticlesHaloEstimate;
  // VerletClusterLists pads every tower to a multiple of the cluster size (ClusterTower.h:83-143)
  if (h->cfg.container == APB_CONTAINER_VERLET_CLUSTER_LISTS) slots += slots / 8 + 1024;
  return apbReserveSlots(h, slots);
}

extern "C" int apb_get_alloc_count(apb_handle h, int64_t *out_count) {
  APB_ENTRY(h);
  if (!out_count) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_alloc_count: null argument");
  *out_count = h->allocCount;
  return APB_OK;
}

__global__ void kDeleteHalo(int64_t n, int32_t *own) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && own[i] == APB_OWN_HALO) own[i] = APB_OWN_DUMMY;
}

extern "C" int apb_delete_halo_particles(apb_handle h) {
  APB_ENTRY(h);
  if (h->noHalos) return APB_OK;  // nothing was added since the last call (apb_migrate followed by apb_exchange_halos)
  h->noHalos = true;
  if (h->nslots > 0) {
    ++h->launchCount, kDeleteHalo<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(h->nslots, h->own);
    APB_CUDA(cudaGetLastError());
    if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->countsValid = false;
  h->ownDirty = true;
  return APB_OK;
}

// halo update by id: sort-free lookup through a temporary open-addressing hash table of the halo slots.
__device__ __forceinline__ unsigned long long apbHash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
__global__ void kHashInsertHalo(int64_t n, const int64_t *id, const int32_t *own, long long *keys, int *vals,
                                unsigned long long mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_HALO) return;
  unsigned long long slot = apbHash(static_cast<unsigned long long>(id[i])) & mask;
  while (true) {
    const long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(&keys[slot]), ~0ULL,
                                     static_cast<unsigned long long>(id[i]));
    if (prev == -1LL) {
      vals[slot] = static_cast<int>(i);
      return;
    }
    slot = (slot + 1) & mask;  // duplicate halo ids (periodic images) occupy several slots
  }
}
__global__ void kHashUpdateHalo(int64_t m, const int64_t *ids, const double *nx, const double *ny, const double *nz,
                                const long long *keys, const int *vals, unsigned long long mask, double *x, double *y,
                                double *z, double maxDist2, unsigned long long *notFound) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m) return;
  unsigned long long slot = apbHash(static_cast<unsigned long long>(ids[q])) & mask;
  // several periodic images may share one id: like LinkedCells::updateHaloParticle (LinkedCells.h:95-105) pick the
  // image that lies within the search radius of the new position
  while (true) {
    const long long k = keys[slot];
    if (k == -1LL) break;
    if (k == ids[q]) {
      const int s = vals[slot];
      const double dx = x[s] - nx[q], dy = y[s] - ny[q], dz = z[s] - nz[q];
      if (dx * dx + dy * dy + dz * dz <= maxDist2) {
        x[s] = nx[q];
        y[s] = ny[q];
        z[s] = nz[q];
        return;
      }
    }
    slot = (slot + 1) & mask;
  }
  atomicAdd(notFound, 1ULL);
}

extern "C" int apb_update_halo_particles(apb_handle h, int64_t n, const int64_t *ids, const double *x, const double *y,
                                         const double *z, int64_t *out_not_found) {
  APB_ENTRY(h);
  if (n < 0 || (n > 0 && (!ids || !x || !y || !z)))
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_update_halo_particles: null argument");
  if (out_not_found) *out_not_found = 0;
  if (n == 0) return APB_OK;
  unsigned long long tableSize = 1024;
  while (tableSize < static_cast<unsigned long long>(2 * h->nslots + 2)) tableSize <<= 1;
  // scratch: keys (8B) + vals (4B) per table slot; queries ids + 3 doubles
  APB_CHECK(apbEnsure(h, h->sortK1, tableSize * 8));
  APB_CHECK(apbEnsure(h, h->sortV, tableSize * 4));
  APB_CHECK(apbEnsure(h, h->sortK2, static_cast<size_t>(n) * 32 + 64));
  long long *keys = static_cast<long long *>(h->sortK1.p);
  int *vals = static_cast<int *>(h->sortV.p);
  char *q = static_cast<char *>(h->sortK2.p);
  int64_t *dIds = reinterpret_cast<int64_t *>(q);
  double *dX = reinterpret_cast<double *>(q + 8 * n), *dY = reinterpret_cast<double *>(q + 16 * n),
         *dZ = reinterpret_cast<double *>(q + 24 * n);
  unsigned long long *dNotFound = reinterpret_cast<unsigned long long *>(q + 32 * n);
  APB_CUDA(cudaMemsetAsync(keys, 0xFF, tableSize * 8, h->stream));
  APB_CUDA(cudaMemsetAsync(dNotFound, 0, 8, h->stream));
  APB_CUDA(cudaMemcpyAsync(dIds, ids, 8 * n, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(dX, x, 8 * n, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(dY, y, 8 * n, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(dZ, z, 8 * n, cudaMemcpyHostToDevice, h->stream));
  if (h->nslots > 0)
    ++h->launchCount, kHashInsertHalo<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(h->nslots, h->id, h->own, keys, vals, tableSize - 1);
  // search radius: the halo copy may have moved by at most skin since the lists were built (LinkedCells.h:95-105
  // searches +-skin around the new position; VerletClusterLists.h:199 +-skin/2)
  const double r = h->cfg.skin;
  ++h->launchCount, kHashUpdateHalo<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, dIds, dX, dY, dZ, keys, vals, tableSize - 1,
                                                           h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z],
                                                           3. * r * r, dNotFound);
  APB_CUDA(cudaGetLastError());
  unsigned long long nf = 0;
  APB_CUDA(cudaMemcpyAsync(&nf, dNotFound, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (out_not_found) *out_not_found = static_cast<int64_t>(nf);
  return APB_OK;
}

// grid-stride, one atomic pair per block: 37 k same-address atomics (one per warp) cost 28 us at 1.2 M slots
__global__ void __launch_bounds__(256) kCountOwnership(int64_t n, const int32_t *own, unsigned long long *counts) {
  __shared__ unsigned so[8], sh[8];
  unsigned owned = 0, halo = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int o = own[i];
    owned += o == APB_OWN_OWNED;
    halo += o == APB_OWN_HALO;
  }
  owned = __reduce_add_sync(0xffffffffu, owned);
  halo = __reduce_add_sync(0xffffffffu, halo);
  if ((threadIdx.x & 31) == 0) {
    so[threadIdx.x >> 5] = owned;
    sh[threadIdx.x >> 5] = halo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long to = 0, th = 0;
    for (int w = 0; w < 8; ++w) {
      to += so[w];
      th += sh[w];
    }
    if (to) atomicAdd(&counts[0], to);
    if (th) atomicAdd(&counts[1], th);
  }
}

extern "C" int apb_get_num_particles(apb_handle h, int64_t *out_owned, int64_t *out_halo) {
  APB_ENTRY(h);
  if (!h->countsValid) {
    unsigned long long counts[2] = {0, 0};
    if (h->nslots > 0) {
      // scratch words behind the result struct (allocated in apb_create)
      unsigned long long *d =
          reinterpret_cast<unsigned long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 64);
      APB_CUDA(cudaMemsetAsync(d, 0, 16, h->stream));
      ++h->launchCount, kCountOwnership<<<static_cast<unsigned>(std::min<int64_t>(apbDivUp(h->nslots, 256), 1184)), 256, 0, h->stream>>>(h->nslots, h->own, d);
      APB_CUDA(cudaGetLastError());
      APB_CUDA(cudaMemcpyAsync(counts, d, 16, cudaMemcpyDeviceToHost, h->stream));
      APB_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->numOwned = static_cast<int64_t>(counts[0]);
    h->numHalo = static_cast<int64_t>(counts[1]);
    h->countsValid = true;
    h->ownedKnown = true;
    h->ownedCount = h->numOwned;
  }
  if (out_owned) *out_owned = h->numOwned;
  if (out_halo) *out_halo = h->numHalo;
  return APB_OK;
}

extern "C" int apb_get_num_slots(apb_handle h, int64_t *out_slots) {
  APB_ENTRY(h);
  if (!out_slots) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_num_slots: null argument");
  *out_slots = h->nslots;
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// host mirror transfers
// ------------------------------------------------------------------------------------------------------------------
static int checkColumn(apb_handle h, int32_t column) {
  if (column < 0 || column >= APB_NUM_COLUMNS || !h->active[column])
    return h->fail(APB_ERR_INVALID_ARGUMENT, "column " + std::to_string(column) + " is not stored for this particle kind");
  return APB_OK;
}

extern "C" int apb_download_column(apb_handle h, int32_t column, double *dst) {
  APB_ENTRY(h);
  APB_CHECK(checkColumn(h, column));
  if (h->nslots == 0) return APB_OK;
  if (!dst) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_download_column: null destination");
  APB_CUDA(cudaMemcpyAsync(dst, h->col[column], sizeof(double) * h->nslots, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_upload_column(apb_handle h, int32_t column, const double *src) {
  APB_ENTRY(h);
  h->ownedInsideBox = false;  // positions / ownership may change: the one-pass halo images need apb_migrate first
  h->noHalos = false;
  APB_CHECK(checkColumn(h, column));
  if (h->nslots == 0) return APB_OK;
  if (!src) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_upload_column: null source");
  APB_CUDA(cudaMemcpyAsync(h->col[column], src, sizeof(double) * h->nslots, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_download_ids(apb_handle h, int64_t *ids, int32_t *types, int32_t *ownership) {
  APB_ENTRY(h);
  if (h->nslots == 0) return APB_OK;
  if (ids) APB_CUDA(cudaMemcpyAsync(ids, h->id, sizeof(int64_t) * h->nslots, cudaMemcpyDeviceToHost, h->stream));
  if (types) APB_CUDA(cudaMemcpyAsync(types, h->type, sizeof(int32_t) * h->nslots, cudaMemcpyDeviceToHost, h->stream));
  if (ownership)
    APB_CUDA(cudaMemcpyAsync(ownership, h->own, sizeof(int32_t) * h->nslots, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_upload_ownership(apb_handle h, const int32_t *ownership) {
  APB_ENTRY(h);
  h->ownedKnown = false;  // the number of owned particles may change
  h->ownedInsideBox = false;  // positions / ownership may change: the one-pass halo images need apb_migrate first
  h->noHalos = false;
  if (h->nslots == 0) return APB_OK;
  if (!ownership) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_upload_ownership: null source");
  APB_CUDA(cudaMemcpyAsync(h->own, ownership, sizeof(int32_t) * h->nslots, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  h->countsValid = false;
  h->ownDirty = true;
  apbForgetHaloLinks(h);
  return APB_OK;
}

// Three columns through one pinned staging buffer. The host->pinned memcpy of column k+1 overlaps the DMA of column k.
static int transfer3(apb_handle h, int firstCol, const double *const src[3], double *const dst[3]) {
  const int64_t n = h->nslots;
  if (n == 0) return APB_OK;
  const size_t bytes = sizeof(double) * n;
  // caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister) are used for DMA directly
  bool pinnedUser = true;
  for (int d = 0; d < 3; ++d) {
    cudaPointerAttributes attr;
    const void *p = src ? static_cast<const void *>(src[d]) : static_cast<const void *>(dst[d]);
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess || attr.type != cudaMemoryTypeHost) pinnedUser = false;
  }
  cudaGetLastError();
  if (pinnedUser) {
    for (int d = 0; d < 3; ++d) {
      if (src)
        APB_CUDA(cudaMemcpyAsync(h->col[firstCol + d], src[d], bytes, cudaMemcpyHostToDevice, h->stream));
      else
        APB_CUDA(cudaMemcpyAsync(dst[d], h->col[firstCol + d], bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    APB_CUDA(cudaStreamSynchronize(h->stream));
    return APB_OK;
  }
  APB_CHECK(apbEnsurePinned(h, 3 * bytes));
  char *pin = static_cast<char *>(h->pinned);
  if (src) {
    for (int d = 0; d < 3; ++d) {
      std::memcpy(pin + d * bytes, src[d], bytes);
      APB_CUDA(cudaMemcpyAsync(h->col[firstCol + d], pin + d * bytes, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    APB_CUDA(cudaStreamSynchronize(h->stream));
  } else {
    cudaEvent_t ev[3];
    for (int d = 0; d < 3; ++d) {
      APB_CUDA(cudaEventCreateWithFlags(&ev[d], cudaEventDisableTiming));
      APB_CUDA(cudaMemcpyAsync(pin + d * bytes, h->col[firstCol + d], bytes, cudaMemcpyDeviceToHost, h->stream));
      APB_CUDA(cudaEventRecord(ev[d], h->stream));
    }
    for (int d = 0; d < 3; ++d) {
      APB_CUDA(cudaEventSynchronize(ev[d]));
      std::memcpy(dst[d], pin + d * bytes, bytes);
      APB_CUDA(cudaEventDestroy(ev[d]));
    }
  }
  return APB_OK;
}

extern "C" int apb_upload_positions(apb_handle h, const double *x, const double *y, const double *z) {
  APB_ENTRY(h);
  h->ownedInsideBox = false;  // positions / ownership may change: the one-pass halo images need apb_migrate first
  h->noHalos = false;
  if (h->nslots > 0 && (!x || !y || !z)) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_upload_positions: null source");
  const double *src[3] = {x, y, z};
  return transfer3(h, APB_COL_X, src, nullptr);
}

extern "C" int apb_download_forces(apb_handle h, double *fx, double *fy, double *fz) {
  APB_ENTRY(h);
  if (h->nslots > 0 && (!fx || !fy || !fz)) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_download_forces: null destination");
  double *dst[3] = {fx, fy, fz};
  return transfer3(h, APB_COL_FX, nullptr, dst);
}

// ---- transfers through host arrays indexed by particle id -------------------------------------------------------------
// The host side of a force step whose particle data lives in host memory (the C++ shim's mirror, md-flexible's own
// arrays): only owned particles travel, and the host never has to learn the storage order.
__global__ void kScatterPositionsById(int64_t n, const int32_t *__restrict__ own, const int64_t *__restrict__ id,
                                      int64_t idBegin, int64_t numIds, const double *__restrict__ sx,
                                      const double *__restrict__ sy, const double *__restrict__ sz, double *x, double *y,
                                      double *z) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  const int64_t k = id[i] - idBegin;
  if (k < 0 || k >= numIds) return;
  x[i] = sx[k];
  y[i] = sy[k];
  z[i] = sz[k];
}

__global__ void kGatherForcesById(int64_t n, const int32_t *__restrict__ own, const int64_t *__restrict__ id,
                                  int64_t idBegin, int64_t numIds, const double *__restrict__ fx,
                                  const double *__restrict__ fy, const double *__restrict__ fz, double *sx, double *sy,
                                  double *sz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  const int64_t k = id[i] - idBegin;
  if (k < 0 || k >= numIds) return;
  sx[k] = fx[i];
  sy[k] = fy[i];
  sz[k] = fz[i];
}

static bool isPinnedHost(const void *p) {
  cudaPointerAttributes attr;
  const bool ok = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return ok;
}

extern "C" int apb_upload_positions_by_id(apb_handle h, int64_t idBegin, int64_t numIds, const double *x, const double *y,
                                          const double *z) {
  APB_ENTRY(h);
  h->ownedInsideBox = false;  // positions / ownership may change: the one-pass halo images need apb_migrate first
  h->noHalos = false;
  if (numIds < 0 || (numIds > 0 && (!x || !y || !z)))
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_upload_positions_by_id: bad argument");
  if (numIds == 0 || h->nslots == 0) return APB_OK;
  const size_t bytes = sizeof(double) * numIds;
  APB_CHECK(apbEnsure(h, h->idStage, 3 * bytes));
  char *stage = static_cast<char *>(h->idStage.p);
  const double *src[3] = {x, y, z};
  const bool pinnedUser = isPinnedHost(x) && isPinnedHost(y) && isPinnedHost(z);
  if (!pinnedUser) APB_CHECK(apbEnsurePinned(h, 3 * bytes));
  for (int d = 0; d < 3; ++d) {
    const void *from = src[d];
    if (!pinnedUser) {
      std::memcpy(static_cast<char *>(h->pinned) + d * bytes, src[d], bytes);
      from = static_cast<char *>(h->pinned) + d * bytes;
    }
    APB_CUDA(cudaMemcpyAsync(stage + d * bytes, from, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  ++h->launchCount, kScatterPositionsById<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
      h->nslots, h->own, h->id, idBegin, numIds, reinterpret_cast<const double *>(stage),
      reinterpret_cast<const double *>(stage + bytes), reinterpret_cast<const double *>(stage + 2 * bytes), h->col[APB_COL_X],
      h->col[APB_COL_Y], h->col[APB_COL_Z]);
  APB_CUDA(cudaGetLastError());
  // the caller may reuse its buffers on return; inside apb_force_step_by_id the step's final synchronisation covers it
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_download_forces_by_id(apb_handle h, int64_t idBegin, int64_t numIds, double *fx, double *fy, double *fz) {
  APB_ENTRY(h);
  if (numIds < 0 || (numIds > 0 && (!fx || !fy || !fz)))
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_download_forces_by_id: bad argument");
  if (numIds == 0) return APB_OK;
  const size_t bytes = sizeof(double) * numIds;
  APB_CHECK(apbEnsure(h, h->idStage, 3 * bytes));
  char *stage = static_cast<char *>(h->idStage.p);
  // ids without an owned particle on this device read as zero
  APB_CUDA(cudaMemsetAsync(stage, 0, 3 * bytes, h->stream));
  if (h->nslots > 0) {
    ++h->launchCount, kGatherForcesById<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
        h->nslots, h->own, h->id, idBegin, numIds, h->col[APB_COL_FX], h->col[APB_COL_FY], h->col[APB_COL_FZ],
        reinterpret_cast<double *>(stage), reinterpret_cast<double *>(stage + bytes),
        reinterpret_cast<double *>(stage + 2 * bytes));
    APB_CUDA(cudaGetLastError());
  }
  double *dst[3] = {fx, fy, fz};
  if (isPinnedHost(fx) && isPinnedHost(fy) && isPinnedHost(fz)) {
    for (int d = 0; d < 3; ++d)
      APB_CUDA(cudaMemcpyAsync(dst[d], stage + d * bytes, bytes, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    return APB_OK;
  }
  APB_CHECK(apbEnsurePinned(h, 3 * bytes));
  APB_CUDA(cudaMemcpyAsync(h->pinned, stage, 3 * bytes, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  for (int d = 0; d < 3; ++d) std::memcpy(dst[d], static_cast<char *>(h->pinned) + d * bytes, bytes);
  return APB_OK;
}

__global__ void kFill3(int64_t n, double *a, double *b, double *c, double va, double vb, double vc) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    a[i] = va;
    b[i] = vb;
    c[i] = vc;
  }
}

extern "C" int apb_reset_forces(apb_handle h, double fx, double fy, double fz) {
  APB_ENTRY(h);
  if (h->nslots == 0) return APB_OK;
  ++h->launchCount, kFill3<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(h->nslots, h->col[APB_COL_FX], h->col[APB_COL_FY],
                                                          h->col[APB_COL_FZ], fx, fy, fz);
  APB_CUDA(cudaGetLastError());
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// updateContainer
// ------------------------------------------------------------------------------------------------------------------
// flag: 0 keep, 1 leaver (owned, not in box), 2 drop (halo / dummy)
__global__ void kClassify(int64_t n, const double *x, const double *y, const double *z, const int32_t *own, double lx,
                          double ly, double lz, double hx, double hy, double hz, int *flag, int *keep, int *leaver) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f;
  if (own[i] != APB_OWN_OWNED) {
    f = 2;
  } else {
    const bool in = x[i] >= lx && x[i] < hx && y[i] >= ly && y[i] < hy && z[i] >= lz && z[i] < hz;
    f = in ? 0 : 1;
  }
  flag[i] = f;
  keep[i] = f == 0;
  leaver[i] = f == 1;
}

__global__ void kMarkAndCollect(int64_t n, const int *flag, const int *leaverPos, int32_t *own, int *leaverIdx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = flag[i];
  if (f == 1) leaverIdx[leaverPos[i]] = static_cast<int>(i);
  if (f != 0) own[i] = APB_OWN_DUMMY;
}

__global__ void kGatherD(int64_t m, const int *idx, const double *src, double *dst) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < m) dst[q] = src[idx[q]];
}
__global__ void kGatherI64(int64_t m, const int *idx, const int64_t *src, int64_t *dst) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < m) dst[q] = src[idx[q]];
}
__global__ void kGatherI32(int64_t m, const int *idx, const int32_t *src, int32_t *dst) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < m) dst[q] = src[idx[q]];
}
__global__ void kKeepPerm(int64_t n, const int *flag, const int *keepPos, int *perm) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && flag[i] == 0) perm[keepPos[i]] = static_cast<int>(i);
}

struct GatherArgs {
  const double *src[APB_NUM_COLUMNS];
  double *dst[APB_NUM_COLUMNS];
  int ncols;
};
// gather all active columns by permutation; perm < 0 -> zero
__global__ void kGatherAll(int64_t m, const int *__restrict__ perm, GatherArgs a, const int64_t *idSrc, int64_t *idDst,
                           const int32_t *typeSrc, int32_t *typeDst, const int32_t *ownSrc, int32_t *ownDst) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int s = perm[q];
  if (s >= 0) {
    for (int c = 0; c < a.ncols; ++c) a.dst[c][q] = a.src[c][s];
    idDst[q] = idSrc[s];
    typeDst[q] = typeSrc[s];
    ownDst[q] = ownSrc[s];
  } else {
    for (int c = 0; c < a.ncols; ++c) a.dst[c][q] = 0.;
    idDst[q] = -1;  // std::numeric_limits<size_t>::max() in the reference (ClusterTower.h:159)
    typeDst[q] = 0;
    ownDst[q] = APB_OWN_DUMMY;
  }
}

// permute the whole storage: slot q of the new order takes old slot perm[q]; swaps the double buffers
int apbPermuteStorage(apb_handle h, const int *perm, int64_t newSlots) {
  APB_CHECK(apbReserveSlots(h, newSlots));
  GatherArgs a;
  a.ncols = 0;
  for (int c = 0; c < APB_NUM_COLUMNS; ++c) {
    if (!h->active[c]) continue;
    a.src[a.ncols] = h->col[c];
    a.dst[a.ncols] = h->colTmp[c];
    ++a.ncols;
  }
  if (newSlots > 0) {
    ++h->launchCount, kGatherAll<<<apbDivUp(newSlots, 256), 256, 0, h->stream>>>(newSlots, perm, a, h->id, h->idTmp, h->type, h->typeTmp,
                                                               h->own, h->ownTmp);
    APB_CUDA(cudaGetLastError());
  }
  for (int c = 0; c < APB_NUM_COLUMNS; ++c) std::swap(h->col[c], h->colTmp[c]);
  std::swap(h->id, h->idTmp);
  std::swap(h->type, h->typeTmp);
  std::swap(h->own, h->ownTmp);
  h->nslots = newSlots;
  return APB_OK;
}

extern "C" int apb_update_container(apb_handle h, int32_t keep, int64_t *out_num_leavers) {
  APB_ENTRY(h);
  h->ownedKnown = false;  // the number of owned particles may change
  h->numLeavers = 0;
  for (auto &v : h->leaverCols) v.clear();
  h->leaverIds.clear();
  h->leaverTypes.clear();
  if (out_num_leavers) *out_num_leavers = 0;
  const int64_t n = h->nslots;
  if (n == 0) {
    if (!keep) h->structureValid = false;
    return APB_OK;
  }
  APB_CHECK(apbEnsure(h, h->key, sizeof(int) * n));    // flag
  APB_CHECK(apbEnsure(h, h->rank, sizeof(int) * n));   // keep flags -> keepPos
  APB_CHECK(apbEnsure(h, h->perm, sizeof(int) * n));   // leaver flags -> leaverPos
  APB_CHECK(apbEnsure(h, h->sortV, sizeof(int) * n));  // scan output / perm
  APB_CHECK(apbEnsure(h, h->leaverIdx, sizeof(int) * n + 64));
  int *flag = static_cast<int *>(h->key.p), *keepF = static_cast<int *>(h->rank.p),
      *leavF = static_cast<int *>(h->perm.p);
  const auto &c = h->cfg;
  ++h->launchCount, kClassify<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->own,
                                                     c.box_min[0], c.box_min[1], c.box_min[2], c.box_max[0],
                                                     c.box_max[1], c.box_max[2], flag, keepF, leavF);
  APB_CUDA(cudaGetLastError());
  // totals live behind the result struct
  long long *totals = reinterpret_cast<long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result));
  int *leavPos = static_cast<int *>(h->sortV.p);
  APB_CHECK(apbExclusiveScan(h, leavF, leavPos, n, totals));
  int *leaverIdx = static_cast<int *>(h->leaverIdx.p);
  long long hostTotals[2] = {0, 0};
  if (keep) {
    ++h->launchCount, kMarkAndCollect<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, flag, leavPos, h->own, leaverIdx);
    APB_CUDA(cudaGetLastError());
    APB_CUDA(cudaMemcpyAsync(hostTotals, totals, 8, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
  } else {
    // collect leavers first (own[] still intact apart from the marking, which only concerns non-kept slots)
    ++h->launchCount, kMarkAndCollect<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, flag, leavPos, h->own, leaverIdx);
    APB_CUDA(cudaGetLastError());
    APB_CUDA(cudaMemcpyAsync(hostTotals, totals, 8, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
  }
  const int64_t nl = hostTotals[0];
  h->numLeavers = nl;
  if (nl > 0) {
    // the leavers' ownership was set to dummy above; their data is still in place. Gather every active column, ids and
    // types into one staging buffer and copy it out with one transfer.
    int numActive = 0;
    for (int k = 0; k < APB_NUM_COLUMNS; ++k) numActive += h->active[k] ? 1 : 0;
    APB_CHECK(apbEnsure(h, h->sortK1, sizeof(double) * nl * (numActive + 2)));
    double *tmp = static_cast<double *>(h->sortK1.p);
    std::vector<double> stage(static_cast<size_t>(nl) * (numActive + 2));
    int q = 0;
    for (int k = 0; k < APB_NUM_COLUMNS; ++k)
      if (h->active[k])
        ++h->launchCount, kGatherD<<<apbDivUp(nl, 256), 256, 0, h->stream>>>(nl, leaverIdx, h->col[k], tmp + static_cast<size_t>(q++) * nl);
    ++h->launchCount, kGatherI64<<<apbDivUp(nl, 256), 256, 0, h->stream>>>(nl, leaverIdx, h->id, reinterpret_cast<int64_t *>(tmp + static_cast<size_t>(numActive) * nl));
    ++h->launchCount, kGatherI32<<<apbDivUp(nl, 256), 256, 0, h->stream>>>(nl, leaverIdx, h->type, reinterpret_cast<int32_t *>(tmp + static_cast<size_t>(numActive + 1) * nl));
    APB_CUDA(cudaGetLastError());
    APB_CUDA(cudaMemcpyAsync(stage.data(), tmp, sizeof(double) * stage.size(), cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    q = 0;
    for (int k = 0; k < APB_NUM_COLUMNS; ++k)
      if (h->active[k]) {
        h->leaverCols[k].assign(stage.begin() + static_cast<size_t>(q) * nl, stage.begin() + static_cast<size_t>(q + 1) * nl);
        ++q;
      }
    h->leaverIds.resize(nl);
    h->leaverTypes.resize(nl);
    std::memcpy(h->leaverIds.data(), stage.data() + static_cast<size_t>(numActive) * nl, 8 * nl);
    std::memcpy(h->leaverTypes.data(), stage.data() + static_cast<size_t>(numActive + 1) * nl, 4 * nl);
  }
  if (!keep) {
    // compact the kept (owned, in box) particles, preserving order
    int *keepPos = leavPos;  // reuse
    APB_CHECK(apbExclusiveScan(h, keepF, keepPos, n, totals + 1));
    APB_CUDA(cudaMemcpyAsync(hostTotals + 1, totals + 1, 8, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    const int64_t nk = hostTotals[1];
    int *perm = static_cast<int *>(h->perm.p);  // leaver flags no longer needed
    ++h->launchCount, kKeepPerm<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, flag, keepPos, perm);
    APB_CUDA(cudaGetLastError());
    APB_CHECK(apbPermuteStorage(h, perm, nk));
    if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
    h->structureValid = false;
    h->prunedValid = false;
    h->haloLinksValid = false;
  }
  h->countsValid = false;
  h->ownDirty = true;
  if (out_num_leavers) *out_num_leavers = nl;
  return APB_OK;
}

extern "C" int apb_get_leavers(apb_handle h, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                               int64_t *ids, int32_t *types) {
  APB_ENTRY(h);
  const int64_t nl = h->numLeavers;
  double *dst[6] = {x, y, z, vx, vy, vz};
  for (int k = 0; k < 6; ++k)
    if (dst[k] && nl > 0) std::memcpy(dst[k], h->leaverCols[k].data(), sizeof(double) * nl);
  if (ids && nl > 0) std::memcpy(ids, h->leaverIds.data(), 8 * nl);
  if (types && nl > 0) std::memcpy(types, h->leaverTypes.data(), 4 * nl);
  return APB_OK;
}

extern "C" int apb_get_leaver_column(apb_handle h, int32_t column, double *dst) {
  APB_ENTRY(h);
  if (column < 0 || column >= APB_NUM_COLUMNS || !h->active[column])
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_leaver_column: column not part of this particle kind");
  if (!dst) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_leaver_column: null destination");
  if (h->numLeavers > 0) std::memcpy(dst, h->leaverCols[column].data(), sizeof(double) * h->numLeavers);
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// host-side functor helpers
// ------------------------------------------------------------------------------------------------------------------
extern "C" double apb_lj_calc_shift6(double epsilon24, double sigmaSquared, double cutoffSquared) {
  // ParticlePropertiesLibrary::calcShift6 (ParticlePropertiesLibrary.h:576-582), same operation order
  const double s2 = sigmaSquared / cutoffSquared;
  const double s6 = s2 * s2 * s2;
  return epsilon24 * (s6 - s6 * s6);
}

extern "C" int apb_make_lj_mixing_table(int32_t T, const double *eps, const double *sig, double cutoff, double *out) {
  if (T <= 0 || !eps || !sig || !out) return APB_ERR_INVALID_ARGUMENT;
  const double rc2 = cutoff * cutoff;
  for (int i = 0; i < T; ++i) {
    for (int j = 0; j < T; ++j) {
      const double e24 = 24 * std::sqrt(eps[i] * eps[j]);
      const double s = (sig[i] + sig[j]) / 2.0;
      const double s2 = s * s;
      double *o = out + 3 * (static_cast<size_t>(i) * T + j);
      o[0] = e24;
      o[1] = s2;
      o[2] = apb_lj_calc_shift6(e24, s2, rc2);
    }
  }
  return APB_OK;
}

extern "C" void apb_lj_end_traversal(const apb_traversal_result *raw, double *upot, double *virial) {
  // LJFunctor::endTraversal (LJFunctor.h:661-685): sums * 0.5, Upot additionally / 6; getVirial = x + y + z
  double u = raw->upot_sum;
  u *= 0.5;
  u /= 6.;
  const double vx = raw->virial_sum[0] * 0.5, vy = raw->virial_sum[1] * 0.5, vz = raw->virial_sum[2] * 0.5;
  if (upot) *upot = u;
  if (virial) *virial = vx + vy + vz;
}

extern "C" uint64_t apb_lj_num_flops(const apb_traversal_result *r, int32_t applyShift) {
  // LJFunctor::getNumFLOPs (LJFunctor.h:776-789)
  const uint64_t gN3 = applyShift ? 13 : 12, gNoN3 = applyShift ? 9 : 8;
  return r->num_dist_calls * 8 + r->num_kernel_calls_n3 * 18 + r->num_kernel_calls_no_n3 * 15 +
         r->num_global_calcs_n3 * gN3 + r->num_global_calcs_no_n3 * gNoN3;
}
