// Internal state of one container handle and helpers shared by the translation units of libautopas_b200.so.
// Nothing here is visible through the C ABI (include/autopas_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "autopas_b200.h"

#define APB_OWN_DUMMY 0
#define APB_OWN_OWNED 1
#define APB_OWN_HALO 2

#define APB_MAX_STENCIL 512

// ---- geometry passed to kernels by value -----------------------------------------------------------------------
// LinkedCells grid: restates CellBlock3D::rebuild (containers/CellBlock3D.h:360-426); filled on the host.
struct LCGeom {
  double boxMin[3], boxMax[3];
  double haloBoxMin[3], haloBoxMax[3];
  double cellLength[3], cellLengthReciprocal[3];
  int cellsPerDim[3];  // incl. halo
  int cellsPerInteractionLength;
  int numCells;
};

// VerletClusterLists tower grid: restates ClusterTowerBlock2D::estimateOptimalGridSideLength / resize
// (containers/verletClusterLists/ClusterTowerBlock2D.h:89-114, 140-168); filled on the host (uses std::cbrt).
struct VCLGeom {
  double boxMin[3], boxMax[3];
  double haloBoxMin[3], haloBoxMax[3];
  double side[2], sideReciprocal[2];
  double interactionLength, interactionLengthSqr;
  int towersPerDim[2];
  int numTowersPerInteractionLength;
  int numTowers;
  int clusterSize;
};

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

// slots recorded by the halo exchange of one (dimension, side): which of my slots feed the message to that neighbour
// and where the particles received from that neighbour live
struct HaloLink {
  DevBuf sendIdx, recvSlot;
  long long nSend = 0, nRecv = 0;
  double shift = 0.;
};

struct apb_handle_s {
  apb_config cfg{};
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // copy stream for overlapped transfers
  std::string err;
  bool poisoned = false;

  // ---- SoA particle storage (slots [0, nslots)), double-buffered for the sort ----
  int64_t nslots = 0;
  int64_t cap = 0;
  bool active[APB_NUM_COLUMNS]{};
  double *col[APB_NUM_COLUMNS]{};
  double *colTmp[APB_NUM_COLUMNS]{};
  int64_t *id = nullptr, *idTmp = nullptr;
  int32_t *type = nullptr, *typeTmp = nullptr;
  int32_t *own = nullptr, *ownTmp = nullptr;

  // ---- structure ----
  bool structureValid = false;  // cells / towers+lists match the storage order
  int builtNewton3 = -1;        // VCL lists: which newton3 mode they were built for
  bool ownDirty = false;        // ownership column modified since the structure was built (deleted particles)
  LCGeom lc{};
  VCLGeom vcl{};
  int64_t numCells = 0;  // cells or towers
  DevBuf key, rank, count, start, perm, slotCell, scanTmp, sortK1, sortK2, sortV;
  // gpuLinkedCells: per-slot partner lists within cutoff + skin, built on first use after a rebuild (functors.cu)
  long long structureVersion = 0, lcListVersion = -1;
  int lcListHalf = -1, lcListCap = 0;
  unsigned long long scanEpoch = 0;  // chained scan (apbExclusiveScan): entries of earlier scans read as not ready
  int64_t scanTiles = 0;             // tiles the status array holds; the ticket counter sits behind them
  int stencilN = 0;  // LC neighbour-cell offsets (incl. self at index 0)
  int stencil[APB_MAX_STENCIL][3];
  DevBuf stencilDev;

  // VCL
  int64_t numClusters = 0;
  int64_t numPairs = 0;
  DevBuf clBoxMin, clBoxMax;  // 3 doubles per cluster each (SoA: [3][numClusters])
  DevBuf clHasOwned, clIsHalo, clTower;
  DevBuf twFirstCluster, twNumClusters, twFirstOwned, twFirstTailHalo;
  DevBuf nbrCount, nbrStart, nbrList;
  // per-particle lists for APB_TRAVERSAL_GPUVCL_PRUNED (pruned.cu)
  bool prunedValid = false;
  int prunedNewton3 = 0;    // which newton3 mode the per-particle lists were built for
  int prunedTiles = 0;
  int prunedMaxStaged = 0;  // clusters
  int prunedMaxCompact = 0; // particles staged by the force kernel (largest tile)
  int vclMaxTowerCount = 0; // particles of the fullest tower (set by the VCL rebuild)
  long long prunedRows = 0; // list rows of 32 entries
  unsigned long long prunedEntries = 0;  // real (non padding) list entries = distance evaluations per force call
  int prunedWarps = 0;
  DevBuf prNumStaged, prStagedStart, prStaged, prWarpLen, prWarpStart, prLists, prTileFirst, prTileNum, prTileWarp;
  DevBuf prMasks, prUsed, prCbase, prNumCompact, prCompactSlot;
  DevBuf prEntryLo;
  DevBuf prStageEarly;  // fixed-stride rows of staged sets written by the counting pass (kPrunedStage<false>)
  DevBuf prTileHalo, prTileOrder;  // per tile: stages a halo copy; tiles ordered interior first (+ the interior count)
  // apb_run_steps with gpuvcl_pruned (newton3 off): kLJPruned stores global force + pair sum, the integrator skips its reset
  bool forceOverwrite = false;
  double forceG[3] = {0., 0., 0.};
  int prunedPart = 0;              // 0: whole traversal; 1 / 2: interior / boundary half of a split step (apb_run_steps)
  int prunedCap = 0;  // staged particles incl. the 16 sentinel slots the lists were built for (1280, 2048 or 4096)
  cudaEvent_t evSplit[2] = {nullptr, nullptr};

  // ---- reductions / results ----
  DevBuf partials, partials2;
  DevBuf result;  // apb_traversal_result on device
  DevBuf mixDev;
  std::vector<double> mixHostCache;
  int mixHostShift = -1;  // APPLY_SHIFT bit the derived table was built with

  // ---- leavers (library-owned, valid until the next update_container) ----
  int64_t numLeavers = 0;
  DevBuf leaverIdx;
  DevBuf idStage;  // 3 x numIds doubles: staging of the by-id transfers
  std::vector<double> leaverCols[APB_NUM_COLUMNS];  // every active column of the leavers
  std::vector<int64_t> leaverIds;
  std::vector<int32_t> leaverTypes;

  // pinned staging for host transfers
  void *pinned = nullptr;
  size_t pinnedCap = 0;

  // ---- decomposition / exchange (dynamics.cu) ----
  int nranks = 1, myRank = 0;
  void *comm = nullptr;  // ncclComm_t
  bool decompositionSet = false;
  double globalMin[3]{}, globalMax[3]{};
  bool periodic[3]{};
  int neighbor[3][2]{};
  // peer-memory halo refresh (non-rebuild steps, ranks of one node): every rank owns an arena of receive regions
  // [3 dims][2 sides][2 parities] x p2pCap doubles followed by one 64-bit sequence flag per (dim, side); neighbours map
  // it through CUDA IPC and write refreshed halo positions straight into it over NVLink
  void *p2pArena = nullptr;
  size_t p2pCap = 0;                       // doubles per region
  void *p2pPeer[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // neighbour arenas (mapped)
  std::vector<std::pair<int, void *>> p2pOpened;  // rank -> mapped base (one mapping per distinct neighbour)
  unsigned long long p2pSeq = 0, p2pCountSeq[3] = {0, 0, 0};
  int p2pState = 0;                        // 0 not set up yet, 1 ready, -1 not applicable
  int *p2pCounters = nullptr;              // 3 block counters (last-block-signals pattern)
  bool haloLinksValid = false;
  // single rank, all dimensions periodic: every halo copy is a direct image of an owned particle (one pass instead of
  // three forwarding rounds); haloAllSrc / Dst / Code = source slot, halo slot, shift code (base 3: 0 none, 1 +L, 2 -L)
  DevBuf haloAllSrc, haloAllDst, haloAllCode;
  long long haloAllN = 0;
  bool haloAllMode = false;
  bool noHalos = false;         // set by apb_delete_halo_particles, cleared when halo copies may exist again
  bool ownedInsideBox = false;  // set by apb_migrate, cleared by whatever moves or adds particles
  HaloLink link[3][2];
  DevBuf invPerm, xbuf[4], massDev;
  std::vector<double> massHost;

  // phase timing of apb_run_steps (perf.cu)
  struct LoopEvent {
    void *ev;
    int phase;
    bool begin;
  };
  bool loopTiming = false;
  std::vector<LoopEvent> loopEvents;
  std::vector<void *> eventPool;
  long long launchCount = 0;                      // kernels launched through this handle
  long long allocCount = 0;                       // device allocations made through this handle (growth events)
  bool deferSync = false;                          // apb_run_steps: do not block after every call
  apb_traversal_result *asyncResultDev = nullptr;  // apb_run_steps: where the reduced accumulators of this step go
  DevBuf loopResults;

  // ---- loop control (control.cu): thermostat, dynamic-rebuild trigger, remainder staging ----
  DevBuf thermoDev, rAtRebuild, remBuf;
  // checkpoint record (vtk.cu): control words, libm tables, owned flags + row index, row lengths + offsets, the text
  DevBuf vtkCtl, vtkTables, vtkFlag, vtkLen, vtkOut;
  bool vtkTablesReady = false;
  bool thermostatOn = false;
  int thermostatInterval = 1;
  double thermostatTarget = 0., thermostatDelta = 0.;
  bool dynamicRebuild = false, rAtRebuildValid = false;
  int64_t rAtRebuildSlots = 0, rAtRebuildStride = 0;
  long long dynamicRebuildCount = 0;  // rebuilds triggered by displacement (apb_run_steps)
  int stepsSinceRebuild = 0;  // LogicHandler::_stepsSinceLastListRebuild (used when the dynamic trigger is on)

  int64_t numOwned = 0, numHalo = 0;  // refreshed lazily
  bool countsValid = false;
  // owned count as last counted; stays valid while nothing can add / remove owned particles, so that the single-rank
  // rebuild chain (in-place wrap + one-pass halo images, whose number the host knows) needs no counting kernel
  bool ownedKnown = false, countsTrusted = false;
  int64_t ownedCount = 0;

  int fail(int code, const std::string &msg) {
    err = msg;
    return code;
  }
  int failCuda(cudaError_t e, const char *what, const char *file, int line) {
    poisoned = true;
    err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " at " + file + ":" +
          std::to_string(line);
    return APB_ERR_CUDA;
  }
};

#define APB_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) return h->failCuda(e_, #call, __FILE__, __LINE__); \
  } while (0)

#define APB_CHECK(expr)              \
  do {                               \
    int rc_ = (expr);                \
    if (rc_ != APB_OK) return rc_;   \
  } while (0)

#define APB_ENTRY(h)                                                                       \
  if (!(h)) return APB_ERR_INVALID_ARGUMENT;                                               \
  if ((h)->poisoned) return APB_ERR_CUDA;                                                  \
  {                                                                                        \
    cudaError_t e_ = cudaSetDevice((h)->cfg.device);                                       \
    if (e_ != cudaSuccess) return (h)->failCuda(e_, "cudaSetDevice", __FILE__, __LINE__);  \
  }

// grow-only device buffer
int apbEnsure(apb_handle h, DevBuf &b, size_t bytes);
// grow particle storage to at least `slots` slots, preserving [0, nslots)
int apbReserveSlots(apb_handle h, int64_t slots);
int apbEnsurePinned(apb_handle h, size_t bytes);
void apbForgetHaloLinks(apb_handle h);
// opt-in shared-memory sizes of every kernel that needs more than 48 KB, set once per handle (pruned.cu, build.cu)
int apbInitKernelAttributes(apb_handle h);
int apbInitPrunedAttributes(apb_handle h);
int apbInitBuildAttributes(apb_handle h);
// permute the whole storage: slot q of the new order takes old slot perm[q] (perm < 0: dummy); swaps double buffers
int apbPermuteStorage(apb_handle h, const int *perm, int64_t newSlots);

// exclusive scan of n int32 (device), returns total through *totalDev (device int64) if non-null
int apbExclusiveScan(apb_handle h, const int *in, int *out, int64_t n, long long *totalDev);

// build.cu
int apbRebuildLinkedCells(apb_handle h);
int apbRebuildVCL(apb_handle h, int newton3);
int apbBuildPruned(apb_handle h, int newton3);
void apbComputeLCGeom(const apb_config &cfg, LCGeom &g);
int apbComputeStencil(apb_handle h);

// perf.cu
int apbLoopTimingRecord(apb_handle h, int phase, bool begin);

// dynamics.cu
int apbRemapHaloLinks(apb_handle h, const int *perm, int64_t nOld, int64_t nNew);
void apbCommDestroy(apb_handle h);

// control.cu
int apbSnapshotRebuildPositions(apb_handle h);
// dynamics.cu: in-place ncclAllReduce of `count` doubles / int32 on the handle's stream (sum or max); no-op for one rank
int apbAllReduce(apb_handle h, void *dev, int count, int isDouble, int isMax);

// lj.cu
int apbComputeLJ(apb_handle h, int traversal, const apb_functor *f, int newton3, apb_traversal_result *out);

static inline int apbDivUp(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// ---- device helpers ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// CellBlock3D::get3DIndexOfPosition (containers/CellBlock3D.h:321-349) + threeToOneD
// (utils/ThreeDimensionalMapping.h:29-32). No product-sum appears, so no FMA contraction can change the result.
__host__ __device__ inline int apbCellIndexLC(const LCGeom &g, double px, double py, double pz) {
  const double pos[3] = {px, py, pz};
  int idx[3];
  for (int d = 0; d < 3; ++d) {
    const long long value =
        static_cast<long long>(floor((pos[d] - g.boxMin[d]) * g.cellLengthReciprocal[d])) + g.cellsPerInteractionLength;
    long long v = value < 0 ? 0 : value;
    if (v > g.cellsPerDim[d] - 1) v = g.cellsPerDim[d] - 1;
    int c = static_cast<int>(v);
    if (pos[d] >= g.boxMax[d]) {
      const int firstUpperHalo = g.cellsPerDim[d] - g.cellsPerInteractionLength;
      c = c > firstUpperHalo ? c : firstUpperHalo;
    } else if (pos[d] < g.boxMin[d] && c == g.cellsPerInteractionLength) {
      --c;
    } else if (pos[d] < g.boxMax[d] && c == g.cellsPerDim[d] - g.cellsPerInteractionLength) {
      --c;
    }
    idx[d] = c;
  }
  return (idx[2] * g.cellsPerDim[1] + idx[1]) * g.cellsPerDim[0] + idx[0];
}

// ClusterTowerBlock2D::getTowerIndex2DAtPosition (ClusterTowerBlock2D.h:220-245), 1-D = x + y*nx (:258-262)
__host__ __device__ inline int apbTowerIndex(const VCLGeom &g, double px, double py) {
  const double pos[2] = {px, py};
  int idx[2];
  for (int d = 0; d < 2; ++d) {
    const long long value = static_cast<long long>(floor((pos[d] - g.boxMin[d]) * g.sideReciprocal[d])) +
                            g.numTowersPerInteractionLength;
    long long v = value < 0 ? 0 : value;
    if (v > g.towersPerDim[d] - 1) v = g.towersPerDim[d] - 1;
    int c = static_cast<int>(v);
    if (pos[d] >= g.haloBoxMax[d]) {
      c = g.towersPerDim[d] - 1;
    } else if (pos[d] < g.haloBoxMin[d]) {
      c = 0;
    }
    idx[d] = c;
  }
  return idx[0] + idx[1] * g.towersPerDim[0];
}

// a cell can hold owned particles iff it lies in the non-halo block (CellBlock3D::cellCanContainOwnedParticles)
__host__ __device__ inline bool apbCellCanOwn(const LCGeom &g, int cx, int cy, int cz) {
  const int o = g.cellsPerInteractionLength;
  return cx >= o && cx < g.cellsPerDim[0] - o && cy >= o && cy < g.cellsPerDim[1] - o && cz >= o &&
         cz < g.cellsPerDim[2] - o;
}
// Accumulators a functor did not ask for read as zero, whichever entry point produced them (the statistics kernels run
// if either flag is set): one rule for apb_compute_interactions, apb_force_step_by_id and apb_run_steps.
inline void apbMaskResultByFlags(apb_traversal_result &r, int32_t flags) {
  if (!(flags & APB_FUNCTOR_CALC_GLOBALS)) {
    r.upot_sum = 0.;
    r.virial_sum[0] = r.virial_sum[1] = r.virial_sum[2] = 0.;
    r.num_global_calcs_n3 = r.num_global_calcs_no_n3 = 0;
  }
  if (!(flags & APB_FUNCTOR_COUNT_FLOPS)) {
    r.num_dist_calls = r.num_kernel_calls_n3 = r.num_kernel_calls_no_n3 = 0;
    r.num_global_calcs_n3 = r.num_global_calcs_no_n3 = 0;
  }
}

// gpuLinkedCells kernel variants: 0 = one thread per slot walking the stencil cells (round 1), 1 = one warp per slot
// (lc_warp.cuh), 3 = one thread per slot over cached partner lists (SPH functors). Default: 1 below 16 384 slots (one
// thread per slot cannot fill 148 SMs), `large` above. APB_LC_KERNEL=thread|warp|list overrides (A/B runs; measured in
// profiles/r02_lc_kernels.txt).
inline int apbLCKernelVariant(int64_t numSlots, int large) {
  static const int forced = [] {
    const char *e = getenv("APB_LC_KERNEL");
    if (!e) return -1;
    return e[0] == 't' ? 0 : (e[0] == 'w' ? 1 : 3);
  }();
  if (forced >= 0) return forced;
  return numSlots < 16384 ? 1 : large;
}
#endif
