// Device-resident simulation-loop pieces around the force step, so that a time step never round-trips through the host
// mirror: Störmer-Verlet integration (examples/md-flexible/src/TimeDiscretization.cpp:16-68, 152-165), particle
// migration and halo exchange with md-flexible's regular-grid protocol
// (examples/md-flexible/src/domainDecomposition/RegularGridDecomposition.cpp:159-301): per dimension x, y, z in
// sequence, left + right neighbour, already received halos are forwarded so that edges and corners propagate with only
// six peers. Between ranks the packed buffers travel with NCCL send/recv (NVLink); a rank that is its own neighbour
// (single GPU, periodic box) short-circuits to a device-local copy. All kernels are HBM-bound select / pack / scatter.
#include <dlfcn.h>

#include <algorithm>

#include <chrono>

#include "internal.cuh"

// ---- NCCL, loaded lazily (torch has usually loaded libnccl.so.2 already; the library must also load without it) ----
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
} g_nccl;

bool loadNccl() {
  if (g_nccl.lib) return true;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return false;
#define LOADSYM(field, name)                                          \
  *reinterpret_cast<void **>(&g_nccl.field) = dlsym(lib, name);       \
  if (!g_nccl.field) return false;
  LOADSYM(GetUniqueId, "ncclGetUniqueId")
  LOADSYM(CommInitRank, "ncclCommInitRank")
  LOADSYM(CommDestroy, "ncclCommDestroy")
  LOADSYM(Send, "ncclSend")
  LOADSYM(Recv, "ncclRecv")
  LOADSYM(GroupStart, "ncclGroupStart")
  LOADSYM(GroupEnd, "ncclGroupEnd")
  LOADSYM(AllReduce, "ncclAllReduce")
  LOADSYM(GetErrorString, "ncclGetErrorString")
#undef LOADSYM
  g_nccl.lib = lib;
  return true;
}
}  // namespace

#define APB_NCCL(call)                                                                                        \
  do {                                                                                                        \
    int r_ = (call);                                                                                          \
    if (r_ != 0) {                                                                                            \
      h->poisoned = true;                                                                                     \
      return h->fail(APB_ERR_NCCL, std::string("NCCL error: ") + g_nccl.GetErrorString(r_) + " in " #call);   \
    }                                                                                                         \
  } while (0)

extern "C" int apb_comm_get_unique_id(void *out128) {
  if (!out128) return APB_ERR_INVALID_ARGUMENT;
  if (!loadNccl()) return APB_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return APB_ERR_NCCL;
  std::memcpy(out128, &id, 128);
  return APB_OK;
}

extern "C" int apb_comm_init(apb_handle h, int32_t nranks, int32_t rank, const void *uniqueId128) {
  APB_ENTRY(h);
  if (nranks < 1 || rank < 0 || rank >= nranks) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_comm_init: bad rank");
  h->nranks = nranks;
  h->myRank = rank;
  if (nranks == 1) return APB_OK;
  if (!uniqueId128) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_comm_init: null unique id");
  if (!loadNccl()) return h->fail(APB_ERR_NCCL, "apb_comm_init: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  std::memcpy(&id, uniqueId128, 128);
  ncclComm_t comm = nullptr;
  APB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  h->comm = comm;
  return APB_OK;
}

void apbCommDestroy(apb_handle h) {
  for (auto &o : h->p2pOpened)
    if (o.second) cudaIpcCloseMemHandle(o.second);
  h->p2pOpened.clear();
  if (h->p2pArena) cudaFree(h->p2pArena);
  if (h->p2pCounters) cudaFree(h->p2pCounters);
  h->p2pArena = nullptr;
  h->p2pCounters = nullptr;
  h->p2pState = 0;
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(h->comm));
  h->comm = nullptr;
}

extern "C" int apb_set_decomposition(apb_handle h, const double *globalMin, const double *globalMax,
                                     const int32_t *neighbors6, const int32_t *periodic3) {
  APB_ENTRY(h);
  if (!globalMin || !globalMax || !neighbors6 || !periodic3)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_set_decomposition: null argument");
  for (int d = 0; d < 3; ++d) {
    h->globalMin[d] = globalMin[d];
    h->globalMax[d] = globalMax[d];
    h->periodic[d] = periodic3[d] != 0;
    for (int s = 0; s < 2; ++s) {
      const int nb = neighbors6[2 * d + s];
      if (nb < 0 || nb >= h->nranks) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_set_decomposition: neighbour rank out of range");
      h->neighbor[d][s] = nb;
    }
  }
  h->decompositionSet = true;
  return APB_OK;
}

int apbAllReduce(apb_handle h, void *dev, int count, int isDouble, int isMax) {
  if (h->nranks == 1) return APB_OK;
  APB_NCCL(g_nccl.AllReduce(dev, dev, static_cast<size_t>(count), isDouble ? ncclFloat64 : ncclInt32, isMax ? ncclMax : ncclSum,
                            static_cast<ncclComm_t>(h->comm), h->stream));
  return APB_OK;
}

extern "C" int apb_allreduce_globals(apb_handle h, apb_traversal_result *inout) {
  APB_ENTRY(h);
  if (!inout) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_allreduce_globals: null argument");
  if (h->nranks == 1) return APB_OK;
  // Simulation.cpp:319-322 reduces potential energy and virial with MPI_Reduce(SUM); counters are summed likewise
  double *d = reinterpret_cast<double *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 128);
  double host[9] = {inout->upot_sum, inout->virial_sum[0], inout->virial_sum[1], inout->virial_sum[2],
                    static_cast<double>(inout->num_dist_calls), static_cast<double>(inout->num_kernel_calls_n3),
                    static_cast<double>(inout->num_kernel_calls_no_n3), static_cast<double>(inout->num_global_calcs_n3),
                    static_cast<double>(inout->num_global_calcs_no_n3)};
  APB_CUDA(cudaMemcpyAsync(d, host, sizeof(host), cudaMemcpyHostToDevice, h->stream));
  APB_NCCL(g_nccl.AllReduce(d, d, 9, ncclFloat64, ncclSum, static_cast<ncclComm_t>(h->comm), h->stream));
  APB_CUDA(cudaMemcpyAsync(host, d, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  inout->upot_sum = host[0];
  for (int k = 0; k < 3; ++k) inout->virial_sum[k] = host[1 + k];
  inout->num_dist_calls = static_cast<uint64_t>(host[4]);
  inout->num_kernel_calls_n3 = static_cast<uint64_t>(host[5]);
  inout->num_kernel_calls_no_n3 = static_cast<uint64_t>(host[6]);
  inout->num_global_calcs_n3 = static_cast<uint64_t>(host[7]);
  inout->num_global_calcs_no_n3 = static_cast<uint64_t>(host[8]);
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// time integration
// ------------------------------------------------------------------------------------------------------------------
// calculatePositionsAndResetForces (TimeDiscretization.cpp:16-68): oldF = f; f = globalForce;
// r += v*dt + f*dt^2/(2m)   (owned particles only)
__global__ void kIntegratePositions(int64_t n, const int32_t *__restrict__ own, const int32_t *__restrict__ type,
                                    const double *__restrict__ massOfType, int numTypes, double dt, double gx, double gy,
                                    double gz, double *x, double *y, double *z, const double *vx, const double *vy,
                                    const double *vz, double *fx, double *fy, double *fz, double *ofx, double *ofy,
                                    double *ofz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  const int t = type[i];
  const double m = massOfType[t < numTypes ? t : 0];
  const double s = dt * dt / (2 * m);
  const double Fx = fx[i], Fy = fy[i], Fz = fz[i];
  ofx[i] = Fx;
  ofy[i] = Fy;
  ofz[i] = Fz;
  fx[i] = gx;
  fy[i] = gy;
  fz[i] = gz;
  x[i] += vx[i] * dt + Fx * s;
  y[i] += vy[i] * dt + Fy * s;
  z[i] += vz[i] * dt + Fz * s;
}

// calculateVelocities (TimeDiscretization.cpp:152-165): v += (f + oldF) * dt/(2m)
__global__ void kIntegrateVelocities(int64_t n, const int32_t *__restrict__ own, const int32_t *__restrict__ type,
                                     const double *__restrict__ massOfType, int numTypes, double dt, double *vx,
                                     double *vy, double *vz, const double *fx, const double *fy, const double *fz,
                                     const double *ofx, const double *ofy, const double *ofz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  const int t = type[i];
  const double m = massOfType[t < numTypes ? t : 0];
  const double s = dt / (2 * m);
  vx[i] += (fx[i] + ofx[i]) * s;
  vy[i] += (fy[i] + ofy[i]) * s;
  vz[i] += (fz[i] + ofz[i]) * s;
}

// calculateVelocities of step s followed by calculatePositionsAndResetForces of step s + 1 in one pass over the
// particles (same arithmetic, expression by expression, as the two kernels above): 24 instead of 30 column passes.
// RESET = false: the force column is not reset here because the next force kernel stores instead of adding (kLJPruned,
// PrunedForceArgs::overwrite), and oldForce = force is not copied: the caller swaps the two column pointers (the forces
// just used become the old forces; the column that held the old forces is overwritten by the next force kernel - for
// owned particles; halo copies carry zeros in both, dummies are nobody's business). 18 column passes.
template <bool RESET>
__global__ void kIntegrateVelocitiesPositions(int64_t n, const int32_t *__restrict__ own, const int32_t *__restrict__ type,
                                              const double *__restrict__ massOfType, int numTypes, double dt, double gx,
                                              double gy, double gz, double *x, double *y, double *z, double *vx, double *vy,
                                              double *vz, double *fx, double *fy, double *fz, double *ofx, double *ofy,
                                              double *ofz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  const int t = type[i];
  const double m = massOfType[t < numTypes ? t : 0];
  const double sv = dt / (2 * m), sx = dt * dt / (2 * m);
  const double Fx = fx[i], Fy = fy[i], Fz = fz[i];
  const double Vx = vx[i] + (Fx + ofx[i]) * sv, Vy = vy[i] + (Fy + ofy[i]) * sv, Vz = vz[i] + (Fz + ofz[i]) * sv;
  vx[i] = Vx;
  vy[i] = Vy;
  vz[i] = Vz;
  if (RESET) {  // (without the reset the host swaps the force and oldForce columns instead of copying one into the other)
    ofx[i] = Fx;
    ofy[i] = Fy;
    ofz[i] = Fz;
  }
  if (RESET) {
    fx[i] = gx;
    fy[i] = gy;
    fz[i] = gz;
  }
  x[i] += Vx * dt + Fx * sx;
  y[i] += Vy * dt + Fy * sx;
  z[i] += Vz * dt + Fz * sx;
}

static int uploadMasses(apb_handle h, const double *mass, int numTypes) {
  if (numTypes <= 0 || !mass) return h->fail(APB_ERR_INVALID_ARGUMENT, "integrate: need at least one type mass");
  if (h->massHost.size() != static_cast<size_t>(numTypes) ||
      std::memcmp(h->massHost.data(), mass, sizeof(double) * numTypes) != 0) {
    APB_CHECK(apbEnsure(h, h->massDev, sizeof(double) * numTypes));
    APB_CUDA(cudaMemcpyAsync(h->massDev.p, mass, sizeof(double) * numTypes, cudaMemcpyHostToDevice, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    h->massHost.assign(mass, mass + numTypes);
  }
  return APB_OK;
}

extern "C" int apb_integrate_positions(apb_handle h, double dt, const double *massOfType, int32_t numTypes,
                                       const double *globalForce) {
  APB_ENTRY(h);
  if (!h->active[APB_COL_OLDFX]) return h->fail(APB_ERR_NOT_APPLICABLE, "particle kind has no oldF columns");
  APB_CHECK(uploadMasses(h, massOfType, numTypes));
  h->ownedInsideBox = false;
  if (h->nslots == 0) return APB_OK;
  const double g[3] = {globalForce ? globalForce[0] : 0., globalForce ? globalForce[1] : 0.,
                       globalForce ? globalForce[2] : 0.};
  ++h->launchCount, kIntegratePositions<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
      h->nslots, h->own, h->type, static_cast<const double *>(h->massDev.p), numTypes, dt, g[0], g[1], g[2],
      h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->col[APB_COL_VX], h->col[APB_COL_VY], h->col[APB_COL_VZ],
      h->col[APB_COL_FX], h->col[APB_COL_FY], h->col[APB_COL_FZ], h->col[APB_COL_OLDFX], h->col[APB_COL_OLDFY],
      h->col[APB_COL_OLDFZ]);
  APB_CUDA(cudaGetLastError());
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// velocities of the finished step + positions of the next one (apb_run_steps between two steps)
static int integrateVelocitiesPositions(apb_handle h, double dt, const double *massOfType, int32_t numTypes,
                                        const double *globalForce, bool resetForces = true) {
  if (!h->active[APB_COL_OLDFX]) return h->fail(APB_ERR_NOT_APPLICABLE, "particle kind has no oldF columns");
  APB_CHECK(uploadMasses(h, massOfType, numTypes));
  h->ownedInsideBox = false;
  if (h->nslots == 0) return APB_OK;
  const double g[3] = {globalForce ? globalForce[0] : 0., globalForce ? globalForce[1] : 0.,
                       globalForce ? globalForce[2] : 0.};
  ++h->launchCount;
  if (resetForces)
    kIntegrateVelocitiesPositions<true><<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
        h->nslots, h->own, h->type, static_cast<const double *>(h->massDev.p), numTypes, dt, g[0], g[1], g[2],
        h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->col[APB_COL_VX], h->col[APB_COL_VY], h->col[APB_COL_VZ],
        h->col[APB_COL_FX], h->col[APB_COL_FY], h->col[APB_COL_FZ], h->col[APB_COL_OLDFX], h->col[APB_COL_OLDFY],
        h->col[APB_COL_OLDFZ]);
  else
    kIntegrateVelocitiesPositions<false><<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
        h->nslots, h->own, h->type, static_cast<const double *>(h->massDev.p), numTypes, dt, g[0], g[1], g[2],
        h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->col[APB_COL_VX], h->col[APB_COL_VY], h->col[APB_COL_VZ],
        h->col[APB_COL_FX], h->col[APB_COL_FY], h->col[APB_COL_FZ], h->col[APB_COL_OLDFX], h->col[APB_COL_OLDFY],
        h->col[APB_COL_OLDFZ]);
  APB_CUDA(cudaGetLastError());
  if (!resetForces)
    for (int d = 0; d < 3; ++d) std::swap(h->col[APB_COL_FX + d], h->col[APB_COL_OLDFX + d]);
  return APB_OK;
}

extern "C" int apb_integrate_velocities(apb_handle h, double dt, const double *massOfType, int32_t numTypes) {
  APB_ENTRY(h);
  if (!h->active[APB_COL_OLDFX]) return h->fail(APB_ERR_NOT_APPLICABLE, "particle kind has no oldF columns");
  APB_CHECK(uploadMasses(h, massOfType, numTypes));
  if (h->nslots == 0) return APB_OK;
  ++h->launchCount, kIntegrateVelocities<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
      h->nslots, h->own, h->type, static_cast<const double *>(h->massDev.p), numTypes, dt, h->col[APB_COL_VX],
      h->col[APB_COL_VY], h->col[APB_COL_VZ], h->col[APB_COL_FX], h->col[APB_COL_FY], h->col[APB_COL_FZ],
      h->col[APB_COL_OLDFX], h->col[APB_COL_OLDFY], h->col[APB_COL_OLDFZ]);
  APB_CUDA(cudaGetLastError());
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// exchange machinery shared by halo exchange and migration
// ------------------------------------------------------------------------------------------------------------------
// Selection for one dimension. mode 0 = halo selection, mode 1 = migration. Ordered compaction without N-sized scans:
// kSelectCount (per-block counts) -> kScanBlockCounts (one block) -> kSelectWrite (slot indices in ascending order).
struct SelectArgs {
  int64_t n;
  int mode;
  const double *pos;
  const int32_t *own;
  double lmin, lmax, il, skin;
  int sendLeft, sendRight;
};

__device__ __forceinline__ void selectFlags(const SelectArgs &a, int64_t i, int &l, int &r) {
  l = r = 0;
  if (i >= a.n) return;
  const int o = a.own[i];
  const double p = a.pos[i];
  if (a.mode == 0) {
    if (o == APB_OWN_OWNED) {
      // collectHaloParticlesForLeft/RightNeighbor (:509-545): owned particles in [lmin, lmin+il) / [lmax-il, lmax)
      l = p >= a.lmin && p < a.lmin + a.il;
      r = p >= a.lmax - a.il && p < a.lmax;
    } else if (o == APB_OWN_HALO) {
      // forwarding of already received halos (:204-225): [lmin-skin, lmin+il) else [lmax-il, lmax+skin)
      if (p >= a.lmin - a.skin && p < a.lmin + a.il)
        l = 1;
      else if (p >= a.lmax - a.il && p < a.lmax + a.skin)
        r = 1;
    }
  } else if (o == APB_OWN_OWNED) {
    // categorizeParticlesIntoLeftAndRightNeighbor (:545-593): below the local box -> left, above or at max -> right
    l = p < a.lmin;
    r = p >= a.lmax;
  }
  l = l && a.sendLeft;
  r = r && a.sendRight;
}

#define SEL_BLOCK 256
__global__ void __launch_bounds__(SEL_BLOCK) kSelectCount(SelectArgs a, int2 *__restrict__ blockCounts) {
  __shared__ int sl[SEL_BLOCK / 32], sr[SEL_BLOCK / 32];
  int l, r;
  selectFlags(a, static_cast<int64_t>(blockIdx.x) * SEL_BLOCK + threadIdx.x, l, r);
  const int cl = __popc(__ballot_sync(0xffffffffu, l)), cr = __popc(__ballot_sync(0xffffffffu, r));
  if ((threadIdx.x & 31) == 0) {
    sl[threadIdx.x >> 5] = cl;
    sr[threadIdx.x >> 5] = cr;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tl = 0, tr = 0;
    for (int w = 0; w < SEL_BLOCK / 32; ++w) {
      tl += sl[w];
      tr += sr[w];
    }
    blockCounts[blockIdx.x] = make_int2(tl, tr);
  }
}

// exclusive scan of the per-block counts by one block; totals[0..1] = number selected for the left / right neighbour
__global__ void __launch_bounds__(1024) kScanBlockCounts(int numBlocks, int2 *__restrict__ counts, long long *totals) {
  __shared__ int wl[32], wr[32];
  __shared__ int carryL, carryR;
  if (threadIdx.x == 0) carryL = carryR = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < numBlocks; base += 1024) {
    const int b = base + threadIdx.x;
    const int2 c = b < numBlocks ? counts[b] : make_int2(0, 0);
    int il = c.x, ir = c.y;
    for (int o = 1; o < 32; o <<= 1) {
      const int tl = __shfl_up_sync(0xffffffffu, il, o), tr = __shfl_up_sync(0xffffffffu, ir, o);
      if (lane >= o) {
        il += tl;
        ir += tr;
      }
    }
    if (lane == 31) {
      wl[warp] = il;
      wr[warp] = ir;
    }
    __syncthreads();
    if (warp == 0) {
      int vl = wl[lane], vr = wr[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int tl = __shfl_up_sync(0xffffffffu, vl, o), tr = __shfl_up_sync(0xffffffffu, vr, o);
        if (lane >= o) {
          vl += tl;
          vr += tr;
        }
      }
      wl[lane] = vl;
      wr[lane] = vr;
    }
    __syncthreads();
    const int offL = carryL + (warp ? wl[warp - 1] : 0), offR = carryR + (warp ? wr[warp - 1] : 0);
    if (b < numBlocks) counts[b] = make_int2(offL + il - c.x, offR + ir - c.y);
    __syncthreads();
    if (threadIdx.x == 1023) {
      carryL = offL + il;
      carryR = offR + ir;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] = carryL;
    totals[1] = carryR;
  }
}

__global__ void __launch_bounds__(SEL_BLOCK) kSelectWrite(SelectArgs a, const int2 *__restrict__ blockOffsets,
                                                          int *__restrict__ idxL, int *__restrict__ idxR) {
  __shared__ int sl[SEL_BLOCK / 32], sr[SEL_BLOCK / 32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SEL_BLOCK + threadIdx.x;
  int l, r;
  selectFlags(a, i, l, r);
  const unsigned bl = __ballot_sync(0xffffffffu, l), br = __ballot_sync(0xffffffffu, r);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sl[warp] = __popc(bl);
    sr[warp] = __popc(br);
  }
  __syncthreads();
  int offL = blockOffsets[blockIdx.x].x, offR = blockOffsets[blockIdx.x].y;
  for (int w = 0; w < warp; ++w) {
    offL += sl[w];
    offR += sr[w];
  }
  const unsigned below = (1u << lane) - 1u;
  if (l) idxL[offL + __popc(bl & below)] = static_cast<int>(i);
  if (r) idxR[offR + __popc(br & below)] = static_cast<int>(i);
}

struct PackArgs {
  const int *idx;  // selected slots, ascending
  int ncols;
  const double *src[APB_NUM_COLUMNS];
  const int64_t *id;
  const int32_t *type;
  int32_t *own;
  int markDummy;  // migration: the packed particle leaves this rank
  int shiftCol;   // index (within the packed columns) of the coordinate that gets the periodic shift
  double shift;
  double wrapMin, wrapMax;  // migration: clamp like the reference's nextafter guard
  int clamp;
  int64_t count;  // number of selected particles (row length of the packed buffer)
  double *outCols;  // [ncols][count]
  int64_t *outId;
  int32_t *outType;
  int *outIdx;  // selected slot indices (may be null)
};

__global__ void kPack(PackArgs a) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= a.count) return;
  const int i = a.idx[q];
  for (int c = 0; c < a.ncols; ++c) {
    double v = a.src[c][i];
    if (c == a.shiftCol) {
      v += a.shift;
      if (a.clamp) {
        // RegularGridDecomposition.cpp:575-590: a wrapped coordinate must end up inside [globalMin, globalMax)
        if (v >= a.wrapMax) v = nextafter(a.wrapMax, a.wrapMin);
        if (v < a.wrapMin) v = a.wrapMin;
      }
    }
    a.outCols[static_cast<size_t>(c) * a.count + q] = v;
  }
  a.outId[q] = a.id[i];
  a.outType[q] = a.type[i];
  if (a.outIdx) a.outIdx[q] = i;
  if (a.markDummy) a.own[i] = APB_OWN_DUMMY;
}

struct UnpackArgs {
  int64_t count;
  int64_t firstSlot;
  int ncolsPacked;
  const double *inCols;
  const int64_t *inId;
  const int32_t *inType;
  int nAll;
  double *dst[APB_NUM_COLUMNS];
  int packedIndexOfCol[APB_NUM_COLUMNS];  // -1: zero-fill
  int64_t *id;
  int32_t *type, *own;
  int ownership;
  int *recvSlot;  // may be null
};

__global__ void kUnpackAppend(UnpackArgs a) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= a.count) return;
  const int64_t s = a.firstSlot + q;
  for (int c = 0; c < a.nAll; ++c) {
    const int pc = a.packedIndexOfCol[c];
    a.dst[c][s] = pc >= 0 ? a.inCols[static_cast<size_t>(pc) * a.count + q] : 0.;
  }
  a.id[s] = a.inId[q];
  a.type[s] = a.inType[q];
  a.own[s] = a.ownership;
  if (a.recvSlot) a.recvSlot[q] = static_cast<int>(s);
}


// both sides of one dimension in one launch (the refresh is launch-latency bound: a few thousand particles per message)
__global__ void kGatherPositions2(int64_t m0, int64_t m1, const int *__restrict__ idx0, const int *__restrict__ idx1,
                                  const double *x, const double *y, const double *z, int dim, double shift0,
                                  double shift1, double *out0, double *out1) {
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m0 + m1) return;
  const bool second = q >= m0;
  if (second) q -= m0;
  const int64_t m = second ? m1 : m0;
  const int s = second ? idx1[q] : idx0[q];
  double *out = second ? out1 : out0;
  const double shift = second ? shift1 : shift0;
  double px = nan(""), py = 0., pz = 0.;
  if (s >= 0) {
    px = x[s];
    py = y[s];
    pz = z[s];
    if (dim == 0) px += shift;
    if (dim == 1) py += shift;
    if (dim == 2) pz += shift;
  }
  out[q] = px;
  out[m + q] = py;
  out[2 * m + q] = pz;
}
__global__ void kScatterPositions2(int64_t m0, int64_t m1, const int *__restrict__ slot0, const int *__restrict__ slot1,
                                   const double *in0, const double *in1, double *x, double *y, double *z) {
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m0 + m1) return;
  const bool second = q >= m0;
  if (second) q -= m0;
  const int64_t m = second ? m1 : m0;
  const int s = second ? slot1[q] : slot0[q];
  const double *in = second ? in1 : in0;
  if (s < 0 || isnan(in[q])) return;
  x[s] = in[q];
  y[s] = in[m + q];
  z[s] = in[2 * m + q];
}
// ---- halo refresh through peer memory ---------------------------------------------------------------------------------
// kPushHalo: gathers the positions recorded for the two neighbours of dimension `dim` (shifted at a periodic global
// boundary) and stores them directly into the neighbours' arenas (NVLink peer stores through the IPC mapping); the last
// block to finish publishes the sequence number in the neighbours' flags. kPullHalo spins on the own flags until both
// neighbours have published `seq`, then scatters the received positions into the halo slots.
__global__ void kPushHalo(int64_t m0, int64_t m1, const int *__restrict__ idx0, const int *__restrict__ idx1,
                          const double *x, const double *y, const double *z, int dim, double shift0, double shift1,
                          double *out0, double *out1, unsigned long long *flag0, unsigned long long *flag1,
                          unsigned long long seq, int *counter) {
  __shared__ bool isLast;
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < m0 + m1) {
    const bool second = q >= m0;
    if (second) q -= m0;
    const int64_t m = second ? m1 : m0;
    const int s = second ? idx1[q] : idx0[q];
    double *out = second ? out1 : out0;
    const double shift = second ? shift1 : shift0;
    double px = nan(""), py = 0., pz = 0.;
    if (s >= 0) {
      px = x[s];
      py = y[s];
      pz = z[s];
      if (dim == 0) px += shift;
      if (dim == 1) py += shift;
      if (dim == 2) pz += shift;
    }
    out[q] = px;
    out[m + q] = py;
    out[2 * m + q] = pz;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) isLast = atomicAdd(counter, 1) == static_cast<int>(gridDim.x) - 1;
  __syncthreads();
  if (isLast && threadIdx.x == 0) {
    *counter = 0;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(flag0) = seq;
    *reinterpret_cast<volatile unsigned long long *>(flag1) = seq;
    __threadfence_system();
  }
}

// Waits until *flag >= seq. A neighbour that never publishes (it failed, or the ranks call the exchange in different
// orders) must not hang the GPU: after about ten seconds the wait gives up and raises *timedOut, which the host turns
// into an error at its next read-back.
#define P2P_SPIN_LIMIT_CYCLES 20000000000LL
__device__ __forceinline__ void p2pWait(const unsigned long long *flag, unsigned long long seq, int *timedOut) {
  const long long t0 = clock64();
  while (*reinterpret_cast<const volatile unsigned long long *>(flag) < seq) {
    if (clock64() - t0 > P2P_SPIN_LIMIT_CYCLES) {
      *reinterpret_cast<volatile int *>(timedOut) = 1;
      break;
    }
  }
}

__global__ void kPullHalo(int64_t m0, int64_t m1, const int *__restrict__ slot0, const int *__restrict__ slot1,
                          const double *in0, const double *in1, const unsigned long long *flag0,
                          const unsigned long long *flag1, unsigned long long seq, double *x, double *y, double *z,
                          int *timedOut) {
  if (threadIdx.x == 0) {
    p2pWait(flag0, seq, timedOut);
    p2pWait(flag1, seq, timedOut);
    __threadfence_system();
  }
  __syncthreads();
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m0 + m1) return;
  const bool second = q >= m0;
  if (second) q -= m0;
  const int64_t m = second ? m1 : m0;
  const int s = second ? slot1[q] : slot0[q];
  const double *in = second ? in1 : in0;
  if (s < 0) return;
  const double px = __ldcg(in + q);  // written by the peer: read through L2, not a stale L1 line
  if (isnan(px)) return;
  x[s] = px;
  y[s] = __ldcg(in + m + q);
  z[s] = __ldcg(in + 2 * m + q);
}

static inline double *p2pRegion(void *arena, size_t cap, int d, int s, int parity) {
  return static_cast<double *>(arena) + (static_cast<size_t>((d * 2 + s) * 2 + parity)) * cap;
}
static inline unsigned long long *p2pFlag(void *arena, size_t cap, int d, int s) {
  return reinterpret_cast<unsigned long long *>(static_cast<double *>(arena) + 12 * cap) + (d * 2 + s);
}
// count slots of the generating exchange: {count, sequence number} per (dimension, side, parity)
static inline long long *p2pCountSlot(void *arena, size_t cap, int d, int s, int parity) {
  return reinterpret_cast<long long *>(reinterpret_cast<char *>(static_cast<double *>(arena) + 12 * cap) + 64) +
         2 * ((d * 2 + s) * 2 + parity);
}
// my two send counts (device values of the selection) go straight into the neighbours' count slots
__global__ void kPushCounts(const long long *__restrict__ counts, long long *slot0, long long *slot1, long long seq) {
  *reinterpret_cast<volatile long long *>(slot0) = counts[0];
  *reinterpret_cast<volatile long long *>(slot1) = counts[1];
  __threadfence_system();
  *reinterpret_cast<volatile long long *>(slot0 + 1) = seq;
  *reinterpret_cast<volatile long long *>(slot1 + 1) = seq;
  __threadfence_system();
}
__global__ void kPullCounts(const long long *slot0, const long long *slot1, long long seq, long long *out, int *timedOut) {
  p2pWait(reinterpret_cast<const unsigned long long *>(slot0 + 1), static_cast<unsigned long long>(seq), timedOut);
  p2pWait(reinterpret_cast<const unsigned long long *>(slot1 + 1), static_cast<unsigned long long>(seq), timedOut);
  __threadfence_system();
  out[0] = *reinterpret_cast<const volatile long long *>(slot0);
  out[1] = *reinterpret_cast<const volatile long long *>(slot1);
  out[2] = *reinterpret_cast<volatile int *>(timedOut);
}

// One-time set-up: allocate the arena, trade IPC handles with the distinct neighbour ranks (NCCL send / recv) and map
// theirs. Every rank of the decomposition reaches this at the same point (its first halo refresh), so the pairwise
// exchange matches up. APB_NO_P2P_HALO=1 keeps the NCCL send / recv refresh (set it on all ranks).
static int ensureP2P(apb_handle h) {
  if (h->p2pState != 0) return APB_OK;
  h->p2pState = -1;
  if (h->nranks <= 1 || !h->comm || getenv("APB_NO_P2P_HALO")) return APB_OK;
  std::vector<int> peers;
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s) {
      const int r = h->neighbor[d][s];
      if (r != h->myRank && std::find(peers.begin(), peers.end(), r) == peers.end()) peers.push_back(r);
    }
  if (peers.empty()) return APB_OK;
  size_t mb = 24;  // per region; 12 regions
  if (const char *e = getenv("APB_HALO_ARENA_MB")) mb = std::max<long>(1, atol(e));
  h->p2pCap = mb * 1024 * 1024 / sizeof(double);
  const size_t bytes = 12 * h->p2pCap * sizeof(double) + 64 + 192;  // regions | 6 flags | 12 count slots
  APB_CUDA(cudaMalloc(&h->p2pArena, bytes));
  APB_CUDA(cudaMemsetAsync(h->p2pArena, 0, bytes, h->stream));
  APB_CUDA(cudaMalloc(reinterpret_cast<void **>(&h->p2pCounters), 4 * sizeof(int)));
  APB_CUDA(cudaMemsetAsync(h->p2pCounters, 0, 4 * sizeof(int), h->stream));
  cudaIpcMemHandle_t mine;
  APB_CUDA(cudaIpcGetMemHandle(&mine, h->p2pArena));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  APB_CHECK(apbEnsure(h, h->xbuf[0], 64 * (peers.size() + 1) + 256));
  char *dev = static_cast<char *>(h->xbuf[0].p);
  APB_CUDA(cudaMemcpyAsync(dev, &mine, 64, cudaMemcpyHostToDevice, h->stream));
  ncclComm_t comm = static_cast<ncclComm_t>(h->comm);
  APB_NCCL(g_nccl.GroupStart());
  for (size_t k = 0; k < peers.size(); ++k) {
    APB_NCCL(g_nccl.Send(dev, 64, ncclInt8, peers[k], comm, h->stream));
    APB_NCCL(g_nccl.Recv(dev + 64 * (k + 1), 64, ncclInt8, peers[k], comm, h->stream));
  }
  APB_NCCL(g_nccl.GroupEnd());
  std::vector<cudaIpcMemHandle_t> theirs(peers.size());
  APB_CUDA(cudaMemcpyAsync(theirs.data(), dev + 64, 64 * peers.size(), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  for (size_t k = 0; k < peers.size(); ++k) {
    void *base = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&base, theirs[k], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return h->fail(APB_ERR_CUDA, std::string("peer-memory halo refresh: cudaIpcOpenMemHandle for rank ") +
                                       std::to_string(peers[k]) + " failed (" + cudaGetErrorString(e) +
                                       "); set APB_NO_P2P_HALO=1 on all ranks to use NCCL send/recv instead");
    }
    h->p2pOpened.emplace_back(peers[k], base);
  }
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s) {
      h->p2pPeer[d][s] = nullptr;
      for (auto &o : h->p2pOpened)
        if (o.first == h->neighbor[d][s]) h->p2pPeer[d][s] = o.second;
    }
  h->p2pState = 1;
  return APB_OK;
}

// a rank that is its own neighbour in this dimension (periodic, one rank wide): what goes out to the right re-enters from
// the left and vice versa, so the refresh is a direct slot-to-slot copy. srcA -> dstA are the right-going particles
// (received "from the left"), srcB -> dstB the left-going ones.
__global__ void kRefreshSelf(int64_t mA, int64_t mB, const int *__restrict__ srcA, const int *__restrict__ dstA,
                             const int *__restrict__ srcB, const int *__restrict__ dstB, double shiftA, double shiftB,
                             int dim, double *x, double *y, double *z) {
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= mA + mB) return;
  const bool second = q >= mA;
  if (second) q -= mA;
  const int s = second ? srcB[q] : srcA[q];
  const int t = second ? dstB[q] : dstA[q];
  if (s < 0 || t < 0) return;  // dropped by the rebuild: the stale copy stays
  const double shift = second ? shiftB : shiftA;
  double px = x[s], py = y[s], pz = z[s];
  if (dim == 0) px += shift;
  if (dim == 1) py += shift;
  if (dim == 2) pz += shift;
  x[t] = px;
  y[t] = py;
  z[t] = pz;
}

__global__ void kRemap(int64_t m, int *idx, const int *__restrict__ inv, int64_t nOld) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int s = idx[q];
  idx[q] = (s >= 0 && s < nOld) ? inv[s] : -1;
}
__global__ void kInvertPerm(int64_t mNew, const int *__restrict__ perm, int *inv) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= mNew) return;
  const int s = perm[q];
  if (s >= 0) inv[s] = static_cast<int>(q);
}

// called by the rebuilds after the storage permutation: slot indices recorded by the halo exchange follow the sort
int apbRemapHaloLinks(apb_handle h, const int *perm, int64_t nOld, int64_t nNew) {
  if (!h->haloLinksValid) return APB_OK;
  APB_CHECK(apbEnsure(h, h->invPerm, sizeof(int) * std::max<int64_t>(nOld, 1)));
  int *inv = static_cast<int *>(h->invPerm.p);
  APB_CUDA(cudaMemsetAsync(inv, 0xFF, sizeof(int) * std::max<int64_t>(nOld, 1), h->stream));
  if (nNew > 0) ++h->launchCount, kInvertPerm<<<apbDivUp(nNew, 256), 256, 0, h->stream>>>(nNew, perm, inv);
  if (h->haloAllMode && h->haloAllN > 0) {
    ++h->launchCount, kRemap<<<apbDivUp(h->haloAllN, 256), 256, 0, h->stream>>>(h->haloAllN, static_cast<int *>(h->haloAllSrc.p), inv, nOld);
    ++h->launchCount, kRemap<<<apbDivUp(h->haloAllN, 256), 256, 0, h->stream>>>(h->haloAllN, static_cast<int *>(h->haloAllDst.p), inv, nOld);
  }
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s) {
      HaloLink &L = h->link[d][s];
      if (L.nSend > 0) ++h->launchCount, kRemap<<<apbDivUp(L.nSend, 256), 256, 0, h->stream>>>(L.nSend, static_cast<int *>(L.sendIdx.p), inv, nOld);
      if (L.nRecv > 0) ++h->launchCount, kRemap<<<apbDivUp(L.nRecv, 256), 256, 0, h->stream>>>(L.nRecv, static_cast<int *>(L.recvSlot.p), inv, nOld);
    }
  APB_CUDA(cudaGetLastError());
  return APB_OK;
}

// exchange of counts and packed payloads with the two neighbours of one dimension.
// sendBuf[s]: packed bytes for neighbour s (0 = left, 1 = right); what the left neighbour sends to its right arrives
// here as "from left", so recvFrom[0] receives the left neighbour's right-going buffer.
static int exchangeCounts(apb_handle h, int d, const long long sendCount[2], long long recvCount[2]) {
  const int left = h->neighbor[d][0], right = h->neighbor[d][1];
  if (h->nranks == 1 || (left == h->myRank && right == h->myRank)) {
    recvCount[0] = sendCount[1];  // my right-going buffer re-enters from my left
    recvCount[1] = sendCount[0];
    return APB_OK;
  }
  long long *dbuf = reinterpret_cast<long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 64);
  APB_CUDA(cudaMemcpyAsync(dbuf, sendCount, 16, cudaMemcpyHostToDevice, h->stream));
  ncclComm_t comm = static_cast<ncclComm_t>(h->comm);
  APB_NCCL(g_nccl.GroupStart());
  APB_NCCL(g_nccl.Send(dbuf + 0, 1, ncclInt64, left, comm, h->stream));
  APB_NCCL(g_nccl.Send(dbuf + 1, 1, ncclInt64, right, comm, h->stream));
  APB_NCCL(g_nccl.Recv(dbuf + 2, 1, ncclInt64, left, comm, h->stream));   // left neighbour's right-going count
  APB_NCCL(g_nccl.Recv(dbuf + 3, 1, ncclInt64, right, comm, h->stream));  // right neighbour's left-going count
  APB_NCCL(g_nccl.GroupEnd());
  APB_CUDA(cudaMemcpyAsync(recvCount, dbuf + 2, 16, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

static int exchangePayload(apb_handle h, int d, void *const sendBuf[2], const size_t sendBytes[2], void *recvBuf[2],
                           const size_t recvBytes[2]) {
  const int left = h->neighbor[d][0], right = h->neighbor[d][1];
  if (h->nranks == 1 || (left == h->myRank && right == h->myRank)) {
    if (recvBytes[0]) APB_CUDA(cudaMemcpyAsync(recvBuf[0], sendBuf[1], recvBytes[0], cudaMemcpyDeviceToDevice, h->stream));
    if (recvBytes[1]) APB_CUDA(cudaMemcpyAsync(recvBuf[1], sendBuf[0], recvBytes[1], cudaMemcpyDeviceToDevice, h->stream));
    return APB_OK;
  }
  ncclComm_t comm = static_cast<ncclComm_t>(h->comm);
  APB_NCCL(g_nccl.GroupStart());
  if (sendBytes[0]) APB_NCCL(g_nccl.Send(sendBuf[0], sendBytes[0], ncclInt8, left, comm, h->stream));
  if (sendBytes[1]) APB_NCCL(g_nccl.Send(sendBuf[1], sendBytes[1], ncclInt8, right, comm, h->stream));
  if (recvBytes[0]) APB_NCCL(g_nccl.Recv(recvBuf[0], recvBytes[0], ncclInt8, left, comm, h->stream));
  if (recvBytes[1]) APB_NCCL(g_nccl.Recv(recvBuf[1], recvBytes[1], ncclInt8, right, comm, h->stream));
  APB_NCCL(g_nccl.GroupEnd());
  return APB_OK;
}

static inline size_t alignUp(size_t v) { return (v + 255) & ~size_t(255); }
static inline size_t packedBytes(int ncols, long long count) {
  return alignUp(sizeof(double) * ncols * count) + alignUp(8 * count) + alignUp(4 * count);
}

// isNearRel (utils/Math.h) as used by the decomposition to detect global boundaries
static bool nearRel(double a, double b) {
  const double m = std::max(std::fabs(a), std::fabs(b));
  return std::fabs(a - b) <= 1e-9 * (m > 0 ? m : 1.);
}

// One dimension of a generating exchange (halo generation or migration).
// mode 0: halo generation (packed columns: x, y, z); mode 1: migration (all active columns).
static int exchangeDim(apb_handle h, int d, int mode, int64_t *outSent = nullptr) {
  if (mode == 1) h->ownedKnown = false;  // migration between ranks changes the number of owned particles
  if (h->nranks > 1) APB_CHECK(ensureP2P(h));  // every rank passes here in the same order: the handle trade matches up
  const int64_t n = h->nslots;
  const double lmin = h->cfg.box_min[d], lmax = h->cfg.box_max[d];
  const double il = h->cfg.cutoff + h->cfg.skin;
  const bool atMin = nearRel(lmin, h->globalMin[d]), atMax = nearRel(lmax, h->globalMax[d]);
  const bool per = h->periodic[d];
  if (!per && atMin && atMax) return APB_OK;
  const int sendLeft = per || !atMin, sendRight = per || !atMax;
  const double L = h->globalMax[d] - h->globalMin[d];
  const int64_t nn = std::max<int64_t>(n, 1);
  const int numBlocks = apbDivUp(nn, SEL_BLOCK);
  APB_CHECK(apbEnsure(h, h->perm, sizeof(int) * nn));
  APB_CHECK(apbEnsure(h, h->sortV, sizeof(int) * nn));
  APB_CHECK(apbEnsure(h, h->key, sizeof(int2) * numBlocks));
  int *pl = static_cast<int *>(h->perm.p), *pr = static_cast<int *>(h->sortV.p);
  int2 *blockCounts = static_cast<int2 *>(h->key.p);
  long long *totals = reinterpret_cast<long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result));
  long long sendCount[2] = {0, 0}, recvCount[2] = {0, 0};
  if (n > 0) {
    SelectArgs sa;
    sa.n = n;
    sa.mode = mode;
    sa.pos = h->col[APB_COL_X + d];
    sa.own = h->own;
    sa.lmin = lmin;
    sa.lmax = lmax;
    sa.il = il;
    sa.skin = h->cfg.skin;
    sa.sendLeft = sendLeft;
    sa.sendRight = sendRight;
    ++h->launchCount, kSelectCount<<<numBlocks, SEL_BLOCK, 0, h->stream>>>(sa, blockCounts);
    ++h->launchCount, kScanBlockCounts<<<1, 1024, 0, h->stream>>>(numBlocks, blockCounts, totals);
    ++h->launchCount, kSelectWrite<<<numBlocks, SEL_BLOCK, 0, h->stream>>>(sa, blockCounts, pl, pr);
    APB_CUDA(cudaGetLastError());
  } else {
    APB_CUDA(cudaMemsetAsync(totals, 0, 16, h->stream));
  }
  const int left = h->neighbor[d][0], right = h->neighbor[d][1];
  const bool remote = h->nranks > 1 && !(left == h->myRank && right == h->myRank);
  if (remote && h->p2pState == 1 && per && h->p2pPeer[d][0] && h->p2pPeer[d][1]) {
    // counts through the peer arenas: no NCCL round and one host read-back (send and receive counts together)
    // sequence number and slot parity per dimension: consecutive exchanges of a dimension alternate between its two count
    // slots whatever the number of exchanging dimensions (a rank may run one exchange ahead of its neighbour)
    const long long seq = static_cast<long long>(++h->p2pCountSeq[d]);
    const int parity = static_cast<int>(seq & 1);
    const int to0 = left == right ? 0 : 1, to1 = left == right ? 1 : 0;  // same matching as the payload (see the refresh)
    ++h->launchCount, kPushCounts<<<1, 1, 0, h->stream>>>(totals, p2pCountSlot(h->p2pPeer[d][0], h->p2pCap, d, to0, parity),
                                                        p2pCountSlot(h->p2pPeer[d][1], h->p2pCap, d, to1, parity), seq);
    ++h->launchCount, kPullCounts<<<1, 1, 0, h->stream>>>(p2pCountSlot(h->p2pArena, h->p2pCap, d, 0, parity),
                                                        p2pCountSlot(h->p2pArena, h->p2pCap, d, 1, parity), seq, totals + 2,
                                                        h->p2pCounters + 3);
    APB_CUDA(cudaGetLastError());
    long long both[5];
    APB_CUDA(cudaMemcpyAsync(both, totals, 40, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    if (both[4] != 0) {
      h->poisoned = true;
      return h->fail(APB_ERR_STATE, "peer-memory exchange: a neighbour rank did not publish its data within ten seconds "
                                    "(rank failed, or the ranks call the exchange in different orders)");
    }
    sendCount[0] = both[0], sendCount[1] = both[1], recvCount[0] = both[2], recvCount[1] = both[3];
  } else {
    APB_CUDA(cudaMemcpyAsync(sendCount, totals, 16, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    APB_CHECK(exchangeCounts(h, d, sendCount, recvCount));
  }
  if (outSent) *outSent += sendCount[0] + sendCount[1];
  // packed columns
  int ncols = 0;
  int colIds[APB_NUM_COLUMNS];
  if (mode == 0) {
    ncols = 3;
    colIds[0] = APB_COL_X;
    colIds[1] = APB_COL_Y;
    colIds[2] = APB_COL_Z;
  } else {
    for (int c = 0; c < APB_NUM_COLUMNS; ++c)
      if (h->active[c]) colIds[ncols++] = c;
  }
  const size_t sb[2] = {packedBytes(ncols, sendCount[0]), packedBytes(ncols, sendCount[1])};
  const size_t rb[2] = {packedBytes(ncols, recvCount[0]), packedBytes(ncols, recvCount[1])};
  APB_CHECK(apbEnsure(h, h->xbuf[0], sb[0] + 256));
  APB_CHECK(apbEnsure(h, h->xbuf[1], sb[1] + 256));
  APB_CHECK(apbEnsure(h, h->xbuf[2], rb[0] + 256));
  APB_CHECK(apbEnsure(h, h->xbuf[3], rb[1] + 256));
  void *sendBuf[2] = {h->xbuf[0].p, h->xbuf[1].p}, *recvBuf[2] = {h->xbuf[2].p, h->xbuf[3].p};
  for (int s = 0; s < 2; ++s) {
    HaloLink &Lk = h->link[d][s];
    if (mode == 0) {
      APB_CHECK(apbEnsure(h, Lk.sendIdx, sizeof(int) * std::max<long long>(sendCount[s], 1)));
      Lk.nSend = sendCount[s];
      // shift applied by the sender at a global periodic boundary (:509-545)
      Lk.shift = s == 0 ? (atMin ? +L : 0.) : (atMax ? -L : 0.);
    }
    if (sendCount[s] == 0) continue;
    PackArgs a;
    a.idx = s == 0 ? pl : pr;
    a.own = h->own;
    a.markDummy = mode == 1;
    a.ncols = ncols;
    a.shiftCol = -1;
    for (int c = 0; c < ncols; ++c) {
      a.src[c] = h->col[colIds[c]];
      if (colIds[c] == APB_COL_X + d) a.shiftCol = c;
    }
    a.id = h->id;
    a.type = h->type;
    a.shift = s == 0 ? (atMin ? +L : 0.) : (atMax ? -L : 0.);
    a.clamp = mode == 1 && a.shift != 0.;
    a.wrapMin = h->globalMin[d];
    a.wrapMax = h->globalMax[d];
    a.count = sendCount[s];
    char *base = static_cast<char *>(sendBuf[s]);
    a.outCols = reinterpret_cast<double *>(base);
    a.outId = reinterpret_cast<int64_t *>(base + alignUp(sizeof(double) * ncols * sendCount[s]));
    a.outType = reinterpret_cast<int32_t *>(base + alignUp(sizeof(double) * ncols * sendCount[s]) + alignUp(8 * sendCount[s]));
    a.outIdx = mode == 0 ? static_cast<int *>(Lk.sendIdx.p) : nullptr;
    ++h->launchCount, kPack<<<apbDivUp(sendCount[s], 256), 256, 0, h->stream>>>(a);
    APB_CUDA(cudaGetLastError());
  }
  APB_CHECK(exchangePayload(h, d, sendBuf, sb, recvBuf, rb));
  // append what arrived: from the left neighbour first, then from the right one
  const long long nRecvTotal = recvCount[0] + recvCount[1];
  if (nRecvTotal > 0) APB_CHECK(apbReserveSlots(h, n + nRecvTotal));
  int64_t first = n;
  for (int s = 0; s < 2; ++s) {
    HaloLink &Lk = h->link[d][s];
    if (mode == 0) {
      APB_CHECK(apbEnsure(h, Lk.recvSlot, sizeof(int) * std::max<long long>(recvCount[s], 1)));
      Lk.nRecv = recvCount[s];
    }
    if (recvCount[s] == 0) continue;
    UnpackArgs u;
    u.count = recvCount[s];
    u.firstSlot = first;
    u.ncolsPacked = ncols;
    char *base = static_cast<char *>(recvBuf[s]);
    u.inCols = reinterpret_cast<const double *>(base);
    u.inId = reinterpret_cast<const int64_t *>(base + alignUp(sizeof(double) * ncols * recvCount[s]));
    u.inType = reinterpret_cast<const int32_t *>(base + alignUp(sizeof(double) * ncols * recvCount[s]) + alignUp(8 * recvCount[s]));
    u.nAll = 0;
    for (int c = 0; c < APB_NUM_COLUMNS; ++c) {
      if (!h->active[c]) continue;
      u.dst[u.nAll] = h->col[c];
      int pc = -1;
      for (int k = 0; k < ncols; ++k)
        if (colIds[k] == c) pc = k;
      u.packedIndexOfCol[u.nAll] = pc;
      ++u.nAll;
    }
    u.id = h->id;
    u.type = h->type;
    u.own = h->own;
    u.ownership = mode == 0 ? APB_OWN_HALO : APB_OWN_OWNED;
    u.recvSlot = mode == 0 ? static_cast<int *>(Lk.recvSlot.p) : nullptr;
    ++h->launchCount, kUnpackAppend<<<apbDivUp(recvCount[s], 256), 256, 0, h->stream>>>(u);
    APB_CUDA(cudaGetLastError());
    first += recvCount[s];
  }
  h->nslots = n + nRecvTotal;
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// ---- single rank, fully periodic: all periodic images in one pass -----------------------------------------------------
// With every dimension its own neighbour the three forwarding rounds of exchangeHaloParticles
// (RegularGridDecomposition.cpp:159-236) produce exactly the images  r + (sx, sy, sz),  s_d in {0, +L_d if r_d in
// [min, min + il), -L_d if r_d in [max - il, max)}, not all zero, of every owned particle (owned particles lie inside the
// box after apb_migrate, so the widened forwarding ranges of halo copies never matter). One selection pass, one append,
// one host read-back instead of three of each; the refresh is one slot-to-slot kernel.
struct ImageArgs {
  int64_t n;
  const int32_t *own;
  const double *x, *y, *z;
  double lo[3], hi[3], il;
};
__device__ __forceinline__ int imageOptions(const ImageArgs &a, int64_t i, int opt[3]) {
  opt[0] = opt[1] = opt[2] = 0;
  if (i >= a.n || a.own[i] != APB_OWN_OWNED) return 0;
  const double p[3] = {a.x[i], a.y[i], a.z[i]};
  int count = 1;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int nearMin = p[d] >= a.lo[d] && p[d] < a.lo[d] + a.il, nearMax = p[d] >= a.hi[d] - a.il && p[d] < a.hi[d];
    opt[d] = nearMin | (nearMax << 1);
    count *= 1 + nearMin + nearMax;
  }
  return count - 1;
}
__global__ void __launch_bounds__(SEL_BLOCK) kImageCount(ImageArgs a, int *__restrict__ blockCounts) {
  __shared__ int sw[SEL_BLOCK / 32];
  int opt[3];
  const int c = imageOptions(a, static_cast<int64_t>(blockIdx.x) * SEL_BLOCK + threadIdx.x, opt);
  const int w = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < SEL_BLOCK / 32; ++k) t += sw[k];
    blockCounts[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(SEL_BLOCK) kImageWrite(ImageArgs a, const int *__restrict__ blockOffsets, int *__restrict__ src,
                                                        int *__restrict__ code) {
  __shared__ int sw[SEL_BLOCK / 32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SEL_BLOCK + threadIdx.x;
  int opt[3];
  const int c = imageOptions(a, i, opt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sw[warp] = incl;
  __syncthreads();
  int pos = blockOffsets[blockIdx.x] + incl - c;
  for (int k = 0; k < warp; ++k) pos += sw[k];
  if (c == 0) return;
  for (int cz = 0; cz < 3; ++cz)
    for (int cy = 0; cy < 3; ++cy)
      for (int cx = 0; cx < 3; ++cx) {
        if (cx + cy + cz == 0) continue;
        if ((cx && !((opt[0] >> (cx - 1)) & 1)) || (cy && !((opt[1] >> (cy - 1)) & 1)) || (cz && !((opt[2] >> (cz - 1)) & 1)))
          continue;
        src[pos] = static_cast<int>(i);
        code[pos] = cx + 3 * cy + 9 * cz;
        ++pos;
      }
}
__device__ __forceinline__ double imageShift(int digit, double L) { return digit == 1 ? L : (digit == 2 ? -L : 0.); }

struct ImageAppendArgs {
  int64_t count, firstSlot;
  const int *src, *code;
  int *dst;
  double L[3];
  double *x, *y, *z;
  int nOther;
  double *other[APB_NUM_COLUMNS];  // active columns other than x, y, z: zero-filled like kUnpackAppend does
  int64_t *id;
  int32_t *type, *own;
};
__global__ void kImageAppend(ImageAppendArgs a) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= a.count) return;
  const int s = a.src[q], c = a.code[q];
  const int64_t t = a.firstSlot + q;
  a.x[t] = a.x[s] + imageShift(c % 3, a.L[0]);
  a.y[t] = a.y[s] + imageShift((c / 3) % 3, a.L[1]);
  a.z[t] = a.z[s] + imageShift(c / 9, a.L[2]);
  for (int k = 0; k < a.nOther; ++k) a.other[k][t] = 0.;
  a.id[t] = a.id[s];
  a.type[t] = a.type[s];
  a.own[t] = APB_OWN_HALO;
  a.dst[q] = static_cast<int>(t);
}
// refresh of the images (non-rebuild steps): a source or image slot that the rebuild dropped is skipped
__global__ void kImageRefresh(int64_t m, const int *__restrict__ src, const int *__restrict__ dst, const int *__restrict__ code,
                              double Lx, double Ly, double Lz, double *x, double *y, double *z) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int s = src[q], t = dst[q], c = code[q];
  if (s < 0 || t < 0) return;
  x[t] = x[s] + imageShift(c % 3, Lx);
  y[t] = y[s] + imageShift((c / 3) % 3, Ly);
  z[t] = z[s] + imageShift(c / 9, Lz);
}

static bool allDimsSelf(apb_handle h) {
  for (int d = 0; d < 3; ++d)
    if (!h->periodic[d] || h->neighbor[d][0] != h->myRank || h->neighbor[d][1] != h->myRank ||
        !nearRel(h->cfg.box_min[d], h->globalMin[d]) || !nearRel(h->cfg.box_max[d], h->globalMax[d]))
      return false;
  return true;
}

static int generateImagesAllSelf(apb_handle h) {
  const int64_t n = h->nslots;
  h->haloAllN = 0;
  h->haloAllMode = true;
  if (n == 0) return APB_OK;
  ImageArgs a;
  a.n = n;
  a.own = h->own;
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = h->cfg.box_min[d];
    a.hi[d] = h->cfg.box_max[d];
  }
  a.il = h->cfg.cutoff + h->cfg.skin;
  const int numBlocks = apbDivUp(n, SEL_BLOCK);
  APB_CHECK(apbEnsure(h, h->key, sizeof(int) * (numBlocks + 1)));
  APB_CHECK(apbEnsure(h, h->rank, sizeof(int) * (numBlocks + 1)));
  int *blockCounts = static_cast<int *>(h->key.p), *blockOffsets = static_cast<int *>(h->rank.p);
  long long *totals = reinterpret_cast<long long *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result));
  ++h->launchCount, kImageCount<<<numBlocks, SEL_BLOCK, 0, h->stream>>>(a, blockCounts);
  APB_CUDA(cudaGetLastError());
  APB_CUDA(cudaMemsetAsync(blockCounts + numBlocks, 0, sizeof(int), h->stream));
  APB_CHECK(apbExclusiveScan(h, blockCounts, blockOffsets, numBlocks + 1, totals));
  long long total = 0;
  APB_CUDA(cudaMemcpyAsync(&total, totals, 8, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (total == 0) return APB_OK;
  if (n + total > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "halo generation exceeds 2^31 slots");
  APB_CHECK(apbEnsure(h, h->haloAllSrc, sizeof(int) * total));
  APB_CHECK(apbEnsure(h, h->haloAllDst, sizeof(int) * total));
  APB_CHECK(apbEnsure(h, h->haloAllCode, sizeof(int) * total));
  APB_CHECK(apbReserveSlots(h, n + total));
  a.own = h->own;  // the reservation may have moved the columns
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  ++h->launchCount, kImageWrite<<<numBlocks, SEL_BLOCK, 0, h->stream>>>(a, blockOffsets, static_cast<int *>(h->haloAllSrc.p),
                                                                      static_cast<int *>(h->haloAllCode.p));
  ImageAppendArgs u;
  u.count = total;
  u.firstSlot = n;
  u.src = static_cast<const int *>(h->haloAllSrc.p);
  u.code = static_cast<const int *>(h->haloAllCode.p);
  u.dst = static_cast<int *>(h->haloAllDst.p);
  for (int d = 0; d < 3; ++d) u.L[d] = h->globalMax[d] - h->globalMin[d];
  u.x = h->col[APB_COL_X];
  u.y = h->col[APB_COL_Y];
  u.z = h->col[APB_COL_Z];
  u.nOther = 0;
  for (int c = 0; c < APB_NUM_COLUMNS; ++c)
    if (h->active[c] && c != APB_COL_X && c != APB_COL_Y && c != APB_COL_Z) u.other[u.nOther++] = h->col[c];
  u.id = h->id;
  u.type = h->type;
  u.own = h->own;
  ++h->launchCount, kImageAppend<<<apbDivUp(total, 256), 256, 0, h->stream>>>(u);
  APB_CUDA(cudaGetLastError());
  h->nslots = n + total;
  h->haloAllN = total;
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

static int ensureDecomposition(apb_handle h) {
  if (h->decompositionSet) return APB_OK;
  if (h->nranks != 1) return h->fail(APB_ERR_STATE, "apb_set_decomposition must be called on multi-rank runs");
  for (int d = 0; d < 3; ++d) {  // single rank: fully periodic box that is its own neighbour
    h->globalMin[d] = h->cfg.box_min[d];
    h->globalMax[d] = h->cfg.box_max[d];
    h->periodic[d] = true;
    h->neighbor[d][0] = h->neighbor[d][1] = 0;
  }
  h->decompositionSet = true;
  return APB_OK;
}

// RegularGridDecomposition::exchangeMigratingParticles (:238-301) fused with the container update that precedes it in
// the simulation loop (Simulation.cpp:247-263): halos and dummies are dropped, owned particles that left the local box
// travel to the neighbour (wrapped at periodic global boundaries), arrivals are appended as owned.
// Migration in a dimension in which this rank is its own neighbour (periodic, one rank wide): the leaver re-enters
// through the opposite face, i.e. its coordinate is wrapped in place - same arithmetic as kPack's shift + clamp
// (RegularGridDecomposition.cpp:575-590), no selection, no compaction, no append.
__global__ void kWrapSelf(int64_t n, const int32_t *__restrict__ own, double *x, double *y, double *z, int dimMask,
                          double lx, double ly, double lz, double hx, double hy, double hz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  double *col[3] = {x, y, z};
  const double lo[3] = {lx, ly, lz}, hi[3] = {hx, hy, hz};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (!((dimMask >> d) & 1)) continue;
    const double p = col[d][i];
    const double L = hi[d] - lo[d];
    if (p < lo[d] || p >= hi[d]) {
      double v = p < lo[d] ? p + L : p - L;
      if (v >= hi[d]) v = nextafter(hi[d], lo[d]);
      if (v < lo[d]) v = lo[d];
      col[d][i] = v;
    }
  }
}

// ---- refresh of arbitrary columns of the halo copies ----------------------------------------------------------------
// sph-mpi refreshes its halo particles between the density and the hydro-force pass (examples/sph-mpi/sph-main-mpi.cpp:
// 373-414: updateHaloParticles -> density -> setPressure -> updateHaloParticles -> hydro force): the copies need the
// owners' density and pressure. Here the halo copies recorded by the last generating apb_exchange_halos receive the
// current values of the given columns from their source particles, dimension by dimension in the order of the
// generating exchange (a copy of a copy - edge and corner halos - is fed by the already refreshed copy).
struct ColumnSet {
  int n;
  double *p[APB_NUM_COLUMNS];
};
// one launch for both directions of a dimension (or for all images of a rank that is its own neighbour everywhere)
__global__ void kCopyColumns(int64_t mA, int64_t mB, const int *__restrict__ srcA, const int *__restrict__ dstA,
                             const int *__restrict__ srcB, const int *__restrict__ dstB, ColumnSet cs) {
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= mA + mB) return;
  const bool second = q >= mA;
  if (second) q -= mA;
  const int s = second ? srcB[q] : srcA[q], t = second ? dstB[q] : dstA[q];
  if (s < 0 || t < 0) return;  // dropped by the rebuild
  for (int c = 0; c < cs.n; ++c) cs.p[c][t] = cs.p[c][s];
}
__global__ void kGatherColumns2(int64_t m0, int64_t m1, const int *__restrict__ idx0, const int *__restrict__ idx1,
                                ColumnSet cs, double *out0, double *out1) {
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m0 + m1) return;
  const bool second = q >= m0;
  if (second) q -= m0;
  const int64_t m = second ? m1 : m0;
  const int s = second ? idx1[q] : idx0[q];
  double *out = second ? out1 : out0;
  out[q] = s >= 0 ? 1. : 0.;  // validity of the source slot, then the columns
  for (int c = 0; c < cs.n; ++c) out[(c + 1) * m + q] = s >= 0 ? cs.p[c][s] : 0.;
}
__global__ void kScatterColumns2(int64_t m0, int64_t m1, const int *__restrict__ slot0, const int *__restrict__ slot1,
                                 const double *in0, const double *in1, ColumnSet cs) {
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= m0 + m1) return;
  const bool second = q >= m0;
  if (second) q -= m0;
  const int64_t m = second ? m1 : m0;
  const int t = second ? slot1[q] : slot0[q];
  const double *in = second ? in1 : in0;
  if (t < 0 || in[q] == 0.) return;
  for (int c = 0; c < cs.n; ++c) cs.p[c][t] = in[(c + 1) * m + q];
}

static int refreshHaloColumns(apb_handle h, const ColumnSet &cs) {
  if (cs.n == 0) return APB_OK;
  if (h->haloAllMode) {
    if (h->haloAllN > 0) {
      ++h->launchCount, kCopyColumns<<<apbDivUp(h->haloAllN, 256), 256, 0, h->stream>>>(
          h->haloAllN, 0, static_cast<const int *>(h->haloAllSrc.p), static_cast<const int *>(h->haloAllDst.p), nullptr,
          nullptr, cs);
      APB_CUDA(cudaGetLastError());
    }
    return APB_OK;
  }
  for (int d = 0; d < 3; ++d) {
    HaloLink &L0 = h->link[d][0], &L1 = h->link[d][1];
    const int left = h->neighbor[d][0], right = h->neighbor[d][1];
    if (h->nranks == 1 || (left == h->myRank && right == h->myRank)) {
      if (L1.nSend != L0.nRecv || L0.nSend != L1.nRecv)
        return h->fail(APB_ERR_STATE, "apb_refresh_halo_columns: inconsistent self-exchange links");
      const int64_t m = L1.nSend + L0.nSend;
      if (m > 0) {
        ++h->launchCount, kCopyColumns<<<apbDivUp(m, 256), 256, 0, h->stream>>>(
            L1.nSend, L0.nSend, static_cast<const int *>(L1.sendIdx.p), static_cast<const int *>(L0.recvSlot.p),
            static_cast<const int *>(L0.sendIdx.p), static_cast<const int *>(L1.recvSlot.p), cs);
        APB_CUDA(cudaGetLastError());
      }
      continue;
    }
    void *sendBuf[2], *recvBuf[2];
    size_t sb[2], rb[2];
    for (int s = 0; s < 2; ++s) {
      sb[s] = sizeof(double) * (cs.n + 1) * h->link[d][s].nSend;
      rb[s] = sizeof(double) * (cs.n + 1) * h->link[d][s].nRecv;
      APB_CHECK(apbEnsure(h, h->xbuf[s], sb[s] + 256));
      APB_CHECK(apbEnsure(h, h->xbuf[2 + s], rb[s] + 256));
      sendBuf[s] = h->xbuf[s].p;
      recvBuf[s] = h->xbuf[2 + s].p;
    }
    if (L0.nSend + L1.nSend > 0) {
      ++h->launchCount, kGatherColumns2<<<apbDivUp(L0.nSend + L1.nSend, 256), 256, 0, h->stream>>>(
          L0.nSend, L1.nSend, static_cast<const int *>(L0.sendIdx.p), static_cast<const int *>(L1.sendIdx.p), cs,
          static_cast<double *>(sendBuf[0]), static_cast<double *>(sendBuf[1]));
      APB_CUDA(cudaGetLastError());
    }
    APB_CHECK(exchangePayload(h, d, sendBuf, sb, recvBuf, rb));
    if (L0.nRecv + L1.nRecv > 0) {
      ++h->launchCount, kScatterColumns2<<<apbDivUp(L0.nRecv + L1.nRecv, 256), 256, 0, h->stream>>>(
          L0.nRecv, L1.nRecv, static_cast<const int *>(L0.recvSlot.p), static_cast<const int *>(L1.recvSlot.p),
          static_cast<const double *>(recvBuf[0]), static_cast<const double *>(recvBuf[1]), cs);
      APB_CUDA(cudaGetLastError());
    }
  }
  return APB_OK;
}

extern "C" int apb_migrate(apb_handle h, int64_t *out_num_sent, int64_t *out_num_received) {
  APB_ENTRY(h);
  APB_CHECK(ensureDecomposition(h));
  APB_CHECK(apb_delete_halo_particles(h));
  h->haloLinksValid = false;
  const int64_t before = h->nslots;
  int64_t received = 0, sent = 0;
  int selfMask = 0;
  for (int d = 0; d < 3; ++d) {
    const bool self = h->periodic[d] && h->neighbor[d][0] == h->myRank && h->neighbor[d][1] == h->myRank &&
                      nearRel(h->cfg.box_min[d], h->globalMin[d]) && nearRel(h->cfg.box_max[d], h->globalMax[d]);
    if (self) {
      selfMask |= 1 << d;
      continue;
    }
    const int64_t n0 = h->nslots;
    APB_CHECK(exchangeDim(h, d, 1, &sent));
    received += h->nslots - n0;
  }
  // self dimensions last: arrivals of the exchanged dimensions are wrapped as well (the dimensions are independent)
  if (selfMask && h->nslots > 0) {
    ++h->launchCount, kWrapSelf<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
        h->nslots, h->own, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], selfMask, h->globalMin[0],
        h->globalMin[1], h->globalMin[2], h->globalMax[0], h->globalMax[1], h->globalMax[2]);
    APB_CUDA(cudaGetLastError());
    if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  }
  (void)before;
  // Sent particles and the old halos are dummies now. They are not compacted here: the rebuild that must follow drops
  // dummies while it sorts (VerletClusterLists.h:362-397 / LinkedCells.h:152-202 likewise delete dummies on rebuild).
  if (out_num_received) *out_num_received = received;
  if (out_num_sent) *out_num_sent = sent;  // particles this rank handed to its neighbours (all exchanged dimensions)
  h->structureValid = false;
  h->prunedValid = false;
  h->countsValid = false;
  h->ownedInsideBox = true;
  return APB_OK;
}

// RegularGridDecomposition::exchangeHaloParticles (:159-236). If the container structure is invalid (rebuild step)
// halos are selected and appended; otherwise only the positions of the existing halo copies are refreshed through the
// recorded send / receive slots (bulk ParticleContainerInterface::updateHaloParticle, no re-selection).
extern "C" int apb_exchange_halos(apb_handle h) {
  APB_ENTRY(h);
  APB_CHECK(ensureDecomposition(h));
  if (!h->structureValid) {
    APB_CHECK(apb_delete_halo_particles(h));
    for (int d = 0; d < 3; ++d)
      for (int s = 0; s < 2; ++s) h->link[d][s].nSend = h->link[d][s].nRecv = 0;
    h->haloAllMode = false;
    h->haloAllN = 0;
    h->noHalos = false;  // halo copies are appended from here on: a round that fails midway must not hide them
    static const bool noOnePass = getenv("APB_NO_ONEPASS_HALO") != nullptr;
    h->countsValid = false;
    h->countsTrusted = false;
    if (h->ownedInsideBox && !noOnePass && allDimsSelf(h)) {
      APB_CHECK(generateImagesAllSelf(h));
      if (h->ownedKnown) {  // owned count unchanged since it was counted, halo count = number of images just written
        h->numOwned = h->ownedCount;
        h->numHalo = h->haloAllN;
        h->countsValid = h->countsTrusted = true;
      }
    } else {
      for (int d = 0; d < 3; ++d) APB_CHECK(exchangeDim(h, d, 0));
    }
    h->haloLinksValid = true;
    h->noHalos = false;
    if (h->cfg.particle_kind != APB_PARTICLE_LJ) {
      // halo copies of SPH / multi-site particles need their attributes (mass, smoothing length, quaternion ...): the
      // generating exchange carries positions, ids and types; the other active columns follow through the recorded links
      ColumnSet cs;
      cs.n = 0;
      for (int c = APB_COL_VX; c < APB_NUM_COLUMNS; ++c)
        if (h->active[c]) cs.p[cs.n++] = h->col[c];
      APB_CHECK(refreshHaloColumns(h, cs));
      if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return APB_OK;
  }
  if (!h->haloLinksValid) return h->fail(APB_ERR_STATE, "apb_exchange_halos: no halo links recorded; call it once before the rebuild");
  if (h->haloAllMode) {
    if (h->haloAllN > 0) {
      ++h->launchCount, kImageRefresh<<<apbDivUp(h->haloAllN, 256), 256, 0, h->stream>>>(
          h->haloAllN, static_cast<const int *>(h->haloAllSrc.p), static_cast<const int *>(h->haloAllDst.p),
          static_cast<const int *>(h->haloAllCode.p), h->globalMax[0] - h->globalMin[0], h->globalMax[1] - h->globalMin[1],
          h->globalMax[2] - h->globalMin[2], h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z]);
      APB_CUDA(cudaGetLastError());
    }
    if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
    return APB_OK;
  }
  APB_CHECK(ensureP2P(h));
  ++h->p2pSeq;
  for (int d = 0; d < 3; ++d) {
    HaloLink &L0 = h->link[d][0], &L1 = h->link[d][1];
    const int left = h->neighbor[d][0], right = h->neighbor[d][1];
    if (h->nranks == 1 || (left == h->myRank && right == h->myRank)) {
      // my right-going particles (link 1) arrive in the slots recorded for "from the left" (link 0) and vice versa
      if (L1.nSend != L0.nRecv || L0.nSend != L1.nRecv)
        return h->fail(APB_ERR_STATE, "apb_exchange_halos: inconsistent self-exchange links");
      const int64_t m = L1.nSend + L0.nSend;
      if (m > 0) {
        ++h->launchCount, kRefreshSelf<<<apbDivUp(m, 256), 256, 0, h->stream>>>(
            L1.nSend, L0.nSend, static_cast<const int *>(L1.sendIdx.p), static_cast<const int *>(L0.recvSlot.p),
            static_cast<const int *>(L0.sendIdx.p), static_cast<const int *>(L1.recvSlot.p), L1.shift, L0.shift, d,
            h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z]);
        APB_CUDA(cudaGetLastError());
      }
      continue;
    }
    // (a non-periodic dimension has ranks without one of the neighbours: all ranks keep NCCL for it)
    if (h->p2pState == 1 && h->periodic[d] && h->p2pPeer[d][0] && h->p2pPeer[d][1]) {
      // my left-going particles land in the left neighbour's "from the right" region and vice versa
      if (3 * static_cast<size_t>(std::max(std::max(L0.nSend, L1.nSend), std::max(L0.nRecv, L1.nRecv))) > h->p2pCap)
        return h->fail(APB_ERR_NOT_APPLICABLE, "halo refresh: more halo particles per face than the peer-memory arena "
                                               "holds; raise APB_HALO_ARENA_MB (default 24) on all ranks");
      const int parity = static_cast<int>(h->p2pSeq & 1);
      // Which of the neighbour's two receive regions a send list feeds follows the matching of the generating exchange
      // (exchangePayload): with distinct neighbours my left-going particles are what the left neighbour receives "from
      // its right" (side 1); when both neighbours are the same rank (two ranks in this dimension) NCCL pairs the sends
      // and receives of the group in order, so send list 0 feeds its side 0 and send list 1 its side 1.
      const int to0 = left == right ? 0 : 1, to1 = left == right ? 1 : 0;
      const int64_t mS = L0.nSend + L1.nSend, mR = L0.nRecv + L1.nRecv;
      ++h->launchCount, kPushHalo<<<std::max<int64_t>(1, apbDivUp(mS, 256)), 256, 0, h->stream>>>(
          L0.nSend, L1.nSend, static_cast<const int *>(L0.sendIdx.p), static_cast<const int *>(L1.sendIdx.p),
          h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], d, L0.shift, L1.shift,
          p2pRegion(h->p2pPeer[d][0], h->p2pCap, d, to0, parity), p2pRegion(h->p2pPeer[d][1], h->p2pCap, d, to1, parity),
          p2pFlag(h->p2pPeer[d][0], h->p2pCap, d, to0), p2pFlag(h->p2pPeer[d][1], h->p2pCap, d, to1), h->p2pSeq,
          h->p2pCounters + d);
      ++h->launchCount, kPullHalo<<<std::max<int64_t>(1, apbDivUp(mR, 256)), 256, 0, h->stream>>>(
          L0.nRecv, L1.nRecv, static_cast<const int *>(L0.recvSlot.p), static_cast<const int *>(L1.recvSlot.p),
          p2pRegion(h->p2pArena, h->p2pCap, d, 0, parity), p2pRegion(h->p2pArena, h->p2pCap, d, 1, parity),
          p2pFlag(h->p2pArena, h->p2pCap, d, 0), p2pFlag(h->p2pArena, h->p2pCap, d, 1), h->p2pSeq, h->col[APB_COL_X],
          h->col[APB_COL_Y], h->col[APB_COL_Z], h->p2pCounters + 3);
      APB_CUDA(cudaGetLastError());
      continue;
    }
    void *sendBuf[2], *recvBuf[2];
    size_t sb[2], rb[2];
    for (int s = 0; s < 2; ++s) {
      sb[s] = sizeof(double) * 3 * h->link[d][s].nSend;
      rb[s] = sizeof(double) * 3 * h->link[d][s].nRecv;
      APB_CHECK(apbEnsure(h, h->xbuf[s], sb[s] + 256));
      APB_CHECK(apbEnsure(h, h->xbuf[2 + s], rb[s] + 256));
      sendBuf[s] = h->xbuf[s].p;
      recvBuf[s] = h->xbuf[2 + s].p;
    }
    if (L0.nSend + L1.nSend > 0) {
      ++h->launchCount, kGatherPositions2<<<apbDivUp(L0.nSend + L1.nSend, 256), 256, 0, h->stream>>>(
          L0.nSend, L1.nSend, static_cast<const int *>(L0.sendIdx.p), static_cast<const int *>(L1.sendIdx.p),
          h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], d, L0.shift, L1.shift,
          static_cast<double *>(sendBuf[0]), static_cast<double *>(sendBuf[1]));
      APB_CUDA(cudaGetLastError());
    }
    APB_CHECK(exchangePayload(h, d, sendBuf, sb, recvBuf, rb));
    if (L0.nRecv + L1.nRecv > 0) {
      ++h->launchCount, kScatterPositions2<<<apbDivUp(L0.nRecv + L1.nRecv, 256), 256, 0, h->stream>>>(
          L0.nRecv, L1.nRecv, static_cast<const int *>(L0.recvSlot.p), static_cast<const int *>(L1.recvSlot.p),
          static_cast<const double *>(recvBuf[0]), static_cast<const double *>(recvBuf[1]), h->col[APB_COL_X],
          h->col[APB_COL_Y], h->col[APB_COL_Z]);
      APB_CUDA(cudaGetLastError());
    }
  }
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_refresh_halo_columns(apb_handle h, int32_t numColumns, const int32_t *columns) {
  APB_ENTRY(h);
  APB_CHECK(ensureDecomposition(h));
  if (numColumns < 0 || numColumns > APB_NUM_COLUMNS || (numColumns > 0 && !columns))
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_refresh_halo_columns: bad column list");
  if (!h->haloLinksValid)
    return h->fail(APB_ERR_STATE, "apb_refresh_halo_columns: no halo links recorded; call apb_exchange_halos before the rebuild");
  ColumnSet cs;
  cs.n = 0;
  for (int k = 0; k < numColumns; ++k) {
    const int c = columns[k];
    if (c < 0 || c >= APB_NUM_COLUMNS || !h->active[c])
      return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_refresh_halo_columns: column not part of this particle kind");
    if (c <= APB_COL_Z)
      return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_refresh_halo_columns: positions are refreshed by apb_exchange_halos "
                                               "(they are shifted at the periodic boundary)");
    cs.p[cs.n++] = h->col[c];
  }
  APB_CHECK(refreshHaloColumns(h, cs));
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// the simulation loop, device resident
// ------------------------------------------------------------------------------------------------------------------
int apbCheckTraversal(apb_handle h, int traversal, int newton3);

// Simulation::simulate (examples/md-flexible/src/Simulation.cpp:230-351) for the built-in functors, without leaving the
// device: positions -> (every rebuild_frequency steps: migrate, halo exchange, rebuild | else: halo refresh) -> forces
// -> velocities. Work is enqueued asynchronously; the host only blocks on rebuild steps (it needs sizes) and at the end.
// velocities of the finished step, thermostat (Simulation.cpp:313, 539-546: every thermostatInterval iterations, after
// the velocity update), positions of the next step. Without a thermostat action the two integration halves run fused.
static int finishStep(apb_handle h, const apb_loop_params *p, int s, int numSteps, int64_t it) {
  const bool thermo = h->thermostatOn && it % h->thermostatInterval == 0;
  if (!thermo)
    return s + 1 < numSteps ? integrateVelocitiesPositions(h, p->dt, p->mass_of_type, p->num_types, p->global_force,
                                                           !h->forceOverwrite)
                            : apb_integrate_velocities(h, p->dt, p->mass_of_type, p->num_types);
  APB_CHECK(apb_integrate_velocities(h, p->dt, p->mass_of_type, p->num_types));
  APB_CHECK(apb_apply_thermostat(h, p->mass_of_type, p->num_types, h->thermostatTarget, h->thermostatDelta));
  if (s + 1 < numSteps) APB_CHECK(apb_integrate_positions(h, p->dt, p->mass_of_type, p->num_types, p->global_force));
  return APB_OK;
}

extern "C" int apb_run_steps(apb_handle h, const apb_functor *functor, const apb_loop_params *p, int32_t numSteps,
                             int64_t firstIteration, apb_traversal_result *outPerStep) {
  APB_ENTRY(h);
  if (!functor || !p || numSteps < 0) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_run_steps: bad argument");
  if (p->rebuild_frequency < 1) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_run_steps: rebuild_frequency must be >= 1");
  APB_CHECK(apbCheckTraversal(h, p->traversal, p->newton3));
  if (numSteps == 0) return APB_OK;
  APB_CHECK(apbEnsure(h, h->loopResults, sizeof(apb_traversal_result) * numSteps));
  apb_traversal_result *dres = static_cast<apb_traversal_result *>(h->loopResults.p);
  int rc = APB_OK;
  h->deferSync = true;
  // Inside this loop every force evaluation is preceded by a reset of the force column to the global force. With
  // gpuvcl_pruned (newton3 off: one writer per owned particle, every owned particle written) the kernel stores
  // global force + pair sum instead and the fused integrator skips the reset.
  static const bool noOverwrite = getenv("APB_NO_FORCE_OVERWRITE") != nullptr;
  h->forceOverwrite = !noOverwrite && p->traversal == APB_TRAVERSAL_GPUVCL_PRUNED && !p->newton3 &&
                      functor->kind == APB_FUNCTOR_LJ && h->cfg.particle_kind == APB_PARTICLE_LJ;
  for (int d = 0; d < 3; ++d) h->forceG[d] = p->global_force ? p->global_force[d] : 0.;
  // APB_DEBUG_TIMING: host wall time spent inside each phase call (where the host blocks on count read-backs / peers)
  static const bool dbgTiming = getenv("APB_DEBUG_TIMING") != nullptr;
  // Overlapped halo refresh (interior / boundary split of the force step). The two partial force launches cost about
  // 0.02 ms of tail effects; measured on B200s the split pays from two exchanging dimensions on (4 GPUs: 7.4 -> 7.8,
  // 8 GPUs: 13.5 -> 14.3 GFUPs/s), at 2 GPUs (one remote dimension) it is a wash. APB_SPLIT_STEP=0 / 1 overrides.
  // (A single resident kernel for all three refresh rounds was tried as well: no gain over the per-dimension kernels.)
  int remoteDims = 0;
  for (int d = 0; d < 3; ++d) remoteDims += h->neighbor[d][0] != h->myRank || h->neighbor[d][1] != h->myRank;
  const char *splitEnv = getenv("APB_SPLIT_STEP");
  const bool noSplit = splitEnv ? atoi(splitEnv) == 0 : remoteDims < 2;
  double hostMs[5] = {0, 0, 0, 0, 0};
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int s = 0; s < numSteps && rc == APB_OK; ++s) {
    const int64_t it = firstIteration + s;
    if (s == 0) {  // later steps: positions were advanced together with the previous step's velocities
      apbLoopTimingRecord(h, 3, true);
      rc = apb_integrate_positions(h, p->dt, p->mass_of_type, p->num_types, p->global_force);
      apbLoopTimingRecord(h, 3, false);
      if (rc != APB_OK) break;
    }
    // Rebuild cadence (LogicHandler::neighborListsAreValid, LogicHandler.h:987-997): by iteration number, or - with the
    // dynamic trigger on - when rebuild_frequency steps have passed since the last rebuild or a particle has moved skin / 2
    // (one 4-byte read-back per step: the host has to know before it enqueues the rest of the step).
    bool rebuild = (h->dynamicRebuild ? h->stepsSinceRebuild >= p->rebuild_frequency : it % p->rebuild_frequency == 0) ||
                   !h->structureValid;
    if (!rebuild && h->dynamicRebuild) {
      int32_t needed = 0;
      rc = apb_check_dynamic_rebuild(h, &needed);
      if (rc != APB_OK) break;
      rebuild = needed != 0;
      if (rebuild) ++h->dynamicRebuildCount;
    }
    if (rebuild) h->stepsSinceRebuild = 0;
    ++h->stepsSinceRebuild;
    if (rebuild) {
      apbLoopTimingRecord(h, 1, true);
      const double t0 = now();
      rc = apb_migrate(h, nullptr, nullptr);
      const double t1 = now();
      if (rc == APB_OK) rc = apb_exchange_halos(h);
      const double t2 = now();
      if (rc == APB_OK) rc = apb_rebuild_neighbor_lists(h, p->traversal, p->newton3);
      const double t3 = now();
      hostMs[0] += t1 - t0, hostMs[1] += t2 - t1, hostMs[2] += t3 - t2;
      apbLoopTimingRecord(h, 1, false);
    } else if (h->nranks > 1 && p->traversal == APB_TRAVERSAL_GPUVCL_PRUNED && functor->kind == APB_FUNCTOR_LJ &&
               h->prunedValid && !noSplit) {
      // Split step: the halo refresh (NCCL over NVLink) runs on the second stream while the interior tiles - those
      // that stage no halo copy - are evaluated; the boundary tiles follow once the refreshed positions have landed.
      if (!h->evSplit[0]) {
        APB_CUDA(cudaEventCreateWithFlags(&h->evSplit[0], cudaEventDisableTiming));
        APB_CUDA(cudaEventCreateWithFlags(&h->evSplit[1], cudaEventDisableTiming));
      }
      const double t0 = now();
      APB_CUDA(cudaEventRecord(h->evSplit[0], h->stream));
      APB_CUDA(cudaStreamWaitEvent(h->stream2, h->evSplit[0], 0));
      cudaStream_t mainStream = h->stream;
      h->stream = h->stream2;
      apbLoopTimingRecord(h, 2, true);
      rc = apb_exchange_halos(h);
      apbLoopTimingRecord(h, 2, false);
      cudaEventRecord(h->evSplit[1], h->stream2);
      h->stream = mainStream;
      hostMs[3] += now() - t0;
      if (rc != APB_OK) break;
      h->asyncResultDev = dres + s;
      apbLoopTimingRecord(h, 0, true);
      h->prunedPart = 1;
      rc = apb_compute_interactions(h, p->traversal, functor, p->newton3, nullptr);
      if (rc == APB_OK) {
        APB_CUDA(cudaStreamWaitEvent(h->stream, h->evSplit[1], 0));
        h->prunedPart = 2;
        rc = apb_compute_interactions(h, p->traversal, functor, p->newton3, nullptr);
      }
      h->prunedPart = 0;
      apbLoopTimingRecord(h, 0, false);
      h->asyncResultDev = nullptr;
      if (rc != APB_OK) break;
      apbLoopTimingRecord(h, 3, true);
      rc = finishStep(h, p, s, numSteps, it);
      apbLoopTimingRecord(h, 3, false);
      continue;
    } else {
      apbLoopTimingRecord(h, 2, true);
      const double t0 = now();
      rc = apb_exchange_halos(h);
      hostMs[3] += now() - t0;
      apbLoopTimingRecord(h, 2, false);
    }
    if (rc != APB_OK) break;
    h->asyncResultDev = dres + s;
    apbLoopTimingRecord(h, 0, true);
    rc = apb_compute_interactions(h, p->traversal, functor, p->newton3, nullptr);
    apbLoopTimingRecord(h, 0, false);
    h->asyncResultDev = nullptr;
    if (rc != APB_OK) break;
    apbLoopTimingRecord(h, 3, true);
    rc = finishStep(h, p, s, numSteps, it);
    apbLoopTimingRecord(h, 3, false);
  }
  h->deferSync = false;
  h->asyncResultDev = nullptr;
  h->forceOverwrite = false;
  if (rc != APB_OK) return rc;
  if (dbgTiming)
    fprintf(stderr, "[apb] rank %d run_steps(%d): host ms in migrate %.3f, halo build %.3f, rebuild %.3f, halo refresh %.3f\n",
            h->myRank, numSteps, hostMs[0], hostMs[1], hostMs[2], hostMs[3]);
  if (outPerStep) {
    APB_CUDA(cudaMemcpyAsync(outPerStep, dres, sizeof(apb_traversal_result) * numSteps, cudaMemcpyDeviceToHost, h->stream));
  }
  int refreshTimedOut = 0;
  if (h->p2pState == 1)
    APB_CUDA(cudaMemcpyAsync(&refreshTimedOut, h->p2pCounters + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (refreshTimedOut) {
    h->poisoned = true;
    return h->fail(APB_ERR_STATE, "halo refresh: a neighbour rank did not publish its positions within ten seconds");
  }
  if (outPerStep)
    for (int s = 0; s < numSteps; ++s) apbMaskResultByFlags(outPerStep[s], functor->flags);
  return APB_OK;
}

// One AutoPas::computeInteractions call for a caller whose particle data lives in host memory, as one stream-ordered
// batch with a single host synchronisation at the end (rebuild steps additionally block where sizes are read back):
// positions by id host -> device, [rebuild != 0: migration, halo exchange, neighbour-structure rebuild | halo refresh],
// forces = 0, traversal, forces by id device -> host, accumulators.
extern "C" int apb_force_step_by_id(apb_handle h, int32_t traversal, const apb_functor *functor, int32_t newton3,
                                    int32_t rebuild, int64_t idBegin, int64_t numIds, const double *x, const double *y,
                                    const double *z, double *fx, double *fy, double *fz, apb_traversal_result *out) {
  APB_ENTRY(h);
  if (!functor) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_force_step_by_id: null functor");
  APB_CHECK(apbCheckTraversal(h, traversal, newton3));
  APB_CHECK(apbEnsure(h, h->loopResults, sizeof(apb_traversal_result)));
  apb_traversal_result *dres = static_cast<apb_traversal_result *>(h->loopResults.p);
  const bool lj = functor->kind == APB_FUNCTOR_LJ;
  apb_traversal_result host;
  std::memset(&host, 0, sizeof(host));
  h->deferSync = true;
  int rc = apb_upload_positions_by_id(h, idBegin, numIds, x, y, z);
  if (rc == APB_OK && (rebuild || !h->structureValid)) {
    rc = apb_migrate(h, nullptr, nullptr);
    if (rc == APB_OK) rc = apb_exchange_halos(h);
    if (rc == APB_OK) rc = apb_rebuild_neighbor_lists(h, traversal, newton3);
  } else if (rc == APB_OK) {
    rc = apb_exchange_halos(h);
  }
  if (rc == APB_OK) rc = apb_reset_forces(h, 0., 0., 0.);
  if (rc == APB_OK) {
    if (lj) h->asyncResultDev = dres;  // the LJ kernels leave their reduced accumulators on the device
    rc = apb_compute_interactions(h, traversal, functor, newton3, lj ? nullptr : &host);
    h->asyncResultDev = nullptr;
  }
  h->deferSync = false;
  if (rc != APB_OK) return rc;
  if (lj) APB_CUDA(cudaMemcpyAsync(&host, dres, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
  APB_CHECK(apb_download_forces_by_id(h, idBegin, numIds, fx, fy, fz));  // synchronises
  if (lj) apbMaskResultByFlags(host, functor->flags);
  if (out) *out = host;
  return APB_OK;
}

// the stream all work of this handle is enqueued on (for CUDA-event timing by the caller)
extern "C" int apb_get_stream(apb_handle h, void **outStream) {
  APB_ENTRY(h);
  if (!outStream) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_stream: null argument");
  *outStream = h->stream;
  return APB_OK;
}
