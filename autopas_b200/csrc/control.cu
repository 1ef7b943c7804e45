// Pieces of the simulation loop around the force step that the reference keeps in LogicHandler / md-flexible and that
// would otherwise pull the particle data back to the host every iteration:
//  * velocity-scaling thermostat                 examples/md-flexible/src/Thermostat.h:32-150, 228-275
//  * dynamic-rebuild trigger (rAtRebuild check)   src/autopas/LogicHandler.h:955-965, 1000-1016
//  * remainder traversal for buffered particles   src/autopas/remainder/RemainderPairwiseInteractionHandler.h:63-135
#include <algorithm>
#include <cmath>

#include "internal.cuh"
#include "lj_device.cuh"

#define CTRL_MAX_TYPES 32

int apbPrepareLJParams(apb_handle h, const apb_functor *f, LJParams &p);
int apbFinishStats(apb_handle h, int numBlocks, bool stats, const apb_functor *f, apb_traversal_result *out);
int apbAllReduce(apb_handle h, void *dev, int count, int isDouble, int isMax);  // dynamics.cu (NCCL), no-op for one rank

// ------------------------------------------------------------------------------------------------------------------
// thermostat
// ------------------------------------------------------------------------------------------------------------------
// Thermostat::calcTemperatureComponent (Thermostat.h:66-150): per particle type, sum of m v.v and particle count.
// Fixed-order reduction (per-block partials, one block sums them): the result is reproducible run to run.
// Deviation from the reference, on purpose: md-flexible iterates `autopas.begin()`, i.e. owned AND halo particles, so
// periodic images and - after the MPI reduction - the copies on neighbouring ranks are counted twice. Here every
// particle counts once (owned only); for a homogeneous system the two temperatures agree to O(surface / volume).
__global__ void __launch_bounds__(256) kKineticPerType(int64_t n, const int32_t *__restrict__ own, const int32_t *__restrict__ type,
                                                       const double *__restrict__ massOfType, int numTypes,
                                                       const double *__restrict__ vx, const double *__restrict__ vy,
                                                       const double *__restrict__ vz, double *__restrict__ partials) {
  __shared__ double sk[8], sc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = 0; t < numTypes; ++t) {
    double ke = 0., cnt = 0.;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      if (own[i] == APB_OWN_OWNED && type[i] == t) {
        ke += massOfType[t] * (vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]);
        cnt += 1.;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      ke += __shfl_xor_sync(0xffffffffu, ke, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
      sk[warp] = ke;
      sc[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0., b = 0.;
      for (int w = 0; w < 8; ++w) {
        a += sk[w];
        b += sc[w];
      }
      partials[(static_cast<size_t>(blockIdx.x) * numTypes + t) * 2] = a;
      partials[(static_cast<size_t>(blockIdx.x) * numTypes + t) * 2 + 1] = b;
    }
    __syncthreads();
  }
}

// sums[2 t] = sum m v.v, sums[2 t + 1] = count, over the blocks in ascending order
__global__ void kKineticFinish(int numBlocks, int numTypes, const double *__restrict__ partials, double *__restrict__ sums) {
  const int t = threadIdx.x;
  if (t >= numTypes) return;
  double a = 0., b = 0.;
  for (int k = 0; k < numBlocks; ++k) {
    a += partials[(static_cast<size_t>(k) * numTypes + t) * 2];
    b += partials[(static_cast<size_t>(k) * numTypes + t) * 2 + 1];
  }
  sums[2 * t] = a;
  sums[2 * t + 1] = b;
}

// Thermostat::apply (Thermostat.h:228-275): per type, the immediate target moves by at most |deltaTemperature| towards
// the target; velocities are scaled by sqrt(immediateTarget / current). Three translational degrees of freedom.
__global__ void kThermostatScale(int numTypes, const double *__restrict__ sums, double target, double delta,
                                 double *__restrict__ scale) {
  const int t = threadIdx.x;
  if (t >= numTypes) return;
  const double cnt = sums[2 * t + 1];
  double s = 1.;
  if (cnt > 0.) {
    const double current = sums[2 * t] / (cnt * 3.);
    const double ad = fabs(delta);
    const double immediate = current < target ? fmin(current + ad, target) : fmax(current - ad, target);
    if (current > 0.) s = sqrt(immediate / current);
  }
  scale[t] = s;
}

__global__ void kScaleVelocities(int64_t n, const int32_t *__restrict__ own, const int32_t *__restrict__ type, int numTypes,
                                 const double *__restrict__ scale, double *vx, double *vy, double *vz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || own[i] != APB_OWN_OWNED) return;
  const int t = type[i];
  const double s = scale[t < numTypes ? t : 0];
  vx[i] *= s;
  vy[i] *= s;
  vz[i] *= s;
}

// device buffer layout of h->thermoDev: [2 T] sums | [T] scale | [T] masses | per-block partials
static int kineticSums(apb_handle h, const double *massOfType, int numTypes, double **sumsOut, double **scaleOut) {
  if (numTypes <= 0 || numTypes > CTRL_MAX_TYPES || !massOfType)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "thermostat: 1 .. 32 particle types with their masses are needed");
  const int64_t n = h->nslots;
  const int numBlocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(apbDivUp(std::max<int64_t>(n, 1), 256), 592)));
  const size_t head = static_cast<size_t>(numTypes) * 4;
  APB_CHECK(apbEnsure(h, h->thermoDev, sizeof(double) * (head + static_cast<size_t>(numBlocks) * numTypes * 2)));
  double *base = static_cast<double *>(h->thermoDev.p);
  double *sums = base, *scale = base + 2 * numTypes, *mass = base + 3 * numTypes, *partials = base + head;
  APB_CUDA(cudaMemcpyAsync(mass, massOfType, sizeof(double) * numTypes, cudaMemcpyHostToDevice, h->stream));
  ++h->launchCount, kKineticPerType<<<numBlocks, 256, 0, h->stream>>>(n, h->own, h->type, mass, numTypes, h->col[APB_COL_VX],
                                                                     h->col[APB_COL_VY], h->col[APB_COL_VZ], partials);
  ++h->launchCount, kKineticFinish<<<1, CTRL_MAX_TYPES, 0, h->stream>>>(numBlocks, numTypes, partials, sums);
  APB_CUDA(cudaGetLastError());
  // Thermostat.h:124-131: MPI_Allreduce(SUM) of both numbers per type
  APB_CHECK(apbAllReduce(h, sums, 2 * numTypes, 1, 0));
  *sumsOut = sums;
  *scaleOut = scale;
  return APB_OK;
}

extern "C" int apb_calc_temperature(apb_handle h, const double *massOfType, int32_t numTypes, double *outTemperature,
                                    int64_t *outCount) {
  APB_ENTRY(h);
  double *sums = nullptr, *scale = nullptr;
  APB_CHECK(kineticSums(h, massOfType, numTypes, &sums, &scale));
  double host[2 * CTRL_MAX_TYPES];
  APB_CUDA(cudaMemcpyAsync(host, sums, sizeof(double) * 2 * numTypes, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  for (int t = 0; t < numTypes; ++t) {
    if (outTemperature) outTemperature[t] = host[2 * t + 1] > 0. ? host[2 * t] / (host[2 * t + 1] * 3.) : 0.;
    if (outCount) outCount[t] = static_cast<int64_t>(host[2 * t + 1]);
  }
  return APB_OK;
}

extern "C" int apb_apply_thermostat(apb_handle h, const double *massOfType, int32_t numTypes, double targetTemperature,
                                    double deltaTemperature) {
  APB_ENTRY(h);
  double *sums = nullptr, *scale = nullptr;
  APB_CHECK(kineticSums(h, massOfType, numTypes, &sums, &scale));
  ++h->launchCount, kThermostatScale<<<1, CTRL_MAX_TYPES, 0, h->stream>>>(numTypes, sums, targetTemperature, deltaTemperature, scale);
  if (h->nslots > 0)
    ++h->launchCount, kScaleVelocities<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
        h->nslots, h->own, h->type, numTypes, scale, h->col[APB_COL_VX], h->col[APB_COL_VY], h->col[APB_COL_VZ]);
  APB_CUDA(cudaGetLastError());
  if (!h->deferSync) APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

extern "C" int apb_set_thermostat(apb_handle h, int32_t enable, int32_t interval, double targetTemperature,
                                  double deltaTemperature) {
  APB_ENTRY(h);
  if (enable && interval < 1) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_set_thermostat: interval must be >= 1");
  h->thermostatOn = enable != 0;
  h->thermostatInterval = interval;
  h->thermostatTarget = targetTemperature;
  h->thermostatDelta = deltaTemperature;
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// dynamic-rebuild trigger
// ------------------------------------------------------------------------------------------------------------------
// LogicHandler::checkNeighborListsInvalidDoDynamicRebuild (LogicHandler.h:1000-1016): the lists are invalid as soon as
// one owned particle of the container has moved skin / 2 or more from where it was at the last rebuild
// (ParticleBase::calculateDisplacementSinceRebuild, particles/ParticleBase.h:184-193; same >= comparison, dot product as
// (x*x + y*y) + z*z).
__global__ void kDisplacementCheck(int64_t n, const int32_t *__restrict__ own, const double *__restrict__ x,
                                   const double *__restrict__ y, const double *__restrict__ z,
                                   const double *__restrict__ rx, const double *__restrict__ ry,
                                   const double *__restrict__ rz, double halfSkinSquare, int *flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool moved = false;
  if (i < n && own[i] == APB_OWN_OWNED) {
    const double dx = rx[i] - x[i], dy = ry[i] - y[i], dz = rz[i] - z[i];
    moved = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) >= halfSkinSquare;
  }
  if (__any_sync(0xffffffffu, moved) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

extern "C" int apb_set_dynamic_rebuild(apb_handle h, int32_t enable) {
  APB_ENTRY(h);
  h->dynamicRebuild = enable != 0;
  h->rAtRebuildValid = false;
  return APB_OK;
}

extern "C" int apb_get_dynamic_rebuild_count(apb_handle h, int64_t *outCount) {
  APB_ENTRY(h);
  if (!outCount) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_dynamic_rebuild_count: null argument");
  *outCount = h->dynamicRebuildCount;
  return APB_OK;
}

// LogicHandler::updateRebuildPositions (LogicHandler.h:955-965): called by the rebuilds when the trigger is enabled
int apbSnapshotRebuildPositions(apb_handle h) {
  if (!h->dynamicRebuild) return APB_OK;
  const int64_t n = std::max<int64_t>(h->nslots, 1);
  APB_CHECK(apbEnsure(h, h->rAtRebuild, sizeof(double) * 3 * n));
  double *r = static_cast<double *>(h->rAtRebuild.p);
  for (int d = 0; d < 3; ++d)
    if (h->nslots > 0)
      APB_CUDA(cudaMemcpyAsync(r + d * n, h->col[APB_COL_X + d], sizeof(double) * h->nslots, cudaMemcpyDeviceToDevice, h->stream));
  h->rAtRebuildSlots = h->nslots;
  h->rAtRebuildStride = n;
  h->rAtRebuildValid = true;
  return APB_OK;
}

extern "C" int apb_check_dynamic_rebuild(apb_handle h, int32_t *outRebuildNeeded) {
  APB_ENTRY(h);
  if (!outRebuildNeeded) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_check_dynamic_rebuild: null argument");
  *outRebuildNeeded = 0;
  // without a snapshot that matches the storage order the lists are not valid anyway
  if (!h->structureValid || !h->rAtRebuildValid || h->rAtRebuildSlots != h->nslots) {
    *outRebuildNeeded = 1;
    return APB_OK;
  }
  int *flag = reinterpret_cast<int *>(static_cast<char *>(h->result.p) + sizeof(apb_traversal_result) + 200);
  APB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
  if (h->nslots > 0) {
    const double *r = static_cast<const double *>(h->rAtRebuild.p);
    const int64_t s = h->rAtRebuildStride;
    const double skin = h->cfg.skin;
    ++h->launchCount, kDisplacementCheck<<<apbDivUp(h->nslots, 256), 256, 0, h->stream>>>(
        h->nslots, h->own, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], r, r + s, r + 2 * s, skin * skin * 0.25, flag);
    APB_CUDA(cudaGetLastError());
  }
  // all ranks must rebuild in the same iteration (allreduce(max) of the trigger)
  APB_CHECK(apbAllReduce(h, flag, 1, 0, 1));
  int host = 0;
  APB_CUDA(cudaMemcpyAsync(&host, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  *outRebuildNeeded = host ? 1 : 0;
  return APB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// remainder traversal (LJ)
// ------------------------------------------------------------------------------------------------------------------
// RemainderPairwiseInteractionHandler::computeRemainderInteractions (:63-135): particles that LogicHandler keeps in its
// buffers until the next rebuild interact (1) with all container particles, (2) halo-buffer particles with the owned
// container particles, (3) buffer with buffer, (4) buffer with halo buffer - i.e. every pair with at least one buffered
// particle except halo-halo. Each pair is applied to both partners (LJFunctor::AoSFunctor with newton3,
// LJFunctor.h:123-198); globals carry the weight [i owned] + [j owned].
// Buffers are small (tens to a few thousand particles), so the container side is a plain sweep: one thread per container
// slot, the buffered particles staged in shared memory; the force on a buffered particle is reduced over the warp with
// shuffles and leaves through one RED per warp that found a partner.
struct RemArgs {
  int64_t n, nb;
  const double *x, *y, *z;
  double *fx, *fy, *fz;
  const int32_t *type, *own;
  const double *bx, *by, *bz;
  const int32_t *btype, *bown;
  double *bfx, *bfy, *bfz;
  LJParams p;
  LJStats *partials;
};

#define REM_TILE 128
template <bool MIX, bool STATS>
__global__ void __launch_bounds__(256) kLJRemainderContainer(RemArgs a) {
  __shared__ double sx[REM_TILE], sy[REM_TILE], sz[REM_TILE];
  __shared__ int st[REM_TILE], so[REM_TILE];
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int ownJ = j < a.n ? a.own[j] : APB_OWN_DUMMY;
  const double xj = ownJ ? a.x[j] : 0., yj = ownJ ? a.y[j] : 0., zj = ownJ ? a.z[j] : 0.;
  const int tj = (MIX && ownJ) ? a.type[j] : 0;
  double Fx = 0., Fy = 0., Fz = 0.;
  LJStats stt;
  ljStatsZero(stt);
  for (int64_t b0 = 0; b0 < a.nb; b0 += REM_TILE) {
    const int m = static_cast<int>(min(static_cast<int64_t>(REM_TILE), a.nb - b0));
    __syncthreads();
    if (threadIdx.x < m) {
      sx[threadIdx.x] = a.bx[b0 + threadIdx.x];
      sy[threadIdx.x] = a.by[b0 + threadIdx.x];
      sz[threadIdx.x] = a.bz[b0 + threadIdx.x];
      st[threadIdx.x] = a.btype ? a.btype[b0 + threadIdx.x] : 0;
      so[threadIdx.x] = a.bown[b0 + threadIdx.x];
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
      // i = buffered particle, j = container particle: dr = r_i - r_j (LJFunctor.h:146)
      const double drx = sx[k] - xj, dry = sy[k] - yj, drz = sz[k] - zj;
      const double dr2 = ljDist2(drx, dry, drz);
      const bool pair = ownJ != APB_OWN_DUMMY && !(ownJ == APB_OWN_HALO && so[k] == APB_OWN_HALO);
      const bool hit = pair && dr2 <= a.p.cutoff2;
      if (STATS && pair) ++stt.dist;
      if (!__any_sync(0xffffffffu, hit)) continue;
      double fxb = 0., fyb = 0., fzb = 0.;
      if (hit) {
        double upot6;
        const double fac = ljEval<MIX>(a.p, dr2, st[k], tj, upot6);
        fxb = drx * fac;
        fyb = dry * fac;
        fzb = drz * fac;
        Fx -= fxb;
        Fy -= fyb;
        Fz -= fzb;
        if (STATS) {
          const double w = (so[k] == APB_OWN_OWNED ? 1. : 0.) + (ownJ == APB_OWN_OWNED ? 1. : 0.);
          stt.upot += upot6 * w;
          stt.vir[0] += drx * fxb * w;
          stt.vir[1] += dry * fyb * w;
          stt.vir[2] += drz * fzb * w;
          ++stt.kN3;
          ++stt.gN3;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        fxb += __shfl_xor_sync(0xffffffffu, fxb, o);
        fyb += __shfl_xor_sync(0xffffffffu, fyb, o);
        fzb += __shfl_xor_sync(0xffffffffu, fzb, o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(a.bfx + b0 + k, fxb);
        atomicAdd(a.bfy + b0 + k, fyb);
        atomicAdd(a.bfz + b0 + k, fzb);
      }
    }
  }
  if (ownJ != APB_OWN_DUMMY) {
    a.fx[j] += Fx;
    a.fy[j] += Fy;
    a.fz[j] += Fz;
  }
  if (STATS) ljStatsBlockReduce(stt, a.partials);
}

// buffer with buffer: one thread per buffered particle i, all other buffered particles j; each side adds its own force
// and its own share of the globals, which sums to the newton3 weights
template <bool MIX, bool STATS>
__global__ void __launch_bounds__(128) kLJRemainderBuffers(RemArgs a, int partialOffset) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  LJStats stt;
  ljStatsZero(stt);
  if (i < a.nb && a.bown[i] != APB_OWN_DUMMY) {
    const double xi = a.bx[i], yi = a.by[i], zi = a.bz[i];
    const int ti = (MIX && a.btype) ? a.btype[i] : 0, oi = a.bown[i];
    double Fx = 0., Fy = 0., Fz = 0.;
    for (int64_t j = 0; j < a.nb; ++j) {
      const int oj = a.bown[j];
      if (j == i || oj == APB_OWN_DUMMY || (oi == APB_OWN_HALO && oj == APB_OWN_HALO)) continue;
      const double drx = xi - a.bx[j], dry = yi - a.by[j], drz = zi - a.bz[j];
      const double dr2 = ljDist2(drx, dry, drz);
      if (STATS && j > i) ++stt.dist;
      if (dr2 > a.p.cutoff2) continue;
      double upot6;
      const double fac = ljEval<MIX>(a.p, dr2, ti, (MIX && a.btype) ? a.btype[j] : 0, upot6);
      Fx += drx * fac;
      Fy += dry * fac;
      Fz += drz * fac;
      if (STATS) {
        if (oi == APB_OWN_OWNED) {
          stt.upot += upot6;
          stt.vir[0] += drx * drx * fac;
          stt.vir[1] += dry * dry * fac;
          stt.vir[2] += drz * drz * fac;
        }
        if (j > i) {
          ++stt.kN3;
          ++stt.gN3;
        }
      }
    }
    atomicAdd(a.bfx + i, Fx);
    atomicAdd(a.bfy + i, Fy);
    atomicAdd(a.bfz + i, Fz);
  }
  if (STATS) ljStatsBlockReduce(stt, a.partials, partialOffset + static_cast<int>(blockIdx.x));
}

extern "C" int apb_compute_remainder(apb_handle h, const apb_functor *f, int64_t nb, const double *x, const double *y,
                                     const double *z, const int32_t *types, const int32_t *ownership, double *fx,
                                     double *fy, double *fz, apb_traversal_result *out) {
  APB_ENTRY(h);
  if (!f || nb < 0 || (nb > 0 && (!x || !y || !z || !ownership || !fx || !fy || !fz)))
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_compute_remainder: bad argument");
  if (f->kind != APB_FUNCTOR_LJ || h->cfg.particle_kind != APB_PARTICLE_LJ)
    return h->fail(APB_ERR_NOT_APPLICABLE, "apb_compute_remainder: LJFunctor on MoleculeLJ particles only");
  if (out) std::memset(out, 0, sizeof(*out));
  if (nb == 0) return APB_OK;
  LJParams p;
  APB_CHECK(apbPrepareLJParams(h, f, p));
  const bool mix = f->flags & APB_FUNCTOR_USE_MIXING;
  const bool stats = f->flags & (APB_FUNCTOR_CALC_GLOBALS | APB_FUNCTOR_COUNT_FLOPS);
  // staging: 6 double columns + 2 int columns
  const size_t bytes = static_cast<size_t>(nb) * (6 * 8 + 2 * 4) + 64;
  APB_CHECK(apbEnsure(h, h->remBuf, bytes));
  double *d = static_cast<double *>(h->remBuf.p);
  int32_t *di = reinterpret_cast<int32_t *>(d + 6 * nb);
  APB_CUDA(cudaMemcpyAsync(d, x, 8 * nb, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(d + nb, y, 8 * nb, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(d + 2 * nb, z, 8 * nb, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemsetAsync(d + 3 * nb, 0, 3 * 8 * nb, h->stream));
  if (types) APB_CUDA(cudaMemcpyAsync(di, types, 4 * nb, cudaMemcpyHostToDevice, h->stream));
  APB_CUDA(cudaMemcpyAsync(di + nb, ownership, 4 * nb, cudaMemcpyHostToDevice, h->stream));
  RemArgs a;
  a.n = h->nslots;
  a.nb = nb;
  a.x = h->col[APB_COL_X];
  a.y = h->col[APB_COL_Y];
  a.z = h->col[APB_COL_Z];
  a.fx = h->col[APB_COL_FX];
  a.fy = h->col[APB_COL_FY];
  a.fz = h->col[APB_COL_FZ];
  a.type = h->type;
  a.own = h->own;
  a.bx = d;
  a.by = d + nb;
  a.bz = d + 2 * nb;
  a.bfx = d + 3 * nb;
  a.bfy = d + 4 * nb;
  a.bfz = d + 5 * nb;
  a.btype = types ? di : nullptr;
  a.bown = di + nb;
  a.p = p;
  const int gridC = h->nslots > 0 ? apbDivUp(h->nslots, 256) : 0, gridB = apbDivUp(nb, 128);
  APB_CHECK(apbEnsure(h, h->partials, sizeof(LJStats) * (gridC + gridB)));
  a.partials = static_cast<LJStats *>(h->partials.p);
#define REM_LAUNCH(MIXV, STATSV)                                                                              \
  do {                                                                                                        \
    if (gridC > 0) ++h->launchCount, kLJRemainderContainer<MIXV, STATSV><<<gridC, 256, 0, h->stream>>>(a);    \
    ++h->launchCount, kLJRemainderBuffers<MIXV, STATSV><<<gridB, 128, 0, h->stream>>>(a, gridC);              \
  } while (0)
  if (mix && stats) REM_LAUNCH(true, true);
  else if (mix) REM_LAUNCH(true, false);
  else if (stats) REM_LAUNCH(false, true);
  else REM_LAUNCH(false, false);
#undef REM_LAUNCH
  APB_CUDA(cudaGetLastError());
  APB_CUDA(cudaMemcpyAsync(fx, a.bfx, 8 * nb, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(fy, a.bfy, 8 * nb, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaMemcpyAsync(fz, a.bfz, 8 * nb, cudaMemcpyDeviceToHost, h->stream));
  return apbFinishStats(h, gridC + gridB, stats, f, out);  // synchronises
}
