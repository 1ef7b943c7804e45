/*
 * autopas_b200.h — C ABI of the B200-native short-range interaction path.
 *
 * This is the drop-in boundary: plain pointers and sizes, `int` status codes, no exceptions and no torch / CUDA types.
 * Each entry point names the reference interface it replaces (paths relative to the AutoPas reference tree).
 * The C++ shim `autopas_b200/shim/GpuContainers.h` implements `autopas::ParticleContainerInterface<P>` and
 * `autopas::TraversalInterface` on top of these calls; INTEGRATION.md shows the additive edits a maintainer makes.
 *
 * Conventions
 *  - All entry points return APB_OK (0) or a negative error code; `apb_last_error` gives the message.
 *  - Host buffers passed in are borrowed for the duration of the call only.
 *  - Compute / rebuild / update entry points are single-caller and BLOCK until the device work has finished, because
 *    the AutoTuner times them with host timers (src/autopas/LogicHandler.h:1083-1125).
 *  - "storage order" = the order of particle slots in the device SoA. It changes only in apb_rebuild_neighbor_lists /
 *    apb_update_container(keep=0) / apb_set_particles, mirroring iterator invalidation in the reference.
 *  - Ownership encoding is the reference's: dummy 0, owned 1, halo 2 (src/autopas/particles/OwnershipState.h:20-29).
 */
#ifndef AUTOPAS_B200_H
#define AUTOPAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct apb_handle_s *apb_handle;

enum apb_status {
  APB_OK = 0,
  APB_ERR_INVALID_ARGUMENT = -1,
  APB_ERR_CUDA = -2,              /* sticky: the handle is unusable afterwards */
  APB_ERR_NOT_APPLICABLE = -3,    /* configuration rejected, like TraversalInterface::isApplicableToDomain()==false */
  APB_ERR_STATE = -4,             /* e.g. compute before rebuild (VerletClusterLists.h:148-170 throws likewise) */
  APB_ERR_PARTICLE_OUTSIDE = -5,  /* add owned particle outside the box (LogicHandler.h:343-350) */
  APB_ERR_NCCL = -6,
  APB_ERR_OUT_OF_MEMORY = -7
};

/* options/ContainerOption.h:23-69 — new values gpuLinkedCells / gpuVerletClusterLists */
enum apb_container { APB_CONTAINER_LINKED_CELLS = 0, APB_CONTAINER_VERLET_CLUSTER_LISTS = 1 };

/* options/TraversalOption.h:25-190 — new values with unique prefixes gpulc_ / gpuvcl_
 * (containers/CompatibleTraversals.h:29-37 filters by name prefix). */
enum apb_traversal {
  APB_TRAVERSAL_GPULC_C08 = 0,                /* pair set of lc_c08 (LCC08Traversal.h:73-79) */
  APB_TRAVERSAL_GPULC_C18 = 1,                /* pair set of lc_c18 (LCC18Traversal.h:115-211); same set as c08 */
  APB_TRAVERSAL_GPUVCL_CLUSTER_ITERATION = 2, /* VCLClusterIterationTraversal.h:60-66, newton3 off only */
  APB_TRAVERSAL_GPUVCL_C06 = 3,               /* VCLC06Traversal.h:86-150, newton3 on/off */
  APB_TRAVERSAL_GPUVCL_C01_BALANCED = 4,      /* VCLC01BalancedTraversal.h:60-90, newton3 off only */
  APB_TRAVERSAL_GPUVCL_PRUNED = 5             /* B200-native: cluster-pair list refined to per-particle lists at
                                                 rebuild (rc+skin test); newton3 off (full lists, the fast mode) or on
                                                 (each owned-owned pair once, reaction scattered with RED.ADD.F64);
                                                 same forces / globals as the list-faithful traversals */
};

/* Which particle class the container stores (decides the SoA columns). */
enum apb_particle_kind {
  APB_PARTICLE_LJ = 0,        /* mdLib::MoleculeLJ (MoleculeLJ.h:39-70) */
  APB_PARTICLE_MULTISITE = 1, /* mdLib::MultisiteMoleculeLJ (MultisiteMoleculeLJ.h) */
  APB_PARTICLE_SPH = 2        /* sphLib::SPHParticle (SPHParticle.h:382-425) */
};

/* double-precision SoA columns addressable through apb_upload_column / apb_download_column */
enum apb_column {
  APB_COL_X = 0, APB_COL_Y, APB_COL_Z,
  APB_COL_VX, APB_COL_VY, APB_COL_VZ,
  APB_COL_FX, APB_COL_FY, APB_COL_FZ,          /* SPH: acceleration */
  APB_COL_OLDFX, APB_COL_OLDFY, APB_COL_OLDFZ,
  /* multisite */
  APB_COL_Q0, APB_COL_Q1, APB_COL_Q2, APB_COL_Q3,
  APB_COL_TX, APB_COL_TY, APB_COL_TZ,
  /* SPH */
  APB_COL_MASS, APB_COL_SMTH, APB_COL_DENSITY, APB_COL_PRESSURE, APB_COL_SNDSPEED,
  APB_COL_ENGDOT, APB_COL_VSIGMAX,
  APB_NUM_COLUMNS
};

typedef struct {
  double box_min[3];
  double box_max[3];
  double cutoff;
  double skin;
  double cell_size_factor; /* LinkedCells only (ContainerSelectorInfo::cellSizeFactor) */
  int32_t cluster_size;    /* VerletClusterLists only (LogicHandlerInfo.h: verletClusterSize, default 4) */
  int32_t container;       /* enum apb_container */
  int32_t particle_kind;   /* enum apb_particle_kind */
  int32_t device;          /* CUDA device ordinal */
} apb_config;

enum apb_functor_kind {
  APB_FUNCTOR_LJ = 0,          /* mdLib::LJFunctor (LJFunctor.h) */
  APB_FUNCTOR_LJ_MULTISITE = 1,/* mdLib::LJMultisiteFunctor */
  APB_FUNCTOR_ATM = 2,         /* mdLib::AxilrodTellerMutoFunctor (triwise) */
  APB_FUNCTOR_SPH_DENSITY = 3, /* sphLib::SPHCalcDensityFunctor */
  APB_FUNCTOR_SPH_HYDRO = 4    /* sphLib::SPHCalcHydroForceFunctor */
};

/* LJFunctor template flags (LJFunctor.h:39-41) become run-time bits */
enum apb_functor_flags {
  APB_FUNCTOR_APPLY_SHIFT = 1,
  APB_FUNCTOR_USE_MIXING = 2,
  APB_FUNCTOR_CALC_GLOBALS = 4,
  APB_FUNCTOR_COUNT_FLOPS = 8,
  /* LJFunctor exposes only the sum of the three virial components (getVirial, LJFunctor.h:719). With this bit the
   * kernels may return that sum in virial_sum[0] and zero in [1], [2]: only virial_sum[0]+[1]+[2] is meaningful. */
  APB_FUNCTOR_VIRIAL_TRACE = 16
};

typedef struct {
  int32_t kind;  /* enum apb_functor_kind */
  int32_t flags; /* enum apb_functor_flags */
  double cutoff; /* Functor::getCutoff() */
  /* non-mixing parameters: LJFunctor::setParticleProperties(epsilon24, sigmaSquared) (LJFunctor.h:588-596);
   * shift6 is derived by the library with ParticlePropertiesLibrary::calcShift6 when APPLY_SHIFT is set */
  double epsilon24;
  double sigma_squared;
  /* mixing: num_types and a row-major [i*T+j] table of {epsilon24, sigmaSquared, shift6}
   * (ParticlePropertiesLibrary.h:324-328, built by apb_make_lj_mixing_table) */
  int32_t num_types;
  const double *mixing_table;
  double nu; /* ATM, non-mixing: AxilrodTellerMutoFunctor::setParticleProperties(nu) */
  /* LJMultisiteFunctor: site geometry of ParticlePropertiesLibrary::getSitePositions / getSiteTypes
   * (ParticlePropertiesLibrary.h); molecule type m owns sites [site_start[m], site_start[m + 1]). Site types index
   * the mixing table above (num_types = number of site types). */
  int32_t num_mol_types;
  const int32_t *site_start;    /* num_mol_types + 1 entries */
  const double *site_positions; /* 3 doubles per site: unrotated position relative to the centre of mass */
  const int32_t *site_types;    /* per site */
} apb_functor;

/* raw accumulators as the reference functor keeps them before endTraversal (LJFunctor.h:1121-1136, 1145-1195) */
typedef struct {
  double upot_sum;      /* LJ: sum of 6*Upot*weight ; ATM: sum of 3*Upot ... exactly the functor's _potentialEnergySum input */
  double virial_sum[3];
  uint64_t num_dist_calls;
  uint64_t num_kernel_calls_n3;
  uint64_t num_kernel_calls_no_n3;
  uint64_t num_global_calcs_n3;
  uint64_t num_global_calcs_no_n3;
} apb_traversal_result;

typedef struct {
  int64_t cells_per_dim[3];   /* LinkedCells incl. halo (CellBlock3D::getCellsPerDimensionWithHalo) or towers x,y,1 */
  double cell_length[3];      /* cell length, or tower side length x,y and 0 */
  double interaction_length;  /* cutoff + skin */
  int64_t cluster_size;
  int64_t num_slots;          /* particle slots in storage order, incl. VCL padding dummies */
  int64_t num_cells;          /* cells or towers */
  int64_t num_clusters;       /* VCL */
  int64_t num_cluster_pairs;  /* VCL: entries of the cluster-pair neighbour list */
  int64_t towers_per_interaction_length; /* VCL */
} apb_geometry;

/* ---- life cycle ------------------------------------------------------------------------------------------------ */
/* replaces ContainerSelector::generateContainer (tuning/selectors/ContainerSelector.h:44-108) */
int apb_create(const apb_config *config, apb_handle *out_handle);
int apb_destroy(apb_handle h);
/* message of the last failing call on this handle (h == NULL: last failing apb_create) */
const char *apb_last_error(apb_handle h);

/* ---- particle storage ------------------------------------------------------------------------------------------ */
/* bulk form of ParticleContainerInterface::addParticleImpl / addHaloParticleImpl
 * (containers/ParticleContainerInterface.h:112,144). ids / types may be NULL (0..n-1 / 0). Appended particles are
 * staged and sorted into the container at the next rebuild, like VerletClusterLists::_particlesToAdd
 * (VerletClusterLists.h:184-192). check_box != 0: owned particles must lie in the box (half-open), else
 * APB_ERR_PARTICLE_OUTSIDE. */
int apb_add_particles(apb_handle h, int64_t n, const double *x, const double *y, const double *z, const int64_t *ids,
                      const int32_t *types, int32_t ownership, int32_t check_box);
/* ParticleContainerInterface::reserve(numParticles, numParticlesHaloEstimate) (containers/ParticleContainerInterface.h:95):
 * sizes the device SoA once so that later appends / rebuilds do not have to grow it */
int apb_reserve(apb_handle h, int64_t num_particles, int64_t num_particles_halo_estimate);
/* ParticleContainerInterface::deleteAllParticles */
int apb_delete_all_particles(apb_handle h);
/* ParticleContainerInterface::deleteHaloParticles (LinkedCells.h:110): halo slots become dummies */
int apb_delete_halo_particles(apb_handle h);
/* bulk form of ParticleContainerInterface::updateHaloParticle (:152): positions of existing halo slots are replaced,
 * matched by id. *out_not_found = number of ids that had no halo slot (the caller then falls back to add). */
int apb_update_halo_particles(apb_handle h, int64_t n, const int64_t *ids, const double *x, const double *y,
                              const double *z, int64_t *out_not_found);
/* ParticleContainerInterface::getNumberOfParticles(behavior) */
int apb_get_num_particles(apb_handle h, int64_t *out_owned, int64_t *out_halo);
/* number of particle slots in storage order (incl. dummies); sizes of the column transfers below */
int apb_get_num_slots(apb_handle h, int64_t *out_slots);

/* Host mirror transfers (what ContainerIterator / getParticle hand out, iterators/ContainerIterator.h:324),
 * whole columns in storage order. */
int apb_download_column(apb_handle h, int32_t column, double *dst);
int apb_upload_column(apb_handle h, int32_t column, const double *src);
int apb_download_ids(apb_handle h, int64_t *ids, int32_t *types, int32_t *ownership);
int apb_upload_ownership(apb_handle h, const int32_t *ownership); /* deleteParticle = mark dummy */
/* fused transfers for the force step through host buffers: 3 columns each, pinned staging inside */
int apb_upload_positions(apb_handle h, const double *x, const double *y, const double *z);
int apb_download_forces(apb_handle h, double *fx, double *fy, double *fz);
/* The same through host arrays indexed by particle id: entry k belongs to id id_begin + k (0 <= k < num_ids). Only
 * owned particles are touched: positions are scattered to their slots on the device, forces gathered from them (ids
 * without an owned particle on this device read as zero). The host does not need to know the storage order, so
 * nothing has to be re-read after a rebuild; halo slots are refreshed by apb_exchange_halos / apb_update_halo_particles
 * as usual. */
int apb_upload_positions_by_id(apb_handle h, int64_t id_begin, int64_t num_ids, const double *x, const double *y,
                               const double *z);
int apb_download_forces_by_id(apb_handle h, int64_t id_begin, int64_t num_ids, double *fx, double *fy, double *fz);
/* One computeInteractions call for a caller whose particle data lives in host memory (LogicHandler::computeInteractionsPipeline,
 * LogicHandler.h:1258, seen from md-flexible's arrays): positions by id host -> device, [rebuild != 0: migration, halo
 * exchange, apb_rebuild_neighbor_lists | halo refresh], forces = 0, traversal, forces by id device -> host. Enqueued as
 * one stream-ordered batch with a single host synchronisation at the end. */
int apb_force_step_by_id(apb_handle h, int32_t traversal, const apb_functor *functor, int32_t newton3, int32_t rebuild,
                         int64_t id_begin, int64_t num_ids, const double *x, const double *y, const double *z,
                         double *fx, double *fy, double *fz, apb_traversal_result *out_result);
/* set force columns to a constant (TimeDiscretization.cpp:16-68 resets f to globalForce) */
int apb_reset_forces(apb_handle h, double fx, double fy, double fz);

/* ---- wire format ------------------------------------------------------------------------------------------------ */
/* md-flexible's MPI particle record (examples/md-flexible/src/ParticleSerializationTools.cpp:43-146: serializeParticle /
 * deserializeParticles, AttributesSize = 120 bytes for MoleculeLJ): id, position, velocity, force, oldForce, typeId,
 * ownershipState, 8 bytes each, memcpy'd in that order. apb_serialize_particles writes the records of the owned and / or
 * halo particles (ownership_mask bit 0 / bit 1) in storage order into a host buffer; apb_deserialize_particles appends
 * the particles of such a buffer with all their attributes (ParticleContainerInterface::addParticle / addHaloParticle
 * by the record's ownership state). Single-site particles only. */
#define APB_WIRE_RECORD_BYTES 120
int apb_serialize_particles(apb_handle h, int32_t ownership_mask, void *dst, int64_t capacity_records, int64_t *out_num);
int apb_deserialize_particles(apb_handle h, const void *src, int64_t num_records);
/* md-flexible's VTK checkpoint (examples/md-flexible/src/ParallelVtkWriter.cpp:55-201 recordParticleStates): the bytes of
 * one rank's "<session>_Particles_<rank>_<iteration>.vtu" piece - velocities, forces, typeIds, ids, positions of the owned
 * particles as ASCII rows in storage order, doubles as a default std::ostream prints them ("%.6g"), positions next to
 * the upper box corner with the raised precision of writeWithDynamicPrecision (:130-157) - formatted on the device from
 * the SoA columns and copied into `dst` (host). dst == NULL: only the size is returned through out_bytes. Where the
 * reference throws (a position identical to the border up to 15 digits) the call fails with
 * APB_ERR_INVALID_ARGUMENT. Single-site particles only. md-flexible's loader (MDFlexConfig.cpp:91-180) reads the file. */
int apb_vtk_particle_record(apb_handle h, void *dst, int64_t capacity_bytes, int64_t *out_bytes);
/* The same record written to `path` like the reference does (ParallelVtkWriter.cpp:61-70, 200): device -> pinned host
 * pieces -> fwrite, the copy of one piece overlapping the write of the previous one. A file that cannot be opened fails
 * with the reference's message. */
int apb_vtk_write_particle_record(apb_handle h, const char *path, int64_t *out_bytes);
/* The "<session>_Particles_<iteration>.pvtu" index rank 0 writes next to the pieces (ParallelVtkWriter.cpp:308-356,
 * file names as generateFilename :437-441 builds them); host text, no handle. */
int apb_vtk_pvtu_record(const char *session_name, int32_t num_ranks, uint64_t iteration, int32_t digits, char *dst,
                        int64_t capacity_bytes, int64_t *out_bytes);
/* The way back: md-flexible's checkpoint loader for one piece (loadParticlesFromRankRecord,
 * examples/md-flexible/src/configuration/MDFlexConfig.cpp:91-180). `src` (host) holds the bytes of a ".vtu" piece;
 * NumberOfPoints particles are appended as owned particles with velocities, forces, typeIds, ids and positions converted
 * on the device exactly like `stream >> value` converts them (correctly rounded), oldForce = 0. check_box != 0: a particle
 * outside [box_min, box_max) fails with APB_ERR_PARTICLE_OUTSIDE (AutoPas::addParticle throws there) and nothing is
 * added. A value that is not a decimal number ("inf", "nan": the reference's extraction fails on them too) or a missing
 * data array fails with APB_ERR_INVALID_ARGUMENT. Single-site particles only. */
int apb_vtk_load_particle_record(apb_handle h, const void *src, int64_t num_bytes, int32_t check_box, int64_t *out_num);

/* ---- container maintenance ------------------------------------------------------------------------------------- */
/* ParticleContainerInterface::updateContainer(bool keepNeighborListsValid) (:297);
 * keep != 0: LeavingParticleCollector::collectParticlesAndMarkNonOwnedAsDummy (LeavingParticleCollector.h:85-118);
 * keep == 0: LinkedCells.h:152-202 / VerletClusterLists.h:362-397. Leavers are kept in a library-owned buffer. */
int apb_update_container(apb_handle h, int32_t keep_neighbor_lists_valid, int64_t *out_num_leavers);
int apb_get_leavers(apb_handle h, double *x, double *y, double *z, double *vx, double *vy, double *vz, int64_t *ids,
                    int32_t *types);
/* any other column of the leavers of the last apb_update_container (the returned particles are whole copies,
 * LeavingParticleCollector.h:101-110: forces, quaternion / torque, the SPH attributes), same order as apb_get_leavers */
int apb_get_leaver_column(apb_handle h, int32_t column, double *dst);
/* ParticleContainerInterface::rebuildNeighborLists(TraversalInterface*) (:158): LinkedCells re-binning
 * (counting sort) or VerletClusterLists tower/cluster/pair-list construction (VerletClusterLists.h:779-799). */
int apb_rebuild_neighbor_lists(apb_handle h, int32_t traversal, int32_t newton3);
/* ParticleContainerInterface::getTraversalSelectorInfo (:303) and sizes for the debug dumps */
int apb_get_geometry(apb_handle h, apb_geometry *out);

/* ---- the hot path ---------------------------------------------------------------------------------------------- */
/* ParticleContainerInterface::computeInteractions(TraversalInterface*) (:252) with the functor's initTraversal
 * already applied (accumulators start at zero). `out` receives the raw accumulators the shim deposits into the
 * functor before Functor::endTraversal(newton3). May be NULL. */
int apb_compute_interactions(apb_handle h, int32_t traversal, const apb_functor *functor, int32_t newton3,
                             apb_traversal_result *out);

/* host-side helpers restating functor post-processing, so C callers need no C++:
 * LJFunctor::endTraversal + getPotentialEnergy/getVirial (LJFunctor.h:661-720) */
void apb_lj_end_traversal(const apb_traversal_result *raw, double *out_upot, double *out_virial);
/* LJFunctor::getNumFLOPs (LJFunctor.h:758-789) */
uint64_t apb_lj_num_flops(const apb_traversal_result *raw, int32_t apply_shift);
/* ParticlePropertiesLibrary::calculateMixingCoefficients (ParticlePropertiesLibrary.h:444-473); out: T*T*3 doubles */
int apb_make_lj_mixing_table(int32_t num_types, const double *epsilon, const double *sigma, double cutoff,
                             double *out_table);
/* ParticlePropertiesLibrary::calcShift6 (:576-582) */
double apb_lj_calc_shift6(double epsilon24, double sigma_squared, double cutoff_squared);
/* AxilrodTellerMutoFunctor::endTraversal + getPotentialEnergy / getVirial (AxilrodTellerMutoFunctor.h:360-420):
 * Upot = sum / 9, virial = vx + vy + vz (not scaled) */
void apb_atm_end_traversal(const apb_traversal_result *raw, double *out_upot, double *out_virial);
/* AxilrodTellerMutoFunctor::getNumFLOPs (:476-480): 24 D + 59 K_noN3 + 100 K_N3 + 10 G_noN3 + 24 G_N3 */
uint64_t apb_atm_num_flops(const apb_traversal_result *raw);

/* ---- device-resident simulation loop (SURVEY §8 e, f2): no host round trip between force steps ------------------- */
/* Störmer-Verlet halves of examples/md-flexible/src/TimeDiscretization.cpp:
 * calculatePositionsAndResetForces (:16-68): oldF = f; f = global_force (NULL = 0); r += v dt + f dt^2 / (2 m)
 * calculateVelocities (:152-165): v += (f + oldF) dt / (2 m). mass_of_type[typeId]; owned particles only. */
int apb_integrate_positions(apb_handle h, double dt, const double *mass_of_type, int32_t num_types,
                            const double *global_force);
int apb_integrate_velocities(apb_handle h, double dt, const double *mass_of_type, int32_t num_types);

/* Regular-grid spatial decomposition of examples/md-flexible/src/domainDecomposition/RegularGridDecomposition.cpp,
 * one rank per GPU, over NVLink instead of MPI (ParticleCommunicator.cpp:38-61): NCCL send/recv for the payload of the
 * generating exchanges (rebuild steps); counts and the per-step halo refresh go through peer memory (arenas mapped with
 * CUDA IPC among the ranks of one node; APB_NO_P2P_HALO=1 on all ranks keeps everything on NCCL).
 * apb_comm_get_unique_id: rank 0 creates the 128-byte NCCL id, the host distributes it (e.g. torch.distributed).
 * apb_comm_init: MPI_Cart_create analogue (:106); nranks == 1 needs no id and no NCCL.
 * apb_set_decomposition: this rank's box is apb_config.box_*; neighbours6 = {left,right} rank per dimension
 * (:137-149), periodic3 = BoundaryTypeOption::periodic per dimension. Defaults for a single rank: global box = local
 * box, periodic in all dimensions, own neighbour. */
int apb_comm_get_unique_id(void *out_128_bytes);
int apb_comm_init(apb_handle h, int32_t nranks, int32_t rank, const void *unique_id_128_bytes);
int apb_set_decomposition(apb_handle h, const double *global_box_min, const double *global_box_max,
                          const int32_t *neighbours6, const int32_t *periodic3);
/* AutoPas::updateContainer + RegularGridDecomposition::exchangeMigratingParticles (:238-301) fused on the device:
 * halos dropped, owned particles outside the local box travel to the neighbour per dimension (periodic wrap at global
 * boundaries), arrivals become owned. Invalidates the structure (a rebuild must follow). In a dimension in which the
 * rank is its own neighbour the coordinate is wrapped in place; the out counts cover the exchanged dimensions only. */
int apb_migrate(apb_handle h, int64_t *out_num_sent, int64_t *out_num_received);
/* RegularGridDecomposition::exchangeHaloParticles (:159-236). Structure invalid (rebuild step): select + append halos,
 * x then y then z, forwarding received halos. Structure valid: refresh the positions of the existing halo copies only
 * (bulk updateHaloParticle), same three-phase order, fixed message sizes. A single fully periodic rank called after
 * apb_migrate generates (and later refreshes) all periodic images in one pass - the same set of halo copies. */
int apb_exchange_halos(apb_handle h);
/* Refresh of other columns of the halo copies from their source particles through the links the last generating
 * apb_exchange_halos recorded: sph-mpi's updateHaloParticles between the density and the hydro-force pass
 * (examples/sph-mpi/sph-main-mpi.cpp:271-315, 373-414: the copies need the owners' density and pressure). Positions are
 * not accepted here (apb_exchange_halos shifts them at the periodic boundary). For SPHParticle / MultisiteMoleculeLJ
 * storage the generating exchange itself copies every attribute column to the new halo copies. */
int apb_refresh_halo_columns(apb_handle h, int32_t num_columns, const int32_t *columns);
/* MPI_Reduce(SUM) of potential energy / virial in Simulation.cpp:319-322, as ncclAllReduce; no-op for one rank */
int apb_allreduce_globals(apb_handle h, apb_traversal_result *inout);

/* Velocity-scaling thermostat of examples/md-flexible/src/Thermostat.h on the device.
 * apb_calc_temperature: Thermostat::calcTemperatureComponent (:66-150): per particle type T = sum(m v.v) / (3 N), k_B = 1,
 * summed over the ranks; every particle counts once (owned only - the reference's iterator also visits halo copies).
 * apb_apply_thermostat: Thermostat::apply (:228-275): per type the temperature moves by at most |delta| towards the
 * target, velocities are scaled by sqrt(T_new / T_current). 1 .. 32 types.
 * apb_set_thermostat: apb_run_steps applies it every `interval` iterations after the velocity update
 * (Simulation::updateThermostat, Simulation.cpp:539-546). */
int apb_calc_temperature(apb_handle h, const double *mass_of_type, int32_t num_types, double *out_temperature,
                         int64_t *out_count);
int apb_apply_thermostat(apb_handle h, const double *mass_of_type, int32_t num_types, double target_temperature,
                         double delta_temperature);
int apb_set_thermostat(apb_handle h, int32_t enable, int32_t interval, double target_temperature,
                       double delta_temperature);
/* Dynamic-rebuild trigger (AUTOPAS_ENABLE_DYNAMIC_CONTAINERS, src/autopas/LogicHandler.h:955-965, 1000-1016). Enabled:
 * every rebuild records the positions (ParticleBase::resetRAtRebuild); apb_check_dynamic_rebuild reports 1 as soon as an
 * owned particle is skin / 2 or more away from its recorded position (max over the ranks), and apb_run_steps rebuilds
 * then, or rebuild_frequency steps after the last rebuild, whichever comes first (LogicHandler.h:987-997). */
int apb_set_dynamic_rebuild(apb_handle h, int32_t enable);
int apb_check_dynamic_rebuild(apb_handle h, int32_t *out_rebuild_needed);
int apb_get_dynamic_rebuild_count(apb_handle h, int64_t *out_count); /* rebuilds apb_run_steps did because of it */
/* RemainderPairwiseInteractionHandler::computeRemainderInteractions (src/autopas/remainder/
 * RemainderPairwiseInteractionHandler.h:63-135) for LJFunctor: particles LogicHandler holds in its particle / halo
 * buffers until the next rebuild (LogicHandler.h:339-391) interact with the container particles and with each other
 * (all pairs with a buffered particle except halo-halo), each pair applied to both partners. Container forces are
 * updated in place; fx, fy, fz receive the forces on the buffered particles from this call; ownership: 1 owned, 2 halo
 * (0 = skip); types may be NULL. out_result: raw accumulators to add to those of apb_compute_interactions. */
int apb_compute_remainder(apb_handle h, const apb_functor *functor, int64_t num_buffered, const double *x,
                          const double *y, const double *z, const int32_t *types, const int32_t *ownership, double *fx,
                          double *fy, double *fz, apb_traversal_result *out_result);

typedef struct {
  double dt;                   /* deltaT */
  const double *mass_of_type;  /* ParticlePropertiesLibrary::getMolMass(typeId) */
  int32_t num_types;
  const double *global_force;  /* 3 doubles or NULL */
  int32_t rebuild_frequency;   /* AutoPas::setVerletRebuildFrequency */
  int32_t traversal;           /* enum apb_traversal */
  int32_t newton3;
} apb_loop_params;
/* Simulation::simulate (examples/md-flexible/src/Simulation.cpp:230-351) for a built-in functor, device resident:
 * per iteration  positions -> [iteration % rebuild_frequency == 0: migrate, halo exchange, rebuild | halo refresh]
 * -> forces -> velocities. Enqueued asynchronously; blocks on rebuild steps and at the end. out_per_step (num_steps
 * entries or NULL) receives each step's raw accumulators (rank-local; see apb_allreduce_globals). */
int apb_run_steps(apb_handle h, const apb_functor *functor, const apb_loop_params *params, int32_t num_steps,
                  int64_t first_iteration, apb_traversal_result *out_per_step);
/* the CUDA stream (cudaStream_t) all device work of this handle is enqueued on, for event timing by the caller */
int apb_get_stream(apb_handle h, void **out_stream);

/* ---- measurement ------------------------------------------------------------------------------------------------- */
/* number of CUDA kernels launched through this handle so far (bench.py's gpu_launches) */
int apb_get_launch_count(apb_handle h, int64_t *out_count);
/* number of device allocations (buffer growth events) made through this handle so far: a rebuild of an unchanged
 * system makes none (tests/test_gpu_dynamics.py) */
int apb_get_alloc_count(apb_handle h, int64_t *out_count);
/* CUDA-event timing of the phases of apb_run_steps on the handle's stream, the analogue of the reference's
 * rebuild / computeInteractions / remainder timers (LogicHandler.h:1083-1125).
 * out_ms[4] / out_counts[4]: 0 force kernels, 1 rebuild (migrate + halo generation + structure build),
 * 2 halo refresh, 3 integration. Reading resets the record. */
int apb_enable_loop_timing(apb_handle h, int32_t enable);
int apb_get_loop_timing(apb_handle h, double *out_ms, int64_t *out_counts);
/* sustained DFMA rate of the device (TFLOP/s, 2 flops per FMA): the FP64 roofline denominator, measured live */
int apb_measure_fp64_peak(int32_t device, int32_t repeats, double *out_tflops, double *out_ms);

/* ---- parity artefacts (tests only; not used by the hot path) --------------------------------------------------- */
/* per slot: 1-D cell index (CellBlock3D::get1DIndexOfPosition) or tower index (ClusterTowerBlock2D) */
int apb_debug_cell_of_slot(apb_handle h, int64_t *out_cell);
/* VCL cluster-pair list as pairs of global cluster indices (cluster c covers slots [c*M, (c+1)*M)) */
int apb_debug_cluster_pairs(apb_handle h, int64_t *out_pairs /* 2*num_cluster_pairs */);

#ifdef __cplusplus
}
#endif
#endif /* AUTOPAS_B200_H */
