"""ctypes binding of the C ABI (include/autopas_b200.h) — the same entry points a cgo / JNI / C++ caller binds.

There is no CPU fallback: importing works without a GPU (so that symbol checks run anywhere), but every compute call
needs the CUDA library and a device, and raises ``ApbError`` otherwise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("APB_LIB_PATH", os.path.join(_HERE, "libautopas_b200.so"))  # override: kernel experiments

APB_OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NOT_APPLICABLE, ERR_STATE, ERR_PARTICLE_OUTSIDE, ERR_NCCL, ERR_OOM = (
    -1, -2, -3, -4, -5, -6, -7)

CONTAINER_LINKED_CELLS, CONTAINER_VERLET_CLUSTER_LISTS = 0, 1
(TRAVERSAL_GPULC_C08, TRAVERSAL_GPULC_C18, TRAVERSAL_GPUVCL_CLUSTER_ITERATION, TRAVERSAL_GPUVCL_C06,
 TRAVERSAL_GPUVCL_C01_BALANCED, TRAVERSAL_GPUVCL_PRUNED) = range(6)
PARTICLE_LJ, PARTICLE_MULTISITE, PARTICLE_SPH = 0, 1, 2
FUNCTOR_LJ, FUNCTOR_LJ_MULTISITE, FUNCTOR_ATM, FUNCTOR_SPH_DENSITY, FUNCTOR_SPH_HYDRO = range(5)
FLAG_APPLY_SHIFT, FLAG_USE_MIXING, FLAG_CALC_GLOBALS, FLAG_COUNT_FLOPS, FLAG_VIRIAL_TRACE = 1, 2, 4, 8, 16
OWN_DUMMY, OWN_OWNED, OWN_HALO = 0, 1, 2
WIRE_RECORD_BYTES = 120  # md-flexible's MPI record of a MoleculeLJ (ParticleSerializationTools.cpp:66)

COLUMNS = ["X", "Y", "Z", "VX", "VY", "VZ", "FX", "FY", "FZ", "OLDFX", "OLDFY", "OLDFZ", "Q0", "Q1", "Q2", "Q3", "TX",
           "TY", "TZ", "MASS", "SMTH", "DENSITY", "PRESSURE", "SNDSPEED", "ENGDOT", "VSIGMAX"]
COL = {name: i for i, name in enumerate(COLUMNS)}

TRAVERSAL_NAMES = {
    "gpulc_c08": TRAVERSAL_GPULC_C08,
    "gpulc_c18": TRAVERSAL_GPULC_C18,
    "gpuvcl_cluster_iteration": TRAVERSAL_GPUVCL_CLUSTER_ITERATION,
    "gpuvcl_c06": TRAVERSAL_GPUVCL_C06,
    "gpuvcl_c01_balanced": TRAVERSAL_GPUVCL_C01_BALANCED,
    "gpuvcl_pruned": TRAVERSAL_GPUVCL_PRUNED,
}
CONTAINER_NAMES = {"gpuLinkedCells": CONTAINER_LINKED_CELLS, "gpuVerletClusterLists": CONTAINER_VERLET_CLUSTER_LISTS}


class ApbError(RuntimeError):
    """Mirrors autopas::utils::ExceptionHandler::AutoPasException (src/autopas/utils/ExceptionHandler.h:116)."""

    def __init__(self, code, message):
        super().__init__(f"[apb {code}] {message}")
        self.code = code


class Config(ctypes.Structure):
    _fields_ = [
        ("box_min", ctypes.c_double * 3),
        ("box_max", ctypes.c_double * 3),
        ("cutoff", ctypes.c_double),
        ("skin", ctypes.c_double),
        ("cell_size_factor", ctypes.c_double),
        ("cluster_size", ctypes.c_int32),
        ("container", ctypes.c_int32),
        ("particle_kind", ctypes.c_int32),
        ("device", ctypes.c_int32),
    ]


class Functor(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("cutoff", ctypes.c_double),
        ("epsilon24", ctypes.c_double),
        ("sigma_squared", ctypes.c_double),
        ("num_types", ctypes.c_int32),
        ("mixing_table", ctypes.c_void_p),
        ("nu", ctypes.c_double),
        ("num_mol_types", ctypes.c_int32),
        ("site_start", ctypes.c_void_p),
        ("site_positions", ctypes.c_void_p),
        ("site_types", ctypes.c_void_p),
    ]


class TraversalResult(ctypes.Structure):
    _fields_ = [
        ("upot_sum", ctypes.c_double),
        ("virial_sum", ctypes.c_double * 3),
        ("num_dist_calls", ctypes.c_uint64),
        ("num_kernel_calls_n3", ctypes.c_uint64),
        ("num_kernel_calls_no_n3", ctypes.c_uint64),
        ("num_global_calcs_n3", ctypes.c_uint64),
        ("num_global_calcs_no_n3", ctypes.c_uint64),
    ]


class LoopParams(ctypes.Structure):
    _fields_ = [
        ("dt", ctypes.c_double),
        ("mass_of_type", ctypes.c_void_p),
        ("num_types", ctypes.c_int32),
        ("global_force", ctypes.c_void_p),
        ("rebuild_frequency", ctypes.c_int32),
        ("traversal", ctypes.c_int32),
        ("newton3", ctypes.c_int32),
    ]


class Geometry(ctypes.Structure):
    _fields_ = [
        ("cells_per_dim", ctypes.c_int64 * 3),
        ("cell_length", ctypes.c_double * 3),
        ("interaction_length", ctypes.c_double),
        ("cluster_size", ctypes.c_int64),
        ("num_slots", ctypes.c_int64),
        ("num_cells", ctypes.c_int64),
        ("num_clusters", ctypes.c_int64),
        ("num_cluster_pairs", ctypes.c_int64),
        ("towers_per_interaction_length", ctypes.c_int64),
    ]


_H = ctypes.c_void_p
_vp = ctypes.c_void_p
_i32, _i64, _f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double

# name -> (restype, argtypes); every symbol declared in include/autopas_b200.h must be listed here
SIGNATURES = {
    "apb_create": (_i32, [ctypes.POINTER(Config), ctypes.POINTER(_H)]),
    "apb_destroy": (_i32, [_H]),
    "apb_last_error": (ctypes.c_char_p, [_H]),
    "apb_add_particles": (_i32, [_H, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32]),
    "apb_reserve": (_i32, [_H, _i64, _i64]),
    "apb_delete_all_particles": (_i32, [_H]),
    "apb_delete_halo_particles": (_i32, [_H]),
    "apb_update_halo_particles": (_i32, [_H, _i64, _vp, _vp, _vp, _vp, ctypes.POINTER(_i64)]),
    "apb_get_num_particles": (_i32, [_H, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "apb_get_num_slots": (_i32, [_H, ctypes.POINTER(_i64)]),
    "apb_download_column": (_i32, [_H, _i32, _vp]),
    "apb_upload_column": (_i32, [_H, _i32, _vp]),
    "apb_download_ids": (_i32, [_H, _vp, _vp, _vp]),
    "apb_upload_ownership": (_i32, [_H, _vp]),
    "apb_upload_positions": (_i32, [_H, _vp, _vp, _vp]),
    "apb_download_forces": (_i32, [_H, _vp, _vp, _vp]),
    "apb_upload_positions_by_id": (_i32, [_H, _i64, _i64, _vp, _vp, _vp]),
    "apb_download_forces_by_id": (_i32, [_H, _i64, _i64, _vp, _vp, _vp]),
    "apb_force_step_by_id": (_i32, [_H, _i32, _vp, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "apb_reset_forces": (_i32, [_H, _f64, _f64, _f64]),
    "apb_update_container": (_i32, [_H, _i32, ctypes.POINTER(_i64)]),
    "apb_get_leavers": (_i32, [_H, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "apb_get_leaver_column": (_i32, [_H, _i32, _vp]),
    "apb_serialize_particles": (_i32, [_H, _i32, _vp, _i64, _vp]),
    "apb_deserialize_particles": (_i32, [_H, _vp, _i64]),
    "apb_vtk_particle_record": (_i32, [_H, _vp, _i64, _vp]),
    "apb_vtk_write_particle_record": (_i32, [_H, ctypes.c_char_p, _vp]),
    "apb_vtk_load_particle_record": (_i32, [_H, _vp, _i64, _i32, _vp]),
    "apb_vtk_pvtu_record": (_i32, [ctypes.c_char_p, _i32, ctypes.c_uint64, _i32, _vp, _i64, _vp]),
    "apb_rebuild_neighbor_lists": (_i32, [_H, _i32, _i32]),
    "apb_get_geometry": (_i32, [_H, ctypes.POINTER(Geometry)]),
    "apb_compute_interactions": (_i32, [_H, _i32, ctypes.POINTER(Functor), _i32, ctypes.POINTER(TraversalResult)]),
    "apb_lj_end_traversal": (None, [ctypes.POINTER(TraversalResult), ctypes.POINTER(_f64), ctypes.POINTER(_f64)]),
    "apb_lj_num_flops": (ctypes.c_uint64, [ctypes.POINTER(TraversalResult), _i32]),
    "apb_make_lj_mixing_table": (_i32, [_i32, _vp, _vp, _f64, _vp]),
    "apb_lj_calc_shift6": (_f64, [_f64, _f64, _f64]),
    "apb_atm_end_traversal": (None, [ctypes.POINTER(TraversalResult), ctypes.POINTER(_f64), ctypes.POINTER(_f64)]),
    "apb_atm_num_flops": (ctypes.c_uint64, [ctypes.POINTER(TraversalResult)]),
    "apb_integrate_positions": (_i32, [_H, _f64, _vp, _i32, _vp]),
    "apb_integrate_velocities": (_i32, [_H, _f64, _vp, _i32]),
    "apb_calc_temperature": (_i32, [_H, _vp, _i32, _vp, _vp]),
    "apb_apply_thermostat": (_i32, [_H, _vp, _i32, _f64, _f64]),
    "apb_set_thermostat": (_i32, [_H, _i32, _i32, _f64, _f64]),
    "apb_set_dynamic_rebuild": (_i32, [_H, _i32]),
    "apb_check_dynamic_rebuild": (_i32, [_H, ctypes.POINTER(_i32)]),
    "apb_get_dynamic_rebuild_count": (_i32, [_H, ctypes.POINTER(_i64)]),
    "apb_compute_remainder": (_i32, [_H, ctypes.POINTER(Functor), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     ctypes.POINTER(TraversalResult)]),
    "apb_comm_get_unique_id": (_i32, [_vp]),
    "apb_comm_init": (_i32, [_H, _i32, _i32, _vp]),
    "apb_set_decomposition": (_i32, [_H, _vp, _vp, _vp, _vp]),
    "apb_migrate": (_i32, [_H, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "apb_refresh_halo_columns": (_i32, [_H, _i32, _vp]),
    "apb_exchange_halos": (_i32, [_H]),
    "apb_allreduce_globals": (_i32, [_H, ctypes.POINTER(TraversalResult)]),
    "apb_run_steps": (_i32, [_H, ctypes.POINTER(Functor), _vp, _i32, _i64, _vp]),
    "apb_get_stream": (_i32, [_H, ctypes.POINTER(_vp)]),
    "apb_get_launch_count": (_i32, [_H, ctypes.POINTER(_i64)]),
    "apb_get_alloc_count": (_i32, [_H, ctypes.POINTER(_i64)]),
    "apb_enable_loop_timing": (_i32, [_H, _i32]),
    "apb_get_loop_timing": (_i32, [_H, _vp, _vp]),
    "apb_measure_fp64_peak": (_i32, [_i32, _i32, ctypes.POINTER(_f64), ctypes.POINTER(_f64)]),
    "apb_debug_cell_of_slot": (_i32, [_H, _vp]),
    "apb_debug_cluster_pairs": (_i32, [_H, _vp]),
}

_lib = None


def load():
    """Load libautopas_b200.so. Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ApbError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `make -C autopas_b200/csrc` "
                                     "(__graft_entry__.build()); there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
