"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracle.

* ``liboracle.so``            : our plain-C restatement (oracle/autopas_oracle.c, oracle/*.c), built by oracle/Makefile
* ``_ref/libautopas_ref.so``  : the unmodified AutoPas reference compiled from /root/reference (oracle/ref_driver.cpp);
                                present only if it was built in the development container (it travels to the GPU box
                                as a binary; /root/reference itself does not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package ``autopas_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libautopas_ref.so")

F_SHIFT, F_MIXING, F_NEWTON3, F_SOA = 1, 2, 4, 8
OWN_DUMMY, OWN_OWNED, OWN_HALO = 0, 1, 2


class Result(ctypes.Structure):
    _fields_ = [
        ("upot_sum", ctypes.c_double),
        ("virial_sum", ctypes.c_double * 3),
        ("num_dist_calls", ctypes.c_uint64),
        ("num_kernel_calls_n3", ctypes.c_uint64),
        ("num_kernel_calls_no_n3", ctypes.c_uint64),
        ("num_global_calcs_n3", ctypes.c_uint64),
        ("num_global_calcs_no_n3", ctypes.c_uint64),
    ]

    def as_dict(self):
        return {
            "upot_sum": self.upot_sum,
            "virial_sum": tuple(self.virial_sum),
            "num_dist_calls": self.num_dist_calls,
            "num_kernel_calls_n3": self.num_kernel_calls_n3,
            "num_kernel_calls_no_n3": self.num_kernel_calls_no_n3,
            "num_global_calcs_n3": self.num_global_calcs_n3,
            "num_global_calcs_no_n3": self.num_global_calcs_no_n3,
        }


def build(with_ref=True):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    target = "all" if with_ref else os.path.join(_HERE, "liboracle.so")
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build(with_ref=False)
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_calc_shift6.restype = ctypes.c_double
        _lib.orc_calc_shift6.argtypes = [ctypes.c_double] * 3
        _lib.orc_lj_num_flops.restype = ctypes.c_uint64
    return _lib


def have_ref():
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF)
    return _ref


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def _flags(shift, mixing, newton3, soa=False):
    return (F_SHIFT if shift else 0) | (F_MIXING if mixing else 0) | (F_NEWTON3 if newton3 else 0) | (F_SOA if soa else 0)


def lj_end_traversal(res):
    """LJFunctor::endTraversal normalisation -> (upot, virial)."""
    u, v = ctypes.c_double(), ctypes.c_double()
    lib().orc_lj_end_traversal(ctypes.byref(res), ctypes.byref(u), ctypes.byref(v))
    return u.value, v.value


def lj_num_flops(res, shift):
    return lib().orc_lj_num_flops(ctypes.byref(res), ctypes.c_int(1 if shift else 0))


def mixing_table(eps, sigma, cutoff):
    eps, sigma = _f64(eps), _f64(sigma)
    T = len(eps)
    out = np.zeros(T * T * 3)
    lib().orc_mixing_table(ctypes.c_int(T), _p(eps), _p(sigma), ctypes.c_double(cutoff), _p(out))
    return out


def lc_cell_indices(box_min, box_max, il, csf, x, y, z):
    x, y, z = _f64(x), _f64(y), _f64(z)
    n = len(x)
    cell = np.zeros(n, dtype=np.int64)
    cpd = np.zeros(3, dtype=np.int64)
    lib().orc_lc_cell_indices(_p(_f64(box_min)), _p(_f64(box_max)), ctypes.c_double(il), ctypes.c_double(csf),
                              ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cell), _p(cpd))
    return cell, cpd


def _common(x, y, z, types, own, eps, sigma):
    x, y, z = _f64(x), _f64(y), _f64(z)
    n = len(x)
    own = _i64(own if own is not None else np.ones(n))
    types = _i64(types if types is not None else np.zeros(n))
    eps, sigma = _f64(np.atleast_1d(eps)), _f64(np.atleast_1d(sigma))
    return x, y, z, types, own, eps, sigma, n


def lj_linkedcells(x, y, z, types, own, box_min, box_max, cutoff, skin, csf=1.0, shift=False, mixing=False,
                   newton3=True, eps=1.0, sigma=1.0):
    x, y, z, types, own, eps, sigma, n = _common(x, y, z, types, own, eps, sigma)
    f = np.zeros((n, 3))
    fscale = np.zeros(n)
    res = Result()
    cell = np.full(n, -1, dtype=np.int64)
    lib().orc_lj_linkedcells(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(types), _p(own), _p(_f64(box_min)),
                             _p(_f64(box_max)), ctypes.c_double(cutoff), ctypes.c_double(skin), ctypes.c_double(csf),
                             ctypes.c_int(_flags(shift, mixing, newton3)), ctypes.c_int(len(eps)), _p(eps), _p(sigma),
                             _p(f), _p(fscale), ctypes.byref(res), _p(cell))
    return {"f": f, "fscale": fscale, "res": res, "cell": cell}


def lc_pair_offsets(cells_per_dim, cell_length, interaction_length):
    """Linear offset differences of the cell pairs of one base cell (self = 0 included), ascending."""
    cpd = np.asarray(cells_per_dim, dtype=np.int64)
    cl = _f64(cell_length)
    out = np.zeros(4096, dtype=np.int64)
    lib().orc_lc_pair_offsets.restype = ctypes.c_int64
    n = lib().orc_lc_pair_offsets(_p(cpd), _p(cl), ctypes.c_double(interaction_length), _p(out), ctypes.c_int64(len(out)))
    return out[:n]


def lj_bruteforce(x, y, z, types, own, cutoff, shift=False, mixing=False, eps=1.0, sigma=1.0):
    x, y, z, types, own, eps, sigma, n = _common(x, y, z, types, own, eps, sigma)
    f = np.zeros((n, 3))
    fscale = np.zeros(n)
    res = Result()
    lib().orc_lj_bruteforce(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(types), _p(own), ctypes.c_double(cutoff),
                            ctypes.c_int(_flags(shift, mixing, True)), ctypes.c_int(len(eps)), _p(eps), _p(sigma),
                            _p(f), _p(fscale), ctypes.byref(res))
    return {"f": f, "fscale": fscale, "res": res}


def lj_vcl(x, y, z, types, own, box_min, box_max, cutoff, skin, cluster_size, shift=False, mixing=False,
           newton3=False, eps=1.0, sigma=1.0):
    x, y, z, types, own, eps, sigma, n = _common(x, y, z, types, own, eps, sigma)
    f = np.zeros((n, 3))
    fscale = np.zeros(n)
    res = Result()
    lib().orc_lj_vcl(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(types), _p(own), _p(_f64(box_min)), _p(_f64(box_max)),
                     ctypes.c_double(cutoff), ctypes.c_double(skin), ctypes.c_int64(cluster_size),
                     ctypes.c_int(_flags(shift, mixing, newton3)), ctypes.c_int(len(eps)), _p(eps), _p(sigma), _p(f),
                     _p(fscale), ctypes.byref(res))
    sizes = np.zeros(5, dtype=np.int64)
    side = np.zeros(2)
    lib().orc_vcl_dump_sizes(_p(sizes), _p(side))
    nslots = int(sizes[0] * sizes[4])
    slot_particle = np.zeros(max(nslots, 1), dtype=np.int64)
    slot_tower = np.zeros(max(nslots, 1), dtype=np.int64)
    pairs = np.zeros((max(int(sizes[1]), 1), 2), dtype=np.int64)
    lib().orc_vcl_dump_copy(_p(slot_particle), _p(slot_tower), _p(pairs))
    return {"f": f, "fscale": fscale, "res": res, "num_clusters": int(sizes[0]), "num_pairs": int(sizes[1]),
            "towers_per_dim": (int(sizes[2]), int(sizes[3])), "tower_side": tuple(side),
            "slot_particle": slot_particle[:nslots], "slot_tower": slot_tower[:nslots], "pairs": pairs[:int(sizes[1])]}


def ref_bench_lj(x, y, z, own, box_min, box_max, cutoff, skin, container="VerletClusterLists", traversal="vcl_c06",
                 cluster_size=4, newton3=True, iters=10, rebuild_freq=10):
    """Time the unmodified reference's force step (OpenMP, all host threads). Returns a dict with the total rebuild and
    compute seconds. Only bench.py's cpu_baseline / --impl reference legs call this."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    own = _i64(own)
    cont = {"LinkedCells": 0, "VerletClusterLists": 1}[container]
    trav = {"lc_c08": 0, "lc_c18": 1, "vcl_cluster_iteration": 0, "vcl_c06": 1, "vcl_c01_balanced": 2}[traversal]
    sec = np.zeros(4)
    rc = ref().ref_bench_lj(ctypes.c_int64(len(x)), _p(x), _p(y), _p(z), _p(own), _p(_f64(box_min)), _p(_f64(box_max)),
                            ctypes.c_double(cutoff), ctypes.c_double(skin), ctypes.c_int(cont), ctypes.c_int(trav),
                            ctypes.c_int64(cluster_size), ctypes.c_int(1 if newton3 else 0), ctypes.c_int(iters),
                            ctypes.c_int(rebuild_freq), _p(sec))
    if rc != 0:
        raise RuntimeError("reference bench run failed")
    return {"rebuild_s": sec[0], "compute_s": sec[1], "num_rebuilds": int(sec[2]), "upot": sec[3],
            "threads": ref().ref_num_threads()}


# ---- reference arm of bench.py: the reference's LJ kernels built like its own Release build (oracle/ref_bench.cpp) -------
_refbench = None


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def refbench_path():
    """The timing library for the highest x86-64 micro-architecture level this host runs (v4 = AVX-512, v3 = AVX2)."""
    flags = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    order = ["v4", "v3"] if v4 else ["v3"]
    for lvl in order:
        path = os.path.join(_HERE, "_ref", f"libautopas_refbench_{lvl}.so")
        if os.path.exists(path):
            return path
    return None


def have_refbench():
    return refbench_path() is not None


def refbench():
    global _refbench
    if _refbench is None:
        _refbench = ctypes.CDLL(refbench_path())
        _refbench.refb_isa.restype = ctypes.c_char_p
    return _refbench


def refbench_lj(x, y, z, own, box_min, box_max, cutoff, skin, functor="LJFunctor", container="LinkedCells",
                traversal="lc_c08", cluster_size=4, newton3=True, warmup=1, iters=10, rebuild_freq=10, threads=None):
    """Time the unmodified reference's force step (OpenMP on `threads` host threads, default: every core this process
    may run on - torchrun's OMP_NUM_THREADS=1 is ignored on purpose). Only bench.py's cpu_baseline / --impl reference
    legs call this."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    own = _i64(own)
    lib_ = refbench()
    if threads is None:
        threads = len(os.sched_getaffinity(0))
    lib_.refb_set_num_threads(ctypes.c_int(int(threads)))
    fun = {"LJFunctor": 0, "LJFunctorHWY": 1}[functor]
    cont = {"LinkedCells": 0, "VerletClusterLists": 1}[container]
    trav = {"lc_c08": 0, "lc_c18": 1, "vcl_cluster_iteration": 0, "vcl_c06": 1, "vcl_c01_balanced": 2}[traversal]
    out = np.zeros(5)
    rc = lib_.refb_bench_lj(ctypes.c_int64(len(x)), _p(x), _p(y), _p(z), _p(own), _p(_f64(box_min)), _p(_f64(box_max)),
                            ctypes.c_double(cutoff), ctypes.c_double(skin), ctypes.c_int(fun), ctypes.c_int(cont),
                            ctypes.c_int(trav), ctypes.c_int64(cluster_size), ctypes.c_int(1 if newton3 else 0),
                            ctypes.c_int(warmup), ctypes.c_int(iters), ctypes.c_int(rebuild_freq), _p(out))
    if rc != 0:
        raise RuntimeError("reference bench run failed")
    return {"rebuild_s": out[0], "compute_s": out[1], "num_rebuilds": int(out[2]), "upot": out[3], "virial": out[4],
            "threads": lib_.refb_num_threads(), "isa": lib_.refb_isa().decode()}


# ---- the unmodified reference (oracle/_ref) --------------------------------------------------------------------
def ref_lj_linkedcells(x, y, z, types, own, box_min, box_max, cutoff, skin, csf=1.0, shift=False, mixing=False,
                       newton3=True, soa=False, traversal="lc_c08", eps=1.0, sigma=1.0):
    x, y, z, types, own, eps, sigma, n = _common(x, y, z, types, own, eps, sigma)
    f = np.zeros((n, 3))
    glob = np.zeros(2)
    flops = ctypes.c_uint64()
    hit = ctypes.c_double()
    cell = np.zeros(n, dtype=np.int64)
    cpd = np.zeros(3, dtype=np.int64)
    rc = ref().ref_lj_linkedcells(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(types), _p(own), _p(_f64(box_min)),
                                  _p(_f64(box_max)), ctypes.c_double(cutoff), ctypes.c_double(skin),
                                  ctypes.c_double(csf), ctypes.c_int(_flags(shift, mixing, newton3, soa)),
                                  ctypes.c_int({"lc_c08": 0, "lc_c18": 1}[traversal]), ctypes.c_int(len(eps)), _p(eps),
                                  _p(sigma), _p(f), _p(glob), ctypes.byref(flops), ctypes.byref(hit), _p(cell), _p(cpd))
    if rc != 0:
        raise RuntimeError("reference LinkedCells run failed")
    return {"f": f, "upot": glob[0], "virial": glob[1], "flops": flops.value, "hit_rate": hit.value, "cell": cell,
            "cells_per_dim": cpd}


def ref_lj_vcl(x, y, z, types, own, box_min, box_max, cutoff, skin, cluster_size, shift=False, mixing=False,
               newton3=False, soa=True, traversal="vcl_cluster_iteration", eps=1.0, sigma=1.0):
    x, y, z, types, own, eps, sigma, n = _common(x, y, z, types, own, eps, sigma)
    f = np.zeros((n, 3))
    glob = np.zeros(2)
    flops = ctypes.c_uint64()
    hit = ctypes.c_double()
    trav = {"vcl_cluster_iteration": 0, "vcl_c06": 1, "vcl_c01_balanced": 2}[traversal]
    rc = ref().ref_lj_vcl(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(types), _p(own), _p(_f64(box_min)),
                          _p(_f64(box_max)), ctypes.c_double(cutoff), ctypes.c_double(skin), ctypes.c_int64(cluster_size),
                          ctypes.c_int(_flags(shift, mixing, newton3, soa)), ctypes.c_int(trav), ctypes.c_int(len(eps)),
                          _p(eps), _p(sigma), _p(f), _p(glob), ctypes.byref(flops), ctypes.byref(hit))
    if rc != 0:
        raise RuntimeError("reference VerletClusterLists run failed")
    sizes = np.zeros(5, dtype=np.int64)
    side = np.zeros(2)
    ref().ref_vcl_dump_sizes(_p(sizes), _p(side))
    ncl, npairs, M = int(sizes[0]), int(sizes[1]), int(sizes[4])
    tower_of_particle = np.zeros(max(n, 1), dtype=np.int64)
    cluster_particles = np.zeros(max(ncl * M, 1), dtype=np.int64)
    cluster_tower = np.zeros(max(ncl, 1), dtype=np.int64)
    pairs = np.zeros((max(npairs, 1), 2), dtype=np.int64)
    ref().ref_vcl_dump_copy(_p(tower_of_particle), _p(cluster_particles), _p(cluster_tower), _p(pairs))
    return {"f": f, "upot": glob[0], "virial": glob[1], "flops": flops.value, "hit_rate": hit.value,
            "num_clusters": ncl, "num_pairs": npairs, "towers_per_dim": (int(sizes[2]), int(sizes[3])),
            "tower_side": tuple(side), "tower_of_particle": tower_of_particle[:n],
            "cluster_particles": cluster_particles[:ncl * M].reshape(ncl, M) if ncl else np.zeros((0, M), np.int64),
            "cluster_tower": cluster_tower[:ncl], "pairs": pairs[:npairs]}


# ---- functors other than single-site LJ (oracle/functors_oracle.c, oracle/ref_driver_extra.cpp, _multisite.cpp) -----
_REF_MS = os.path.join(_HERE, "_ref", "libautopas_ref_ms.so")
_ref_ms = None


def have_ref_ms():
    return os.path.exists(_REF_MS)


def ref_ms():
    global _ref_ms
    if _ref_ms is None:
        _ref_ms = ctypes.CDLL(_REF_MS)
    return _ref_ms


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_D = ctypes.c_double


def sph_W(dr2, h):
    fn = lib().orc_sph_W
    fn.restype = _D
    fn.argtypes = [_D, _D]
    return fn(float(dr2), float(h))


def sph_gradW(dr, h):
    out = np.zeros(3)
    lib().orc_sph_gradW(_p(_f64(dr)), _D(float(h)), _p(out))
    return out


def sph_density(pos, mass, smth, own):
    n = len(pos)
    x, y, z = (_f64(pos[:, d]) for d in range(3))
    rho, scale = np.zeros(n), np.zeros(n)
    lib().orc_sph_density(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(_f64(mass)), _p(_f64(smth)), _p(_i64(own)), _p(rho),
                          _p(scale))
    return rho, scale


def sph_hydro(pos, vel, mass, smth, density, pressure, snd, own, with_eng_scale=False):
    """Returns (acc, engDot, vsigmax, scale of acc[, scale of engDot]): the scales are the sums of the magnitudes of the
    per-pair terms, the yardstick for 1e-12 comparisons of sums whose terms cancel."""
    n = len(pos)
    cols = [_f64(pos[:, d]) for d in range(3)] + [_f64(vel[:, d]) for d in range(3)]
    acc, eng, vsig, scale, scale_e = np.zeros(3 * n), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    lib().orc_sph_hydro(ctypes.c_int64(n), *[_p(c) for c in cols], _p(_f64(mass)), _p(_f64(smth)), _p(_f64(density)),
                        _p(_f64(pressure)), _p(_f64(snd)), _p(_i64(own)), None, _p(acc), _p(eng), _p(vsig), _p(scale),
                        _p(scale_e))
    if with_eng_scale:
        return acc.reshape(n, 3), eng, vsig, scale, scale_e
    return acc.reshape(n, 3), eng, vsig, scale


def atm(pos, types, own, cutoff, nu=None, nu_of_type=None):
    """Returns dict(f, scale, upot3_sum, virial_sum[3], kernel_calls); Upot = upot3_sum / 9 (ATM endTraversal)."""
    n = len(pos)
    x, y, z = (_f64(pos[:, d]) for d in range(3))
    t = _i64(np.zeros(n) if types is None else types)
    f, scale, res = np.zeros(3 * n), np.zeros(n), np.zeros(5)
    if nu_of_type is not None:
        v = np.asarray(nu_of_type, dtype=np.float64)
        mix = _f64(np.cbrt(v[:, None, None] * v[None, :, None] * v[None, None, :]).ravel())
        T, mixp, nu0 = len(v), _p(mix), 0.0
    else:
        T, mixp, nu0 = 0, None, float(nu)
    lib().orc_atm(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(t), _p(_i64(own)), _D(float(cutoff)), _D(nu0),
                  ctypes.c_int(T), mixp, _p(f), _p(scale), _p(res))
    return {"f": f.reshape(n, 3), "scale": scale, "upot3_sum": res[0], "virial_sum": res[1:4].copy(),
            "kernel_calls": int(res[4])}


def multisite(pos, quat, mol_type, own, cutoff, shift, eps, sigma, site_start, site_pos, site_type):
    n = len(pos)
    x, y, z = (_f64(pos[:, d]) for d in range(3))
    mix = mixing_table(eps, sigma, cutoff)
    f, tq, scale, res = np.zeros(3 * n), np.zeros(3 * n), np.zeros(n), np.zeros(4)
    keep = (_f64(np.asarray(quat).ravel()), _i64(mol_type), _i64(own), _f64(mix), _i32(site_start),
            _f64(np.asarray(site_pos).ravel()), _i32(site_type))
    lib().orc_multisite(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(keep[0]), _p(keep[1]), _p(keep[2]),
                        _D(float(cutoff)), ctypes.c_int(1 if shift else 0), ctypes.c_int(len(eps)), _p(keep[3]),
                        _p(keep[4]), _p(keep[5]), _p(keep[6]), _p(f), _p(tq), _p(scale), _p(res))
    return {"f": f.reshape(n, 3), "torque": tq.reshape(n, 3), "scale": scale, "upot6_sum": res[0],
            "virial_sum": res[1:4].copy()}


def ref_sph(pos, vel, mass, smth, density, pressure, snd, own, box_min, box_max, cutoff, skin, which, newton3):
    n = len(pos)
    cols = [_f64(pos[:, d]) for d in range(3)] + [_f64(vel[:, d]) for d in range(3)]
    rho, acc, eng, vsig = np.zeros(n), np.zeros(3 * n), np.zeros(n), np.zeros(n)
    rc = ref().ref_sph(ctypes.c_int64(n), *[_p(c) for c in cols], _p(_f64(mass)), _p(_f64(smth)), _p(_f64(density)),
                       _p(_f64(pressure)), _p(_f64(snd)), _p(_i64(own)), _p(_f64(box_min)), _p(_f64(box_max)),
                       _D(float(cutoff)), _D(float(skin)), ctypes.c_int(which), ctypes.c_int(1 if newton3 else 0),
                       _p(rho), _p(acc), _p(eng), _p(vsig))
    if rc != 0:
        raise RuntimeError("reference SPH run failed")
    return {"density": rho, "acc": acc.reshape(n, 3), "engdot": eng, "vsigmax": vsig}


def ref_atm(pos, types, own, box_min, box_max, cutoff, skin, nu=None, nu_of_type=None):
    n = len(pos)
    x, y, z = (_f64(pos[:, d]) for d in range(3))
    t = _i64(np.zeros(n) if types is None else types)
    f, g, flops = np.zeros(3 * n), np.zeros(2), ctypes.c_uint64(0)
    nt = 0 if nu_of_type is None else len(nu_of_type)
    nup = None if nu_of_type is None else _p(_f64(nu_of_type))
    keep = _f64(nu_of_type) if nu_of_type is not None else None
    nup = None if keep is None else _p(keep)
    rc = ref().ref_atm(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(t), _p(_i64(own)), _p(_f64(box_min)),
                       _p(_f64(box_max)), _D(float(cutoff)), _D(float(skin)), _D(0.0 if nu is None else float(nu)),
                       ctypes.c_int(nt), nup, _p(f), _p(g), ctypes.byref(flops))
    if rc != 0:
        raise RuntimeError("reference ATM run failed")
    return {"f": f.reshape(n, 3), "upot": g[0], "virial": g[1], "flops": flops.value}


def ref_multisite(pos, quat, mol_type, own, box_min, box_max, cutoff, skin, shift, newton3, eps, sigma, site_start,
                  site_pos, site_type):
    n = len(pos)
    x, y, z = (_f64(pos[:, d]) for d in range(3))
    f, tq, g = np.zeros(3 * n), np.zeros(3 * n), np.zeros(2)
    keep = (_f64(np.asarray(quat).ravel()), _i64(mol_type), _i64(own), _f64(eps), _f64(sigma), _i32(site_start),
            _f64(np.asarray(site_pos).ravel()), _i32(site_type))
    rc = ref_ms().ref_multisite(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(keep[0]), _p(keep[1]), _p(keep[2]),
                                _p(_f64(box_min)), _p(_f64(box_max)), _D(float(cutoff)), _D(float(skin)),
                                ctypes.c_int(1 if shift else 0), ctypes.c_int(1 if newton3 else 0),
                                ctypes.c_int(len(eps)), _p(keep[3]), _p(keep[4]), ctypes.c_int(len(site_start) - 1),
                                _p(keep[5]), _p(keep[6]), _p(keep[7]), _p(f), _p(tq), _p(g))
    if rc != 0:
        raise RuntimeError("reference multisite run failed")
    return {"f": f.reshape(n, 3), "torque": tq.reshape(n, 3), "upot": g[0], "virial": g[1]}


# ---- md-flexible's MPI wire format -------------------------------------------------------------------------------------
WIRE_RECORD_BYTES = 120
_REF_WIRE = os.path.join(_HERE, "_ref", "libautopas_ref_wire.so")
_WIRE_DTYPE = np.dtype([("id", "<u8"), ("r", "<f8", 3), ("v", "<f8", 3), ("f", "<f8", 3), ("oldf", "<f8", 3), ("type", "<u8"),
                        ("own", "<i8")])


def wire_serialize(ids, r, v, f, oldf, types, own):
    """Restatement of ParticleSerializationTools::serializeParticle for MoleculeLJ
    (examples/md-flexible/src/ParticleSerializationTools.cpp:43-58, 76-82, 104-114): the attributes id, posX..Z, velocityX..Z,
    forceX..Z, oldForceX..Z, typeId, ownershipState are memcpy'd back to back, 8 bytes each = 120 bytes per particle
    (AttributesSize, :66). Returns a uint8 array."""
    n = len(ids)
    rec = np.zeros(n, dtype=_WIRE_DTYPE)
    assert _WIRE_DTYPE.itemsize == WIRE_RECORD_BYTES
    rec["id"], rec["type"], rec["own"] = ids, types, own
    rec["r"], rec["v"], rec["f"], rec["oldf"] = r, v, f, oldf
    return rec.view(np.uint8).reshape(-1).copy()


def wire_deserialize(data):
    """ParticleSerializationTools::deserializeParticles (:132-141): a dict of arrays."""
    rec = np.ascontiguousarray(data, dtype=np.uint8).view(_WIRE_DTYPE)
    return {"id": rec["id"].astype(np.int64), "r": rec["r"].copy(), "v": rec["v"].copy(), "f": rec["f"].copy(),
            "oldf": rec["oldf"].copy(), "type": rec["type"].astype(np.int64), "own": rec["own"].copy()}


def have_ref_wire():
    return os.path.exists(_REF_WIRE)


def ref_wire_serialize(ids, r, v, f, oldf, types, own):
    """The unmodified reference (oracle/_ref/libautopas_ref_wire.so) on the same particles."""
    lib_ = ctypes.CDLL(_REF_WIRE)
    lib_.ref_wire_serialize.restype = ctypes.c_int64
    n = len(ids)
    cols = [_f64(a[:, d]) for a in (r, v, f, oldf) for d in range(3)]
    ptrs = (ctypes.c_void_p * 12)(*[c.ctypes.data for c in cols])
    out = np.zeros(n * WIRE_RECORD_BYTES, dtype=np.uint8)
    ids_, types_, own_ = _i64(ids), _i64(types), _i64(own)
    nb = lib_.ref_wire_serialize(ctypes.c_int64(n), ptrs, _p(ids_), _p(types_), _p(own_), _p(out))
    assert nb == n * WIRE_RECORD_BYTES
    return out


def ref_wire_deserialize(data):
    lib_ = ctypes.CDLL(_REF_WIRE)
    lib_.ref_wire_deserialize.restype = ctypes.c_int64
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = len(data) // WIRE_RECORD_BYTES
    cols = [np.zeros(n) for _ in range(12)]
    ptrs = (ctypes.c_void_p * 12)(*[c.ctypes.data for c in cols])
    ids, types, own = (np.zeros(n, dtype=np.int64) for _ in range(3))
    m = lib_.ref_wire_deserialize(ctypes.c_int64(len(data)), _p(data), ptrs, _p(ids), _p(types), _p(own))
    assert m == n
    c = np.stack(cols, axis=1)
    return {"id": ids, "r": c[:, 0:3], "v": c[:, 3:6], "f": c[:, 6:9], "oldf": c[:, 9:12], "type": types, "own": own}


# ---- md-flexible's VTK checkpoint record --------------------------------------------------------------------------------
_REF_VTK = os.path.join(_HERE, "_ref", "vtk_ref_writer")


def vtk_particle_record(ids, r, v, f, types, box_max):
    """oracle/vtk_oracle.c: the bytes of one rank's `<session>_Particles_<rank>_<iteration>.vtu` piece for the particles
    in the given order (ParallelVtkWriter.cpp:55-201). Raises ValueError where the reference would throw."""
    fn = lib().vtk_oracle_particle_record
    fn.restype = ctypes.c_int64
    n = len(ids)
    r_, v_, f_, ids_, types_, bm = _f64(r).reshape(-1), _f64(v).reshape(-1), _f64(f).reshape(-1), _i64(ids), _i64(types), _f64(box_max)
    args = (ctypes.c_int64(n), _p(r_), _p(v_), _p(f_), _p(ids_), _p(types_), _p(bm))
    need = fn(*args, None, ctypes.c_int64(0))
    if need < 0:
        raise ValueError("a position cannot be told from the box border within 15 digits")
    out = np.zeros(need, dtype=np.uint8)
    got = fn(*args, _p(out), ctypes.c_int64(need))
    assert got == need
    return out


def vtk_position_precision(position, border):
    fn = lib().vtk_oracle_position_precision
    fn.restype = ctypes.c_int
    return fn(ctypes.c_double(position), ctypes.c_double(border))


def vtk_pvtu_record(session, num_ranks, iteration, digits):
    """The `.pvtu` index of rank 0 (ParallelVtkWriter.cpp:308-356)."""
    fn = lib().vtk_oracle_pvtu_record
    fn.restype = ctypes.c_int64
    args = (session.encode(), ctypes.c_int(num_ranks), ctypes.c_uint64(iteration), ctypes.c_int(digits))
    need = fn(*args, None, ctypes.c_int64(0))
    out = np.zeros(need, dtype=np.uint8)
    fn(*args, _p(out), ctypes.c_int64(need))
    return out


def have_ref_vtk():
    return os.path.exists(_REF_VTK)


ref_vtk_seconds = {}  # seconds the unmodified writer / loader took in the last ref_vtk_records / ref_vtk_load call


def ref_vtk_records(ids, r, v, f, types, box_min, box_max, session="ref", iteration=7, digits=6):
    """The unmodified ParallelVtkWriter (oracle/_ref/vtk_ref_writer, oracle/ref_driver_vtk.cpp) on a stock
    AutoPas<MoleculeLJ> holding these particles -> (bytes of the .vtu piece, bytes of the .pvtu index). The piece lists
    the particles in the iteration order of the reference's container; its `ids` array tells which."""
    import tempfile
    n = len(ids)
    rec = np.zeros(n, dtype=[("r", "<f8", 3), ("v", "<f8", 3), ("f", "<f8", 3), ("id", "<i8"), ("type", "<i8")])
    rec["r"], rec["v"], rec["f"], rec["id"], rec["type"] = r, v, f, ids, types
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in.bin"), "wb") as fh:
            fh.write(np.int64(n).tobytes())
            fh.write(_f64(box_min).tobytes())
            fh.write(_f64(box_max).tobytes())
            fh.write(rec.tobytes())
        os.makedirs(os.path.join(d, "out"))
        done = subprocess.run([_REF_VTK, os.path.join(d, "in.bin"), os.path.join(d, "out"), session, str(iteration), str(digits)],
                              check=True, capture_output=True, text=True, cwd=d)
        ref_vtk_seconds["write"] = float(done.stdout.split()[1])  # the writer alone, without filling the container
        it = str(iteration).zfill(digits)
        piece = np.fromfile(os.path.join(d, "out", session, "data", f"{session}_Particles_0_{it}.vtu"), dtype=np.uint8)
        index = np.fromfile(os.path.join(d, "out", session, f"{session}_Particles_{it}.pvtu"), dtype=np.uint8)
    return piece, index


def vtk_load_particle_record(piece):
    """Restatement of md-flexible's checkpoint loader for one piece (loadParticlesFromRankRecord,
    examples/md-flexible/src/configuration/MDFlexConfig.cpp:91-180): NumberOfPoints, then for velocities, forces, typeIds,
    ids, positions: go behind the word (findWord :54-65), skip the rest of that line, read NumberOfPoints x {3, 3, 1, 1, 3}
    whitespace-separated values with operator>> (= strtod / strtoul: float() and int() round the same way).
    Returns dict(ids, r, v, f, types). Pinned by particles of the unmodified loader: tests/golden/fn_vtk_load.npz, live."""
    text = bytes(piece).decode()
    at = text.index("NumberOfPoints")
    at = text.index('"', at) + 1
    n = int(text[at:text.index('"', at)])
    if n == 0:
        raise ValueError("Could not determine the number of particles in the checkpoint file")
    out = {}
    for word, key, k, conv in (("velocities", "v", 3, float), ("forces", "f", 3, float), ("typeIds", "types", 1, int),
                               ("ids", "ids", 1, int), ("positions", "r", 3, float)):
        at = text.index('"' + word + '"', at)  # (findWord compares whole words between separators; '"' is one)
        at = text.index("\n", at) + 1
        end = text.index("<", at)
        tokens = text[at:end].split()[: n * k]
        if len(tokens) < n * k:
            raise ValueError(f"data array {word} holds fewer values than NumberOfPoints asks for")
        a = np.array([conv(t) for t in tokens], dtype=np.float64 if conv is float else np.int64)
        out[key] = a.reshape(n, 3) if k == 3 else a
    return out


def ref_vtk_load(pvtu_path, rank=0, num_ranks=1):
    """The unmodified loader (MDFlexConfig::loadParticlesFromCheckpoint through oracle/_ref/vtk_ref_writer --load) on a
    checkpoint on disk -> dict(ids, r, v, f, types) in the order it produced the particles."""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "particles.bin")
        done = subprocess.run([_REF_VTK, "--load", str(pvtu_path), str(rank), str(num_ranks), out], check=True, capture_output=True, text=True)
        ref_vtk_seconds["load"] = float(done.stdout.split()[1])
        raw = np.fromfile(out, dtype=np.uint8)
    n = int(raw[:8].view(np.int64)[0])
    rec = raw[8:].view(np.dtype([("r", "<f8", 3), ("v", "<f8", 3), ("f", "<f8", 3), ("id", "<i8"), ("type", "<i8")]))
    assert len(rec) == n
    return {"ids": rec["id"].copy(), "r": rec["r"].copy(), "v": rec["v"].copy(), "f": rec["f"].copy(), "types": rec["type"].copy()}


def vtk_parse_ids(piece):
    """The `ids` DataArray of a .vtu piece as the checkpoint loader reads it (MDFlexConfig.cpp:158-160)."""
    text = bytes(piece).decode()
    body = text.split('Name="ids"', 1)[1].split(">\n", 1)[1].split("</DataArray>", 1)[0]
    return np.array([int(t) for t in body.split()], dtype=np.int64)


def ref_set_timing_reps(reps):
    """The extra-functor drivers (ref_sph, ref_atm) time their traversal; with reps > 1 they repeat it (outputs then only
    serve timing)."""
    ref().ref_set_timing_reps(ctypes.c_int(int(reps)))


def ref_last_compute_seconds():
    f = ref().ref_last_compute_seconds
    f.restype = ctypes.c_double
    return float(f())
