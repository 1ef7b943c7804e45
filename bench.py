#!/usr/bin/env python
"""bench.py — MFUPs/s of the LJ fp64 force step (BASELINE.json metric) on N B200s of one node.

A "step" is one iteration of the md-flexible simulation loop restricted to the hot path (Simulation.cpp:230-351):
positions -> [every `rebuild` steps: migration, halo exchange, neighbour-structure rebuild | halo position refresh]
-> LJ force kernel (shift + globals + FLOP counters) -> velocities.

Workloads:
  c3 (default, the north_star target)  BASELINE configs[2]: 252^3 = 16 003 008-particle spinodal-decomposition box
      (spacing 1.5, T = 1.4, skin 0.5), the TOTAL fixed and split over the ranks by the regular-grid decomposition
      = strong scaling; one GPU holds all of it (about 6 GB).
  c2  BASELINE configs[1]: 1M-particle LJ liquid (rho* = 0.8442, jittered 100^3 lattice, skin 0.3) per GPU = weak
      scaling. The default run measures it as well and reports it under the key "c2" of the same JSON line.

`value`        device-resident loop (apb_run_steps), CUDA events on the library's stream, max over ranks
`e2e`          the same step through the C ABI with HOST buffers: positions uploaded and forces downloaded every step
`roofline`     the dominant kernel (LJ force) against the FP64 DFMA peak measured live (MEASURED_PEAKS.json has no FP64)
`cpu_baseline` the unmodified reference (oracle/_ref, OpenMP on all host cores, LJFunctor and LJFunctorHWY) on a bounded
               sample of the same workload
--impl reference prints the reference arm in the same JSON shape.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CUTOFF, REBUILD = 2.5, 10
METRIC = "MFUPs/s LJ fp64 force step"
# BASELINE.json configs[1]: 100^3 simple-cubic lattice at rho* = 0.8442, every coordinate jittered by
# U(-0.15, 0.15) from std::mt19937(42); skin 0.3
C2 = {"n_per_dim": 100, "rho": 0.8442, "jitter": 0.15, "skin": 0.3, "dt": 0.002, "temperature": 1.0}
# BASELINE.json configs[2]: spinodal-decomposition start, 252^3 simple-cubic lattice at spacing 1.5,
# Maxwell-Boltzmann velocities at T = 1.4, skin 0.5, deltaT 0.00182367; the TOTAL is fixed (strong scaling)
C3 = {"n_per_dim": 252, "spacing": 1.5, "skin": 0.5, "dt": 0.00182367, "temperature": 1.4}
# the CPU arm runs a bounded sample: a periodic sub-box of the same lattice / density (MFUPs/s is per particle)
CPU_SAMPLE_N_PER_DIM = {"c2": 100, "c3": 126}


def decomposition(n):
    """DomainTools::generateDecomposition (examples/md-flexible/src/domainDecomposition/DomainTools.cpp:26-71):
    prime factors folded into three dimensions."""
    dims = [1, 1, 1]
    f, m, primes = 2, n, []
    while m > 1:
        while m % f == 0:
            primes.append(f)
            m //= f
        f += 1
    for i, p in enumerate(sorted(primes, reverse=True)):
        dims[i % 3] *= p
    return dims


def rank_coords(rank, dims):
    return [rank % dims[0], (rank // dims[0]) % dims[1], rank // (dims[0] * dims[1])]


def coords_rank(c, dims):
    return (c[2] % dims[2] * dims[1] + c[1] % dims[1]) * dims[0] + c[0] % dims[0]


def mt19937_uniform(seed, count, lo, hi):
    """`count` draws of std::uniform_real_distribution<double>(lo, hi) on std::mt19937(seed), bit for bit (libstdc++:
    generate_canonical<double, 53> takes two 32-bit outputs, (d1 + d2 * 2^32) / 2^64, then scales). numpy's legacy
    RandomState seeds MT19937 with the same init_genrand as the C++ engine."""
    raw = np.random.RandomState(seed)._bit_generator.random_raw(2 * count).astype(np.float64)
    canon = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    canon = np.minimum(canon, np.nextafter(1.0, 0.0))
    return (hi - lo) * canon + lo


def make_workload(workload, n_per_dim, rank, dims, seed=42):
    """Positions, velocities, local box and global box of one rank.
    c2: n_per_dim^3 jittered lattice per rank (weak scaling); c3: the n_per_dim^3 lattice split over the ranks."""
    c = np.array(rank_coords(rank, dims), dtype=float)
    rng = np.random.default_rng(seed + rank)
    if workload == "c3":
        spacing = C3["spacing"]
        counts = [n_per_dim // d for d in dims]
        if any(n_per_dim % d for d in dims):
            raise SystemExit(f"c3: {n_per_dim} lattice points per dimension do not split over {dims}")
        Ls = np.array(counts, dtype=float) * spacing
        lo = c * Ls
        gs = [(np.arange(k) + 0.5) * spacing for k in counts]
        zz, yy, xx = np.meshgrid(gs[2], gs[1], gs[0], indexing="ij")
        pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1) + lo
        del xx, yy, zz
        vel = rng.normal(0.0, np.sqrt(C3["temperature"]), pos.shape)
        return pos, vel, lo, lo + Ls, np.zeros(3), np.array(dims, dtype=float) * Ls
    spacing = C2["rho"] ** (-1.0 / 3.0)
    L = n_per_dim * spacing
    lo = c * L
    g = (np.arange(n_per_dim) + 0.5) * spacing
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1) + lo
    pos += mt19937_uniform(seed + rank, pos.size, -C2["jitter"], C2["jitter"]).reshape(pos.shape)
    vel = rng.normal(0.0, np.sqrt(C2["temperature"]), pos.shape)
    vel -= vel.mean(axis=0)
    return pos, vel, lo, lo + L, np.zeros(3), np.array(dims, dtype=float) * L


def workload_text(workload, n_per_dim):
    if workload == "c3":
        return (f"C3 spinodal-decomposition LJ box, {n_per_dim ** 3} particles in total ({n_per_dim}^3 lattice, "
                f"spacing 1.5, T=1.4), cutoff 2.5, skin 0.5, rebuild every 10 steps, periodic, split over the GPUs")
    return (f"C2 LJ liquid rho*=0.8442, {n_per_dim ** 3} particles per GPU ({n_per_dim}^3 lattice jittered by "
            f"U(-0.15,0.15) mt19937(42)), cutoff 2.5, skin 0.3, rebuild every 10 steps, periodic")


def periodic_images(pos, box_min, box_max, width):
    L = box_max - box_min
    out = []
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for c in (-1, 0, 1):
                if (a, b, c) == (0, 0, 0):
                    continue
                p = pos + np.array([a, b, c]) * L
                m = np.all((p >= box_min - width) & (p < box_max + width), axis=1)
                out.append(p[m])
    return np.vstack(out)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        if self.index is None:
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if parts[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def checkpoint_cpu_reference(n=250_000):
    """cpu_baseline leg for the checkpoint entry: the UNMODIFIED md-flexible writer and loader (oracle/_ref/vtk_ref_writer:
    ParallelVtkWriter::recordParticleStates and MDFlexConfig::loadParticlesFromCheckpoint, single-threaded like in
    md-flexible) on a bounded sample of the same kind of box; seconds of the writer / loader alone, scaled to a million
    particles. None when oracle/_ref did not travel."""
    import tempfile
    import oracle
    if not oracle.have_ref_vtk():
        return None
    rng = np.random.default_rng(3)
    L = float(np.ceil((n / 0.8442) ** (1 / 3)))
    piece, index = oracle.ref_vtk_records(np.arange(n), rng.uniform(0, L, (n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3)),
                                          np.zeros(n, dtype=np.int64), [0, 0, 0], [L, L, L], "bench", 0, 6)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "bench", "data"))
        piece.tofile(os.path.join(d, "bench", "data", "bench_Particles_0_000000.vtu"))
        index.tofile(os.path.join(d, "bench", "bench_Particles_000000.pvtu"))
        oracle.ref_vtk_load(os.path.join(d, "bench", "bench_Particles_000000.pvtu"))
    return {"kind": "reference", "cores": 1, "sample": f"{n} particles, unmodified ParallelVtkWriter / MDFlexConfig loader",
            "ms_write_per_million_particles": oracle.ref_vtk_seconds["write"] * 1e3 * 1e6 / n,
            "ms_load_per_million_particles": oracle.ref_vtk_seconds["load"] * 1e3 * 1e6 / n}


def other_configs_cpu_reference(lines, bf):
    """cpu_baseline leg for the side measurements: the unmodified reference (oracle/_ref/libautopas_ref.so: LinkedCells,
    lc_c08 for SPH / lc_c01 for the three-body functor, AoS, newton3 off, OpenMP on all host cores) on the inputs of
    tools/bench_functors.py C4 / C5, traversal time only. Skipped when oracle/_ref did not travel."""
    import oracle
    if not oracle.have_ref():
        return
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    what = f"unmodified reference, LinkedCells, AoS, newton3 off, OpenMP {cores} threads, traversal only"
    try:
        oracle.ref_set_timing_reps(1)
        pos, L = bf.lattice(64, 1.2, 0.1, 4)
        halo = bf.images(pos, L, 2.7)
        allpos = np.vstack([pos, halo])
        own = np.r_[np.ones(len(pos)), 2 * np.ones(len(halo))].astype(np.int64)
        oracle.ref_atm(allpos, None, own, [0, 0, 0], [L, L, L], 2.5, 0.2, nu=0.073)
        ref_ms = {"C4": oracle.ref_last_compute_seconds() * 1e3}
        d = 0.4
        h = 1.2 * d
        cutoff = 2.5 * h
        pos, L = bf.lattice(128, d, 0.05, 5)
        halo = bf.images(pos, L, 1.1 * cutoff)
        allpos = np.vstack([pos, halo])
        n = len(allpos)
        own = np.r_[np.ones(len(pos)), 2 * np.ones(len(halo))].astype(np.int64)
        rng = np.random.default_rng(0)
        vel = rng.normal(0, 0.1, (n, 3))
        args_ = (allpos, vel, np.full(n, d ** 3), np.full(n, h), np.full(n, 1.0), np.full(n, 1.0), np.full(n, 1.2), own,
                 [0, 0, 0], [L, L, L], cutoff, 0.1 * cutoff)
        oracle.ref_sph(*args_, 0, False)
        ref_ms["C5 SPH density"] = oracle.ref_last_compute_seconds() * 1e3
        oracle.ref_sph(*args_, 1, False)
        ref_ms["C5 SPH hydro"] = oracle.ref_last_compute_seconds() * 1e3
        for ln in lines:
            for key, ms in ref_ms.items():
                if ln["config"].startswith(key):
                    ln["cpu_reference"] = {"ms_per_call": ms, "cores": cores, "kind": "reference", "sample": what}
    except Exception as exc:  # the side measurement must not cost the headline line
        lines.append({"cpu_reference_error": repr(exc)})


def reference_arm(workload, iters, warmup, n_per_dim=0):
    """The unmodified reference (oracle/_ref, built like its own Release build: -O3, the host's vector ISA) on every host
    core this process may use, on a bounded sample of the workload: a periodic sub-box of the same lattice with its
    periodic halo images, rebuild every 10 iterations, SoA, newton3, shift + globals. Both LJ kernels md-flexible offers
    are timed - mdLib::LJFunctor (auto-vectorised) and mdLib::LJFunctorHWY (Highway) - on LinkedCells / lc_c08 and on
    VerletClusterLists / vcl_c06 (cluster size 4, the reference default); the fastest is `value`, as the AutoTuner would
    pick it, and all are listed."""
    import oracle
    if not oracle.have_refbench():
        return None
    npd = min(n_per_dim, CPU_SAMPLE_N_PER_DIM[workload]) if n_per_dim else CPU_SAMPLE_N_PER_DIM[workload]
    skin = C3["skin"] if workload == "c3" else C2["skin"]
    pos, _, bmin, bmax, _, _ = make_workload(workload, npd, 0, [1, 1, 1])
    halo = periodic_images(pos, bmin, bmax, CUTOFF + skin)
    allpos = np.vstack([pos, halo])
    own = np.r_[np.ones(len(pos)), 2 * np.ones(len(halo))].astype(np.int64)
    threads = len(os.sched_getaffinity(0))  # not OMP_NUM_THREADS: torchrun sets it to 1 for its workers
    runs = []
    for functor in ("LJFunctor", "LJFunctorHWY"):
        for container, traversal in (("LinkedCells", "lc_c08"), ("VerletClusterLists", "vcl_c06")):
            r = oracle.refbench_lj(allpos[:, 0], allpos[:, 1], allpos[:, 2], own, bmin, bmax, CUTOFF, skin,
                                   functor=functor, container=container, traversal=traversal, cluster_size=4,
                                   newton3=True, warmup=warmup, iters=iters, rebuild_freq=REBUILD, threads=threads)
            total = r["rebuild_s"] + r["compute_s"]
            runs.append({"functor": functor, "container": container, "traversal": traversal,
                         "value": len(pos) * iters / total * 1e-6, "seconds": total, "rebuild_s": r["rebuild_s"],
                         "compute_s": r["compute_s"], "threads": r["threads"], "num_rebuilds": r["num_rebuilds"],
                         "isa": r["isa"], "upot": r["upot"]})
    best = max(runs, key=lambda q: q["value"])
    best_autovec = max((q for q in runs if q["functor"] == "LJFunctor"), key=lambda q: q["value"])
    return {"value": best["value"], "unit": "MFUPs/s", "cores": best["threads"], "kind": "reference",
            "iters": iters, "seconds": best["seconds"], "rebuild_s": best["rebuild_s"], "compute_s": best["compute_s"],
            "functor": best["functor"], "container": best["container"], "traversal": best["traversal"],
            "isa": best["isa"], "value_ljfunctor_autovec": best_autovec["value"],
            "configurations": [{k: q[k] for k in ("functor", "container", "traversal", "value")} for q in runs],
            "particles": len(pos),
            "sample": f"{iters} force iterations + {best['num_rebuilds']} rebuild(s) ({warmup} warm-up) of a periodic "
                      f"{npd}^3 = {len(pos)}-particle box of the {workload.upper()} workload (same lattice, density, "
                      f"cutoff, skin); unmodified reference, -O3 {best['isa']}, OpenMP {best['threads']} threads, SoA, "
                      f"newton3, shift + globals; fastest of LJFunctor / LJFunctorHWY x LinkedCells lc_c08 / "
                      f"VerletClusterLists vcl_c06: {best['functor']} on {best['container']}/{best['traversal']}"}


def measure(workload, n_per_dim, args, ctx, with_e2e=True):
    """One workload on this process's GPU: device-resident timed loop, roofline of the force kernel, end-to-end leg."""
    import torch

    from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor, capi

    rank, world, local_rank, dims = ctx["rank"], ctx["world"], ctx["local_rank"], ctx["dims"]
    barrier, max_over_ranks, sum_over_ranks = ctx["barrier"], ctx["max"], ctx["sum"]
    skin, dt = (C3["skin"], C3["dt"]) if workload == "c3" else (C2["skin"], C2["dt"])
    pos, vel, bmin, bmax, gmin, gmax = make_workload(workload, n_per_dim, rank, dims)
    n = len(pos)
    c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, CUTOFF, skin, clusterSize=args.cluster_size,
                             device=local_rank)
    if world > 1:
        c.commInit(world, rank, ctx["nccl_id"]())
        me = rank_coords(rank, dims)
        nb = []
        for d in range(3):
            lo, hi = list(me), list(me)
            lo[d] -= 1
            hi[d] += 1
            nb += [coords_rank(lo, dims), coords_rank(hi, dims)]
        c.setDecomposition(gmin, gmax, nb, (1, 1, 1))
    # ParticleContainerInterface::reserve: owned particles + an estimate of the halo shell
    il = CUTOFF + skin
    halo_est = int(n * (np.prod((bmax - bmin + 2 * il) / (bmax - bmin)) - 1.0) * 1.1) + 1024
    c.reserve(n, halo_est)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n) + rank * n)
    for d, name in enumerate(("VX", "VY", "VZ")):
        c.uploadColumn(name, vel[:, d])
    del pos, vel
    functor = LJFunctor(CUTOFF, applyShift=True, calculateGlobals=True, countFLOPs=True,
                        virialTraceOnly=not args.virial_components)
    functor.setParticleProperties(24.0, 1.0)
    trav = GpuTraversal(args.traversal, functor, bool(args.newton3))
    mass = [1.0]

    stream = torch.cuda.ExternalStream(c.getStream(), device=torch.device("cuda", local_rank))
    steps = max(1, args.steps)
    warm = max(3, args.warmup)
    # any window of K consecutive iterations holds K / 10 rebuilds (iteration % 10 == 0) when K is a multiple of 10,
    # whatever its phase: the warm-up does not have to be rounded to a rebuild period
    c.runSteps(trav, warm, 0, dt, mass, REBUILD, wantResults=False)

    # ---- device-resident timed region ----
    c.enableLoopTiming(True)
    c.getLoopTiming()
    sampler = ClockSampler(local_rank if rank == 0 else None)  # one nvidia-smi poller per job, not per rank
    launches0, allocs0 = c.getLaunchCount(), c.getAllocCount()
    # the poller is spawned BEFORE the barrier: forking nvidia-smi takes rank 0 several milliseconds, and a rank that starts
    # late stalls its neighbours inside their timed region (their exchanges wait for it)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    res = c.runSteps(trav, steps, warm, dt, mass, REBUILD)
    e1.record(stream)
    e1.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    launches = c.getLaunchCount() - launches0
    allocs = c.getAllocCount() - allocs0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    timing = c.getLoopTiming()
    c.enableLoopTiming(False)
    owned_total = sum_over_ranks(float(c.getNumberOfParticles("owned")))
    value = owned_total * steps / (ms_total * 1e-3) * 1e-6
    rebuilds = sum(1 for it in range(warm, warm + steps) if it % REBUILD == 0)

    # ---- roofline of the dominant kernel (LJ force): reference FLOP model on the kernel's own counters ----
    force_ms, force_launches = timing["force"]
    flops_per_step = np.mean([8 * r.num_dist_calls + 15 * r.num_kernel_calls_no_n3 + 18 * r.num_kernel_calls_n3 +
                              9 * r.num_global_calcs_no_n3 + 13 * r.num_global_calcs_n3 for r in res])
    dist_per_step = np.mean([r.num_dist_calls for r in res])
    hit_rate = np.mean([(r.num_kernel_calls_no_n3 + r.num_kernel_calls_n3) / max(r.num_dist_calls, 1) for r in res])
    kernel_ms = force_ms / max(steps, 1)  # force phase per step (a split step launches the kernel twice)
    achieved_tf = flops_per_step / (kernel_ms * 1e-3) / 1e12
    peak_tf = ctx["fp64_peak"]()
    kernel_name = ("kLJPrunedN3" if args.newton3 else "kLJPruned") if args.traversal == "gpuvcl_pruned" else "kLJClusterPairs"
    # bound: the FP64 FMA pipe (north_star asks for the fraction of the B200 FP64 peak; the kernel is neither HBM- nor
    # tensor-bound: ~1.4 TB/s of DRAM traffic, and the path is not a contraction)
    n_local = c.getNumberOfParticles("owned")
    roofline = {"bound": "fp64", "kernel": kernel_name, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": None,
                "kernel_ms": kernel_ms, "flops_per_launch": flops_per_step, "hit_rate": hit_rate,
                "peak_source": "FP64 DFMA microbenchmark measured live by bench.py (apb_measure_fp64_peak); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "share_of_step": force_ms / ms_total,
                "algorithmic_bytes_per_launch": int(n_local * 48 + dist_per_step * 2)}
    try:  # DRAM bytes per launch of this kernel on this workload, from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "kLJPruned_traffic.json")))
        key = f"{workload}_n{world}_{args.traversal}"
        if key in tr:
            roofline["traffic"] = tr[key]["dram_bytes_per_launch"]
            roofline["traffic_source"] = tr[key].get("source")
    except (OSError, ValueError, KeyError):
        pass
    try:
        roofline["hbm_peak_gbs_measured"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except (OSError, ValueError):
        roofline["hbm_peak_gbs_measured"] = None
    phases = {k: v[0] / steps for k, v in timing.items()}
    phases["rebuild_ms_per_rebuild"] = timing["rebuild"][0] / max(rebuilds, 1)

    out = {"value": value, "steps": steps, "warmup": warm, "ms_per_step": ms_total / steps, "roofline": roofline,
           "phases_ms_per_step": phases, "gpu_launches": int(launches), "device_allocations_in_timed_region": int(allocs),
           "rebuilds_in_timed_region": rebuilds, "clocks": clocks, "particles_total": int(owned_total),
           "host_wall_ms_per_step": wall / steps * 1e3, "upot_last": res[steps - 1].upot_sum * 0.5 / 6.0}
    g = c.getTraversalSelectorInfo()
    out["num_clusters"], out["num_cluster_pairs"] = int(g.num_clusters), int(g.num_cluster_pairs)

    # ---- end to end through the C ABI with host buffers ----
    # The host owns x, y, z and fx, fy, fz indexed by particle id (pinned). Every step: positions host -> device, [every
    # 10th: migration, halo exchange, rebuild | halo refresh], force kernel, forces device -> host, Upot / virial read back.
    if with_e2e:
        e2e_steps = max(1, args.e2e_steps)
        host = {k: torch.empty(n, dtype=torch.float64).pin_memory().numpy() for k in ("x", "y", "z", "fx", "fy", "fz")}
        c.migrate()  # positions back into the (periodic) box before the host takes its copy
        c.exchangeHalos()
        c.rebuildNeighborLists(trav)
        ids_s, _, own_s = c.downloadIds()
        m_owned = own_s == capi.OWN_OWNED
        k_owned = ids_s[m_owned] - rank * n
        by_id = bool(((k_owned >= 0) & (k_owned < n)).all()) and len(k_owned) == n
        by_id = bool(ctx["min"](1.0 if by_id else 0.0))
        h2d = d2h = 0
        if by_id:
            for d, col in enumerate(("X", "Y", "Z")):
                host["xyz"[d]][k_owned] = c.downloadColumn(col)[m_owned]
        else:
            # particles migrated between ranks during the device-resident run: transfers in storage order instead
            cap = int(c.numSlots() * 1.3) + 4096
            host = {k: torch.empty(cap, dtype=torch.float64).pin_memory().numpy() for k in ("x", "y", "z", "fx", "fy", "fz")}

            def pull_positions():
                lib = capi.load()
                for k, col in (("x", "X"), ("y", "Y"), ("z", "Z")):
                    lib.apb_download_column(c._h, capi.COL[col], host[k].ctypes.data)
                return c.numSlots()

            ns = pull_positions()
        del ids_s, own_s, m_owned
        barrier()
        t0 = time.perf_counter()
        for it in range(e2e_steps):
            if by_id:
                # one C-ABI call per step: apb_force_step_by_id (upload, [rebuild chain | halo refresh], forces, download)
                functor.initTraversal()
                raw = c.forceStepById(trav, host["x"], host["y"], host["z"], host["fx"], host["fy"], host["fz"],
                                      rebuild=it % REBUILD == 0, idBegin=rank * n)
                functor.endTraversal(bool(args.newton3))
                h2d += 3 * 8 * n
                d2h += 3 * 8 * n + ctypes.sizeof(raw)
                continue
            c.uploadPositions(host["x"][:ns], host["y"][:ns], host["z"][:ns])
            h2d += 3 * 8 * ns
            if it % REBUILD == 0:
                c.migrate()
                c.exchangeHalos()
                c.rebuildNeighborLists(trav)
                ns = pull_positions()
                d2h += 3 * 8 * ns
            else:
                c.exchangeHalos()
            c.resetForces()
            functor.initTraversal()
            raw = c.computeInteractions(trav)
            functor.endTraversal(bool(args.newton3))
            c.downloadForces(host["fx"][:ns], host["fy"][:ns], host["fz"][:ns])
            d2h += 3 * 8 * ns + ctypes.sizeof(raw)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        out["e2e"] = {"value": owned_total * e2e_steps / e2e_s * 1e-6, "unit": "MFUPs/s",
                      "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                      "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                      "rebuilds": sum(1 for it in range(e2e_steps) if it % REBUILD == 0),
                      "upot_last": functor.getPotentialEnergy(),
                      "note": ("per rank: positions host->device and forces device->host every step, pinned host arrays "
                               "indexed by particle id, one apb_force_step_by_id call per step, Upot/virial read back"
                               if by_id else
                               "per rank: positions host->device and forces device->host every step in storage order "
                               "(pinned), Upot/virial read back")}
        del host
    c.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3"],
                    help="c3 (default): BASELINE configs[2], the 16M-particle spinodal box split over the GPUs, strong "
                         "scaling, with the c2 line added under the key 'c2'; c2: configs[1] only, 1M-particle LJ "
                         "liquid per GPU, weak scaling")
    ap.add_argument("--n-per-dim", type=int, default=0, help="lattice points per dimension (c2: per rank; c3: in total)")
    ap.add_argument("--no-c2", action="store_true", help="skip the additional c2 measurement of the default run")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the C1 / C4 / C5 functor timings the default one-GPU run appends under 'other_configs'")
    ap.add_argument("--cluster-size", type=int, default=32)
    ap.add_argument("--traversal", default="gpuvcl_pruned")
    ap.add_argument("--newton3", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--virial-components", action="store_true",
                    help="accumulate the virial per component instead of its sum (LJFunctor::getVirial returns the sum)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    dims = decomposition(world)
    npd = args.n_per_dim or (C3["n_per_dim"] if args.workload == "c3" else C2["n_per_dim"])
    scaling = "strong" if args.workload == "c3" else "weak"

    if args.impl == "reference":
        if rank != 0:
            return
        iters = max(1, args.steps)
        ref = reference_arm(args.workload, iters, max(1, min(args.warmup, 3)), args.n_per_dim)
        if ref is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libautopas_refbench_*.so was not built"}),
                  file=json_out, flush=True)
            return
        line = {"impl": "reference", "metric": METRIC, "value": ref["value"], "unit": "MFUPs/s", "n_gpus": n_gpus,
                "steps": iters, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ref["seconds"] / iters * 1e3,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(args.workload, npd), "container": ref["container"],
                           "traversal": ref["traversal"], "functor": ref["functor"], "newton3": True, "cluster_size": 4,
                           "host_threads": ref["cores"], "isa": ref["isa"], "sample_particles": ref["particles"],
                           "value_ljfunctor_autovec": ref["value_ljfunctor_autovec"],
                           "configurations_timed": ref["configurations"]},
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "MFUPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist

    from autopas_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_over_ranks(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def nccl_id():
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            if capi.load().apb_comm_get_unique_id(raw) != 0:
                raise SystemExit("apb_comm_get_unique_id failed")
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        return bytes(idbuf.cpu().numpy().tobytes())

    peak_cache = {}

    def fp64_peak():
        if "v" not in peak_cache:
            tf, ms = ctypes.c_double(), ctypes.c_double()
            capi.load().apb_measure_fp64_peak(local_rank, 3, ctypes.byref(tf), ctypes.byref(ms))
            peak_cache["v"] = tf.value
        return peak_cache["v"]

    ctx = {"rank": rank, "world": world, "local_rank": local_rank, "dims": dims, "barrier": barrier,
           "max": lambda v: reduce_over_ranks(v, dist.ReduceOp.MAX), "sum": lambda v: reduce_over_ranks(v, dist.ReduceOp.SUM),
           "min": lambda v: reduce_over_ranks(v, dist.ReduceOp.MIN), "nccl_id": nccl_id, "fp64_peak": fp64_peak}

    main_res = measure(args.workload, npd, args, ctx)
    c2_res = None
    if args.workload == "c3" and not args.no_c2:
        c2_res = measure("c2", C2["n_per_dim"], args, ctx)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_arm(args.workload, REBUILD, 1, args.n_per_dim)
            if ref is not None:
                cpu = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample", "value_ljfunctor_autovec",
                                           "configurations")}
        except Exception as exc:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "MFUPs/s", "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}

    if rank == 0:
        def config_of(workload, n_per_dim, r):
            return {"workload": workload_text(workload, n_per_dim), "container": "gpuVerletClusterLists",
                    "traversal": args.traversal, "newton3": bool(args.newton3), "cluster_size": args.cluster_size,
                    "functor": "LJFunctor shift+globals+flop counters" + (", virial per component" if args.virial_components else ""),
                    "decomposition": dims, "particles_total": r["particles_total"], "num_clusters": r["num_clusters"],
                    "num_cluster_pairs": r["num_cluster_pairs"],
                    "l2_policy": "inputs larger than L2: per-particle lists + SoA columns streamed every step exceed "
                                 "the 126 MB L2",
                    "host_wall_ms_per_step": r["host_wall_ms_per_step"],
                    "rebuilds_in_timed_region": r["rebuilds_in_timed_region"],
                    "device_allocations_in_timed_region": r["device_allocations_in_timed_region"]}

        r = main_res
        rl = dict(r["roofline"])
        rl["phases_ms_per_step"] = r["phases_ms_per_step"]
        line = {"metric": METRIC, "value": r["value"], "unit": "MFUPs/s", "n_gpus": n_gpus, "steps": r["steps"],
                "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args.workload, npd, r),
                "roofline": rl, "cpu_baseline": cpu, "e2e": r.get("e2e"), "gpu_launches": r["gpu_launches"],
                "clocks": r["clocks"], "phases_ms_per_step": r["phases_ms_per_step"], "upot_last": r["upot_last"]}
        if c2_res is not None:
            q = c2_res
            rl2 = dict(q["roofline"])
            rl2["phases_ms_per_step"] = q["phases_ms_per_step"]
            line["c2"] = {"value": q["value"], "unit": "MFUPs/s", "scaling": "weak", "steps": q["steps"],
                          "warmup": q["warmup"], "ms_per_step": q["ms_per_step"],
                          "config": config_of("c2", C2["n_per_dim"], q), "roofline": rl2, "e2e": q.get("e2e"),
                          "gpu_launches": q["gpu_launches"], "clocks": q["clocks"], "upot_last": q["upot_last"]}
        if world == 1 and not args.no_other_configs and args.workload == "c3" and not args.no_c2:
            # BASELINE configs[0], [3], [4] (parity-test configurations, not bench lines) timed on the same box, so that
            # the driver-run record carries them: tools/bench_functors.py through the same C ABI
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_functors
                line["other_configs"] = bench_functors.c1(False) + bench_functors.c4(False) + bench_functors.c5(False)
                if not args.no_cpu_baseline:
                    other_configs_cpu_reference(line["other_configs"], bench_functors)
            except Exception as exc:  # never lose the headline line over the side measurements
                line["other_configs"] = {"error": repr(exc)}
            try:  # md-flexible's checkpoint of a 4 M-particle box written from the device SoA
                import bench_vtk
                line["checkpoint"] = bench_vtk.record_timing(4_000_000)
                if not args.no_cpu_baseline:
                    try:
                        line["checkpoint"]["cpu_reference"] = checkpoint_cpu_reference()
                    except Exception as exc:
                        line["checkpoint"]["cpu_reference"] = {"error": repr(exc)}
            except Exception as exc:
                line["checkpoint"] = {"error": repr(exc)}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
