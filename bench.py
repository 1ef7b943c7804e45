#!/usr/bin/env python
"""bench.py — MFUPs/s of the LJ fp64 force step (BASELINE.json metric) on N B200s of one node.

A "step" is one iteration of the md-flexible simulation loop restricted to the hot path (Simulation.cpp:230-351):
positions -> [every `rebuild` steps: migration, halo exchange, neighbour-structure rebuild | halo position refresh]
-> LJ force kernel (shift + globals) -> velocities, on the BASELINE `configs[1]` workload (1M-particle LJ liquid,
VerletClusterLists, skin 0.3, rebuild every 10 steps) per GPU.  N > 1: regular-grid decomposition, one sub-box of the
same size per rank (weak scaling), NCCL halo exchange / migration over NVLink.

`value`       device-resident loop (apb_run_steps), CUDA events on the library's stream, max over ranks
`e2e`         the same step through the C ABI with HOST buffers: positions uploaded and forces downloaded every step
`roofline`    the dominant kernel (LJ force) against the FP64 DFMA peak measured live (MEASURED_PEAKS.json has no FP64)
`cpu_baseline` the unmodified reference (oracle/_ref, OpenMP on the host cores) on the same 1M-particle workload
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

RHO = 0.8442
CUTOFF, SKIN, REBUILD, DT = 2.5, 0.3, 10, 0.002
METRIC = "MFUPs/s LJ fp64 force step"
# BASELINE.json configs[2]: spinodal-decomposition start, 252^3 simple-cubic lattice at spacing 1.5,
# Maxwell-Boltzmann velocities at T = 1.4, skin 0.5, deltaT 0.00182367; the TOTAL is fixed (strong scaling)
C3 = {"n_per_dim": 252, "spacing": 1.5, "skin": 0.5, "dt": 0.00182367, "temperature": 1.4}


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


def make_workload(n_per_dim, rank, dims, seed=42, workload="c2"):
    """c2: 100^3 simple-cubic lattice at rho* = 0.8442, jittered, one sub-box of n_per_dim^3 per rank
    (weak scaling). c3: the 252^3 lattice at spacing 1.5 split over the ranks (strong scaling), Brownian velocities."""
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
        vel = rng.normal(0.0, np.sqrt(C3["temperature"]), pos.shape)
        return pos, vel, lo, lo + Ls, np.zeros(3), np.array(dims, dtype=float) * Ls
    spacing = RHO ** (-1.0 / 3.0)
    L = n_per_dim * spacing
    lo = c * L
    g = (np.arange(n_per_dim) + 0.5) * spacing
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1) + lo
    pos += rng.uniform(-0.1, 0.1, pos.shape)
    vel = rng.normal(0.0, 1.0, pos.shape)
    vel -= vel.mean(axis=0)
    return pos, vel, lo, lo + L, np.zeros(3), np.array(dims, dtype=float) * L


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


def reference_arm(args, pos, box_min, box_max):
    """The unmodified reference (oracle/_ref) on the host cores, same particles + periodic halo images, rebuild every 10
    iterations, LJFunctor SoA with shift + globals, newton3. Two of the reference's own configurations are timed -
    VerletClusterLists / vcl_c06 (cluster size 4, the reference default; the container BASELINE configs[1] names) and
    LinkedCells / lc_c08 - and the faster one is reported, as the AutoTuner would pick it; both are listed."""
    import oracle
    if not oracle.have_ref():
        return None
    halo = periodic_images(pos, box_min, box_max, CUTOFF + SKIN)
    allpos = np.vstack([pos, halo])
    own = np.r_[np.ones(len(pos)), 2 * np.ones(len(halo))].astype(np.int64)
    iters = max(REBUILD, (args.steps // REBUILD) * REBUILD) if args.impl == "reference" else REBUILD
    runs = []
    for container, traversal in (("VerletClusterLists", "vcl_c06"), ("LinkedCells", "lc_c08")):
        kw = dict(container=container, traversal=traversal, cluster_size=4, newton3=True, rebuild_freq=REBUILD)
        if args.impl == "reference" and args.warmup > 0:
            oracle.ref_bench_lj(allpos[:, 0], allpos[:, 1], allpos[:, 2], own, box_min, box_max, CUTOFF, SKIN, iters=1, **kw)
        r = oracle.ref_bench_lj(allpos[:, 0], allpos[:, 1], allpos[:, 2], own, box_min, box_max, CUTOFF, SKIN,
                                iters=iters, **kw)
        total = r["rebuild_s"] + r["compute_s"]
        runs.append({"container": container, "traversal": traversal, "value": len(pos) * iters / total * 1e-6,
                     "seconds": total, "rebuild_s": r["rebuild_s"], "compute_s": r["compute_s"],
                     "threads": r["threads"], "num_rebuilds": r["num_rebuilds"]})
    best = max(runs, key=lambda q: q["value"])
    return {"value": best["value"], "unit": "MFUPs/s", "cores": best["threads"], "kind": "reference",
            "iters": iters, "seconds": best["seconds"], "rebuild_s": best["rebuild_s"], "compute_s": best["compute_s"],
            "container": best["container"], "traversal": best["traversal"],
            "configurations": [{k: q[k] for k in ("container", "traversal", "value")} for q in runs],
            "sample": f"{iters} force iterations + {best['num_rebuilds']} rebuild(s) of the full {len(pos)}-particle "
                      f"workload; fastest of the reference's {' and '.join(q['container'] + '/' + q['traversal'] for q in runs)}"
                      f" (SoA, newton3, LJFunctor shift + globals): {best['container']}/{best['traversal']}, "
                      f"OpenMP {best['threads']} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-dim", type=int, default=100, help="lattice points per dimension per rank (100 -> 1M)")
    ap.add_argument("--cluster-size", type=int, default=32)
    ap.add_argument("--traversal", default="gpuvcl_pruned")
    ap.add_argument("--newton3", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 (default): BASELINE configs[1], 1M-particle LJ liquid per GPU, weak scaling; "
                         "c3: configs[2], the 16M-particle spinodal box split over the GPUs, strong scaling")
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
    global SKIN, DT
    if args.workload == "c3":
        SKIN, DT = C3["skin"], C3["dt"]
        if args.n_per_dim == 100:
            args.n_per_dim = C3["n_per_dim"]
        workload = (f"C3 spinodal-decomposition LJ box, {args.n_per_dim ** 3} particles in total ({args.n_per_dim}^3 "
                    f"lattice, spacing 1.5, T=1.4), cutoff 2.5, skin 0.5, rebuild every 10 steps, periodic")
    else:
        n_local = args.n_per_dim ** 3
        workload = (f"C2 LJ liquid rho*=0.8442, {n_local} particles per GPU (jittered {args.n_per_dim}^3 lattice), cutoff 2.5, "
                    f"skin 0.3, rebuild every 10 steps, periodic")

    if args.impl == "reference":
        if rank != 0:
            return
        pos, vel, bmin, bmax, gmin, gmax = make_workload(args.n_per_dim, 0, [1, 1, 1], workload=args.workload)
        ref = reference_arm(args, pos, bmin, bmax)
        if ref is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libautopas_ref.so was not built"}), file=json_out, flush=True)
            return
        line = {"impl": "reference", "metric": METRIC, "value": ref["value"], "unit": "MFUPs/s", "n_gpus": n_gpus,
                "steps": ref["iters"], "warmup": min(args.warmup, 1), "ms_per_step": ref["seconds"] / ref["iters"] * 1e3,
                "higher_is_better": True, "scaling": "strong" if args.workload == "c3" else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "container": ref["container"], "traversal": ref["traversal"],
                           "newton3": True, "cluster_size": 4, "host_threads": ref["cores"],
                           "configurations_timed": ref["configurations"]},
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "MFUPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist

    from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pos, vel, bmin, bmax, gmin, gmax = make_workload(args.n_per_dim, rank, dims, workload=args.workload)
    n = len(pos)
    c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, CUTOFF, SKIN, clusterSize=args.cluster_size,
                             device=local_rank)
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            rc = capi.load().apb_comm_get_unique_id(raw)
            if rc != 0:
                raise SystemExit("apb_comm_get_unique_id failed")
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        c.commInit(world, rank, bytes(idbuf.cpu().numpy().tobytes()))
        me = rank_coords(rank, dims)
        nb = []
        for d in range(3):
            lo, hi = list(me), list(me)
            lo[d] -= 1
            hi[d] += 1
            nb += [coords_rank(lo, dims), coords_rank(hi, dims)]
        c.setDecomposition(gmin, gmax, nb, (1, 1, 1))
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n) + rank * n)
    for d, name in enumerate(("VX", "VY", "VZ")):
        c.uploadColumn(name, vel[:, d])
    functor = LJFunctor(CUTOFF, applyShift=True, calculateGlobals=True, countFLOPs=True,
                       virialTraceOnly=not args.virial_components)
    functor.setParticleProperties(24.0, 1.0)
    trav = GpuTraversal(args.traversal, functor, bool(args.newton3))
    mass = [1.0]

    stream = torch.cuda.ExternalStream(c.getStream(), device=torch.device("cuda", local_rank))
    steps = max(REBUILD, (args.steps // REBUILD) * REBUILD)  # whole rebuild periods
    warm = max(3, args.warmup)
    warm = ((warm + REBUILD - 1) // REBUILD) * REBUILD
    c.runSteps(trav, warm, 0, DT, mass, REBUILD, wantResults=False)

    # ---- device-resident timed region ----
    c.enableLoopTiming(True)
    c.getLoopTiming()
    sampler = ClockSampler(local_rank if rank == 0 else None)  # one nvidia-smi poller per job, not per rank
    launches0 = c.getLaunchCount()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    res = c.runSteps(trav, steps, warm, DT, mass, REBUILD)
    e1.record(stream)
    e1.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    launches = c.getLaunchCount() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    timing = c.getLoopTiming()
    c.enableLoopTiming(False)
    owned_total = sum_over_ranks(float(c.getNumberOfParticles("owned")))
    value = owned_total * steps / (ms_total * 1e-3) * 1e-6

    # ---- roofline of the dominant kernel (LJ force): reference FLOP model on the kernel's own counters ----
    force_ms, force_launches = timing["force"]
    flops_per_step = np.mean([8 * r.num_dist_calls + 15 * r.num_kernel_calls_no_n3 + 18 * r.num_kernel_calls_n3 +
                              9 * r.num_global_calcs_no_n3 + 13 * r.num_global_calcs_n3 for r in res])
    hit_rate = np.mean([(r.num_kernel_calls_no_n3 + r.num_kernel_calls_n3) / max(r.num_dist_calls, 1) for r in res])
    peak_tf, peak_ms = ctypes.c_double(), ctypes.c_double()
    capi.load().apb_measure_fp64_peak(local_rank, 3, ctypes.byref(peak_tf), ctypes.byref(peak_ms))
    kernel_ms = force_ms / max(force_launches, 1)
    achieved_tf = flops_per_step / (kernel_ms * 1e-3) / 1e12
    # bound: the FP64 FMA pipe (north_star asks for the fraction of the B200 FP64 peak; the kernel is neither HBM- nor
    # tensor-bound: 268 MB of DRAM traffic per 0.19 ms launch = 1.4 TB/s, and the path is not a contraction)
    roofline = {"bound": "fp64", "kernel": "kLJPruned" if args.traversal == "gpuvcl_pruned" else "kLJClusterPairs",
                "achieved": achieved_tf, "peak": peak_tf.value, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf.value if peak_tf.value else None, "traffic": None,
                "kernel_ms": kernel_ms, "flops_per_launch": flops_per_step, "hit_rate": hit_rate,
                "peak_source": "FP64 DFMA microbenchmark measured live by bench.py (apb_measure_fp64_peak); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "share_of_step": force_ms / ms_total,
                "phases_ms_per_step": {k: v[0] / steps for k, v in timing.items()}}
    if args.workload == "c2" and args.traversal == "gpuvcl_pruned" and args.n_per_dim == 100:
        try:  # DRAM bytes per launch of this kernel on this workload, from the committed ncu --set full capture
            roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "kLJPruned_traffic.json")))["dram_bytes_per_launch"]
            roofline["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)"
            roofline["algorithmic_bytes_per_launch"] = int(n * 48 + np.mean([r.num_dist_calls for r in res]) * 2)
        except (OSError, ValueError, KeyError):
            pass
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        roofline["hbm_peak_gbs_measured"] = peaks.get("hbm_gbs")
    except (OSError, ValueError):
        roofline["hbm_peak_gbs_measured"] = None

    # ---- end to end through the C ABI with host buffers ----
    # The host owns x, y, z and fx, fy, fz indexed by particle id (pinned). Every step: positions host -> device, [every
    # 10th: migration, halo exchange, rebuild | halo refresh], force kernel, forces device -> host, Upot / virial read back.
    e2e_steps = max(REBUILD, (args.e2e_steps // REBUILD) * REBUILD)
    host = {k: torch.empty(n, dtype=torch.float64).pin_memory().numpy() for k in ("x", "y", "z", "fx", "fy", "fz")}
    c.migrate()  # positions back into the (periodic) box before the host takes its copy
    c.exchangeHalos()
    c.rebuildNeighborLists(trav)
    ids_s, _, own_s = c.downloadIds()
    m_owned = own_s == capi.OWN_OWNED
    k_owned = ids_s[m_owned] - rank * n
    if not ((k_owned >= 0) & (k_owned < n)).all():
        k_owned = None  # particles migrated between ranks during the device-resident run: fall back to slot order
    for d, col in enumerate(("X", "Y", "Z")):
        colv = c.downloadColumn(col)
        if k_owned is not None:
            host["xyz"[d]][k_owned] = colv[m_owned]
    h2d = d2h = 0
    upot_e2e = []
    if k_owned is None:
        cap = int(c.numSlots() * 1.3) + 4096
        host = {k: torch.empty(cap, dtype=torch.float64).pin_memory().numpy() for k in ("x", "y", "z", "fx", "fy", "fz")}

        def pull_positions():
            lib = capi.load()
            for k, col in (("x", "X"), ("y", "Y"), ("z", "Z")):
                lib.apb_download_column(c._h, capi.COL[col], host[k].ctypes.data)
            return c.numSlots()

        ns = pull_positions()
    barrier()
    t0 = time.perf_counter()
    for it in range(e2e_steps):
        if k_owned is not None:
            # one C-ABI call per step: apb_force_step_by_id (upload, [rebuild chain | halo refresh], forces, download)
            functor.initTraversal()
            raw = c.forceStepById(trav, host["x"], host["y"], host["z"], host["fx"], host["fy"], host["fz"],
                                  rebuild=it % REBUILD == 0, idBegin=rank * n)
            functor.endTraversal(bool(args.newton3))
            h2d += 3 * 8 * n
            d2h += 3 * 8 * n + ctypes.sizeof(raw)
            upot_e2e.append(functor.getPotentialEnergy())
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
        upot_e2e.append(functor.getPotentialEnergy())
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": owned_total * e2e_steps / e2e_s * 1e-6, "unit": "MFUPs/s",
           "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps,
           "ms_per_step": e2e_s / e2e_steps * 1e3,
           "note": ("positions host->device and forces device->host every step, pinned host arrays indexed by particle "
                    "id, one apb_force_step_by_id call per step, Upot/virial read back"
                    if k_owned is not None else
                    "positions host->device and forces device->host every step in storage order (pinned), Upot/virial "
                    "read back")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_arm(args, pos, bmin, bmax)
            if ref is not None:
                cpu = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "MFUPs/s", "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}

    if rank == 0:
        g = c.getTraversalSelectorInfo()
        line = {"metric": METRIC, "value": value, "unit": "MFUPs/s", "n_gpus": n_gpus, "steps": steps, "warmup": warm,
                "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong" if args.workload == "c3" else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "container": "gpuVerletClusterLists", "traversal": args.traversal,
                           "newton3": bool(args.newton3), "cluster_size": args.cluster_size,
                           "functor": "LJFunctor shift+globals+flop counters" + (", virial per component" if args.virial_components else ""), "decomposition": dims,
                           "particles_total": int(owned_total), "num_clusters": int(g.num_clusters),
                           "num_cluster_pairs": int(g.num_cluster_pairs),
                           "l2_policy": "inputs larger than L2: per-particle lists + SoA columns streamed every step "
                                        "exceed the 126 MB L2",
                           "host_wall_ms_per_step": wall / steps * 1e3},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "upot_last": res[steps - 1].upot_sum * 0.5 / 6.0}
        print(json.dumps(line), file=json_out, flush=True)
    c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
