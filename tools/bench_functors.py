"""Throughput of the kernels behind BASELINE.json configs[0], [3] and [4] (the parity-test configurations that are not
bench.py lines), timed on one B200 through the C ABI. One JSON object per line.

  C1  32 768-particle cubic grid (spacing 1.1225), LJFunctor on gpuLinkedCells / gpulc_c08, newton3 on and off
  C4  262 144-particle argon box (spacing 1.2, jittered), AxilrodTellerMutoFunctor (nu 0.073) on gpuLinkedCells, newton3 off
  C5  2 097 152 SPH particles (128^3, h = 1.2 d, support 2.5 h), density then hydro-force functor, newton3 off

usage: python tools/bench_functors.py [c1] [c4] [c5] [--small]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from autopas_b200 import (AxilrodTellerMutoFunctor, GpuParticleContainer, GpuTraversal, LJFunctor,  # noqa: E402
                          SPHCalcDensityFunctor, SPHCalcHydroForceFunctor, capi)


def lattice(n_per_dim, spacing, jitter, seed):
    rng = np.random.default_rng(seed)
    g = (np.arange(n_per_dim) + 0.5) * spacing
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    if jitter:
        pos = pos + rng.uniform(-jitter, jitter, pos.shape)
    L = n_per_dim * spacing
    return np.clip(pos, 0, np.nextafter(L, 0)), L


def images(pos, L, width):
    out = []
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for c in (-1, 0, 1):
                if (a, b, c) == (0, 0, 0):
                    continue
                p = pos + np.array([a, b, c]) * L
                out.append(p[np.all((p >= -width) & (p < L + width), axis=1)])
    return np.vstack(out)


def timed(fn, reps):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def container(pos, L, cutoff, skin, kind):
    halo = images(pos, L, cutoff + skin)
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], cutoff, skin, particleKind=kind)
    n = len(pos)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    c.addHaloParticles(halo[:, 0], halo[:, 1], halo[:, 2], np.arange(len(halo)) + n)
    return c, n, len(halo)


def c1(small):
    out = []
    npd = 16 if small else 32
    pos, L = lattice(npd, 1.1225, 0.0, 1)
    for n3 in (True, False):
        c, n, nh = container(pos, L, 2.5, 0.2, capi.PARTICLE_LJ)
        f = LJFunctor(2.5, applyShift=True, calculateGlobals=True, countFLOPs=True)
        f.setParticleProperties(24.0, 1.0)
        t = GpuTraversal("gpulc_c08", f, n3)
        c.rebuildNeighborLists(t)

        def call():
            c.resetForces()
            f.initTraversal()
            c.computeInteractions(t)
            f.endTraversal(n3)
        s = timed(call, 10)
        out.append({"config": "C1 LJ gpuLinkedCells/gpulc_c08", "newton3": n3, "particles": n, "halo": nh,
                    "ms_per_call": s * 1e3, "MFUPs_per_s": n / s * 1e-6, "flops_reference_model": f.getNumFLOPs(),
                    "TFLOP_per_s": f.getNumFLOPs() / s * 1e-12, "hit_rate": f.getHitRate()})
        c.close()
    return out


def c4(small):
    npd = 24 if small else 64
    pos, L = lattice(npd, 1.2, 0.1, 4)
    c, n, nh = container(pos, L, 2.5, 0.2, capi.PARTICLE_LJ)
    f = AxilrodTellerMutoFunctor(2.5, calculateGlobals=True, countFLOPs=True)
    f.setParticleProperties(0.073)
    t = GpuTraversal("gpulc_c08", f, False)
    c.rebuildNeighborLists(t)

    def call():
        c.resetForces()
        f.initTraversal()
        c.computeInteractions(t)
        f.endTraversal(False)
    s = timed(call, 3)
    out = [{"config": "C4 AxilrodTellerMuto gpuLinkedCells/gpulc_c08", "newton3": False, "particles": n, "halo": nh,
            "ms_per_call": s * 1e3, "MFUPs_per_s": n / s * 1e-6, "kernel_calls_per_particle": f._raw.num_kernel_calls_no_n3 / n,
            "flops_reference_model": f.getNumFLOPs(), "TFLOP_per_s": f.getNumFLOPs() / s * 1e-12}]
    c.close()
    return out


def c5(small):
    npd = 32 if small else 128
    d = 0.4
    h = 1.2 * d
    pos, L = lattice(npd, d, 0.05, 5)
    cutoff = 2.5 * h
    c, n, nh = container(pos, L, cutoff, 0.1 * cutoff, capi.PARTICLE_SPH)
    dens, hyd = SPHCalcDensityFunctor(), SPHCalcHydroForceFunctor()
    td, th = GpuTraversal("gpulc_c08", dens, False), GpuTraversal("gpulc_c08", hyd, False)
    c.rebuildNeighborLists(td)
    ns = c.numSlots()
    rng = np.random.default_rng(0)
    c.uploadColumn("MASS", np.full(ns, d ** 3))
    c.uploadColumn("SMTH", np.full(ns, h))
    for k in ("VX", "VY", "VZ"):
        c.uploadColumn(k, rng.normal(0, 0.1, ns))
    c.uploadColumn("PRESSURE", np.full(ns, 1.0))
    c.uploadColumn("SNDSPEED", np.full(ns, 1.2))

    def call_d():
        c.uploadColumn("DENSITY", np.zeros(ns)) if False else None
        dens.initTraversal()
        c.computeInteractions(td)
        dens.endTraversal(False)
    call_d()  # first call ever: list buffers are allocated
    c.rebuildNeighborLists(td)
    t0 = time.perf_counter()
    call_d()  # first call after a rebuild: includes the partner-list build of the list variant
    first = time.perf_counter() - t0
    sd = timed(call_d, 3)
    c.uploadColumn("DENSITY", np.full(ns, 1.0))

    def call_h():
        hyd.initTraversal()
        c.computeInteractions(th)
        hyd.endTraversal(False)
    sh = timed(call_h, 3)
    out = [{"config": f"C5 SPH {name} gpuLinkedCells/gpulc_c08", "newton3": False, "particles": n, "halo": nh,
            "ms_per_call": s * 1e3, "MFUPs_per_s": n / s * 1e-6, "ms_first_density_call_after_rebuild": first * 1e3}
           for name, s in (("density", sd), ("hydro force", sh))]
    c.close()
    return out


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c4", "c5"]
    small = "--small" in sys.argv
    for w in which:
        for line in {"c1": c1, "c4": c4, "c5": c5}[w](small):
            print(json.dumps(line), flush=True)
