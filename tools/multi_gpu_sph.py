"""BASELINE.json configs[4] on several GPUs: SPH density + hydro-force pass of a periodic box split over the ranks
(regular grid decomposition, the flow of examples/sph-mpi/sph-main-mpi.cpp:373-414: halo exchange -> density ->
pressure -> halo refresh of density and pressure -> hydro force). Run under torchrun, one rank per GPU (or plainly for
one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/multi_gpu_sph.py [n_per_dim=128] [--check]

--check: a 40^3 system; density, acceleration, engDot and vSigMax of every particle are compared with a single-GPU run
of the whole periodic system on rank 0 (1e-12 of the per-particle term scale). Without it: the 2 097 152-particle box
(128^3) is timed, one JSON line from rank 0."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from autopas_b200 import GpuParticleContainer, GpuTraversal, SPHCalcDensityFunctor, SPHCalcHydroForceFunctor, capi  # noqa: E402

D = 0.4          # lattice spacing
H = 1.2 * D      # smoothing length; kernel support 2.5 h
CUTOFF = 2.5 * H
SKIN = 0.1 * CUTOFF


def attributes(ids):
    """Deterministic per-particle attributes from the global id (any rank can evaluate them)."""
    g = ids.astype(np.float64)
    vel = 0.1 * np.stack([np.sin(0.37 * g), np.cos(0.91 * g), np.sin(1.3 * g + 1.0)], axis=1)
    mass = D ** 3 * (1.0 + 0.05 * np.sin(0.11 * g))
    snd = 1.2 + 0.1 * np.cos(0.23 * g)
    return vel, mass, snd


def local_particles(npd, rank, dims):
    counts = [npd // d for d in dims]
    if any(npd % d for d in dims):
        raise SystemExit(f"{npd} lattice points per dimension do not split over {dims}")
    c = bench.rank_coords(rank, dims)
    ix = [np.arange(counts[d]) + c[d] * counts[d] for d in range(3)]
    zz, yy, xx = np.meshgrid(ix[2], ix[1], ix[0], indexing="ij")
    gid = (zz.ravel() * npd + yy.ravel()) * npd + xx.ravel()
    base = (np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1) + 0.5) * D
    g = gid.astype(np.float64)
    jit = 0.05 * np.stack([np.sin(12.9898 * g), np.sin(78.233 * g + 1.0), np.sin(37.719 * g + 2.0)], axis=1)
    lo = np.array([c[d] * counts[d] * D for d in range(3)])
    hi = lo + np.array(counts) * D
    pos = np.clip(base + jit, lo, np.nextafter(hi, lo))
    return pos, gid, lo, hi


def sph_pass(c, timing=None):
    """halo exchange -> density -> pressure -> halo refresh -> hydro force; returns per-slot ids / ownership"""
    dens, hyd = SPHCalcDensityFunctor(), SPHCalcHydroForceFunctor()
    td, th = GpuTraversal("gpulc_c08", dens, False), GpuTraversal("gpulc_c08", hyd, False)
    c.exchangeHalos()
    c.rebuildNeighborLists(td)

    def density():
        dens.initTraversal()
        c.computeInteractions(td)
        dens.endTraversal(False)

    def hydro():
        hyd.initTraversal()
        c.computeInteractions(th)
        hyd.endTraversal(False)
    density()
    ids, _, own = c.downloadIds()
    rho = c.downloadColumn("DENSITY")
    c.uploadColumn("PRESSURE", np.where(own == 1, 0.4 * rho, 0.0))  # ideal-gas-like closure on the owned particles
    c.refreshHaloColumns(["DENSITY", "PRESSURE"])
    hydro()
    if timing is not None:
        # steady state: refresh + density + refresh + hydro force, max over ranks
        for name, fn in (("density", density), ("hydro", hydro), ("refresh", lambda: c.refreshHaloColumns(["DENSITY", "PRESSURE"]))):
            best = 1e9
            for _ in range(4):
                torch.cuda.synchronize()
                if dist.is_initialized():
                    dist.barrier()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            t = torch.tensor([best], dtype=torch.float64, device="cuda")
            if dist.is_initialized():
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            timing[name] = float(t.item())
    return ids, own


def build(npd, rank, world, dims, local, nccl_id):
    pos, gid, lo, hi = local_particles(npd, rank, dims)
    c = GpuParticleContainer("gpuLinkedCells", lo, hi, CUTOFF, SKIN, particleKind=capi.PARTICLE_SPH, device=local)
    if world > 1:
        c.commInit(world, rank, nccl_id)
    me = bench.rank_coords(rank, dims)
    nb = []
    for d in range(3):
        a, b = list(me), list(me)
        a[d] -= 1
        b[d] += 1
        nb += [bench.coords_rank(a, dims), bench.coords_rank(b, dims)]
    L = npd * D
    c.setDecomposition([0, 0, 0], [L, L, L], nb, (1, 1, 1))
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], gid)
    vel, mass, snd = attributes(gid)
    ns = c.numSlots()
    assert ns == len(gid)
    for k, v in (("VX", vel[:, 0]), ("VY", vel[:, 1]), ("VZ", vel[:, 2]), ("MASS", mass), ("SMTH", np.full(ns, H)),
                 ("SNDSPEED", snd)):
        c.uploadColumn(k, v)
    return c, len(gid)


def by_id(c, ids, own):
    m = own == 1
    out = {"id": ids[m]}
    for k in ("DENSITY", "FX", "FY", "FZ", "ENGDOT", "VSIGMAX"):
        out[k] = c.downloadColumn(k)[m]
    return out


def main():
    check = "--check" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    npd = int(args[0]) if args else (40 if check else 128)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            assert capi.load().apb_comm_get_unique_id(raw) == 0
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        nccl_id = idbuf.cpu().numpy().tobytes()
    dims = bench.decomposition(world)
    c, n_local = build(npd, rank, world, dims, local, nccl_id)
    timing = None if check else {}
    ids, own = sph_pass(c, timing)
    ok = True
    if check:
        mine = by_id(c, ids, own)
        gathered = [mine]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
        if rank == 0:
            s, _ = build(npd, 0, 1, [1, 1, 1], local, None)
            sids, sown = sph_pass(s)
            ref = by_id(s, sids, sown)
            s.close()
            o1 = np.argsort(ref["id"])
            allid = np.concatenate([g["id"] for g in gathered])
            assert len(allid) == npd ** 3 and len(np.unique(allid)) == npd ** 3, "particles lost or duplicated"
            o2 = np.argsort(allid)
            worst = {}
            for k in ("DENSITY", "FX", "FY", "FZ", "ENGDOT", "VSIGMAX"):
                a = np.concatenate([g[k] for g in gathered])[o2]
                b = ref[k][o1]
                worst[k] = float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))
            print(f"ranks {world} dims {dims} particles {npd ** 3}: max deviation from the one-GPU run relative to the "
                  f"column maximum: " + "  ".join(f"{k} {v:.2e}" for k, v in worst.items()))
            ok = all(v <= 1e-12 for v in worst.values())
            print("MULTI_GPU_SPH_CHECK", "PASS" if ok else "FAIL")
    elif rank == 0:
        n = npd ** 3
        total = timing["density"] + timing["hydro"] + 2 * timing["refresh"]
        print(json.dumps({"config": "C5 SPH density + hydro force, periodic box split over the GPUs (gpuLinkedCells/gpulc_c08)",
                          "n_gpus": world, "decomposition": dims, "particles": n, "particles_per_gpu": n_local,
                          "ms_density": timing["density"] * 1e3, "ms_hydro": timing["hydro"] * 1e3,
                          "ms_halo_refresh": timing["refresh"] * 1e3, "ms_pass": total * 1e3,
                          "MFUPs_per_s": n / total * 1e-6}), flush=True)
    c.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
