"""Multi-rank parity check of the NCCL halo exchange / migration path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank owns one sub-box of a periodic global box (regular grid decomposition like md-flexible). The decomposed run
(halo exchange, rebuild, force step, several time steps with migration) is compared against a single-GPU run of the
whole system on rank 0: forces by particle id to 1e-12 of the pair-force scale, potential energy and virial (after
apb_allreduce_globals) to 1e-12 relative, and particle conservation after migration."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor, capi  # noqa: E402

RC, SKIN, DT, REBUILD = 2.5, 0.3, 0.002, 5


def functor():
    f = LJFunctor(RC, applyShift=True, calculateGlobals=True)
    f.setParticleProperties(24.0, 1.0)
    return f


def state_by_id(c, cols):
    ids, _, own = c.downloadIds()
    m = own == 1
    out = {"id": ids[m]}
    for k in cols:
        out[k] = c.downloadColumn(k)[m]
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = bench.decomposition(world)
    npd = 24
    pos, vel, bmin, bmax, gmin, gmax = bench.make_workload("c2", npd, rank, dims, seed=7)
    n = len(pos)
    c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, RC, SKIN, clusterSize=32, device=local)
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (ctypes.c_ubyte * 128)()
        assert capi.load().apb_comm_get_unique_id(raw) == 0
        idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    c.commInit(world, rank, idbuf.cpu().numpy().tobytes())
    me = bench.rank_coords(rank, dims)
    nb = []
    for d in range(3):
        lo, hi = list(me), list(me)
        lo[d] -= 1
        hi[d] += 1
        nb += [bench.coords_rank(lo, dims), bench.coords_rank(hi, dims)]
    c.setDecomposition(gmin, gmax, nb, (1, 1, 1))
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n) + rank * n)
    for d, name in enumerate(("VX", "VY", "VZ")):
        c.uploadColumn(name, vel[:, d])
    f = functor()
    t = GpuTraversal("gpuvcl_pruned", f, False)
    steps = 12
    res = c.runSteps(t, steps, 0, DT, [1.0], REBUILD)
    raw = capi.TraversalResult()
    raw.upot_sum = res[steps - 1].upot_sum
    for k in range(3):
        raw.virial_sum[k] = res[steps - 1].virial_sum[k]
    c.allreduceGlobals(raw)
    mine = state_by_id(c, ("X", "Y", "Z", "FX", "FY", "FZ"))
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v for k, v in mine.items()})
    ok = True
    if rank == 0:
        # single-GPU run of the whole system
        allpos, allvel, allid = [], [], []
        for r in range(world):
            p, v, *_ = bench.make_workload("c2", npd, r, dims, seed=7)
            allpos.append(p)
            allvel.append(v)
            allid.append(np.arange(len(p)) + r * len(p))
        P, V, I = np.vstack(allpos), np.vstack(allvel), np.concatenate(allid)
        s = GpuParticleContainer("gpuVerletClusterLists", gmin, gmax, RC, SKIN, clusterSize=32, device=local)
        s.addParticles(P[:, 0], P[:, 1], P[:, 2], I)
        for d, name in enumerate(("VX", "VY", "VZ")):
            s.uploadColumn(name, V[:, d])
        f1 = functor()
        t1 = GpuTraversal("gpuvcl_pruned", f1, False)
        res1 = s.runSteps(t1, steps, 0, DT, [1.0], REBUILD)
        ref = state_by_id(s, ("X", "Y", "Z", "FX", "FY", "FZ"))
        order = np.argsort(ref["id"])
        ids = np.concatenate([g["id"] for g in gathered])
        assert len(ids) == len(I) and len(np.unique(ids)) == len(I), "particles lost or duplicated by migration"
        o2 = np.argsort(ids)
        L = np.asarray(gmax) - np.asarray(gmin)
        worst_f, worst_x = 0.0, 0.0
        fscale = np.abs(np.stack([ref[k][order] for k in ("FX", "FY", "FZ")], 1)).max() + 1.0
        for k in ("X", "Y", "Z", "FX", "FY", "FZ"):
            a = np.concatenate([g[k] for g in gathered])[o2]
            b = ref[k][order]
            if k in ("X", "Y", "Z"):
                d = "XYZ".index(k)
                diff = np.abs((a - b + 0.5 * L[d]) % L[d] - 0.5 * L[d])
                worst_x = max(worst_x, diff.max())
            else:
                worst_f = max(worst_f, np.abs(a - b).max() / fscale)
        u_dec, u_one = raw.upot_sum, res1[steps - 1].upot_sum
        v_dec = sum(raw.virial_sum)
        v_one = sum(res1[steps - 1].virial_sum)
        print(f"ranks {world} dims {dims}: max |dx| {worst_x:.3e}  max |dF|/Fmax {worst_f:.3e}  "
              f"Upot rel {abs(u_dec - u_one) / abs(u_one):.3e}  virial rel {abs(v_dec - v_one) / abs(v_one):.3e}")
        ok = worst_x < 1e-9 and worst_f < 1e-9 and abs(u_dec - u_one) <= 1e-11 * abs(u_one)
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
        s.close()
    c.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
