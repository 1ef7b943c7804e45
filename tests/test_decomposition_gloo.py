"""CPU, world_size 2, gloo: the host side of the multi-GPU path (bench.py / tools/multi_gpu_check.py): the regular-grid
decomposition of DomainTools::generateDecomposition, neighbour tables, broadcast of the 128-byte communicator id, and the
three-phase (x, then y, then z, forwarding received halos) halo selection protocol of
RegularGridDecomposition::exchangeHaloParticles (RegularGridDecomposition.cpp:159-236) that the device kernels
(kSelectCount / kSelectWrite / kPack in autopas_b200/csrc/dynamics.cu) implement. The protocol is restated in numpy, run
across two processes with gloo send/recv, and checked against a brute-force construction of every rank's halo set from the
global periodic system."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

IL, SKIN = 2.8, 0.3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _select(pos, own, d, lmin, lmax):
    """kSelect, mode 0: owned in [lmin, lmin+il) -> left, [lmax-il, lmax) -> right; received halos are forwarded from
    [lmin-skin, lmin+il) else [lmax-il, lmax+skin)."""
    p = pos[:, d]
    owned, halo = own == 1, own == 2
    left = (owned & (p >= lmin) & (p < lmin + IL)) | (halo & (p >= lmin - SKIN) & (p < lmin + IL))
    right = (owned & (p >= lmax - IL) & (p < lmax)) | (halo & ~((p >= lmin - SKIN) & (p < lmin + IL)) &
                                                        (p >= lmax - IL) & (p < lmax + SKIN))
    return left, right


def _exchange(rank, world, x, dst, src):
    """send x to dst while receiving from src (sizes first), like ncclSend / ncclRecv grouped per dimension"""
    if dst == rank and src == rank:
        return x.copy()
    n_out = torch.tensor([len(x)], dtype=torch.int64)
    n_in = torch.zeros(1, dtype=torch.int64)
    reqs = [dist.isend(n_out, dst), dist.irecv(n_in, src)]
    for r in reqs:
        r.wait()
    out = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1))
    buf = torch.zeros(int(n_in.item()) * 4, dtype=torch.float64)
    reqs = [dist.isend(out, dst), dist.irecv(buf, src)]
    for r in reqs:
        r.wait()
    return buf.numpy().reshape(-1, 4)


def _worker(rank, world, port, n_per_dim, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = bench.decomposition(world)
        assert dims == [2, 1, 1]
        pos, vel, bmin, bmax, gmin, gmax = bench.make_workload("c2", n_per_dim, rank, dims, seed=3)
        n = len(pos)
        # communicator id: rank 0 makes 128 bytes, everybody gets the same ones
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idbuf = torch.from_numpy(np.random.default_rng(0).integers(0, 256, 128, dtype=np.uint8))
        dist.broadcast(idbuf, 0)
        ids = [torch.zeros(128, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(ids, idbuf)
        assert all(torch.equal(ids[0], t) for t in ids)
        # neighbour table, symmetric: my right neighbour's left neighbour is me
        me = bench.rank_coords(rank, dims)
        nb = []
        for d in range(3):
            lo, hi = list(me), list(me)
            lo[d] -= 1
            hi[d] += 1
            nb += [bench.coords_rank(lo, dims), bench.coords_rank(hi, dims)]
        table = [torch.zeros(6, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(table, torch.tensor(nb, dtype=torch.int64))
        for d in range(3):
            assert int(table[nb[2 * d + 1]][2 * d]) == rank and int(table[nb[2 * d]][2 * d + 1]) == rank
        # boxes tile the global box
        boxes = [torch.zeros(6, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(boxes, torch.tensor(np.concatenate([bmin, bmax])))
        vol = sum(float(np.prod(b[3:].numpy() - b[:3].numpy())) for b in boxes)
        assert vol == pytest.approx(float(np.prod(gmax - gmin)), rel=1e-14)
        # ---- three-phase halo exchange, numpy restatement of exchangeDim(mode 0) ----
        L = gmax - gmin
        P = np.concatenate([pos, (np.arange(n) + rank * n)[:, None].astype(float)], axis=1)  # x, y, z, id
        own = np.ones(n, dtype=np.int64)
        links = []  # per dimension: (send slots to the left, to the right, receive slots from the left, from the right)
        for d in range(3):
            left, right = _select(P[:, :3], own, d, bmin[d], bmax[d])
            at_min, at_max = abs(bmin[d] - gmin[d]) < 1e-9, abs(bmax[d] - gmax[d]) < 1e-9
            to_left, to_right = P[left].copy(), P[right].copy()
            if at_min:
                to_left[:, d] += L[d]   # the sender applies the periodic shift (:509-545)
            if at_max:
                to_right[:, d] -= L[d]
            from_right = _exchange(rank, world, to_left, nb[2 * d], nb[2 * d + 1])   # what my right neighbour sent left
            from_left = _exchange(rank, world, to_right, nb[2 * d + 1], nb[2 * d])
            first = len(P)
            links.append((np.flatnonzero(left), np.flatnonzero(right), np.arange(first, first + len(from_left)),
                          np.arange(first + len(from_left), first + len(from_left) + len(from_right))))
            P = np.concatenate([P, from_left, from_right])
            own = np.concatenate([own, np.full(len(from_left) + len(from_right), 2)])
        halos = P[own == 2]
        # ---- apb_refresh_halo_columns (dynamics.cu: refreshHaloColumns): a column that changed on the owners reaches
        # every copy through the recorded links, dimension by dimension (copies of copies are fed by the refreshed copy)
        col = np.where(own == 1, np.cos(P[:, 3]) + 2.0, np.nan)  # new owner values by id; copies stale (NaN)
        for d in range(3):
            sl, sr, rl, rr = links[d]
            pack = lambda idx: np.stack([col[idx]] * 4, axis=1)  # noqa: E731  (4 doubles per entry, like _exchange expects)
            col[rr] = _exchange(rank, world, pack(sl), nb[2 * d], nb[2 * d + 1])[:, 0]
            col[rl] = _exchange(rank, world, pack(sr), nb[2 * d + 1], nb[2 * d])[:, 0]
        assert not np.isnan(col).any()
        np.testing.assert_array_equal(col, np.cos(P[:, 3]) + 2.0)  # every copy carries its owner's value
        # ---- brute force: every periodic image of every particle of the global system inside my halo shell ----
        allpos = [torch.zeros((n, 3), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allpos, torch.from_numpy(np.ascontiguousarray(pos)))
        G = np.concatenate([t.numpy() for t in allpos])
        gid = np.arange(len(G)).astype(float)
        want = []
        for a in (-1, 0, 1):
            for b in (-1, 0, 1):
                for c in (-1, 0, 1):
                    q = G + np.array([a, b, c]) * L
                    inside = np.all((q >= bmin) & (q < bmax), axis=1)
                    shell = np.all((q >= bmin - IL) & (q < bmax + IL), axis=1) & ~inside
                    want.append(np.concatenate([q[shell], gid[shell, None]], axis=1))
        want = np.concatenate(want)
        key = lambda A: sorted((int(r[3]), round(r[0], 9), round(r[1], 9), round(r[2], 9)) for r in A)  # noqa: E731
        assert key(halos) == key(want), (len(halos), len(want))
        open(os.path.join(result_dir, f"ok{rank}"), "w").write(str(len(halos)))
    finally:
        dist.destroy_process_group()


def test_two_rank_decomposition_and_halo_protocol(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 8, str(tmp_path)), nprocs=world, join=True)
    counts = [int(open(tmp_path / f"ok{r}").read()) for r in range(world)]
    assert all(c > 0 for c in counts)


def test_decomposition_matches_generate_decomposition():
    # DomainTools::generateDecomposition (DomainTools.cpp:26-71): 1 -> 1x1x1, 2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2
    assert [bench.decomposition(n) for n in (1, 2, 4, 8, 6, 12)] == [[1, 1, 1], [2, 1, 1], [2, 2, 1], [2, 2, 2], [3, 2, 1],
                                                                      [3, 2, 2]]
