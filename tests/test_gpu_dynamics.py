"""Device-resident loop pieces (halo exchange, migration, integration, apb_run_steps) against host-built references.
Needs a B200."""
import numpy as np
import pytest

import oracle
from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor, capi
from scenarios import grid_lattice, periodic_images

pytestmark = pytest.mark.gpu


def _functor(rc):
    f = LJFunctor(rc, applyShift=True, calculateGlobals=True)
    f.setParticleProperties(24.0, 1.0)
    return f


def _host_forces(pos, bmin, bmax, rc, skin):
    """Forces of the periodic system from the oracle with host-built periodic images."""
    hpos, _ = periodic_images(pos, bmin, bmax, rc + skin)
    allpos = np.vstack([pos, hpos])
    own = np.r_[np.ones(len(pos)), 2 * np.ones(len(hpos))].astype(np.int64)
    o = oracle.lj_linkedcells(allpos[:, 0], allpos[:, 1], allpos[:, 2], None, own, bmin, bmax, rc, skin, shift=True,
                              newton3=True)
    return o, len(hpos)


@pytest.mark.parametrize("cont,trav,n3,M", [("gpuLinkedCells", "gpulc_c08", True, 0),
                                            ("gpuVerletClusterLists", "gpuvcl_pruned", False, 32),
                                            ("gpuVerletClusterLists", "gpuvcl_c06", True, 4)])
def test_device_halo_exchange_equals_host_periodic_images(cont, trav, n3, M):
    """RegularGridDecomposition::exchangeHaloParticles with one rank = all periodic images within cutoff+skin,
    including edges and corners through the x -> y -> z forwarding."""
    rc, skin = 2.5, 0.3
    pos, bmin, bmax = grid_lattice(14, 1.1, jitter=0.12, seed=1)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    n = len(pos)
    o, nhalo = _host_forces(pos, bmin, bmax, rc, skin)
    c = GpuParticleContainer(cont, bmin, bmax, rc, skin, clusterSize=max(M, 1))
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    c.exchangeHalos()
    assert c.getNumberOfParticles("halo") == nhalo  # same halo set as the 26-image construction
    f = _functor(rc)
    t = GpuTraversal(trav, f, n3)
    c.rebuildNeighborLists(t)
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(n3)
    ids, _, own = c.downloadIds()
    F = np.zeros((n, 3))
    m = own == 1
    for d, name in enumerate(("FX", "FY", "FZ")):
        F[ids[m], d] = c.downloadColumn(name)[m]
    err = np.abs(F - o["f"][:n]).max(axis=1)
    assert np.all(err <= 1e-12 * o["fscale"][:n] + 1e-300)
    u, v = oracle.lj_end_traversal(o["res"])
    assert f.getPotentialEnergy() == pytest.approx(u, rel=1e-12)
    assert f.getVirial() == pytest.approx(v, rel=1e-12)
    c.close()


def test_halo_refresh_and_migration_follow_moving_particles():
    """Move particles (< skin/2) after the rebuild: refreshed halo copies must equal fresh periodic images; then let
    particles cross the box faces: apb_migrate wraps them (exchangeMigratingParticles with one periodic rank)."""
    rc, skin = 2.5, 0.4
    pos, bmin, bmax = grid_lattice(12, 1.15, jitter=0.1, seed=3)
    L = bmax - bmin
    pos = bmin + np.mod(pos + 0.52 - bmin, L)  # puts a lattice plane within reach of the box faces
    n = len(pos)
    c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, rc, skin, clusterSize=8)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    c.exchangeHalos()
    f = _functor(rc)
    t = GpuTraversal("gpuvcl_pruned", f, False)
    c.rebuildNeighborLists(t)
    rng = np.random.default_rng(0)
    ids, _, own = c.downloadIds()
    shift = rng.uniform(-1, 1, (n, 3))
    shift *= 0.19 / np.linalg.norm(shift, axis=1, keepdims=True)
    newpos = pos + shift  # may leave the box slightly: still owned until the next container update
    for d, name in enumerate("XYZ"):
        col = c.downloadColumn(name)
        m = own == 1
        col[m] = newpos[ids[m], d]
        c.uploadColumn(name, col)
    c.exchangeHalos()  # refresh only
    x, y, z = (c.downloadColumn(k) for k in "XYZ")
    ids, _, own = c.downloadIds()
    hal = own == 2
    got = np.stack([x[hal], y[hal], z[hal]], axis=1)
    src = newpos[ids[hal]]
    k = np.round((got - src) / L)
    np.testing.assert_allclose(got, src + k * L, rtol=0, atol=1e-12)  # each halo is an exact image of its source
    assert np.all(np.abs(k) <= 1) and np.all(np.abs(k).sum(axis=1) >= 1)
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(False)
    F = np.zeros((n, 3))
    m = own == 1
    for d, name in enumerate(("FX", "FY", "FZ")):
        F[ids[m], d] = c.downloadColumn(name)[m]
    wrapped = bmin + np.mod(newpos - bmin, L)
    o, _ = _host_forces(wrapped, bmin, bmax, rc, skin)
    err = np.abs(F - o["f"][:n]).max(axis=1)
    assert np.all(err <= 1e-12 * o["fscale"][:n] + 1e-300)
    # migration: everything back inside, nothing lost
    outside = np.any((newpos < bmin) | (newpos >= bmax), axis=1).sum()
    assert outside > 0
    c.migrate()
    assert c.getNumberOfParticles("owned") == n and c.getNumberOfParticles("halo") == 0
    x, y, z = (c.downloadColumn(k) for k in "XYZ")
    ids, _, own = c.downloadIds()
    live = own == 1  # sent particles and the old halos stay behind as dummies until the rebuild drops them
    P, ids = np.stack([x, y, z], axis=1)[live], ids[live]
    assert np.all((P >= bmin) & (P < bmax))
    np.testing.assert_allclose(P[np.argsort(ids)], wrapped, rtol=0, atol=1e-12)
    c.close()


def test_run_steps_matches_stepwise_calls_and_conserves_energy():
    """apb_run_steps (async device loop) == the same loop driven call by call; total energy of the NVE run is conserved
    to the accuracy of velocity Verlet."""
    rc, skin, dt, rebuild = 2.5, 0.3, 0.002, 5
    pos, bmin, bmax = grid_lattice(12, 1.2, jitter=0.05, seed=5)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    n = len(pos)
    rng = np.random.default_rng(1)
    vel = rng.normal(0, 0.8, (n, 3))
    vel -= vel.mean(axis=0)

    def setup():
        c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, rc, skin, clusterSize=32)
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        for d, name in enumerate(("VX", "VY", "VZ")):
            c.uploadColumn(name, vel[:, d])
        return c

    steps = 23
    f = _functor(rc)
    t = GpuTraversal("gpuvcl_pruned", f, False)
    a = setup()
    res = a.runSteps(t, steps, 0, dt, [1.0], rebuild)
    b = setup()
    upots = []
    for it in range(steps):
        b.integratePositions(dt, [1.0])
        if it % rebuild == 0:
            b.migrate()
            b.exchangeHalos()
            b.rebuildNeighborLists(t)
        else:
            b.exchangeHalos()
        f.initTraversal()
        b.computeInteractions(t)
        f.endTraversal(False)
        upots.append(f.getPotentialEnergy())
        b.integrateVelocities(dt, [1.0])

    def state(c):
        ids, _, own = c.downloadIds()
        m = own == 1
        order = np.argsort(ids[m])
        return {k: c.downloadColumn(k)[m][order] for k in ("X", "Y", "Z", "VX", "VY", "VZ", "FX", "FY", "FZ")}

    sa, sb = state(a), state(b)
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k])  # same kernels, same order: bit-identical trajectories
    up_async = [r.upot_sum * 0.5 / 6.0 for r in res]
    np.testing.assert_allclose(up_async, upots, rtol=1e-14)
    ke = 0.5 * (sa["VX"] ** 2 + sa["VY"] ** 2 + sa["VZ"] ** 2).sum()
    ke0 = 0.5 * (vel ** 2).sum()
    e_end = ke + upots[-1]
    # energy at iteration 0 (after the first force evaluation, before the first velocity half-step completes) is
    # approximated by ke0 + upot[0]; velocity Verlet keeps the drift small at this time step
    e0 = ke0 + upots[0]
    assert abs(e_end - e0) <= 5e-3 * abs(ke0)
    a.close()
    b.close()


def test_transfers_by_particle_id():
    """apb_upload_positions_by_id / apb_download_forces_by_id: host arrays indexed by id, only owned particles travel."""
    rng = np.random.default_rng(5)
    n, nh, L = 3000, 300, 9.0
    pos = rng.uniform(0, L, (n, 3))
    halo = rng.uniform(-0.9, L + 0.9, (nh, 3))
    halo = halo[((halo < 0) | (halo >= L)).any(axis=1)]
    c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [L, L, L], 1.0, 0.2, clusterSize=8)
    base = 1000
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n) + base)
    c.addHaloParticles(halo[:, 0], halo[:, 1], halo[:, 2], np.arange(len(halo)) + base + n)
    f = LJFunctor(1.0, applyShift=True, calculateGlobals=True)
    f.setParticleProperties(24.0, 1.0)
    t = GpuTraversal("gpuvcl_pruned", f, False)
    c.rebuildNeighborLists(t)
    moved = pos + rng.uniform(-0.04, 0.04, pos.shape)  # < skin / 2
    moved = np.clip(moved, 0, np.nextafter(L, 0))
    c.uploadPositionsById(np.ascontiguousarray(moved[:, 0]), np.ascontiguousarray(moved[:, 1]),
                          np.ascontiguousarray(moved[:, 2]), idBegin=base)
    ids, _, own = c.downloadIds()
    m = own == 1
    for d, name in enumerate("XYZ"):
        col = c.downloadColumn(name)
        np.testing.assert_array_equal(col[m], moved[ids[m] - base, d])
        hm = own == 2
        np.testing.assert_array_equal(np.sort(col[hm]), np.sort(halo[:, d]))  # halo slots untouched
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(False)
    fx, fy, fz = np.full(n + 5, 7.0), np.full(n + 5, 7.0), np.full(n + 5, 7.0)
    c.downloadForcesById(fx, fy, fz, idBegin=base)
    for d, (name, arr) in enumerate((("FX", fx), ("FY", fy), ("FZ", fz))):
        col = c.downloadColumn(name)
        np.testing.assert_array_equal(arr[ids[m] - base], col[m])
        np.testing.assert_array_equal(arr[n:], np.zeros(5))  # ids without an owned particle read as zero
    c.close()


def test_force_step_by_id_equals_separate_calls():
    """apb_force_step_by_id (one stream-ordered batch) against the same step through the individual entry points."""
    rng = np.random.default_rng(11)
    n, L = 4000, 10.0
    pos = rng.uniform(0, L, (n, 3))
    out = []
    for fused in (False, True):
        c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [L, L, L], 1.0, 0.2, clusterSize=32)
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        f = LJFunctor(1.0, applyShift=True, calculateGlobals=True, countFLOPs=True)
        f.setParticleProperties(24.0, 1.0)
        t = GpuTraversal("gpuvcl_pruned", f, False)
        x, y, z = (np.ascontiguousarray(pos[:, d]) for d in range(3))
        fx, fy, fz = np.zeros(n), np.zeros(n), np.zeros(n)
        res = []
        for it in range(3):
            moved = np.clip(pos + 0.01 * it, 0, np.nextafter(L, 0))  # displacement < skin / 2
            x, y, z = (np.ascontiguousarray(moved[:, d]) for d in range(3))
            f.initTraversal()
            if fused:
                c.forceStepById(t, x, y, z, fx, fy, fz, rebuild=it == 0)
            else:
                c.uploadPositionsById(x, y, z) if it > 0 else None
                if it == 0:
                    c.migrate()
                    c.exchangeHalos()
                    c.rebuildNeighborLists(t)
                else:
                    c.exchangeHalos()
                c.resetForces()
                c.computeInteractions(t)
                c.downloadForcesById(fx, fy, fz)
            f.endTraversal(False)
            res.append((np.stack([fx, fy, fz], 1).copy(), f.getPotentialEnergy(), f.getVirial(), f.getNumFLOPs()))
        out.append(res)
        c.close()
    for a, b in zip(*out):
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1:] == b[1:]
    assert np.abs(out[0][-1][0]).max() > 0


def test_rebuilds_after_motion_cost_the_same_and_allocate_nothing():
    """LogicHandler times every rebuild, the first ones after particles start moving included (LogicHandler.h:1066-1141).
    Three consecutive rebuild periods of a moving 125k-particle liquid: the rebuild phase of each must lie within 2x
    of the others, and from the second period on no device buffer may grow (stream-ordered pool + headroom)."""
    import ctypes
    import time

    rc, skin, dt = 2.5, 0.3, 0.002
    pos, bmin, bmax = grid_lattice(50, 1.0581, jitter=0.1, seed=5)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    n = len(pos)
    rng = np.random.default_rng(1)
    vel = rng.normal(0, 1, (n, 3))
    c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, rc, skin, clusterSize=32)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    for d, name in enumerate(("VX", "VY", "VZ")):
        c.uploadColumn(name, vel[:, d])
    f = _functor(rc)
    t = GpuTraversal("gpuvcl_pruned", f, False)
    c.enableLoopTiming(True)
    per_period, allocs = [], []
    for period in range(4):
        c.getLoopTiming()
        a0 = c.getAllocCount()
        c.runSteps(t, 10, 10 * period, dt, [1.0], 10, wantResults=False)
        tm = c.getLoopTiming()
        assert tm["rebuild"][1] == 1
        per_period.append(tm["rebuild"][0])
        allocs.append(c.getAllocCount() - a0)
    c.close()
    # period 0 builds everything for the first time (particles at rest in a fresh handle); 1..3 follow moving particles
    moving = per_period[1:]
    assert max(moving) <= 2.0 * min(moving), per_period
    assert allocs[2] == 0 and allocs[3] == 0, allocs
    # (period 0 also warms the process-wide memory pool, i.e. pays the driver's page allocations once)


def test_handle_reuse_after_delete_all_regenerates_halos():
    """deleteAllParticles followed by a refill and a rebuild must not refresh through the halo links of the old
    particle set (stale slot lists would overwrite live particles): the next exchange generates halos again."""
    rc, skin = 2.5, 0.3
    c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [0, 0, 0] + np.array([15.4, 15.4, 15.4]), rc, skin,
                             clusterSize=32)
    for seed, npd in ((1, 14), (2, 13)):
        pos, bmin, bmax = grid_lattice(npd, 15.4 / npd, jitter=0.1, seed=seed)
        bmin, bmax = np.zeros(3), np.full(3, 15.4)
        pos = np.mod(pos - pos.min(axis=0) + 0.3, 15.4)
        n = len(pos)
        o, nhalo = _host_forces(pos, bmin, bmax, rc, skin)
        c.deleteAllParticles()
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        f = _functor(rc)
        t = GpuTraversal("gpuvcl_pruned", f, False)
        if seed == 2:
            # a rebuild without a generating exchange in between: the old links must already be gone
            c.rebuildNeighborLists(t)
            with pytest.raises(capi.ApbError, match="no halo links"):
                c.exchangeHalos()  # structure valid + links of the OLD particle set would have been the refresh branch
        c.migrate()
        c.exchangeHalos()
        assert c.getNumberOfParticles("halo") == nhalo
        c.rebuildNeighborLists(t)
        c.exchangeHalos()  # refresh branch: must leave positions untouched (nothing moved)
        f.initTraversal()
        c.computeInteractions(t)
        f.endTraversal(False)
        ids, _, own = c.downloadIds()
        F = np.zeros((n, 3))
        m = own == 1
        for d, name in enumerate(("FX", "FY", "FZ")):
            F[ids[m], d] = c.downloadColumn(name)[m]
        err = np.abs(F - o["f"][:n]).max(axis=1)
        assert np.all(err <= 1e-12 * o["fscale"][:n] + 1e-300)
    c.close()
