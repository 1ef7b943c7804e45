"""Loop-control pieces on the device: velocity-scaling thermostat (md-flexible Thermostat.h),
dynamic-rebuild trigger (LogicHandler.h:955-1016) and the remainder traversal for buffered particles
(RemainderPairwiseInteractionHandler.h:63-135). Needs a B200."""
import numpy as np
import pytest

import oracle
from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor
from scenarios import grid_lattice, uniform_with_halo

pytestmark = pytest.mark.gpu


def _functor(rc, mixing=False, eps=None, sigma=None):
    if mixing:
        from autopas_b200 import ParticlePropertiesLibrary
        ppl = ParticlePropertiesLibrary(rc)
        for t, (e, s) in enumerate(zip(eps, sigma)):
            ppl.addSiteType(t, 1.0)
            ppl.addLJParametersToSite(t, e, s)
        ppl.calculateMixingCoefficients()
        return LJFunctor(rc, ppl, applyShift=True, useMixing=True, calculateGlobals=True, countFLOPs=True)
    f = LJFunctor(rc, applyShift=True, calculateGlobals=True, countFLOPs=True)
    f.setParticleProperties(24.0, 1.0)
    return f


# ---- thermostat -----------------------------------------------------------------------------------------------------
def _thermo_setup(seed=3):
    rng = np.random.default_rng(seed)
    pos, bmin, bmax = grid_lattice(10, 1.2, jitter=0.05, seed=seed)
    n = len(pos)
    types = rng.integers(0, 2, n).astype(np.int32)
    vel = rng.normal(0, 1.0, (n, 3)) * np.where(types == 0, 1.0, 0.4)[:, None]
    c = GpuParticleContainer("gpuLinkedCells", bmin, bmax, 2.5, 0.3)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n), types)
    for d, name in enumerate(("VX", "VY", "VZ")):
        c.uploadColumn(name, vel[:, d])
    return c, types, vel


def test_calc_temperature_matches_formula():
    """Thermostat::calcTemperatureComponent: T_type = sum(m v.v) / (3 N_type)."""
    c, types, vel = _thermo_setup()
    mass = np.array([1.0, 2.5])
    t, cnt = c.calcTemperature(mass)
    for ty in range(2):
        m = types == ty
        assert cnt[ty] == m.sum()
        assert t[ty] == pytest.approx(mass[ty] * (vel[m] ** 2).sum() / (3 * m.sum()), rel=1e-13)
    c.close()


@pytest.mark.parametrize("target,delta", [(1.4, 0.1), (0.05, 0.02), (0.9, 10.0)])
def test_apply_thermostat_scales_like_the_reference(target, delta):
    """Thermostat::apply (Thermostat.h:228-275): per type the temperature moves by at most |delta| towards the target
    and velocities are scaled by sqrt(immediateTarget / current); with a large delta the target is reached at once
    (ThermostatTest.cpp: 'testApplyAndCalcTemperature')."""
    c, types, vel = _thermo_setup(seed=5)
    mass = np.array([1.0, 2.5])
    t0, _ = c.calcTemperature(mass)
    c.applyThermostat(mass, target, -delta)  # the reference takes |delta|
    t1, _ = c.calcTemperature(mass)
    ids, _, _ = c.downloadIds()
    v1 = np.stack([c.downloadColumn(k) for k in ("VX", "VY", "VZ")], axis=1)
    for ty in range(2):
        cur = t0[ty]
        imm = min(cur + delta, target) if cur < target else max(cur - delta, target)
        assert t1[ty] == pytest.approx(imm, rel=1e-13)
        m = types[ids] == ty
        np.testing.assert_allclose(v1[m], vel[ids][m] * np.sqrt(imm / cur), rtol=1e-14)
    c.close()


def test_run_steps_applies_the_thermostat_every_interval():
    """apb_run_steps with apb_set_thermostat == the stepwise loop with Thermostat::apply after the velocity update of
    every `interval`-th iteration (Simulation.cpp:313, 539-546)."""
    rc, skin, dt = 2.5, 0.3, 0.002
    pos, bmin, bmax = grid_lattice(12, 1.2, jitter=0.05, seed=5)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    n = len(pos)
    vel = np.random.default_rng(1).normal(0, 0.8, (n, 3))
    out = []
    for device_loop in (True, False):
        c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, rc, skin, clusterSize=32)
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        for d, name in enumerate(("VX", "VY", "VZ")):
            c.uploadColumn(name, vel[:, d])
        f = _functor(rc)
        t = GpuTraversal("gpuvcl_pruned", f, False)
        if device_loop:
            c.setThermostat(True, 3, 1.4, 0.05)
            c.runSteps(t, 12, 0, dt, [1.0], 5, wantResults=False)
        else:
            for it in range(12):
                c.integratePositions(dt, [1.0])
                if it % 5 == 0:
                    c.migrate()
                    c.exchangeHalos()
                    c.rebuildNeighborLists(t)
                else:
                    c.exchangeHalos()
                f.initTraversal()
                c.computeInteractions(t)
                f.endTraversal(False)
                c.integrateVelocities(dt, [1.0])
                if it % 3 == 0:
                    c.applyThermostat([1.0], 1.4, 0.05)
        temp, _ = c.calcTemperature([1.0])
        ids, _, own = c.downloadIds()
        m = own == 1
        v = np.zeros((n, 3))
        for d, name in enumerate(("VX", "VY", "VZ")):
            v[ids[m], d] = c.downloadColumn(name)[m]
        out.append((temp[0], v))
        c.close()
    assert out[0][0] == out[1][0]
    np.testing.assert_array_equal(out[0][1], out[1][1])
    assert out[0][0] > (vel ** 2).sum() / (3 * n)  # heated towards the target


# ---- dynamic-rebuild trigger ------------------------------------------------------------------------------------------
def test_dynamic_rebuild_trigger_threshold():
    """LogicHandler::checkNeighborListsInvalidDoDynamicRebuild: displacement^2 >= (skin / 2)^2 of any owned particle."""
    rc, skin = 2.5, 0.4
    pos, bmin, bmax = grid_lattice(8, 1.2, jitter=0.05, seed=2)
    n = len(pos)
    c = GpuParticleContainer("gpuLinkedCells", bmin, bmax, rc, skin)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    f = _functor(rc)
    t = GpuTraversal("gpulc_c08", f, True)
    c.setDynamicRebuild(True)
    assert c.checkDynamicRebuild()  # no lists yet
    c.rebuildNeighborLists(t)
    assert not c.checkDynamicRebuild()
    x = c.downloadColumn("X")
    x0 = x.copy()
    k = n // 2
    x[k] = np.nextafter(x0[k] + skin / 2, -np.inf) - 1e-12
    c.uploadColumn("X", x)
    assert not c.checkDynamicRebuild()  # just below skin / 2
    x[k] = x0[k] + skin / 2 + 1e-12
    c.uploadColumn("X", x)
    assert c.checkDynamicRebuild()
    c.rebuildNeighborLists(t)  # records the new positions (ParticleBase::resetRAtRebuild)
    assert not c.checkDynamicRebuild()
    c.close()


def test_run_steps_rebuilds_early_when_particles_move_fast():
    """With the trigger on, a hot system rebuilds before rebuild_frequency steps have passed, and the forces at the end
    still equal the oracle's on the final positions (no interaction was missed)."""
    rc, skin, dt = 2.5, 0.2, 0.004
    pos, bmin, bmax = grid_lattice(12, 1.2, jitter=0.05, seed=7)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    n = len(pos)
    vel = np.random.default_rng(2).normal(0, 2.5, (n, 3))  # ~0.01 per step: skin / 2 after about ten steps at 1 sigma
    c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, rc, skin, clusterSize=32)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    for d, name in enumerate(("VX", "VY", "VZ")):
        c.uploadColumn(name, vel[:, d])
    f = _functor(rc)
    t = GpuTraversal("gpuvcl_pruned", f, False)
    c.setDynamicRebuild(True)
    c.runSteps(t, 40, 0, dt, [1.0], 20, wantResults=False)
    assert c.getDynamicRebuildCount() >= 2
    # forces of the last step against the oracle on the positions they were computed for
    ids, _, own = c.downloadIds()
    P = np.stack([c.downloadColumn(k) for k in "XYZ"], axis=1)
    F = np.stack([c.downloadColumn(k) for k in ("FX", "FY", "FZ")], axis=1)
    live = own != 0
    o = oracle.lj_bruteforce(P[live, 0], P[live, 1], P[live, 2], None, own[live].astype(np.int64), rc, shift=True)
    m = own[live] == 1
    err = np.abs(F[live][m] - o["f"][m]).max(axis=1)
    assert np.all(err <= 1e-12 * o["fscale"][m] + 1e-300)
    c.close()


# ---- remainder traversal ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cont,trav,n3,M", [("gpuLinkedCells", "gpulc_c08", True, 0),
                                            ("gpuVerletClusterLists", "gpuvcl_pruned", False, 32)])
@pytest.mark.parametrize("mixing", [False, True])
def test_remainder_equals_one_big_container(cont, trav, n3, M, mixing):
    """Container traversal + remainder traversal of the buffered particles == the oracle on the union of both sets:
    forces on container and buffered particles, potential energy and virial."""
    L, rc, skin = 8.0, 1.0, 0.2
    pos, own, types = uniform_with_halo(1500, 300, [L, L, L], rc, seed=11, ntypes=2)
    rng = np.random.default_rng(4)
    nb, nbh = 40, 15
    bpos = np.vstack([rng.uniform(0, L, (nb, 3)), rng.uniform(-rc, L + rc, (nbh, 3))])
    bown = np.r_[np.ones(nb), 2 * np.ones(nbh)].astype(np.int64)
    inside = np.all((bpos[nb:] >= 0) & (bpos[nb:] < L), axis=1)
    bpos[nb:][inside, 0] = -0.5 * rc  # halo-buffer particles live outside the box
    btypes = rng.integers(0, 2, nb + nbh).astype(np.int64)
    if not mixing:
        types, btypes = np.zeros_like(types), np.zeros_like(btypes)
    kw = dict(shift=True, mixing=mixing, eps=[1.0, 1.2], sigma=[1.0, 0.95]) if mixing else dict(shift=True)
    allpos, allown, alltypes = np.vstack([pos, bpos]), np.r_[own, bown], np.r_[types, btypes]
    o = oracle.lj_bruteforce(allpos[:, 0], allpos[:, 1], allpos[:, 2], alltypes, allown, rc, **kw)

    c = GpuParticleContainer(cont, [0, 0, 0], [L, L, L], rc, skin, clusterSize=max(M, 1))
    mo, mh = own == 1, own == 2
    ids = np.arange(len(pos))
    c.addParticles(pos[mo, 0], pos[mo, 1], pos[mo, 2], ids[mo], types[mo].astype(np.int32))
    c.addHaloParticles(pos[mh, 0], pos[mh, 1], pos[mh, 2], ids[mh], types[mh].astype(np.int32))
    f = _functor(rc, mixing, [1.0, 1.2], [1.0, 0.95])
    t = GpuTraversal(trav, f, n3)
    c.rebuildNeighborLists(t)
    f.initTraversal()
    c.computeInteractions(t)
    fb, raw = c.computeRemainder(f, bpos[:, 0], bpos[:, 1], bpos[:, 2], bown, btypes)
    f.endTraversal(n3)
    F = np.vstack([c.forcesById(len(pos)), fb])
    owned = allown == 1
    err = np.abs(F[owned] - o["f"][owned]).max(axis=1)
    assert np.all(err <= 1e-12 * o["fscale"][owned] + 1e-300)
    u, v = oracle.lj_end_traversal(o["res"])
    assert f.getPotentialEnergy() == pytest.approx(u, rel=1e-12)
    assert f.getVirial() == pytest.approx(v, rel=1e-12)
    assert raw.num_kernel_calls_n3 > 0
    c.close()
