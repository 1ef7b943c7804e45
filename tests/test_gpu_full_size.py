"""Parity at BASELINE.json's full sizes for the configurations the oracle cannot run whole in seconds: C3 (16 003 008-particle
LJ box), C4 (262 144-particle three-body box) and C5 (2 097 152 SPH particles), each as one periodic system on one B200.
The GPU result of the whole system is compared with the oracle on sampled regions - every particle of a region together
with everything within the cutoff of it, evaluated by the brute-force oracle (forces / densities of the region's inner
particles depend on nothing else) - and through size-independent properties: total force of the periodic system zero to
rounding, newton3 on == off, the counters of both modes related as the functor rules say."""
import numpy as np
import pytest

import bench
import oracle
from autopas_b200 import (AxilrodTellerMutoFunctor, GpuParticleContainer, GpuTraversal, LJFunctor, SPHCalcDensityFunctor,
                          SPHCalcHydroForceFunctor, capi)

pytestmark = pytest.mark.gpu


def _jittered_lattice(npd, spacing, jitter, seed):
    rng = np.random.default_rng(seed)
    g = (np.arange(npd) + 0.5) * spacing
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    pos += rng.uniform(-jitter, jitter, pos.shape)
    L = npd * spacing
    return np.mod(pos, L), L


def _regions(pos, L, centres, half, reach):
    """For every centre c (with half <= c <= L - half per axis: the inner box lies inside the periodic box, possibly
    touching its faces): the particles inside [c - half, c + half) at their own coordinates (inner), followed by every
    periodic image pos + k L, k in {-1, 0, 1}^3, inside the box widened by `reach` that is not an inner particle. The
    image coordinates are computed as the device computes its halo copies (one addition of +-L), so the oracle sees
    bit-identical positions. Returns (number of inner, particle index of every entry, positions)."""
    out = []
    for c in centres:
        c = np.clip(c, half, L - half)
        lo, hi = c - half - reach, c + half + reach
        per_axis = []
        for a in range(3):
            ks = {}
            for k in (-1, 0, 1):
                q = pos[:, a] + k * L
                m = (q >= lo[a]) & (q < hi[a])
                if m.any():
                    ks[k] = m
            per_axis.append(ks)
        inner_mask = np.all((pos >= c - half) & (pos < c + half), axis=1)
        idx, coords = [np.flatnonzero(inner_mask)], [pos[inner_mask]]
        for kx, mx in per_axis[0].items():
            for ky, my in per_axis[1].items():
                for kz, mz in per_axis[2].items():
                    m = mx & my & mz
                    if (kx, ky, kz) == (0, 0, 0):
                        m = m & ~inner_mask
                    j = np.flatnonzero(m)
                    if len(j):
                        idx.append(j)
                        coords.append(pos[j] + np.array([kx, ky, kz]) * L)
        out.append((int(inner_mask.sum()), np.concatenate(idx), np.vstack(coords)))
    return out


def _by_id(c, names, n):
    ids, _, own = c.downloadIds()
    m = own == capi.OWN_OWNED
    out = []
    for k in names:
        a = np.zeros(n)
        a[ids[m]] = c.downloadColumn(k)[m]
        out.append(a)
    return out


def test_c3_full_size_lj_box_against_sampled_oracle_and_properties():
    npd, spacing, rc, skin = bench.C3["n_per_dim"], bench.C3["spacing"], 2.5, bench.C3["skin"]
    pos, L = _jittered_lattice(npd, spacing, 0.25, 3)
    n = len(pos)
    assert n == 16003008
    results = {}
    for n3 in (False, True):
        c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [L, L, L], rc, skin, clusterSize=32)
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        c.exchangeHalos()
        f = LJFunctor(rc, applyShift=True, calculateGlobals=True, countFLOPs=True)
        f.setParticleProperties(24.0, 1.0)
        t = GpuTraversal("gpuvcl_pruned", f, n3)
        c.rebuildNeighborLists(t)
        f.initTraversal()
        c.computeInteractions(t)
        f.endTraversal(n3)
        F = np.stack(_by_id(c, ("FX", "FY", "FZ"), n), axis=1)
        results[n3] = (F, f.getPotentialEnergy(), f.getVirial(), f._raw.num_kernel_calls_no_n3, f._raw.num_kernel_calls_n3)
        c.close()
    F, upot, virial, k_no, _ = results[False]
    F3, upot3, virial3, k3_no, k3 = results[True]
    fmax = np.abs(F).max()
    assert np.abs(F.sum(axis=0)).max() <= 1e-9 * fmax * np.sqrt(n)  # momentum conservation of the periodic system
    assert np.abs(F - F3).max() <= 1e-11 * fmax
    assert upot3 == pytest.approx(upot, rel=1e-12) and virial3 == pytest.approx(virial, rel=1e-12)
    # every owned-owned pair within the cutoff is one newton3 kernel call or two calls without newton3; pairs with a halo
    # copy are evaluated once, from the owned side, in both modes
    assert k3_no == 0 and k3 < k_no < 2 * k3
    rng = np.random.default_rng(5)
    centres = np.vstack([rng.uniform(0, L, (4, 3)), [[0.3, 0.2, L - 0.1]], [[L / 2, 0.1, 0.4]]])  # bulk, a corner, an edge
    for ni, order, rpos in _regions(pos, L, centres, 6.0, rc):
        own = np.r_[np.ones(ni), 2 * np.ones(len(order) - ni)].astype(np.int64)
        o = oracle.lj_bruteforce(rpos[:, 0], rpos[:, 1], rpos[:, 2], None, own, rc, shift=True)
        err = np.abs(F[order[:ni]] - o["f"][:ni]).max(axis=1)
        assert ni > 100 and np.all(err <= 1e-12 * o["fscale"][:ni] + 1e-300), (ni, np.max(err / o["fscale"][:ni]))


def test_c4_full_size_three_body_box_against_sampled_oracle():
    rc, skin, nu = 2.5, 0.2, 0.073
    pos, L = _jittered_lattice(64, 1.2, 0.1, 4)
    n = len(pos)
    assert n == 262144
    hpos = bench.periodic_images(pos, np.zeros(3), np.full(3, L), rc + skin)
    results = {}
    for n3 in (False, True):
        c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], rc, skin)
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        c.addHaloParticles(hpos[:, 0], hpos[:, 1], hpos[:, 2], n + np.arange(len(hpos)))
        f = AxilrodTellerMutoFunctor(rc, calculateGlobals=True, countFLOPs=True)
        f.setParticleProperties(nu)
        t = GpuTraversal("gpulc_c08", f, n3)
        c.rebuildNeighborLists(t)
        f.initTraversal()
        c.computeInteractions(t)
        f.endTraversal(n3)
        results[n3] = (np.stack(_by_id(c, ("FX", "FY", "FZ"), n), axis=1), f.getPotentialEnergy())
        c.close()
    F, upot = results[False]
    F3, upot3 = results[True]
    fmax = np.abs(F).max()
    assert np.abs(F - F3).max() <= 1e-11 * fmax and upot3 == pytest.approx(upot, rel=1e-12)
    assert np.abs(F.sum(axis=0)).max() <= 1e-9 * fmax * np.sqrt(n)
    rng = np.random.default_rng(6)
    centres = np.vstack([rng.uniform(0, L, (2, 3)), [[0.2, L - 0.3, 0.1]]])
    for ni, order, rpos in _regions(pos, L, centres, 2.4, 2 * rc):  # a triplet reaches two cutoffs from its members
        own = np.r_[np.ones(ni), 2 * np.ones(len(order) - ni)].astype(np.int64)
        o = oracle.atm(rpos, None, own, rc, nu=nu)
        err = np.abs(F[order[:ni]] - o["f"][:ni]).max(axis=1)
        assert ni > 30 and np.all(err <= 1e-12 * o["scale"][:ni] + 1e-300), (ni, np.max(err / o["scale"][:ni]))


def test_c5_full_size_sph_box_against_sampled_oracle():
    d = 0.4
    h = 1.2 * d
    cutoff = 2.5 * h
    pos, L = _jittered_lattice(128, d, 0.05, 5)
    n = len(pos)
    assert n == 2097152
    rng = np.random.default_rng(7)
    vel = rng.normal(0, 0.1, (n, 3))
    mass, smth, snd = d ** 3 * rng.uniform(0.9, 1.1, n), h * rng.uniform(0.92, 1.0, n), rng.uniform(1.0, 1.4, n)
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], cutoff, 0.1 * cutoff, particleKind=capi.PARTICLE_SPH)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    for k, a in (("VX", vel[:, 0]), ("VY", vel[:, 1]), ("VZ", vel[:, 2]), ("MASS", mass), ("SMTH", smth), ("SNDSPEED", snd)):
        c.uploadColumn(k, a)
    c.exchangeHalos()  # periodic images carry the attributes of their owners
    dens, hyd = SPHCalcDensityFunctor(), SPHCalcHydroForceFunctor()
    td, th = GpuTraversal("gpulc_c08", dens, False), GpuTraversal("gpulc_c08", hyd, False)
    c.rebuildNeighborLists(td)
    dens.initTraversal()
    c.computeInteractions(td)
    dens.endTraversal(False)
    (rho,) = _by_id(c, ("DENSITY",), n)
    pressure = 0.4 * rho
    ids, _, own = c.downloadIds()
    col = c.downloadColumn("PRESSURE")
    m = own == capi.OWN_OWNED
    col[m] = pressure[ids[m]]
    c.uploadColumn("PRESSURE", col)
    c.refreshHaloColumns(["DENSITY", "PRESSURE"])
    hyd.initTraversal()
    c.computeInteractions(th)
    hyd.endTraversal(False)
    ax, ay, az, eng = _by_id(c, ("FX", "FY", "FZ", "ENGDOT"), n)
    acc = np.stack([ax, ay, az], axis=1)
    c.close()
    centres = np.vstack([rng.uniform(0, L, (3, 3)), [[0.1, 0.2, L - 0.1]]])
    for ni, order, rpos in _regions(pos, L, centres, 2.0, cutoff):
        own_r = np.r_[np.ones(ni), 2 * np.ones(len(order) - ni)].astype(np.int64)
        o_rho, sc = oracle.sph_density(rpos, mass[order], smth[order], own_r)
        assert ni > 500 and np.all(np.abs(rho[order[:ni]] - o_rho[:ni]) <= 1e-12 * sc[:ni])
        o_acc, o_eng, _, sca, sce = oracle.sph_hydro(rpos, vel[order], mass[order], smth[order], rho[order], pressure[order],
                                                     snd[order], own_r, with_eng_scale=True)
        assert np.all(np.abs(acc[order[:ni]] - o_acc[:ni]).max(axis=1) <= 1e-12 * sca[:ni])
        assert np.all(np.abs(eng[order[:ni]] - o_eng[:ni]) <= 1e-12 * sce[:ni])
