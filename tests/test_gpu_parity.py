"""Parity of the CUDA path (through the C ABI) against the oracle and the reference fixtures. Needs a B200.

Tolerances: cell indices, tower/cluster membership, cluster-pair sets and FLOP counters are bit-exact; forces are held
to |dF_i|_inf <= 1e-12 * sum_j(|f_ij,x|+|f_ij,y|+|f_ij,z|) (the reference's own cross-configuration tolerance is 1e-10
relative, TraversalComparison.cpp:245-246); potential energy and virial to 1e-12 relative."""
import os

import numpy as np
import pytest

import oracle
from autopas_b200 import ApbError, GpuParticleContainer, GpuTraversal, LJFunctor, ParticlePropertiesLibrary, capi
from scenarios import grid_lattice, periodic_images, uniform_with_halo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FTOL = 1e-12


def make_functor(cutoff, shift, mixing, eps, sigma, globals_=True, flops=True):
    eps, sigma = np.atleast_1d(eps), np.atleast_1d(sigma)
    if mixing:
        ppl = ParticlePropertiesLibrary(cutoff)
        for t in range(len(eps)):
            ppl.addSiteType(t, 1.0)
            ppl.addLJParametersToSite(t, eps[t], sigma[t])
        ppl.calculateMixingCoefficients()
        return LJFunctor(cutoff, ppl, applyShift=shift, useMixing=True, calculateGlobals=globals_, countFLOPs=flops)
    f = LJFunctor(cutoff, applyShift=shift, calculateGlobals=globals_, countFLOPs=flops)
    f.setParticleProperties(24.0 * eps[0], sigma[0] * sigma[0])
    return f


def run_gpu(container_opt, traversal_opt, pos, own, types, box_min, box_max, cutoff, skin, newton3, shift=False,
            mixing=False, eps=1.0, sigma=1.0, csf=1.0, M=4, globals_=True, flops=True):
    """TraversalComparison::calculateForcesImpl (TraversalComparison.cpp:135-199) against the GPU container."""
    c = GpuParticleContainer(container_opt, box_min, box_max, cutoff, skin, cellSizeFactor=csf, clusterSize=M)
    n = len(pos)
    ids = np.arange(n)
    mo, mh = own == 1, own == 2
    c.addParticles(pos[mo, 0], pos[mo, 1], pos[mo, 2], ids[mo], types[mo] if types is not None else None)
    if mh.any():
        c.addHaloParticles(pos[mh, 0], pos[mh, 1], pos[mh, 2], ids[mh], types[mh] if types is not None else None)
    functor = make_functor(cutoff, shift, mixing, eps, sigma, globals_, flops)
    trav = GpuTraversal(traversal_opt, functor, newton3)
    c.rebuildNeighborLists(trav)
    functor.initTraversal()
    c.computeInteractions(trav)
    functor.endTraversal(newton3)
    return c, functor


def check_forces(f_gpu, f_ref, fscale, own):
    m = own == 1
    err = np.abs(f_gpu[m] - f_ref[m]).max(axis=1)
    bad = err > FTOL * fscale[m] + 1e-300
    assert not bad.any(), f"max rel err {np.max(err / np.maximum(fscale[m], 1e-300))}"


# ---- reference golden literal (LJFunctorTestNoGlobals.h:27-31) through every GPU traversal ----------------------
@pytest.mark.parametrize("cont,trav,n3", [
    ("gpuLinkedCells", "gpulc_c08", True), ("gpuLinkedCells", "gpulc_c08", False), ("gpuLinkedCells", "gpulc_c18", True),
    ("gpuVerletClusterLists", "gpuvcl_cluster_iteration", False), ("gpuVerletClusterLists", "gpuvcl_c06", True),
    ("gpuVerletClusterLists", "gpuvcl_c06", False), ("gpuVerletClusterLists", "gpuvcl_c01_balanced", False),
    ("gpuVerletClusterLists", "gpuvcl_pruned", False), ("gpuVerletClusterLists", "gpuvcl_pruned", True)])
def test_lj_golden_literal(cont, trav, n3):
    pos = np.array([[1.0, 1.0, 1.0], [1.1, 1.2, 1.3]])
    own = np.array([1, 1])
    c, f = run_gpu(cont, trav, pos, own, None, [0, 0, 0], [5, 5, 5], 1.0, 0.2, n3)
    F = c.forcesById(2)
    expected = np.array([-4547248.8989645941, -9094497.7979291882, -13641746.696893783])
    np.testing.assert_allclose(F[0], expected, atol=1e-7, rtol=0)
    np.testing.assert_allclose(F[1], -expected, atol=1e-7, rtol=0)
    c.close()


# ---- fixtures generated from the unmodified reference ---------------------------------------------------------------
def _golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("fn_", "c1_"))) if os.path.isdir(GOLDEN) else []


@pytest.mark.parametrize("fname", _golden_files())
def test_gpu_matches_reference_fixture(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    cfg = {k: g[k].item() for k in ("cutoff", "skin", "shift", "mixing", "newton3", "cluster_size", "csf")}
    pos, own, types = g["pos"], g["own"], g["types"]
    lc = str(g["container"]) == "LinkedCells"
    n3 = bool(cfg["newton3"])
    M = int(cfg["cluster_size"])
    kw = dict(shift=bool(cfg["shift"]), mixing=bool(cfg["mixing"]), eps=g["eps"], sigma=g["sigma"])
    if lc:
        o = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], types, own, g["box_min"], g["box_max"], cfg["cutoff"],
                                  cfg["skin"], cfg["csf"], newton3=n3, **kw)
        c, f = run_gpu("gpuLinkedCells", "gpulc_c08", pos, own, types, g["box_min"], g["box_max"], cfg["cutoff"],
                       cfg["skin"], n3, csf=cfg["csf"], **kw)
        ids, _, _ = c.downloadIds()
        np.testing.assert_array_equal(c.debugCellOfSlot(), g["ref_cell"][ids])  # bit-exact cell assignment
    else:
        o = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, g["box_min"], g["box_max"], cfg["cutoff"],
                          cfg["skin"], M, newton3=n3, **kw)
        c, f = run_gpu("gpuVerletClusterLists", "gpuvcl_c06" if n3 else "gpuvcl_cluster_iteration", pos, own, types,
                       g["box_min"], g["box_max"], cfg["cutoff"], cfg["skin"], n3, M=M, **kw)
        ids, _, _ = c.downloadIds()
        np.testing.assert_array_equal(ids.reshape(-1, M), g["ref_cluster_particles"])  # bit-exact clusters
        pairs = c.debugClusterPairs()
        assert len(pairs) == len(g["ref_pairs"])
        assert {tuple(p) for p in pairs} == {tuple(p) for p in g["ref_pairs"]}  # bit-exact cluster-pair set
        info = c.getTraversalSelectorInfo()
        assert (info.cells_per_dim[0], info.cells_per_dim[1]) == tuple(g["ref_towers_per_dim"])
    check_forces(c.forcesById(len(pos)), g["ref_f"], o["fscale"], own)
    assert f.getPotentialEnergy() == pytest.approx(g["ref_upot"].item(), rel=1e-12)
    assert f.getVirial() == pytest.approx(g["ref_virial"].item(), rel=1e-12)
    assert f.getNumFLOPs() == int(g["ref_flops"])  # identical distance / kernel / globals counters
    c.close()


@pytest.mark.parametrize("fname", [f for f in _golden_files() if f.startswith("vcl")])
@pytest.mark.parametrize("n3", [False, True])
def test_pruned_matches_reference_fixture(fname, n3):
    """gpuvcl_pruned (newton3 off and on) on the VerletClusterLists fixtures of the unmodified reference: forces, Upot
    and virial of vcl_c06 / vcl_cluster_iteration are mode independent, so every fixture checks both modes."""
    g = np.load(os.path.join(GOLDEN, fname))
    cfg = {k: g[k].item() for k in ("cutoff", "skin", "shift", "mixing", "newton3", "cluster_size", "csf")}
    M = int(cfg["cluster_size"])
    if M & (M - 1) or M > 32:
        pytest.skip("gpuvcl_pruned needs a power-of-two cluster size")
    pos, own, types = g["pos"], g["own"], g["types"]
    kw = dict(shift=bool(cfg["shift"]), mixing=bool(cfg["mixing"]), eps=g["eps"], sigma=g["sigma"])
    o = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, g["box_min"], g["box_max"], cfg["cutoff"], cfg["skin"], M,
                      newton3=False, **kw)
    c, f = run_gpu("gpuVerletClusterLists", "gpuvcl_pruned", pos, own, types, g["box_min"], g["box_max"], cfg["cutoff"],
                   cfg["skin"], n3, M=M, **kw)
    check_forces(c.forcesById(len(pos)), g["ref_f"], o["fscale"], own)
    assert f.getPotentialEnergy() == pytest.approx(g["ref_upot"].item(), rel=1e-12)
    assert f.getVirial() == pytest.approx(g["ref_virial"].item(), rel=1e-12)
    c.close()


# ---- every GPU configuration against the oracle on a fresh random scenario (TraversalComparison style) --------------
CONFIGS = [("gpuLinkedCells", "gpulc_c08", n3, 0) for n3 in (True, False)] + \
          [("gpuLinkedCells", "gpulc_c18", n3, 0) for n3 in (True, False)] + \
          [("gpuVerletClusterLists", "gpuvcl_cluster_iteration", False, M) for M in (1, 2, 4, 8, 16, 32)] + \
          [("gpuVerletClusterLists", "gpuvcl_c06", n3, M) for n3 in (True, False) for M in (4, 32)] + \
          [("gpuVerletClusterLists", "gpuvcl_c01_balanced", False, 8)] + \
          [("gpuVerletClusterLists", "gpuvcl_pruned", n3, M) for n3 in (False, True) for M in (4, 8, 32)]


@pytest.mark.parametrize("cont,trav,n3,M", CONFIGS)
@pytest.mark.parametrize("n,nh,L", [(100, 200, 3.0), (2000, 200, 10.0)])
def test_gpu_matches_oracle(cont, trav, n3, M, n, nh, L):
    pos, own, types = uniform_with_halo(n, nh, [L, L, L], 1.0, seed=n + 17, ntypes=2)
    kw = dict(shift=True, mixing=True, eps=[1.0, 1.2], sigma=[1.0, 0.95])
    bmin, bmax = [0, 0, 0], [L, L, L]
    if cont == "gpuLinkedCells":
        o = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], types, own, bmin, bmax, 1.0, 0.1, newton3=n3, **kw)
    else:
        o = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, bmin, bmax, 1.0, 0.1, M, newton3=n3, **kw)
    c, f = run_gpu(cont, trav, pos, own, types, bmin, bmax, 1.0, 0.1, n3, M=max(M, 1), **kw)
    check_forces(c.forcesById(len(pos)), o["f"], o["fscale"], own)
    u, v = oracle.lj_end_traversal(o["res"])
    assert f.getPotentialEnergy() == pytest.approx(u, rel=1e-12)
    assert f.getVirial() == pytest.approx(v, rel=1e-12)
    if trav != "gpuvcl_pruned":  # the pruned traversal evaluates fewer distances by design
        assert f.getNumFLOPs() == oracle.lj_num_flops(o["res"], True)
        assert f.getHitRate() == pytest.approx(
            (o["res"].num_kernel_calls_n3 + o["res"].num_kernel_calls_no_n3) / o["res"].num_dist_calls, rel=1e-15)
    if cont == "gpuVerletClusterLists":
        ids, _, _ = c.downloadIds()
        np.testing.assert_array_equal(ids.reshape(-1, max(M, 1)), o["slot_particle"].reshape(-1, max(M, 1)))
        if trav == "gpuvcl_pruned" and n3:
            # gpuvcl_pruned refines the newton3-off cluster-pair list in both modes: that is the list it keeps
            o = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, bmin, bmax, 1.0, 0.1, M, newton3=False, **kw)
            r = f._raw  # every in-cutoff pair once, counted as a newton3 kernel call (LJFunctor.h:518-531)
            assert r.num_kernel_calls_no_n3 == 0 and r.num_global_calcs_n3 == r.num_kernel_calls_n3 > 0
        assert {tuple(p) for p in c.debugClusterPairs()} == {tuple(p) for p in o["pairs"]}
    else:
        ids, _, _ = c.downloadIds()
        np.testing.assert_array_equal(c.debugCellOfSlot(), o["cell"][ids])
    c.close()


def test_cell_size_factor_half():
    pos, own, types = uniform_with_halo(600, 150, [6., 6., 6.], 1.0, seed=5)
    for n3 in (True, False):
        o = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [6, 6, 6], 1.0, 0.2, 0.5,
                                  newton3=n3)
        c, f = run_gpu("gpuLinkedCells", "gpulc_c08", pos, own, None, [0, 0, 0], [6, 6, 6], 1.0, 0.2, n3, csf=0.5)
        check_forces(c.forcesById(len(pos)), o["f"], o["fscale"], own)
        assert f.getNumFLOPs() == oracle.lj_num_flops(o["res"], False)
        c.close()


# ---- edge cases the reference tests ---------------------------------------------------------------------------------
@pytest.mark.parametrize("cont,trav", [("gpuLinkedCells", "gpulc_c08"), ("gpuVerletClusterLists", "gpuvcl_cluster_iteration"),
                                       ("gpuVerletClusterLists", "gpuvcl_pruned")])
def test_empty_and_single_particle(cont, trav):
    c = GpuParticleContainer(cont, [0, 0, 0], [5, 5, 5], 1.0, 0.2)
    f = make_functor(1.0, True, False, 1.0, 1.0)
    t = GpuTraversal(trav, f, False)
    c.rebuildNeighborLists(t)
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(False)
    assert f.getPotentialEnergy() == 0.0 and c.getNumberOfParticles("ownedOrHalo") == 0
    c.addParticles([2.5], [2.5], [2.5])
    c.rebuildNeighborLists(t)
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(False)
    assert c.getNumberOfParticles("owned") == 1
    np.testing.assert_array_equal(c.forcesById(1), np.zeros((1, 3)))
    c.close()


def test_add_outside_box_throws_and_compute_needs_rebuild():
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [5, 5, 5], 1.0, 0.2)
    with pytest.raises(ApbError) as e:  # LogicHandler.h:343-350
        c.addParticles([5.0], [1.0], [1.0])
    assert e.value.code == capi.ERR_PARTICLE_OUTSIDE
    c.addParticles([1.0], [1.0], [1.0])
    f = make_functor(1.0, False, False, 1.0, 1.0)
    with pytest.raises(ApbError) as e:
        c.computeInteractions(GpuTraversal("gpulc_c08", f, True))
    assert e.value.code == capi.ERR_STATE
    with pytest.raises(ApbError) as e:  # traversal of the other container family
        c.rebuildNeighborLists(GpuTraversal("gpuvcl_c06", f, True))
    assert e.value.code == capi.ERR_NOT_APPLICABLE
    c.close()
    v = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [5, 5, 5], 1.0, 0.2)
    with pytest.raises(ApbError) as e:  # CompatibleTraversals.h:142-151
        v.rebuildNeighborLists(GpuTraversal("gpuvcl_cluster_iteration", f, True))
    assert e.value.code == capi.ERR_NOT_APPLICABLE
    v.close()


def test_functor_end_traversal_twice_throws():
    f = make_functor(1.0, False, False, 1.0, 1.0)
    f.initTraversal()
    f.endTraversal(True)
    with pytest.raises(ApbError):  # LJFunctor.h:664-667
        f.endTraversal(True)


@pytest.mark.parametrize("cont,trav,n3", [("gpuLinkedCells", "gpulc_c08", True),
                                          ("gpuVerletClusterLists", "gpuvcl_c06", True),
                                          ("gpuVerletClusterLists", "gpuvcl_cluster_iteration", False),
                                          ("gpuVerletClusterLists", "gpuvcl_pruned", False)])
def test_deleted_particles_and_shift_after_list_build(cont, trav, n3):
    """TraversalComparison.cpp:162-177: mark 30 % deleted before and after the list build, shift by < skin/2 after it;
    forces must equal a fresh evaluation of the final state."""
    rng = np.random.default_rng(9)
    pos, own, _ = uniform_with_halo(1500, 200, [8., 8., 8.], 1.0, seed=21)
    n = len(pos)
    c = GpuParticleContainer(cont, [0, 0, 0], [8, 8, 8], 1.0, 0.2, clusterSize=8)
    ids = np.arange(n)
    mo, mh = own == 1, own == 2
    c.addParticles(pos[mo, 0], pos[mo, 1], pos[mo, 2], ids[mo])
    c.addHaloParticles(pos[mh, 0], pos[mh, 1], pos[mh, 2], ids[mh])
    f = make_functor(1.0, True, False, 1.0, 1.0)
    t = GpuTraversal(trav, f, n3)
    deleted = np.zeros(n, dtype=bool)

    def delete_some(frac):
        sid, _, sown = c.downloadIds()
        kill = (rng.uniform(size=len(sid)) < frac) & (sown != 0)
        sown[kill] = 0
        deleted[sid[kill]] = True
        c.uploadOwnership(sown)

    delete_some(0.15)
    c.rebuildNeighborLists(t)
    sid, _, sown = c.downloadIds()
    shift = rng.uniform(-1, 1, (n, 3))
    shift *= (0.1 * rng.uniform(size=(n, 1))) / np.linalg.norm(shift, axis=1, keepdims=True)  # |shift| < skin/2
    newpos = pos + shift
    live = sid >= 0
    for d, name in enumerate("XYZ"):
        col = c.downloadColumn(name)
        col[live & (sown != 0)] = newpos[sid[live & (sown != 0)], d]
        c.uploadColumn(name, col)
    delete_some(0.15)
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(n3)
    own_final = own.copy()
    own_final[deleted] = 0
    bf = oracle.lj_bruteforce(newpos[:, 0], newpos[:, 1], newpos[:, 2], None, own_final, 1.0, shift=True)
    check_forces(c.forcesById(n), bf["f"], bf["fscale"], own_final)
    u, v = oracle.lj_end_traversal(bf["res"])
    assert f.getPotentialEnergy() == pytest.approx(u, rel=1e-12)
    assert f.getVirial() == pytest.approx(v, rel=1e-12)
    c.close()


@pytest.mark.parametrize("cont", ["gpuLinkedCells", "gpuVerletClusterLists"])
def test_update_container_returns_leavers(cont):
    """LeavingParticleCollector.h:85-118 (keep lists) and LinkedCells.h:152-202 / VerletClusterLists.h:362-397."""
    pos, own, _ = uniform_with_halo(500, 100, [6., 6., 6.], 1.0, seed=2)
    n = len(pos)
    for keep in (True, False):
        c = GpuParticleContainer(cont, [0, 0, 0], [6, 6, 6], 1.0, 0.4, clusterSize=4)
        ids = np.arange(n)
        mo, mh = own == 1, own == 2
        c.addParticles(pos[mo, 0], pos[mo, 1], pos[mo, 2], ids[mo])
        c.addHaloParticles(pos[mh, 0], pos[mh, 1], pos[mh, 2], ids[mh])
        f = make_functor(1.0, False, False, 1.0, 1.0)
        t = GpuTraversal("gpulc_c08" if cont == "gpuLinkedCells" else "gpuvcl_cluster_iteration", f, False)
        c.rebuildNeighborLists(t)
        sid, _, sown = c.downloadIds()
        x = c.downloadColumn("X")
        movers = (sown == 1) & (x > 5.9)
        x[movers] += 0.15  # leaves through the +x face
        c.uploadColumn("X", x)
        c.uploadColumn("FY", sid * 0.25)     # leavers are whole copies: any other column travels with them
        c.uploadColumn("OLDFZ", sid * -0.5)
        expect = set(sid[movers & (x >= 6.0)].tolist())
        leavers = c.updateContainer(keep)
        assert set(leavers["id"].tolist()) == expect
        assert np.all(leavers["x"] >= 6.0)
        assert np.array_equal(c.leaverColumn("X"), leavers["x"]) and np.array_equal(c.leaverColumn("VZ"), leavers["vz"])
        assert np.array_equal(c.leaverColumn("FY"), leavers["id"] * 0.25)
        assert np.array_equal(c.leaverColumn("OLDFZ"), leavers["id"] * -0.5)
        with pytest.raises(ApbError):
            c.leaverColumn("DENSITY")  # not a column of MoleculeLJ storage
        assert c.getNumberOfParticles("halo") == 0
        assert c.getNumberOfParticles("owned") == int(mo.sum()) - len(expect)
        if keep:
            assert c.numSlots() == len(sid)  # storage is not reshuffled
            f.initTraversal()
            c.computeInteractions(t)  # lists stay usable
        else:
            with pytest.raises(ApbError):
                c.computeInteractions(t)
        c.close()


def test_update_halo_particles_by_id():
    pos, own, _ = uniform_with_halo(300, 120, [5., 5., 5.], 1.0, seed=4)
    n = len(pos)
    c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [5, 5, 5], 1.0, 0.3, clusterSize=4)
    ids = np.arange(n)
    mo, mh = own == 1, own == 2
    c.addParticles(pos[mo, 0], pos[mo, 1], pos[mo, 2], ids[mo])
    c.addHaloParticles(pos[mh, 0], pos[mh, 1], pos[mh, 2], ids[mh])
    f = make_functor(1.0, False, False, 1.0, 1.0)
    t = GpuTraversal("gpuvcl_cluster_iteration", f, False)
    c.rebuildNeighborLists(t)
    newh = pos[mh] + 0.05
    missing = c.updateHaloParticles(np.r_[ids[mh], [10 ** 6]], np.r_[newh[:, 0], [0.]], np.r_[newh[:, 1], [0.]],
                                    np.r_[newh[:, 2], [0.]])
    assert missing == 1  # the unknown id has no slot: caller falls back to addHaloParticle (LogicHandler.h:379-389)
    sid, _, sown = c.downloadIds()
    x = c.downloadColumn("X")
    hal = sown == 2
    np.testing.assert_array_equal(x[hal], (pos[:, 0] + 0.05)[sid[hal]])
    c.close()


# ---- size-independent properties at the BASELINE sizes --------------------------------------------------------------
def _periodic_system(n_per_dim, spacing, jitter, cutoff, skin, seed=42):
    pos, bmin, bmax = grid_lattice(n_per_dim, spacing, jitter, seed)
    L = bmax - bmin
    pos = bmin + np.mod(pos - bmin, L)
    hpos, hsrc = periodic_images(pos, bmin, bmax, cutoff + skin)
    return pos, hpos, hsrc, bmin, bmax


@pytest.mark.parametrize("n_per_dim,spacing,cutoff,skin", [(32, 1.1225, 2.5, 0.2), (100, 1.0581, 2.5, 0.3)])
def test_full_size_properties(n_per_dim, spacing, cutoff, skin):
    """C1 (32^3) and C2 (100^3 = 1M) at full size: every container / traversal against the CPU oracle (forces 1e-12 of
    the per-particle sum of pair-force magnitudes, Upot and virial 1e-12), against the committed full-size fixture of
    the unmodified reference (C1) and - when oracle/_ref travelled to the box - against the unmodified reference run
    live on the same particles; plus the size-independent properties: total force of the periodic system vanishes,
    newton3 on / off agree."""
    pos, hpos, hsrc, bmin, bmax = _periodic_system(n_per_dim, spacing, 0.1, cutoff, skin)
    n = len(pos)
    allpos = np.vstack([pos, hpos])
    allown = np.r_[np.ones(n), 2 * np.ones(len(hpos))].astype(np.int64)
    o = oracle.lj_linkedcells(allpos[:, 0], allpos[:, 1], allpos[:, 2], None, allown, bmin, bmax, cutoff, skin, shift=True,
                              newton3=True)
    ou, ov = oracle.lj_end_traversal(o["res"])
    refs = [("oracle", o["f"][:n], ou, ov)]
    if n_per_dim == 32:
        g = np.load(os.path.join(GOLDEN, "c1_full_size.npz"))
        assert int(g["n"]) == n and int(g["num_halo"]) == len(hpos) and float(g["pos_checksum"]) == float(allpos.sum())
        refs.append(("reference fixture", g["ref_f"], float(g["ref_upot"]), float(g["ref_virial"])))
    if oracle.have_ref():
        r = oracle.ref_lj_linkedcells(allpos[:, 0], allpos[:, 1], allpos[:, 2], None, allown, bmin, bmax, cutoff, skin, 1.0,
                                      shift=True, newton3=True, soa=True)
        refs.append(("reference live", r["f"][:n], r["upot"], r["virial"]))
    results = {}
    for cont, trav, n3, M in [("gpuLinkedCells", "gpulc_c08", False, 0), ("gpuLinkedCells", "gpulc_c18", True, 0),
                              ("gpuVerletClusterLists", "gpuvcl_cluster_iteration", False, 4),
                              ("gpuVerletClusterLists", "gpuvcl_c06", True, 32),
                              ("gpuVerletClusterLists", "gpuvcl_pruned", False, 32),
                              ("gpuVerletClusterLists", "gpuvcl_pruned", True, 32)]:
        c = GpuParticleContainer(cont, bmin, bmax, cutoff, skin, clusterSize=max(M, 1))
        c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
        c.addHaloParticles(hpos[:, 0], hpos[:, 1], hpos[:, 2], n + np.arange(len(hpos)))
        f = make_functor(cutoff, True, False, 1.0, 1.0, flops=False)
        t = GpuTraversal(trav, f, n3)
        c.rebuildNeighborLists(t)
        f.initTraversal()
        c.computeInteractions(t)
        f.endTraversal(n3)
        F = c.forcesById(n + len(hpos))[:n]
        results[(cont, trav, n3)] = (F, f.getPotentialEnergy(), f.getVirial())
        assert c.getNumberOfParticles("owned") == n
        c.close()
        for name, rf, ru, rv in refs:
            err = np.abs(F - rf).max(axis=1)
            assert np.all(err <= FTOL * o["fscale"][:n] + 1e-300), (cont, trav, n3, name, np.max(err / o["fscale"][:n]))
            assert f.getPotentialEnergy() == pytest.approx(ru, rel=1e-12), (cont, trav, n3, name)
            assert f.getVirial() == pytest.approx(rv, rel=1e-12), (cont, trav, n3, name)
    keys = list(results)
    F0, u0, v0 = results[keys[0]]
    fmax = np.abs(F0).max()
    assert np.abs(F0.sum(axis=0)).max() <= 1e-9 * fmax * np.sqrt(n)  # momentum conservation of the periodic system
    for k in keys[1:]:
        F, u, v = results[k]
        assert np.abs(F - F0).max() <= 1e-11 * fmax, k
        assert u == pytest.approx(u0, rel=1e-12), k
        assert v == pytest.approx(v0, rel=1e-12), k
    assert u0 < 0  # a liquid-density LJ lattice is bound


# ---- gpuvcl_pruned specifics ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mixing", [False, True])
def test_pruned_virial_trace_only(mixing):
    """APB_FUNCTOR_VIRIAL_TRACE: only the sum of the virial components (all LJFunctor::getVirial returns,
    LJFunctor.h:719) is accumulated; Upot, virial, forces and counters must equal the per-component run."""
    pos, own, types = uniform_with_halo(3000, 400, [9., 9., 9.], 1.0, seed=77, ntypes=2)
    out = []
    for trace in (False, True):
        c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [9, 9, 9], 1.0, 0.15, clusterSize=32)
        ids = np.arange(len(pos))
        mo, mh = own == 1, own == 2
        c.addParticles(pos[mo, 0], pos[mo, 1], pos[mo, 2], ids[mo], types[mo])
        c.addHaloParticles(pos[mh, 0], pos[mh, 1], pos[mh, 2], ids[mh], types[mh])
        f = make_functor(1.0, True, mixing, [1.0, 1.3], [1.0, 0.9])
        f.virialTraceOnly = trace
        t = GpuTraversal("gpuvcl_pruned", f, False)
        c.rebuildNeighborLists(t)
        f.initTraversal()
        c.computeInteractions(t)
        f.endTraversal(False)
        out.append((c.forcesById(len(pos)), f.getPotentialEnergy(), f.getVirial(), f.getNumFLOPs()))
        c.close()
    np.testing.assert_array_equal(out[0][0], out[1][0])
    assert out[0][1] == out[1][1]
    assert out[1][2] == pytest.approx(out[0][2], rel=1e-12)
    assert out[0][3] == out[1][3]
    kw = dict(shift=True, mixing=mixing, eps=[1.0, 1.3], sigma=[1.0, 0.9])
    o = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types if mixing else None, own, [0, 0, 0], [9, 9, 9], 1.0, 0.15, 32,
                      newton3=False, **kw)
    u, v = oracle.lj_end_traversal(o["res"])
    assert out[1][1] == pytest.approx(u, rel=1e-12)
    assert out[1][2] == pytest.approx(v, rel=1e-12)


@pytest.mark.parametrize("cutoff", [1.0, 2.5, 1.7])
def test_pruned_cutoff_decision_is_bit_exact(cutoff):
    """Pairs whose squared distance lies within a few ulp of cutoff^2: the pruned kernel evaluates dr2 with FMAs and
    re-evaluates such pairs with separately rounded products and sums, so the number of kernel calls (pairs with
    dr2 <= cutoff^2, LJFunctor.h:149 / :494) must equal the oracle's exactly. Scenario: 8 centres, each with 500
    particles on the sphere of radius cutoff around it (coordinates of the order of the cutoff, so dr2 scatters by a
    few ulp around cutoff^2)."""
    rng = np.random.default_rng(int(cutoff * 10))
    k = np.arange(500) + 0.5
    phi, theta = np.arccos(1 - 2 * k / 500), np.pi * (1 + 5 ** 0.5) * k
    sphere = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)
    centres = (np.stack(np.meshgrid(*[np.arange(2)] * 3, indexing="ij"), axis=-1).reshape(-1, 3) * 4.0 + 2.0) * cutoff
    centres = centres + rng.uniform(-0.01, 0.01, centres.shape)
    parts, near, inside, outside = [centres], 0, 0, 0
    for cpos in centres:
        p = cpos + sphere * (cutoff * (1.0 + rng.integers(-3, 4, 500) * 1.1e-16))[:, None]
        d = cpos - p
        dr2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        near += int((np.abs(dr2 - cutoff * cutoff) <= 4 * np.spacing(cutoff * cutoff)).sum())
        inside += int((dr2 <= cutoff * cutoff).sum())
        outside += int((dr2 > cutoff * cutoff).sum())
        parts.append(p)
    assert near > 1000 and inside > 400 and outside > 400, (near, inside, outside)
    pos = np.vstack(parts)
    L = 8.0 * cutoff
    own = np.ones(len(pos), dtype=np.int64)
    bf = oracle.lj_bruteforce(pos[:, 0], pos[:, 1], pos[:, 2], None, own, cutoff, shift=True)
    c, f = run_gpu("gpuVerletClusterLists", "gpuvcl_pruned", pos, own, None, [0, 0, 0], [L, L, L], cutoff, 0.2 * cutoff,
                   False, shift=True, M=32)
    assert f._raw.num_kernel_calls_no_n3 == 2 * bf["res"].num_kernel_calls_n3  # newton3 off: both directions
    check_forces(c.forcesById(len(pos)), bf["f"], bf["fscale"], own)
    c.close()
    c, f = run_gpu("gpuVerletClusterLists", "gpuvcl_pruned", pos, own, None, [0, 0, 0], [L, L, L], cutoff, 0.2 * cutoff,
                   True, shift=True, M=32)
    assert f._raw.num_kernel_calls_n3 == bf["res"].num_kernel_calls_n3  # newton3: every pair once
    check_forces(c.forcesById(len(pos)), bf["f"], bf["fscale"], own)
    u, v = oracle.lj_end_traversal(bf["res"])
    assert f.getPotentialEnergy() == pytest.approx(u, rel=1e-12) and f.getVirial() == pytest.approx(v, rel=1e-12)
    c.close()


@pytest.mark.parametrize("newton3", [True, False])
def test_gpu_flop_counter_literals(newton3):
    """LJFunctorFlopCounterTest.cpp:35-170 through gpuLinkedCells / gpulc_c08: the counters of the four-molecule
    scenario equal the reference's literals (6 distance calls, 2 newton3 kernel calls | 10, 1 + 2)."""
    pos = np.array([[0.9, 0.3, 0.9], [0.9, 0.9, 0.9], [0.9, 1.5, 0.9], [0.1, 2.4, 0.1]])
    own = np.ones(4, dtype=np.int64)
    c, f = run_gpu("gpuLinkedCells", "gpulc_c08", pos, own, None, [0, 0, 0], [3, 3, 3], 1.1, 0.2, newton3, shift=True)
    r = f._raw
    dist, kn3, knon3 = (6, 2, 0) if newton3 else (10, 1, 2)
    assert (r.num_dist_calls, r.num_kernel_calls_n3, r.num_kernel_calls_no_n3) == (dist, kn3, knon3)
    assert (r.num_global_calcs_n3, r.num_global_calcs_no_n3) == (kn3, knon3)
    assert f.getNumFLOPs() == 8 * dist + 18 * kn3 + 15 * knon3 + 13 * kn3 + 9 * knon3
    assert f.getHitRate() == pytest.approx((kn3 + knon3) / dist, abs=1e-14)
    c.close()


@pytest.mark.parametrize("scale", [1e-3, 1.0, 400.0])
def test_pruned_reciprocal_is_accurate_over_the_whole_range(scale):
    """kLJPruned replaces the division 1 / dr2 by MUFU.RCP64H + one cubic refinement step (pruned.cu: prRcp). Pairs at
    distances from 1e-3 cutoff up to the cutoff, in units that put dr2 between 1e-12 and 1e+5 (tiny and huge arguments
    of the reciprocal): forces against the oracle's IEEE division, 1e-12 of the sum of pair-force magnitudes."""
    rng = np.random.default_rng(int(scale * 1000) + 1)
    cutoff = 2.5 * scale
    n_c = 40
    centres = rng.uniform(0.2, 0.8, (n_c, 3)) * 40 * cutoff
    parts = [centres]
    for r_rel in (1e-3, 3e-3, 0.03, 0.2, 0.5, 0.9, 0.999):
        d = rng.normal(size=(n_c, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        parts.append(centres + d * r_rel * cutoff)
    pos = np.vstack(parts)
    L = 40 * cutoff
    own = np.ones(len(pos), dtype=np.int64)
    # sigma = scale: the potential keeps its shape, only the units change
    kw = dict(shift=True, mixing=True, eps=[1.0, 0.8], sigma=[scale, 0.9 * scale])
    types = rng.integers(0, 2, len(pos)).astype(np.int64)
    bf = oracle.lj_bruteforce(pos[:, 0], pos[:, 1], pos[:, 2], types, own, cutoff, **kw)
    for n3 in (False, True):
        c, f = run_gpu("gpuVerletClusterLists", "gpuvcl_pruned", pos, own, types, [0, 0, 0], [L, L, L], cutoff, 0.1 * cutoff,
                       n3, M=32, **kw)
        check_forces(c.forcesById(len(pos)), bf["f"], bf["fscale"], own)
        u, v = oracle.lj_end_traversal(bf["res"])
        assert f.getPotentialEnergy() == pytest.approx(u, rel=1e-12) and f.getVirial() == pytest.approx(v, rel=1e-12)
        c.close()


def test_pruned_capacity_limits_on_dense_blobs():
    """Inhomogeneous systems (droplets): a blob whose tiles stay within the staging capacity of gpuvcl_pruned must give
    the oracle's forces; one that exceeds it (more than 4080 particles referenced by one tile) must be rejected with
    APB_ERR_NOT_APPLICABLE - the tuner then drops the configuration (TraversalSelector.h:353-356) - while the
    list-faithful traversal of the same container still runs."""
    rng = np.random.default_rng(8)
    L, rc, skin = 12.0, 2.5, 0.3

    def blob(n, radius):
        d = rng.normal(size=(n, 3))
        d *= (radius * rng.uniform(0, 1, n) ** (1 / 3) / np.linalg.norm(d, axis=1))[:, None]
        gas = rng.uniform(0, L, (300, 3))
        return np.vstack([d + L / 2, gas])

    # (a) dense but within capacity: 2500 particles inside radius 2 (density 75: every particle sees ~2000 others)
    pos = blob(2500, 2.0)
    own = np.ones(len(pos), dtype=np.int64)
    bf = oracle.lj_bruteforce(pos[:, 0], pos[:, 1], pos[:, 2], None, own, rc, shift=True)
    c, f = run_gpu("gpuVerletClusterLists", "gpuvcl_pruned", pos, own, None, [0, 0, 0], [L, L, L], rc, skin, False, shift=True,
                   M=32)
    check_forces(c.forcesById(len(pos)), bf["f"], bf["fscale"], own)
    assert f._raw.num_kernel_calls_no_n3 == 2 * bf["res"].num_kernel_calls_n3
    c.close()
    # (b) beyond capacity: 7000 particles inside radius 2.2
    pos = blob(7000, 2.2)
    own = np.ones(len(pos), dtype=np.int64)
    c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [L, L, L], rc, skin, clusterSize=32)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(len(pos)))
    f = make_functor(rc, True, False, 1.0, 1.0)
    with pytest.raises(ApbError) as e:
        c.rebuildNeighborLists(GpuTraversal("gpuvcl_pruned", f, False))
    assert e.value.code == capi.ERR_NOT_APPLICABLE
    t = GpuTraversal("gpuvcl_c06", f, False)  # same handle, list-faithful traversal: still applicable
    c.rebuildNeighborLists(t)
    f.initTraversal()
    c.computeInteractions(t)
    f.endTraversal(False)
    bf = oracle.lj_bruteforce(pos[:, 0], pos[:, 1], pos[:, 2], None, own, rc, shift=True)
    check_forces(c.forcesById(len(pos)), bf["f"], bf["fscale"], own)
    c.close()
