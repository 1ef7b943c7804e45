"""Pins the CPU oracle (oracle/autopas_oracle.c) against the reference's own golden vectors, against fixtures generated
from the unmodified reference (tests/golden/make_golden.py) and — when oracle/_ref was built — against the reference live.
No GPU needed."""
import os

import numpy as np
import pytest

import oracle
from scenarios import uniform_with_halo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- LJFunctorTestNoGlobals.h:27-31 (abs tolerance 1e-7 as in the reference test) -------------------------------
EXPECTED_FORCE = np.array([-4547248.8989645941, -9094497.7979291882, -13641746.696893783])
EXPECTED_FORCE_MIXING = np.array([-835415983.7676939964294, -1670831967.5353879928588, -2506247951.3030819892883])


@pytest.mark.parametrize("newton3", [True, False])
def test_lj_golden_force(newton3):
    x, y, z = np.array([1.0, 1.1]), np.array([1.0, 1.2]), np.array([1.0, 1.3])
    r = oracle.lj_linkedcells(x, y, z, None, None, [0, 0, 0], [5, 5, 5], 1.0, 0.2, newton3=newton3)
    np.testing.assert_allclose(r["f"][0], EXPECTED_FORCE, atol=1e-7, rtol=0)
    np.testing.assert_allclose(r["f"][1], -EXPECTED_FORCE, atol=1e-7, rtol=0)


def test_lj_golden_force_mixing():
    # LJFunctorTestNoGlobals.cpp:19-32: types (eps, sigma) = (1, 1) and (2, 2); p2 has type 1
    x, y, z = np.array([1.0, 1.1]), np.array([1.0, 1.2]), np.array([1.0, 1.3])
    r = oracle.lj_linkedcells(x, y, z, np.array([0, 1]), None, [0, 0, 0], [5, 5, 5], 1.0, 0.2, mixing=True,
                              eps=[1.0, 2.0], sigma=[1.0, 2.0])
    np.testing.assert_allclose(r["f"][0], EXPECTED_FORCE_MIXING, atol=1e-5, rtol=1e-15)
    np.testing.assert_allclose(r["f"][1], -EXPECTED_FORCE_MIXING, atol=1e-5, rtol=1e-15)


def _lj_potential(pi, pj, cutoff, sigma, eps):
    # tests/testAutopas/testingHelpers/LJPotential.h:10-44 (closed form used by LJFunctorTestGlobals.cpp)
    d = np.linalg.norm(pi - pj)
    if d > cutoff:
        return 0.0
    sdr = sigma / d
    lj6 = sdr ** 6
    return 4.0 * eps * (lj6 * lj6 - lj6)


def _lj_virial(pi, pj, cutoff, sigma, eps):
    d = pi - pj
    r = np.linalg.norm(d)
    if r > cutoff:
        return 0.0
    f = 24 * eps * (2 * (sigma / r) ** 12 - (sigma / r) ** 6) / r ** 2 * d
    return float(np.dot(d, f))


@pytest.mark.parametrize("where", ["inside", "boundary"])
def test_lj_globals_closed_form(where):
    # LJFunctorTestGlobals.cpp:67-72, 353-359: owned/owned pair counts fully, owned/halo pair counts half
    p1 = np.array([1.0, 1.0, 1.0])
    p2 = np.array([1.1, 1.2, 1.3]) if where == "inside" else np.array([-0.1, 1.2, 1.3])
    pos = np.vstack([p1, p2])
    own = np.array([1, 1 if where == "inside" else 2])
    factor = 1.0 if where == "inside" else 0.5
    for n3 in (True, False):
        r = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [5, 5, 5], 2.0, 0.2, newton3=n3,
                                  eps=1.0, sigma=1.0)
        upot, virial = oracle.lj_end_traversal(r["res"])
        assert upot == pytest.approx(factor * _lj_potential(p1, p2, 2.0, 1.0, 1.0), rel=1e-13)
        assert virial == pytest.approx(factor * _lj_virial(p1, p2, 2.0, 1.0, 1.0), rel=1e-13)


# ---- CellBlock3DTest.cpp:42-141 ----------------------------------------------------------------------------------
CELLBLOCKS = {
    "1x1x1": ([0, 0, 0], [10, 10, 10], 10.0, 1.0),
    "1x1x1_cs2": ([0, 0, 0], [10, 10, 10], 5.0, 2.0),
    "2x2x2": ([0, 0, 0], [10, 10, 10], 5.0, 1.0),
    "2x2x2_cs05": ([0, 0, 0], [10, 10, 10], 10.0, 0.5),
    "3x3x3": ([0, 0, 0], [10, 10, 10], 3.0, 1.0),
    "11x4x4": ([2. / 3., 0, 0], [1., .125, .125], 3. / 100., 1.0),
    "19x19x19": ([0, 0, 0], [58.5, 58.5, 58.5], 3.0, 1.0),
}


@pytest.mark.parametrize("name,start,dr,num", [
    ("1x1x1", -5., 10., 3), ("1x1x1_cs2", -5., 10., 3), ("2x2x2", -2.5, 5., 4), ("2x2x2_cs05", -7.5, 5., 6),
    ("3x3x3", -1.6, 3.3, 5)])
def test_cellblock_index_mesh(name, start, dr, num):
    bmin, bmax, il, csf = CELLBLOCKS[name]
    zz, yy, xx = np.meshgrid(np.arange(num), np.arange(num), np.arange(num), indexing="ij")
    x, y, z = start + xx.ravel() * dr, start + yy.ravel() * dr, start + zz.ravel() * dr
    cell, cpd = oracle.lc_cell_indices(bmin, bmax, il, csf, x, y, z)
    np.testing.assert_array_equal(cell, np.arange(num ** 3))
    assert tuple(cpd) == (num, num, num)


@pytest.mark.parametrize("name", list(CELLBLOCKS))
def test_cellblock_boundaries(name):
    bmin, bmax, il, csf = CELLBLOCKS[name]
    bmin, bmax = np.array(bmin, float), np.array(bmax, float)
    shifts = [np.nextafter(bmin, -1.), bmin, np.nextafter(bmax, -1.), bmax]
    _, cpd = oracle.lc_cell_indices(bmin, bmax, il, csf, [bmin[0]], [bmin[1]], [bmin[2]])
    halo = 1 if csf >= 1 else int(np.ceil(1 / csf))
    expected = [lambda d: halo - 1, lambda d: halo, lambda d: cpd[d] - halo - 1, lambda d: cpd[d] - halo]
    for a in range(4):
        for b in range(4):
            for c in range(4):
                p = [shifts[a][0], shifts[b][1], shifts[c][2]]
                cell, _ = oracle.lc_cell_indices(bmin, bmax, il, csf, [p[0]], [p[1]], [p[2]])
                ix = cell[0] % cpd[0]
                iy = (cell[0] // cpd[0]) % cpd[1]
                iz = cell[0] // (cpd[0] * cpd[1])
                assert (ix, iy, iz) == (expected[a](0), expected[b](1), expected[c](2))


# ---- VerletClusterListsTest.cpp:128-256 properties ---------------------------------------------------------------
@pytest.mark.parametrize("M", [4, 32])
def test_vcl_pair_set_covers_brute_force(M):
    pos, own, _ = uniform_with_halo(500, 100, [6., 6., 6.], 1.0, seed=3)
    cutoff, skin = 1.0, 0.2
    r = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [6, 6, 6], cutoff, skin, M, newton3=False)
    slot_particle = r["slot_particle"].reshape(-1, M)
    covered = set()
    for A, B in r["pairs"]:
        for i in slot_particle[A]:
            for j in slot_particle[B]:
                if i >= 0 and j >= 0:
                    covered.add((int(i), int(j)))
    for c in slot_particle:
        for i in c:
            for j in c:
                if i >= 0 and j >= 0 and i != j:
                    covered.add((int(i), int(j)))
    il2 = (cutoff + skin) ** 2
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    need = np.argwhere(d2 <= il2)
    for i, j in need:
        if i == j or (own[i] == 2 and own[j] == 2):
            continue
        if own[i] == 1:  # newton3 off: every owned particle's list must hold all its partners
            assert (int(i), int(j)) in covered
    # newton3 lists hold each interacting cluster pair once: half the non-newton3 entries between owned clusters
    r3 = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [6, 6, 6], cutoff, skin, M, newton3=True)
    und = {tuple(sorted(p)) for p in map(tuple, r3["pairs"])}
    assert len(und) == r3["num_pairs"]
    und_no = {tuple(sorted(p)) for p in map(tuple, r["pairs"])}
    assert und_no <= und


@pytest.mark.parametrize("container", ["lc", "vcl4", "vcl32"])
@pytest.mark.parametrize("newton3", [True, False])
def test_traversals_agree_with_brute_force(container, newton3):
    # TraversalComparison.cpp:229-281: every configuration against one ground truth at 1e-10; we hold 1e-12
    pos, own, types = uniform_with_halo(800, 200, [7., 7., 7.], 1.0, seed=7, ntypes=2)
    kw = dict(shift=True, mixing=True, eps=[1.0, 1.3], sigma=[1.0, 0.9])
    bf = oracle.lj_bruteforce(pos[:, 0], pos[:, 1], pos[:, 2], types, own, 1.0, **kw)
    if container == "lc":
        r = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], types, own, [0, 0, 0], [7, 7, 7], 1.0, 0.1,
                                  newton3=newton3, **kw)
    else:
        r = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, [0, 0, 0], [7, 7, 7], 1.0, 0.1,
                          4 if container == "vcl4" else 32, newton3=newton3, **kw)
    m = own == 1
    err = np.abs(r["f"][m] - bf["f"][m]).max(axis=1)
    assert np.all(err <= 1e-12 * bf["fscale"][m] + 1e-300)
    u, v = oracle.lj_end_traversal(r["res"])
    ub, vb = oracle.lj_end_traversal(bf["res"])
    assert u == pytest.approx(ub, rel=1e-12)
    assert v == pytest.approx(vb, rel=1e-12)


# ---- fixtures generated from the unmodified reference -------------------------------------------------------------
def _golden_files():
    if not os.path.isdir(GOLDEN):
        return []
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("fn_", "c1_")))


@pytest.mark.parametrize("fname", _golden_files())
def test_oracle_matches_reference_fixture(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    cfg = {k: g[k].item() for k in ("cutoff", "skin", "shift", "mixing", "newton3", "cluster_size", "csf")}
    pos, own, types = g["pos"], g["own"], g["types"]
    kw = dict(shift=bool(cfg["shift"]), mixing=bool(cfg["mixing"]), newton3=bool(cfg["newton3"]), eps=g["eps"],
              sigma=g["sigma"])
    if str(g["container"]) == "LinkedCells":
        r = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], types, own, g["box_min"], g["box_max"], cfg["cutoff"],
                                  cfg["skin"], cfg["csf"], **kw)
        np.testing.assert_array_equal(r["cell"][own != 0], g["ref_cell"][own != 0])
    else:
        M = int(cfg["cluster_size"])
        r = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, g["box_min"], g["box_max"], cfg["cutoff"],
                          cfg["skin"], M, **kw)
        np.testing.assert_array_equal(r["slot_particle"].reshape(-1, M), g["ref_cluster_particles"])
        assert {tuple(p) for p in r["pairs"]} == {tuple(p) for p in g["ref_pairs"]}
        assert r["num_pairs"] == len(g["ref_pairs"])
    m = own == 1
    err = np.abs(r["f"][m] - g["ref_f"][m]).max(axis=1)
    assert np.all(err <= 1e-12 * r["fscale"][m] + 1e-300)
    u, v = oracle.lj_end_traversal(r["res"])
    assert u == pytest.approx(g["ref_upot"].item(), rel=1e-12)
    assert v == pytest.approx(g["ref_virial"].item(), rel=1e-12)
    assert oracle.lj_num_flops(r["res"], cfg["shift"]) == int(g["ref_flops"])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref was not built (reference tree absent)")
@pytest.mark.parametrize("newton3", [True, False])
@pytest.mark.parametrize("container", ["lc", "vcl"])
def test_oracle_matches_reference_live(container, newton3):
    pos, own, types = uniform_with_halo(1500, 300, [8., 8., 8.], 1.0, seed=11)
    args = (pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [8, 8, 8], 1.0, 0.1)
    if container == "lc":
        o = oracle.lj_linkedcells(*args, 1.0, shift=True, newton3=newton3)
        r = oracle.ref_lj_linkedcells(*args, 1.0, shift=True, newton3=newton3, soa=True)
        np.testing.assert_array_equal(o["cell"], r["cell"])
    else:
        o = oracle.lj_vcl(*args, 8, shift=True, newton3=newton3)
        r = oracle.ref_lj_vcl(*args, 8, shift=True, newton3=newton3, soa=True,
                              traversal="vcl_c06" if newton3 else "vcl_cluster_iteration")
        np.testing.assert_array_equal(o["slot_particle"].reshape(-1, 8), r["cluster_particles"])
        assert {tuple(p) for p in o["pairs"]} == {tuple(p) for p in r["pairs"]}
    m = own == 1
    err = np.abs(o["f"][m] - r["f"][m]).max(axis=1)
    assert np.all(err <= 1e-12 * o["fscale"][m] + 1e-300)
    u, v = oracle.lj_end_traversal(o["res"])
    assert u == pytest.approx(r["upot"], rel=1e-12)
    assert v == pytest.approx(r["virial"], rel=1e-12)
    assert oracle.lj_num_flops(o["res"], True) == r["flops"]


# ---- cell-pair stencil against the reference's c08 offset tables (LCC08CellHandlerUtilityTest.cpp:27-233) -----------
C08_DIFFS_OVERLAP1 = [0, 1, 11, 12, 13, 131, 132, 133, 143, 144, 145, 155, 156, 157]
C08_DIFFS_OVERLAP2 = [
    0, 1, 2, 10, 11, 12, 13, 14, 22, 23, 24, 25, 26, 118, 119, 120, 121, 122, 130, 131, 132,
    133, 134, 142, 143, 144, 145, 146, 154, 155, 156, 157, 158, 166, 167, 168, 169, 170, 262, 263, 264, 265,
    266, 274, 275, 276, 277, 278, 286, 287, 288, 289, 290, 298, 299, 300, 301, 302, 310, 311, 312, 313, 314]
C08_DIFFS_OVERLAP3 = [
    0, 1, 2, 3, 9, 10, 11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 26, 27, 33, 34, 35,
    36, 37, 38, 39, 105, 106, 107, 108, 109, 110, 111, 117, 118, 119, 120, 121, 122, 123, 129, 130, 131,
    132, 133, 134, 135, 141, 142, 143, 144, 145, 146, 147, 153, 154, 155, 156, 157, 158, 159, 165, 166, 167,
    168, 169, 170, 171, 177, 178, 179, 180, 181, 182, 183, 249, 250, 251, 252, 253, 254, 255, 261, 262, 263,
    264, 265, 266, 267, 273, 274, 275, 276, 277, 278, 279, 285, 286, 287, 288, 289, 290, 291, 297, 298, 299,
    300, 301, 302, 303, 309, 310, 311, 312, 313, 314, 315, 321, 322, 323, 324, 325, 326, 327, 394, 395, 396,
    397, 398, 405, 406, 407, 408, 409, 410, 411, 417, 418, 419, 420, 421, 422, 423, 429, 430, 431, 432, 433,
    434, 435, 441, 442, 443, 444, 445, 446, 447, 453, 454, 455, 456, 457, 458, 459, 466, 467, 468, 469, 470]


@pytest.mark.parametrize("il,expected", [(1.0, C08_DIFFS_OVERLAP1), (2.0, C08_DIFFS_OVERLAP2), (3.0, C08_DIFFS_OVERLAP3)])
def test_cell_pair_offsets_match_c08_tables(il, expected):
    """The reference flattens its c08 offset pairs to sorted offset differences for 12^3 cells of length 1
    (LCC08CellHandlerUtilityTest.h:31-40); the oracle's cell-pair stencil must produce the same tables (14 / 63 / 168
    pairs for overlap 1 / 2 / 3, the latter without the four far corners)."""
    got = oracle.lc_pair_offsets([12, 12, 12], [1.0, 1.0, 1.0], il)
    assert list(got) == expected


def test_vcl_grid_alignment_literals():
    """VerletClusterListsTest.cpp:258-308 (testGridAlignment): box 10^3, cutoff 2, skin 0.05, cluster size 4 and 257
    particles give 5 x 5 owned towers of side 2 with 2 halo towers on each side (9 x 9), and the four corner probes fall
    into towers (2,2), (1,1), (6,6), (7,7)."""
    rng = np.random.default_rng(3)
    probes = np.array([[0., 0., 0.], [-0.1, -0.1, -0.1], [9.9, 9.9, 9.9], [10., 10., 10.]])
    filler = 5.0 + rng.uniform(-1.5, 1.5, (253, 3))  # the reference puts them all at the centre; only the count matters
    pos = np.vstack([probes, filler])
    own = np.ones(len(pos), dtype=np.int64)
    own[[1, 3]] = 2
    o = oracle.lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [10, 10, 10], 2.0, 0.05, 4)
    assert o["towers_per_dim"] == (9, 9)
    assert o["tower_side"] == (2.0, 2.0)
    tower_of = {int(p): int(t) for p, t in zip(o["slot_particle"], o["slot_tower"]) if p >= 0}
    assert [tower_of[i] for i in range(4)] == [2 + 2 * 9, 1 + 1 * 9, 6 + 6 * 9, 7 + 7 * 9]


def test_vcl_newton3_list_is_half_of_the_non_newton3_list():
    """VerletClusterListsTest.cpp:215-256 (testNewton3NeighborList): 2431 uniform particles in a 3^3 box, cutoff 1,
    skin 0.1, cluster size 4, no halos: the non-newton3 list has twice the entries of the newton3 list, and every
    newton3 entry (A, B) appears in the non-newton3 list as (A, B) and as (B, A)."""
    rng = np.random.default_rng(42)
    pos = rng.uniform(0, 3, (2431, 3))
    own = np.ones(len(pos), dtype=np.int64)
    args = (pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [3, 3, 3], 1.0, 0.1, 4)
    non3 = {tuple(p) for p in oracle.lj_vcl(*args, newton3=False)["pairs"]}
    n3 = {tuple(p) for p in oracle.lj_vcl(*args, newton3=True)["pairs"]}
    assert len(non3) == 2 * len(n3)
    for a, b in n3:
        assert (a, b) in non3 and (b, a) in non3


@pytest.mark.parametrize("newton3", [True, False])
@pytest.mark.parametrize("shift", [True, False])
def test_flop_counter_literals(newton3, shift):
    """LJFunctorFlopCounterTest.cpp:35-170, LinkedCells / lc_c08 / SoA: four molecules in a 3^3 box (cutoff 1.1,
    skin 0.2) give 6 distance calls and 2 newton3 kernel calls with newton3, and 10 distance calls, 1 newton3 (inside a
    cell the SoA functor always uses newton3) + 2 non-newton3 kernel calls without; globals are counted per kernel call;
    FLOPs = 8 D + 18 K_n3 + 15 K_non3 + (12 | 13 with shift) G_n3 + (8 | 9) G_non3; hit rate = kernel calls / D."""
    pos = np.array([[0.9, 0.3, 0.9], [0.9, 0.9, 0.9], [0.9, 1.5, 0.9], [0.1, 2.4, 0.1]])
    own = np.ones(4, dtype=np.int64)
    o = oracle.lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], None, own, [0, 0, 0], [3, 3, 3], 1.1, 0.2, 1.0,
                              shift=shift, newton3=newton3)
    r = o["res"]
    dist, kn3, knon3 = (6, 2, 0) if newton3 else (10, 1, 2)
    assert (r.num_dist_calls, r.num_kernel_calls_n3, r.num_kernel_calls_no_n3) == (dist, kn3, knon3)
    assert (r.num_global_calcs_n3, r.num_global_calcs_no_n3) == (kn3, knon3)
    expected = 8 * dist + 18 * kn3 + 15 * knon3 + (13 if shift else 12) * kn3 + (9 if shift else 8) * knon3
    assert oracle.lj_num_flops(r, shift) == expected


def test_oracle_matches_full_size_reference_fixture_c1():
    """The C restatement against the unmodified reference at BASELINE configs[0] size (32 768 particles + periodic
    images): forces 1e-12 of the per-particle sum of pair-force magnitudes, Upot / virial 1e-12, cell indices exact."""
    import os
    from scenarios import grid_lattice, periodic_images
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_full_size.npz"))
    rc, skin = float(g["cutoff"]), float(g["skin"])
    pos, bmin, bmax = grid_lattice(32, 1.1225, 0.1, 42)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    hpos, _ = periodic_images(pos, bmin, bmax, rc + skin)
    allpos = np.vstack([pos, hpos])
    assert float(allpos.sum()) == float(g["pos_checksum"])
    own = np.r_[np.ones(len(pos)), 2 * np.ones(len(hpos))].astype(np.int64)
    o = oracle.lj_linkedcells(allpos[:, 0], allpos[:, 1], allpos[:, 2], None, own, bmin, bmax, rc, skin, shift=True, newton3=True)
    n = len(pos)
    err = np.abs(o["f"][:n] - g["ref_f"]).max(axis=1)
    assert np.all(err <= 1e-12 * o["fscale"][:n])
    u, v = oracle.lj_end_traversal(o["res"])
    assert u == pytest.approx(float(g["ref_upot"]), rel=1e-12) and v == pytest.approx(float(g["ref_virial"]), rel=1e-12)
    np.testing.assert_array_equal(o["cell"], g["ref_cell"])
