"""GPU parity of the functors other than single-site LJ (autopas_b200/csrc/functors.cu) through the C ABI: against the
CPU oracle on seeded scenarios and against fixtures produced by the unmodified reference (tests/golden/fn_*.npz).
Error norm: |gpu - ref| <= 1e-12 * (sum of the magnitudes of the particle's contributions)."""
import os

import numpy as np
import pytest

import oracle
from autopas_b200 import (ApbError, AxilrodTellerMutoFunctor, GpuParticleContainer, GpuTraversal, LJMultisiteFunctor,
                          ParticlePropertiesLibrary, SPHCalcDensityFunctor, SPHCalcHydroForceFunctor, capi)
from functor_scenarios import atm_scenario, multisite_scenario, sph_scenario

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(a, b, scale, tol=1e-12):
    a, b, s = np.asarray(a), np.asarray(b), np.asarray(scale)
    if a.ndim == 2:
        s = s[:, None]
    return np.all(np.abs(a - b) <= tol * s + 1e-300)


def make_container(s, kind, types=None):
    c = GpuParticleContainer("gpuLinkedCells", s["box_min"], s["box_max"], s["cutoff"], s["skin"], particleKind=kind)
    pos, own = s["pos"], s["own"]
    ids = np.arange(len(pos))
    o, h = own == 1, own == 2
    t = None if types is None else np.asarray(types, dtype=np.int32)
    c.addParticles(pos[o, 0], pos[o, 1], pos[o, 2], ids[o], None if t is None else t[o])
    c.addHaloParticles(pos[h, 0], pos[h, 1], pos[h, 2], ids[h], None if t is None else t[h])
    return c


def upload_by_id(c, **columns):
    ids, _, _ = c.downloadIds()
    for name, values in columns.items():
        c.uploadColumn(name, np.asarray(values, dtype=np.float64)[ids])
    return ids


def by_id(c, name, ids, n):
    out = np.zeros(n)
    out[ids] = c.downloadColumn(name)
    return out


def run(c, trav_name, functor, n3):
    t = GpuTraversal(trav_name, functor, n3)
    functor.initTraversal()
    c.computeInteractions(t)
    functor.endTraversal(n3)


# ---- SPH ------------------------------------------------------------------------------------------------------------
def _sph_gpu(s, n3, density_in=None):
    n = len(s["pos"])
    c = make_container(s, capi.PARTICLE_SPH)
    c.rebuildNeighborLists(GpuTraversal("gpulc_c08", SPHCalcDensityFunctor(), n3))
    ids = upload_by_id(c, VX=s["vel"][:, 0], VY=s["vel"][:, 1], VZ=s["vel"][:, 2], MASS=s["mass"], SMTH=s["smth"],
                       PRESSURE=s["pressure"], SNDSPEED=s["snd"])
    run(c, "gpulc_c08", SPHCalcDensityFunctor(), n3)
    rho = by_id(c, "DENSITY", ids, n)
    if density_in is None:
        density_in = np.where(s["own"] == 1, rho, 1.0) + 0.5
    upload_by_id(c, DENSITY=density_in)
    run(c, "gpulc_c18", SPHCalcHydroForceFunctor(), n3)
    acc = np.stack([by_id(c, k, ids, n) for k in ("FX", "FY", "FZ")], axis=1)
    eng, vsig = by_id(c, "ENGDOT", ids, n), by_id(c, "VSIGMAX", ids, n)
    c.close()
    return rho, density_in, acc, eng, vsig


@pytest.mark.parametrize("n3", [False, True])
def test_sph_matches_oracle(n3):
    s = sph_scenario(seed=31, vary_h=not n3)  # newton3 evaluates a pair once with the first particle's support
    rho, dens_in, acc, eng, vsig = _sph_gpu(s, n3)
    owned = s["own"] == 1
    o_rho, sc = oracle.sph_density(s["pos"], s["mass"], s["smth"], s["own"])
    assert close(rho[owned], o_rho[owned], sc[owned])
    o_acc, o_eng, o_vsig, sc, sce = oracle.sph_hydro(s["pos"], s["vel"], s["mass"], s["smth"], dens_in, s["pressure"],
                                                     s["snd"], s["own"], with_eng_scale=True)
    assert close(acc[owned], o_acc[owned], sc[owned])
    assert close(eng[owned], o_eng[owned], sce[owned])  # 1e-12 of the sum of the magnitudes of engDot's own terms
    np.testing.assert_allclose(vsig[owned], o_vsig[owned], rtol=1e-14)


@pytest.mark.parametrize("n3", [0, 1])
def test_sph_matches_reference_fixture(n3):
    g = dict(np.load(os.path.join(GOLDEN, f"fn_sph_n3{n3}.npz")))
    g["cutoff"], g["skin"] = float(g["cutoff"]), float(g["skin"])
    rho, _, acc, eng, vsig = _sph_gpu(g, bool(n3), density_in=g["density_in"])
    owned = g["own"] == 1
    _, sc = oracle.sph_density(g["pos"], g["mass"], g["smth"], g["own"])
    assert close(rho[owned], g["ref_density"][owned], sc[owned])
    _, _, _, sc, sce = oracle.sph_hydro(g["pos"], g["vel"], g["mass"], g["smth"], g["density_in"], g["pressure"], g["snd"],
                                        g["own"], with_eng_scale=True)
    assert close(acc[owned], g["ref_acc"][owned], sc[owned])
    assert close(eng[owned], g["ref_engdot"][owned], sce[owned])
    np.testing.assert_allclose(vsig[owned], g["ref_vsigmax"][owned], rtol=1e-14)


def test_sph_needs_sph_particles_and_linked_cells():
    s = sph_scenario(seed=3)
    c = make_container(s, capi.PARTICLE_LJ)
    t = GpuTraversal("gpulc_c08", SPHCalcDensityFunctor(), False)
    c.rebuildNeighborLists(t)
    with pytest.raises(ApbError):
        c.computeInteractions(t)
    c.close()


# ---- Axilrod-Teller-Muto --------------------------------------------------------------------------------------------
def _atm_gpu(s, nu=None, nu_of_type=None, n3=False):
    n = len(s["pos"])
    c = make_container(s, capi.PARTICLE_LJ, types=s["types"])
    if nu_of_type is not None:
        ppl = ParticlePropertiesLibrary(s["cutoff"])
        for t, v in enumerate(nu_of_type):
            ppl.addSiteType(t, 1.0)
            ppl.addATMParametersToSite(t, v)
        f = AxilrodTellerMutoFunctor(s["cutoff"], ppl, useMixing=True, calculateGlobals=True, countFLOPs=True)
    else:
        f = AxilrodTellerMutoFunctor(s["cutoff"], calculateGlobals=True, countFLOPs=True)
        f.setParticleProperties(nu)
    c.rebuildNeighborLists(GpuTraversal("gpulc_c08", f, n3))
    run(c, "gpulc_c08", f, n3)
    F = c.forcesById(n)
    c.close()
    return F, f


def _atm_virial_scale(s, o):
    """The ATM virial sums f_p * r_p with ABSOLUTE positions (AxilrodTellerMutoFunctor.h:269-284): terms of both signs
    of magnitude |f_p| |r_p| cancel in the sum, so its rounding error scales with the sum of magnitudes, not with the
    result. Tolerance of the virial comparisons: 1e-12 of that magnitude sum (per-particle net forces as a lower bound
    of the per-triplet terms)."""
    owned = s["own"] == 1
    return np.abs(o["f"][owned] * s["pos"][owned]).sum()


@pytest.mark.parametrize("n3", [False, True])
def test_atm_matches_oracle(n3):
    s = atm_scenario(seed=41)
    F, f = _atm_gpu(s, nu=0.073, n3=n3)
    o = oracle.atm(s["pos"], s["types"], s["own"], s["cutoff"], nu=0.073)
    owned = s["own"] == 1
    assert close(F[owned], o["f"][owned], o["scale"][owned])
    assert f.getPotentialEnergy() == pytest.approx(o["upot3_sum"] / 9.0, rel=1e-12)
    assert abs(f.getVirial() - o["virial_sum"].sum()) <= 1e-12 * _atm_virial_scale(s, o)
    if n3:  # every triplet of non-dummy particles once, by its lowest slot (AxilrodTellerMutoFunctor.h:253-259)
        from scipy.spatial import cKDTree
        P = s["pos"][s["own"] != 0]
        nb = cKDTree(P).query_ball_point(P, s["cutoff"])
        triplets = 0
        for i, lst in enumerate(nb):
            hi = [j for j in lst if j > i]
            for a in range(len(hi)):
                d = P[hi[a + 1:]] - P[hi[a]]
                triplets += int(((d * d).sum(axis=1) <= s["cutoff"] ** 2).sum())
        assert f._raw.num_kernel_calls_n3 == triplets and f._raw.num_kernel_calls_no_n3 == 0
        assert 3 * triplets >= o["kernel_calls"]  # the newton3-off count leaves out halo-cell particles as receivers
        assert f.getNumFLOPs() == 24 * f._raw.num_dist_calls + 100 * f._raw.num_kernel_calls_n3 + 24 * f._raw.num_global_calcs_n3
    else:
        assert f._raw.num_kernel_calls_no_n3 == o["kernel_calls"]


@pytest.mark.parametrize("name", ["fn_atm.npz", "fn_atm_mix.npz"])
@pytest.mark.parametrize("n3", [False, True])
def test_atm_matches_reference_fixture(name, n3):
    g = dict(np.load(os.path.join(GOLDEN, name)))
    g["cutoff"], g["skin"] = float(g["cutoff"]), float(g["skin"])
    kw = dict(nu_of_type=g["nu_of_type"]) if "nu_of_type" in g else dict(nu=float(g["nu"]))
    F, f = _atm_gpu(g, n3=n3, **kw)
    o = oracle.atm(g["pos"], g["types"], g["own"], g["cutoff"], **kw)
    owned = g["own"] == 1
    assert close(F[owned], g["ref_f"][owned], o["scale"][owned])
    assert f.getPotentialEnergy() == pytest.approx(float(g["ref_upot"]), rel=1e-12)
    assert abs(f.getVirial() - float(g["ref_virial"])) <= 1e-12 * _atm_virial_scale(g, o)


# ---- LJ multi-site --------------------------------------------------------------------------------------------------
def _multisite_gpu(s, n3, shift=True):
    n = len(s["pos"])
    c = make_container(s, capi.PARTICLE_MULTISITE, types=s["mol_type"])
    ppl = ParticlePropertiesLibrary(s["cutoff"])
    for t, (e, sg) in enumerate(zip(s["eps"], s["sigma"])):
        ppl.addSiteType(t, 1.0)
        ppl.addLJParametersToSite(t, e, sg)
    ss = list(s["site_start"])
    for m in range(len(ss) - 1):
        ppl.addMolType(m, list(np.asarray(s["site_type"])[ss[m]:ss[m + 1]]), np.asarray(s["site_pos"])[ss[m]:ss[m + 1]])
    f = LJMultisiteFunctor(s["cutoff"], ppl, applyShift=shift, useMixing=True, calculateGlobals=True)
    c.rebuildNeighborLists(GpuTraversal("gpulc_c08", f, n3))
    q = np.asarray(s["quat"])
    ids = upload_by_id(c, Q0=q[:, 0], Q1=q[:, 1], Q2=q[:, 2], Q3=q[:, 3])
    run(c, "gpulc_c08", f, n3)
    F = c.forcesById(n)
    T = np.stack([by_id(c, k, ids, n) for k in ("TX", "TY", "TZ")], axis=1)
    c.close()
    return F, T, f


@pytest.mark.parametrize("n3", [False, True])
def test_multisite_matches_oracle(n3):
    s = multisite_scenario(seed=51)
    F, T, f = _multisite_gpu(s, n3)
    o = oracle.multisite(s["pos"], s["quat"], s["mol_type"], s["own"], s["cutoff"], True, s["eps"], s["sigma"],
                         s["site_start"], s["site_pos"], s["site_type"])
    owned = s["own"] == 1
    assert close(F[owned], o["f"][owned], o["scale"][owned])
    assert close(T[owned], o["torque"][owned], o["scale"][owned])
    assert f.getPotentialEnergy() == pytest.approx(o["upot6_sum"] * 0.5 / 6.0, rel=1e-12)
    assert f.getVirial() == pytest.approx(o["virial_sum"].sum() * 0.5, rel=1e-12)


@pytest.mark.parametrize("n3", [0, 1])
def test_multisite_matches_reference_fixture(n3):
    g = dict(np.load(os.path.join(GOLDEN, f"fn_multisite_n3{n3}.npz")))
    g["cutoff"], g["skin"] = float(g["cutoff"]), float(g["skin"])
    F, T, f = _multisite_gpu(g, bool(n3))
    o = oracle.multisite(g["pos"], g["quat"], g["mol_type"], g["own"], g["cutoff"], True, g["eps"], g["sigma"],
                         g["site_start"], g["site_pos"], g["site_type"])
    owned = g["own"] == 1
    assert close(F[owned], g["ref_f"][owned], o["scale"][owned])
    assert close(T[owned], g["ref_torque"][owned], o["scale"][owned])
    assert f.getPotentialEnergy() == pytest.approx(float(g["ref_upot"]), rel=1e-12)
    assert f.getVirial() == pytest.approx(float(g["ref_virial"]), rel=1e-12)


# ---- SPH across a periodic boundary: generating halo exchange + column refresh ------------------------------------------
def test_sph_periodic_halo_exchange_and_column_refresh():
    """The sph-mpi flow (examples/sph-mpi/sph-main-mpi.cpp:373-414) on one periodic rank: updateHaloParticles -> density
    -> pressure -> updateHaloParticles -> hydro force. The generating exchange must give the halo copies every attribute
    of their owners, apb_refresh_halo_columns the owners' new density and pressure; results against the oracle with
    host-built periodic images."""
    from scenarios import grid_lattice, periodic_images
    cutoff, skin = 1.0, 0.1
    pos, bmin, bmax = grid_lattice(9, 0.35, jitter=0.07, seed=5)
    L = bmax - bmin
    pos = bmin + np.mod(pos - bmin, L)
    n = len(pos)
    rng = np.random.default_rng(6)
    vel = rng.normal(0, 0.5, (n, 3))
    mass, smth, snd = rng.uniform(0.8, 1.2, n), rng.uniform(0.30, 0.40, n), rng.uniform(1.0, 1.4, n)
    c = GpuParticleContainer("gpuLinkedCells", bmin, bmax, cutoff, skin, particleKind=capi.PARTICLE_SPH)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    upload_by_id(c, VX=vel[:, 0], VY=vel[:, 1], VZ=vel[:, 2], MASS=mass, SMTH=smth, SNDSPEED=snd)
    c.exchangeHalos()  # generating: copies carry mass, smoothing length, velocity, sound speed
    hpos, hsrc = periodic_images(pos, bmin, bmax, cutoff + skin)
    assert c.getNumberOfParticles("halo") == len(hpos)
    ids, _, own = c.downloadIds()
    hal = own == 2
    for name, ref in (("MASS", mass), ("SMTH", smth), ("SNDSPEED", snd), ("VX", vel[:, 0]), ("VZ", vel[:, 2])):
        assert np.array_equal(c.downloadColumn(name)[hal], ref[ids[hal]]), name
    c.rebuildNeighborLists(GpuTraversal("gpulc_c08", SPHCalcDensityFunctor(), False))
    run(c, "gpulc_c08", SPHCalcDensityFunctor(), False)
    ids, _, own = c.downloadIds()
    owned, hal = own == 1, own == 2
    rho = np.zeros(n)
    rho[ids[owned]] = c.downloadColumn("DENSITY")[owned]
    # oracle: owned + images, attributes of an image = attributes of its source
    src = np.r_[np.arange(n), hsrc]
    apos = np.vstack([pos, hpos])
    aown = np.r_[np.ones(n), 2 * np.ones(len(hpos))].astype(np.int64)
    o_rho, sc = oracle.sph_density(apos, mass[src], smth[src], aown)
    assert close(rho, o_rho[:n], sc[:n])
    pressure = 0.3 * rho * (1.0 + 0.1 * np.sin(np.arange(n)))
    col = c.downloadColumn("PRESSURE")
    col[owned] = pressure[ids[owned]]
    c.uploadColumn("PRESSURE", col)
    c.refreshHaloColumns(["DENSITY", "PRESSURE"])
    assert np.array_equal(c.downloadColumn("DENSITY")[hal], rho[ids[hal]])
    assert np.array_equal(c.downloadColumn("PRESSURE")[hal], pressure[ids[hal]])
    with pytest.raises(ApbError):
        c.refreshHaloColumns(["X"])  # positions go through exchangeHalos (periodic shift)
    run(c, "gpulc_c18", SPHCalcHydroForceFunctor(), False)
    acc = np.zeros((n, 3))
    for d, k in enumerate(("FX", "FY", "FZ")):
        acc[ids[owned], d] = c.downloadColumn(k)[owned]
    eng = np.zeros(n)
    eng[ids[owned]] = c.downloadColumn("ENGDOT")[owned]
    o_acc, o_eng, _, sc, sce = oracle.sph_hydro(apos, vel[src], mass[src], smth[src], rho[src], pressure[src], snd[src], aown,
                                                with_eng_scale=True)
    assert close(acc, o_acc[:n], sc[:n])
    assert close(eng, o_eng[:n], sce[:n])
    c.close()


def test_sph_leavers_carry_their_attributes_and_refresh_needs_links():
    """updateContainer returns whole particles (LeavingParticleCollector.h:101-110): for SPHParticle storage the leavers'
    mass / smoothing length / density ... come back through apb_get_leaver_column; apb_refresh_halo_columns without a
    generating halo exchange is a state error."""
    s = sph_scenario(seed=9)
    n = len(s["pos"])
    owned = s["own"] == 1
    c = GpuParticleContainer("gpuLinkedCells", s["box_min"], s["box_max"], s["cutoff"], s["skin"], particleKind=capi.PARTICLE_SPH)
    ids = np.arange(n)
    c.addParticles(s["pos"][owned, 0], s["pos"][owned, 1], s["pos"][owned, 2], ids[owned])
    with pytest.raises(ApbError):
        c.refreshHaloColumns(["DENSITY"])
    upload_by_id(c, MASS=s["mass"], SMTH=s["smth"], DENSITY=s["mass"] * 3.0, ENGDOT=s["smth"] - 1.0)
    sid, _, _ = c.downloadIds()
    x = c.downloadColumn("X")
    movers = x > s["box_max"][0] - 0.2
    assert movers.any()
    x[movers] += 0.25
    c.uploadColumn("X", x)
    leavers = c.updateContainer(False)
    assert set(leavers["id"].tolist()) == set(sid[movers].tolist())
    lid = leavers["id"]
    assert np.array_equal(c.leaverColumn("MASS"), s["mass"][lid])
    assert np.array_equal(c.leaverColumn("SMTH"), s["smth"][lid])
    assert np.array_equal(c.leaverColumn("DENSITY"), s["mass"][lid] * 3.0)
    assert np.array_equal(c.leaverColumn("ENGDOT"), s["smth"][lid] - 1.0)
    with pytest.raises(ApbError):
        c.leaverColumn("OLDFX")  # SPHParticle has no oldF
    c.close()


def test_sph_partner_lists_skip_particles_deleted_since_the_rebuild():
    """updateContainer(keepNeighborListsValid=true) marks leavers as dummies without a rebuild
    (LeavingParticleCollector.h:85-118): the cached partner lists of gpuLinkedCells still name them, the functor must not see
    them (SPHCalcDensityFunctor.h:46: dummies are skipped)."""
    rng = np.random.default_rng(3)
    npd, d = 24, 0.35  # 13 824 particles... below 16 384 slots the warp kernel would run: force the list kernels
    if os.environ.get("APB_LC_KERNEL", "list")[0] != "l":
        pytest.skip("another kernel variant is forced")
    npd = 28  # 21 952 particles: the partner-list kernels are the default
    g = (np.arange(npd) + 0.5) * d
    pos = np.stack([a.ravel() for a in np.meshgrid(g, g, g, indexing="ij")], axis=1) + rng.uniform(-0.06, 0.06, (npd ** 3, 3))
    n = len(pos)
    L = npd * d
    mass, smth = rng.uniform(0.8, 1.2, n), rng.uniform(0.30, 0.40, n)
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 1.0, 0.2, particleKind=capi.PARTICLE_SPH)
    c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
    upload_by_id(c, MASS=mass, SMTH=smth)
    c.rebuildNeighborLists(GpuTraversal("gpulc_c08", SPHCalcDensityFunctor(), False))
    run(c, "gpulc_c08", SPHCalcDensityFunctor(), False)  # builds the lists
    ids, _, own = c.downloadIds()
    x = c.downloadColumn("X")
    leave = (own == 1) & (x > L - 0.3) & (ids % 3 == 0)
    assert leave.sum() > 20
    x[leave] += 0.4  # out of the box: leavers
    c.uploadColumn("X", x)
    leavers = c.updateContainer(True)
    assert set(leavers["id"].tolist()) == set(ids[leave].tolist())
    c.uploadColumn("DENSITY", np.zeros(c.numSlots()))
    run(c, "gpulc_c08", SPHCalcDensityFunctor(), False)
    ids2, _, own2 = c.downloadIds()
    rho = np.zeros(n)
    m = own2 == 1
    rho[ids2[m]] = c.downloadColumn("DENSITY")[m]
    keep = np.ones(n, dtype=bool)
    keep[ids[leave]] = False
    o_rho, sc = oracle.sph_density(pos[keep], mass[keep], smth[keep], np.ones(keep.sum(), dtype=np.int64))
    assert close(rho[keep], o_rho, sc)
    c.close()
