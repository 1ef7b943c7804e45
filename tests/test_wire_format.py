"""md-flexible's MPI wire format (examples/md-flexible/src/ParticleSerializationTools.cpp:43-146).
CPU: the oracle restatement against bytes produced by the unmodified reference (committed fixture
tests/golden/fn_wire_format.npz, and live when oracle/_ref travelled). GPU: apb_serialize_particles /
apb_deserialize_particles byte for byte against the oracle, round trip, full-size property."""
import os

import numpy as np
import pytest

import oracle
from autopas_b200 import ApbError, GpuParticleContainer, capi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fn_wire_format.npz")


def _golden():
    g = np.load(GOLDEN)
    return (g["ids"], g["r"], g["v"], g["f"], g["oldf"], g["types"], g["own"]), g["ref_bytes"]


def test_oracle_wire_format_matches_reference_bytes():
    p, ref_bytes = _golden()
    assert len(ref_bytes) == len(p[0]) * oracle.WIRE_RECORD_BYTES == len(p[0]) * 120
    mine = oracle.wire_serialize(*p)
    assert np.array_equal(mine, ref_bytes)
    back = oracle.wire_deserialize(ref_bytes)
    for key, want in zip(("id", "r", "v", "f", "oldf", "type", "own"), p):
        assert np.array_equal(back[key].view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), key  # bit patterns (-0.0, denormals)


@pytest.mark.skipif(not oracle.have_ref_wire(), reason="oracle/_ref/libautopas_ref_wire.so not built (reference tree absent)")
def test_oracle_wire_format_matches_reference_live():
    rng = np.random.default_rng(5)
    n = 1000
    p = (rng.permutation(5 * n)[:n].astype(np.int64), rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3)),
         rng.normal(size=(n, 3)), rng.integers(0, 3, n).astype(np.int64), rng.integers(1, 3, n).astype(np.int64))
    assert np.array_equal(oracle.wire_serialize(*p), oracle.ref_wire_serialize(*p))
    back = oracle.ref_wire_deserialize(oracle.wire_serialize(*p))
    assert np.array_equal(back["id"], p[0]) and np.array_equal(back["oldf"], p[4]) and np.array_equal(back["own"], p[6])


def _fill(c, p):
    ids, r, v, f, oldf, types, own = p
    for state, add in ((1, c.addParticles), (2, c.addHaloParticles)):
        m = own == state
        add(r[m, 0], r[m, 1], r[m, 2], ids[m], types[m].astype(np.int32))
    sid, _, _ = c.downloadIds()
    where = {int(i): k for k, i in enumerate(ids)}
    order = np.array([where[int(i)] for i in sid])
    for name, a in (("V", v), ("F", f), ("OLDF", oldf)):
        for d, ax in enumerate("XYZ"):
            c.uploadColumn(name + ax, a[order, d])
    return order


@pytest.mark.gpu
def test_gpu_serialisation_is_byte_exact_and_round_trips():
    p, ref_bytes = _golden()
    big = 200.0  # owned particles of the fixture lie within 10 sigma = 100 of the origin
    c = GpuParticleContainer("gpuLinkedCells", [-big, -big, -big], [big, big, big], 1.0, 0.1)
    order = _fill(c, p)
    data = c.serializeParticles("ownedOrHalo")
    assert len(data) == len(ref_bytes)
    want = ref_bytes.reshape(-1, 120)[order].reshape(-1)  # storage order: owned first, then halo
    assert np.array_equal(data, want)
    assert np.array_equal(c.serializeParticles("owned"), ref_bytes.reshape(-1, 120)[order][p[6][order] == 1].reshape(-1))
    assert np.array_equal(c.serializeParticles("halo"), ref_bytes.reshape(-1, 120)[order][p[6][order] == 2].reshape(-1))
    # the reference's bytes into a second container, and out again
    d = GpuParticleContainer("gpuVerletClusterLists", [-big, -big, -big], [big, big, big], 1.0, 0.1, clusterSize=4)
    d.deserializeParticles(ref_bytes)
    assert d.getNumberOfParticles("owned") == int((p[6] == 1).sum()) and d.getNumberOfParticles("halo") == int((p[6] == 2).sum())
    assert np.array_equal(d.serializeParticles("ownedOrHalo"), ref_bytes)
    with pytest.raises(ApbError):
        d.deserializeParticles(ref_bytes[:119])
    bad = ref_bytes.copy()
    bad[112:120] = 0  # ownership state dummy
    with pytest.raises(ApbError):
        d.deserializeParticles(bad[:120])
    c.close()
    d.close()
    s = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [4, 4, 4], 1.0, 0.1, particleKind=capi.PARTICLE_SPH)
    with pytest.raises(ApbError):
        s.serializeParticles()  # the 120-byte record is MoleculeLJ's
    s.close()


@pytest.mark.gpu
def test_gpu_serialisation_full_size_round_trip():
    """1 M particles: serialise -> deserialise into an empty container -> serialise: identical bytes; the ids survive as a
    set and the records equal the oracle's for the same columns."""
    rng = np.random.default_rng(11)
    n = 1_000_000
    L = 100.0
    ids = rng.permutation(n).astype(np.int64)
    r = rng.uniform(0, L, (n, 3))
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    c.addParticles(r[:, 0], r[:, 1], r[:, 2], ids, rng.integers(0, 2, n).astype(np.int32))
    cols = {k: rng.normal(size=n) for k in ("VX", "VY", "VZ", "FX", "FY", "FZ", "OLDFX", "OLDFY", "OLDFZ")}
    for k, a in cols.items():
        c.uploadColumn(k, a)
    data = c.serializeParticles("owned")
    assert len(data) == n * 120
    sid, stype, sown = c.downloadIds()
    get = lambda *names: np.stack([c.downloadColumn(k) for k in names], axis=1)  # noqa: E731
    want = oracle.wire_serialize(sid, get("X", "Y", "Z"), get("VX", "VY", "VZ"), get("FX", "FY", "FZ"),
                                 get("OLDFX", "OLDFY", "OLDFZ"), stype, sown)
    assert np.array_equal(data, want)
    d = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    d.deserializeParticles(data)
    assert np.array_equal(d.serializeParticles("owned"), data)
    c.close()
    d.close()
