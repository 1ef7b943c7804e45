"""md-flexible's VTK checkpoint record (examples/md-flexible/src/ParallelVtkWriter.cpp:55-201, :308-356).
CPU: the oracle restatement (oracle/vtk_oracle.c) against bytes written by the unmodified reference writer (committed
fixture tests/golden/fn_vtk.npz, and live when oracle/_ref travelled); the product's decimal formatter
(autopas_b200/csrc/vtk_format.cuh, compiled for the host by this test) against the C library's printf.
GPU: apb_vtk_particle_record byte for byte against the oracle and the fixture, error paths, full-size record."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # (when run as the chunked child process)
import oracle  # noqa: E402
from autopas_b200 import ApbError, GpuParticleContainer, ParallelVtkWriter, capi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "fn_vtk.npz")


def _golden():
    g = np.load(GOLDEN)
    return g


def _in_order_of(piece, ids):
    """row order of a piece (its `ids` array) as indices into `ids`"""
    where = {int(i): k for k, i in enumerate(ids)}
    return np.array([where[int(i)] for i in oracle.vtk_parse_ids(piece)])


def test_oracle_vtk_record_matches_reference_bytes():
    g = _golden()
    o = _in_order_of(g["ref_piece"], g["ids"])  # iteration order of the reference's LinkedCells
    assert sorted(o) == list(range(len(g["ids"])))
    mine = oracle.vtk_particle_record(g["ids"][o], g["r"][o], g["v"][o], g["f"][o], g["types"][o], g["box_max"])
    assert np.array_equal(mine, g["ref_piece"])
    text = bytes(mine).decode()
    assert "9.9999999999999 " in text and " inf -inf nan\n" in text and "e-310" in text  # raised precision, specials, denormals
    assert np.array_equal(oracle.vtk_pvtu_record(str(g["session"]), 1, int(g["iteration"]), int(g["digits"])), g["ref_pvtu"])


@pytest.mark.skipif(not oracle.have_ref_vtk(), reason="oracle/_ref/vtk_ref_writer not built (reference tree absent)")
def test_oracle_vtk_record_matches_reference_live():
    rng = np.random.default_rng(9)
    n = 3000
    lo, hi = np.array([-3.0, 0.0, 100.0]), np.array([12.5, 1000.0, 104.0])
    r = lo + rng.uniform(0, 1, (n, 3)) * (hi - lo)
    for d in range(3):
        r[d * 200:(d + 1) * 200, d] = hi[d] - 10.0 ** -rng.uniform(1, 12, 200)
    v = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-12, 12, (n, 3))
    f = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-6, 15, (n, 3))
    ids, types = rng.permutation(7 * n)[:n].astype(np.int64), rng.integers(0, 5, n).astype(np.int64)
    piece, index = oracle.ref_vtk_records(ids, r, v, f, types, lo, hi, "live", 123456, 4)
    o = _in_order_of(piece, ids)
    assert np.array_equal(oracle.vtk_particle_record(ids[o], r[o], v[o], f[o], types[o], hi), piece)
    assert np.array_equal(oracle.vtk_pvtu_record("live", 1, 123456, 4), index)  # iteration wider than the digits


@pytest.fixture(scope="module")
def host_formatter(tmp_path_factory):
    """vtk_format.cuh compiled for the host (tests/vtk/vtk_format_host.cpp): the same functions the kernels call."""
    so = str(tmp_path_factory.mktemp("vtkfmt") / "libvtkfmt.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-x", "c++",
                    "-I" + os.path.join(HERE, "..", "autopas_b200", "csrc"), "-o", so, os.path.join(HERE, "vtk", "vtk_format_host.cpp")],
                   check=True)
    lib = ctypes.CDLL(so)
    lib.check_random.restype = ctypes.c_int64
    lib.position_precision.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.fmt_g.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_char_p]
    lib.fmt_u64.argtypes = [ctypes.c_uint64, ctypes.c_char_p]
    return lib


def test_formatter_equals_printf(host_formatter):
    """Every precision the writer can use (6 ... 15, plus 1, 3, 16, 17) on random bit patterns (all exponents,
    denormals), MD-like magnitudes, exact decimal ties and the neighbourhood of powers of ten."""
    msg = ctypes.create_string_buffer(256)
    for mode, count in ((0, 200_000), (1, 200_000), (2, 200_000), (3, 100_000)):
        bad = host_formatter.check_random(ctypes.c_uint64(1000 + mode), ctypes.c_int64(count), mode, msg)
        assert bad == 0, msg.value.decode()
    buf = ctypes.create_string_buffer(64)
    for v, want in ((0.0, "0"), (-0.0, "-0"), (float("inf"), "inf"), (float("-inf"), "-inf"), (0.5, "0.5"), (1e-5, "1e-05"),
                    (123456.5, "123456"), (123457.5, "123458"), (999999.5, "1e+06"), (0.0001, "0.0001"), (5e-324, "4.94066e-324"),
                    (1.7976931348623157e308, "1.79769e+308"), (100.0, "100"), (2.5e-5, "2.5e-05")):
        n = host_formatter.fmt_g(v, 6, buf)
        assert buf.raw[:n].decode() == want == "%g" % v, (v, buf.raw[:n])
    for v in (0, 7, 10, 4294967296, 2 ** 64 - 1):
        n = host_formatter.fmt_u64(v, buf)
        assert buf.raw[:n].decode() == str(v)


def test_position_precision_equals_oracle(host_formatter):
    """writeWithDynamicPrecision's precision choice (ParallelVtkWriter.cpp:130-157) incl. the cases where it throws."""
    rng = np.random.default_rng(0)
    seen = set()
    for it in range(40_000):
        border = float(rng.choice([10.0, 100.0, 7.5, 1.0, 1000.0, 37.8, 0.3, 1e-3, 12345.678, 1e6, 378.0]))
        kind = it % 4
        if kind == 0:
            pos = border - 10.0 ** (-rng.uniform(0, 17))
        elif kind == 1:
            pos = border - rng.integers(1, 300) * np.spacing(border)
        elif kind == 2:
            pos = rng.uniform(-border, border)
        else:
            pos = border * (1 - 10.0 ** (-rng.integers(1, 16)))
        want = oracle.vtk_position_precision(pos, border)
        assert host_formatter.position_precision(pos, border) == want, (pos.hex(), border)
        seen.add(want)
    assert seen == {-1, *range(6, 16)}


GOLDEN_LOAD = os.path.join(HERE, "golden", "fn_vtk_load.npz")


def _same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_oracle_vtk_loader_matches_reference_particles():
    """oracle.vtk_load_particle_record == the particles the UNMODIFIED loader (MDFlexConfig::loadParticlesFromCheckpoint)
    made of a piece written by the unmodified writer: bit patterns of every value (denormals, 1e300, raised precision)."""
    g = np.load(GOLDEN_LOAD)
    o = oracle.vtk_load_particle_record(g["piece"])
    for mine, ref in (("ids", "ref_ids"), ("types", "ref_types"), ("r", "ref_r"), ("v", "ref_v"), ("f", "ref_f")):
        assert _same_bits(o[mine], g[ref]), mine
    assert 2 ** 40 + 5 in o["ids"] and np.any(np.abs(o["v"]) < 1e-308) and np.any(np.abs(o["v"]) > 1e299)


@pytest.mark.skipif(not oracle.have_ref_vtk(), reason="oracle/_ref/vtk_ref_writer not built (reference tree absent)")
def test_oracle_vtk_loader_matches_reference_live(tmp_path):
    """a checkpoint of two pieces: piece names, piece count and particles as the unmodified loader sees them"""
    rng = np.random.default_rng(12)
    pieces = []
    os.makedirs(tmp_path / "run" / "data")
    for rank in range(2):
        n = 400 + 100 * rank
        ids = (rng.permutation(5 * n)[:n] + 10_000 * rank).astype(np.int64)
        lo, hi = np.array([10.0 * rank, 0.0, 0.0]), np.array([10.0 * rank + 10.0, 8.0, 12.0])
        r = lo + rng.uniform(0, 1, (n, 3)) * (hi - lo)
        r[:30, 0] = hi[0] - 10.0 ** -rng.uniform(1, 12, 30)
        v = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-12, 12, (n, 3))
        f = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-6, 15, (n, 3))
        piece, _ = oracle.ref_vtk_records(ids, r, v, f, rng.integers(0, 3, n), lo, hi, "run", 1200, 5)
        piece.tofile(str(tmp_path / "run" / "data" / f"run_Particles_{rank}_01200.vtu"))
        pieces.append(piece)
    index = tmp_path / "run" / "run_Particles_01200.pvtu"
    oracle.vtk_pvtu_record("run", 2, 1200, 5).tofile(str(index))
    from autopas_b200 import checkpointPieces
    assert checkpointPieces(str(index)) == [str(tmp_path / "run" / "data" / f"run_Particles_{k}_01200.vtu") for k in range(2)]
    for rank in range(2):  # as many ranks as pieces: every rank reads its own
        ref = oracle.ref_vtk_load(index, rank, 2)
        mine = oracle.vtk_load_particle_record(pieces[rank])
        assert all(_same_bits(mine[k], ref[k]) for k in ("ids", "types", "r", "v", "f"))
    ref = oracle.ref_vtk_load(index, 0, 1)  # one rank, two pieces: rank 0 reads both
    both = [oracle.vtk_load_particle_record(p) for p in pieces]
    assert all(_same_bits(np.concatenate([b[k] for b in both]), ref[k]) for k in ("ids", "types", "r", "v", "f"))


def test_parser_equals_strtod(host_formatter):
    """apbParseDouble (the loader's decimal -> binary conversion) against the C library on printed doubles of every
    precision, random digit strings up to 40 digits with exponents over the whole range, exact ties and the borders of
    the subnormal / overflow ranges; malformed tokens are reported."""
    host_formatter.check_parse.restype = ctypes.c_int64
    host_formatter.parse_double.restype = ctypes.c_double
    msg = ctypes.create_string_buffer(256)
    for mode, count in ((0, 100_000), (1, 600_000), (2, 150_000)):
        assert host_formatter.check_parse(ctypes.c_uint64(2000 + mode), ctypes.c_int64(count), mode, msg) == 0, msg.value.decode()
    status = ctypes.c_int()
    for text, want in (("0", 0.0), ("-0", -0.0), ("1e-400", 0.0), ("1e400", float("inf")), ("-1e400", float("-inf")), ("4.94066e-324", 5e-324),
                       ("2.4703282292062327e-324", 0.0), ("2.4703282292062328e-324", 5e-324), (".5", 0.5), ("5.", 5.0), ("+1E+2", 100.0)):
        got = host_formatter.parse_double(text.encode(), len(text), ctypes.byref(status))
        assert status.value == 0 and got == want and np.signbit(got) == np.signbit(want), text
    for text in ("inf", "nan", "-inf", "", "-", "1e", "1e+", "0x10", "1.2.3", "12a", "1" * 41, "."):
        host_formatter.parse_double(text.encode(), len(text), ctypes.byref(status))
        assert status.value == 1, text


def test_checkpoint_piece_selection_follows_the_reference(tmp_path):
    """MDFlexConfig::loadParticlesFromCheckpoint (MDFlexConfig.cpp:648-671): with as many ranks as pieces each rank
    loads its own piece, otherwise rank 0 loads all and the others none; piece names are rebuilt from the index's own
    name (:93-117), not from its Piece entries. Host logic only (a recording stand-in for the container)."""
    from autopas_b200 import checkpointPieces, loadParticlesFromCheckpoint
    os.makedirs(tmp_path / "out" / "spinodal_run" / "data")
    index = tmp_path / "out" / "spinodal_run" / "spinodal_run_Particles_0250000.pvtu"
    oracle.vtk_pvtu_record("spinodal_run", 4, 250000, 7).tofile(str(index))
    names = [str(tmp_path / "out" / "spinodal_run" / "data" / f"spinodal_run_Particles_{k}_0250000.vtu") for k in range(4)]
    assert checkpointPieces(str(index)) == names
    for k, name in enumerate(names):
        np.frombuffer(f"piece {k}".encode(), dtype=np.uint8).tofile(name)

    class Recorder:
        def __init__(self):
            self.seen = []

        def loadVtkParticleRecord(self, data, checkInBox=True):
            self.seen.append(bytes(data).decode())
            return 10

    for rank in range(4):
        rec = Recorder()
        assert loadParticlesFromCheckpoint(str(index), rank, 4, rec) == 10 and rec.seen == [f"piece {rank}"]
    for rank, want in ((0, [f"piece {k}" for k in range(4)]), (1, [])):
        rec = Recorder()
        assert loadParticlesFromCheckpoint(str(index), rank, 2, rec) == 10 * len(want) and rec.seen == want


def test_pvtu_record_through_the_c_abi():
    """host text only (no device): equal to the reference's index file"""
    g = _golden()
    w = ParallelVtkWriter.__new__(ParallelVtkWriter)
    w._session, w._digits, w._rank, w._ranks = str(g["session"]), int(g["digits"]), 0, 1
    assert np.array_equal(w.pvtuRecord(int(g["iteration"])), g["ref_pvtu"])
    w._ranks, w._digits = 3, 2
    assert np.array_equal(w.pvtuRecord(12345), oracle.vtk_pvtu_record(str(g["session"]), 3, 12345, 2))


def _fill(c, ids, r, v, f, types):
    c.addParticles(r[:, 0], r[:, 1], r[:, 2], ids, types.astype(np.int32))
    sid, _, _ = c.downloadIds()
    where = {int(i): k for k, i in enumerate(ids)}
    order = np.array([where[int(i)] for i in sid])
    for name, a in (("V", v), ("F", f)):
        for d, ax in enumerate("XYZ"):
            c.uploadColumn(name + ax, a[order, d])
    return order


@pytest.mark.gpu
@pytest.mark.parametrize("container", ["gpuLinkedCells", "gpuVerletClusterLists"])
def test_gpu_vtk_record_is_byte_exact(container, tmp_path):
    g = _golden()
    ids, r, v, f, types = g["ids"], g["r"], g["v"], g["f"], g["types"]
    c = GpuParticleContainer(container, g["box_min"], g["box_max"], 1.0, 0.2, clusterSize=4)
    order = _fill(c, ids, r, v, f, types)
    want = oracle.vtk_particle_record(ids[order], r[order], v[order], f[order], types[order], g["box_max"])
    got = c.vtkParticleRecord()
    assert len(got) == len(want) and np.array_equal(got, want), bytes(got[:400])
    # the same rows as the reference's file, whatever order its container iterates in
    rows = lambda piece, name: sorted(bytes(piece).decode().split(f'Name="{name}"', 1)[1].split(">\n", 1)[1].split("        </DataArray>", 1)[0].splitlines())  # noqa: E731
    for name in ("velocities", "forces", "typeIds", "ids", "positions"):
        assert rows(got, name) == rows(g["ref_piece"], name), name
    # halo copies and deleted particles are not part of the record (IteratorBehavior::owned)
    c.addHaloParticles(np.array([-0.5]), np.array([0.0]), np.array([1.0]), np.array([999999], dtype=np.int64))
    assert np.array_equal(c.vtkParticleRecord(), want)
    # after a rebuild the rows follow the new storage order
    if container == "gpuVerletClusterLists":
        from autopas_b200 import GpuTraversal, LJFunctor
        fn = LJFunctor(1.0)
        fn.setParticleProperties(1.0, 1.0)
        c.rebuildNeighborLists(GpuTraversal("gpuvcl_cluster_iteration", fn, False))
        sid, _, sown = c.downloadIds()
        where = {int(i): k for k, i in enumerate(ids)}
        o2 = np.array([where[int(i)] for i, s in zip(sid, sown) if s == 1])
        assert np.array_equal(c.vtkParticleRecord(), oracle.vtk_particle_record(ids[o2], r[o2], v[o2], f[o2], types[o2], g["box_max"]))
    # files through the writer mirror: names as the reference's loader expects them (MDFlexConfig.cpp:91-120)
    w = ParallelVtkWriter("sess", str(tmp_path), 6)
    path = w.recordParticleStates(42, c)
    assert path.endswith("/sess/data/sess_Particles_0_000042.vtu")
    assert np.array_equal(np.fromfile(path, dtype=np.uint8), c.vtkParticleRecord())  # device -> pinned pieces -> file
    with pytest.raises(ApbError, match="Failed to open file"):  # the reference's std::runtime_error (ParallelVtkWriter.cpp:68-70)
        c.writeVtkParticleRecord(str(tmp_path / "no" / "such" / "folder" / "x.vtu"))
    assert np.array_equal(np.fromfile(str(tmp_path / "sess" / "sess_Particles_000042.pvtu"), dtype=np.uint8),
                          oracle.vtk_pvtu_record("sess", 1, 42, 6))
    c.close()


def _chunked_record_equals_oracle():
    """run in a child process with APB_VTK_CHUNK_ROWS set (the library reads it once)"""
    g = _golden()
    ids, r, v, f, types = g["ids"], g["r"], g["v"], g["f"], g["types"]
    c = GpuParticleContainer("gpuLinkedCells", g["box_min"], g["box_max"], 1.0, 0.2)
    order = _fill(c, ids, r, v, f, types)
    want = oracle.vtk_particle_record(ids[order], r[order], v[order], f[order], types[order], g["box_max"])
    assert np.array_equal(c.vtkParticleRecord(), want)
    c.close()
    rng = np.random.default_rng(4)
    n, L = 50_000, 40.0
    R = rng.uniform(0, L, (n, 3))
    c = GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [L, L, L], 2.5, 0.3, clusterSize=32)
    c.addParticles(R[:, 0], R[:, 1], R[:, 2], np.arange(n, dtype=np.int64))
    c.addHaloParticles(np.array([-1.0, L + 1.0]), np.array([1.0, 2.0]), np.array([1.0, 2.0]), np.array([n, n + 1], dtype=np.int64))
    for k in ("VX", "VY", "VZ", "FX", "FY", "FZ"):
        c.uploadColumn(k, rng.normal(size=c.numSlots()) * 10.0 ** rng.integers(-6, 7, c.numSlots()))
    sid, stype, sown = c.downloadIds()
    o = sown == 1
    col = lambda *names: np.stack([c.downloadColumn(k) for k in names], axis=1)[o]  # noqa: E731
    want = oracle.vtk_particle_record(sid[o], col("X", "Y", "Z"), col("VX", "VY", "VZ"), col("FX", "FY", "FZ"), stype[o], [L, L, L])
    assert np.array_equal(c.vtkParticleRecord(), want)
    c.close()
    print("chunked record OK")


@pytest.mark.gpu
@pytest.mark.parametrize("chunk_rows", [97, 4096])
def test_gpu_vtk_record_in_chunks_of_rows(chunk_rows):
    """Byte offsets are 32-bit within a chunk of rows (2^24 by default, so that a data array may exceed 2 GB): with small
    chunks the fixture and a 50 000-particle container (several chunks per data array, chunk borders inside a warp's
    stretch of rows) give the same bytes."""
    env = dict(os.environ, APB_VTK_CHUNK_ROWS=str(chunk_rows))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--chunked"], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(HERE))
    assert r.returncode == 0 and "chunked record OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_gpu_vtk_record_error_paths_and_empty_container():
    box = [10.0, 10.0, 10.0]
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], box, 1.0, 0.2)
    assert np.array_equal(c.vtkParticleRecord(), oracle.vtk_particle_record(np.zeros(0, np.int64), np.zeros((0, 3)), np.zeros((0, 3)),
                                                                          np.zeros((0, 3)), np.zeros(0, np.int64), box))
    c.addParticles(np.array([5.0, np.nextafter(10.0, 0.0)]), np.array([5.0, 5.0]), np.array([5.0, 5.0]), np.array([1, 2], dtype=np.int64))
    with pytest.raises(ValueError):
        oracle.vtk_particle_record(np.array([2]), np.array([[np.nextafter(10.0, 0.0), 5.0, 5.0]]), np.zeros((1, 3)), np.zeros((1, 3)),
                                   np.array([0]), box)
    with pytest.raises(ApbError, match="15 digits"):  # the reference throws std::runtime_error (ParallelVtkWriter.cpp:141-149)
        c.vtkParticleRecord()
    c.deleteAllParticles()
    c.addParticles(np.array([5.0]), np.array([5.0]), np.array([5.0]), np.array([1], dtype=np.int64))
    n = ctypes.c_int64()
    lib = capi.load()
    assert lib.apb_vtk_particle_record(c._h, None, 0, ctypes.byref(n)) == capi.APB_OK and n.value > 800
    small = np.zeros(100, dtype=np.uint8)
    assert lib.apb_vtk_particle_record(c._h, small.ctypes.data_as(ctypes.c_void_p), 100, ctypes.byref(n)) == capi.ERR_INVALID_ARGUMENT
    c.close()
    s = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [4, 4, 4], 1.0, 0.1, particleKind=capi.PARTICLE_SPH)
    with pytest.raises(ApbError):
        s.vtkParticleRecord()  # md-flexible's record is MoleculeLJ's
    s.close()


@pytest.mark.gpu
def test_gpu_vtk_record_full_size(tmp_path):
    """1 M particles of a liquid box (positions incl. the band below the upper corner, velocities and forces over many
    decades): the record equals the oracle's for the same columns in storage order, and reads back like md-flexible's
    loader reads it (MDFlexConfig.cpp:122-165: counts, ids, values within the printed precision)."""
    rng = np.random.default_rng(21)
    n, L = 1_000_000, 100.0
    ids = rng.permutation(n).astype(np.int64)
    r = rng.uniform(0, L, (n, 3))
    r[:5000, 0] = L - 10.0 ** -rng.uniform(1, 12, 5000)
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    c.addParticles(r[:, 0], r[:, 1], r[:, 2], ids, rng.integers(0, 3, n).astype(np.int32))
    for k in ("VX", "VY", "VZ", "FX", "FY", "FZ"):
        c.uploadColumn(k, rng.normal(size=n) * 10.0 ** rng.integers(-6, 7, n))
    got = c.vtkParticleRecord()
    sid, stype, _ = c.downloadIds()
    col = lambda *names: np.stack([c.downloadColumn(k) for k in names], axis=1)  # noqa: E731
    R, V, F = col("X", "Y", "Z"), col("VX", "VY", "VZ"), col("FX", "FY", "FZ")
    want = oracle.vtk_particle_record(sid, R, V, F, stype, [L, L, L])
    assert len(got) == len(want) and np.array_equal(got, want)
    path = str(tmp_path / "full.vtu")  # 128 MB: several 32 MB pieces through the two pinned buffers
    assert c.writeVtkParticleRecord(path) == len(want) and np.array_equal(np.fromfile(path, dtype=np.uint8), want)
    text = bytes(got).decode()
    assert f'NumberOfPoints="{n}"' in text
    payload = lambda name: text.split(f'Name="{name}"', 1)[1].split(">\n", 1)[1].split("</DataArray>", 1)[0]  # noqa: E731
    assert np.array_equal(np.array(payload("ids").split(), dtype=np.int64), sid)
    back = np.array(payload("positions").split(), dtype=np.float64).reshape(n, 3)
    assert np.all(np.abs(back - R) <= 5.1e-6 * np.maximum(np.abs(R), 1e-300)) and np.all(back < L)  # no position rounds onto the border
    vb = np.array(payload("velocities").split(), dtype=np.float64).reshape(n, 3)
    assert np.all(np.abs(vb - V) <= 5.1e-6 * np.abs(V))
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("container", ["gpuLinkedCells", "gpuVerletClusterLists"])
def test_gpu_vtk_loader_equals_reference_loader(container, tmp_path):
    """the reference's piece -> device: every attribute has the bits the unmodified reference loader produced; written
    again it is the same file (up to the row order of the reference's container, which the loader keeps)"""
    g = np.load(GOLDEN_LOAD)
    c = GpuParticleContainer(container, g["box_min"], g["box_max"], 1.0, 0.2, clusterSize=4)
    assert c.loadVtkParticleRecord(g["piece"]) == len(g["ref_ids"]) == c.getNumberOfParticles("owned")
    sid, stype, sown = c.downloadIds()
    col = lambda *names: np.stack([c.downloadColumn(k) for k in names], axis=1)  # noqa: E731
    assert np.array_equal(sid, g["ref_ids"]) and np.array_equal(stype, g["ref_types"]) and np.all(sown == 1)
    assert _same_bits(col("X", "Y", "Z"), g["ref_r"]) and _same_bits(col("VX", "VY", "VZ"), g["ref_v"]) and _same_bits(col("FX", "FY", "FZ"), g["ref_f"])
    assert not np.any(col("OLDFX", "OLDFY", "OLDFZ"))
    assert np.array_equal(c.vtkParticleRecord(), g["piece"])  # write(load(reference file)) == reference file
    # a second piece appends; through the file-name logic of the loader
    w = ParallelVtkWriter("again", str(tmp_path), 3)
    w.recordParticleStates(7, c)
    d = GpuParticleContainer(container, g["box_min"], g["box_max"], 1.0, 0.2, clusterSize=4)
    from autopas_b200 import loadParticlesFromCheckpoint
    assert loadParticlesFromCheckpoint(str(tmp_path / "again" / "again_Particles_007.pvtu"), 0, 1, d) == len(sid)
    assert np.array_equal(d.vtkParticleRecord(), g["piece"])
    c.close()
    d.close()


@pytest.mark.gpu
def test_gpu_vtk_loader_error_paths():
    g = np.load(GOLDEN_LOAD)
    text = bytes(g["piece"])
    small = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [5, 5, 5], 1.0, 0.2)
    with pytest.raises(ApbError) as e:  # AutoPas::addParticle throws for a particle outside the box
        small.loadVtkParticleRecord(g["piece"])
    assert e.value.code == capi.ERR_PARTICLE_OUTSIDE and small.getNumberOfParticles("owned") == 0
    assert small.loadVtkParticleRecord(g["piece"], checkInBox=False) == len(g["ref_ids"])
    small.close()
    c = GpuParticleContainer("gpuLinkedCells", g["box_min"], g["box_max"], 1.0, 0.2)
    with pytest.raises(ApbError, match="not a decimal number"):  # the fixture of the writer test holds "inf -inf nan"
        c.loadVtkParticleRecord(_golden()["ref_piece"])
    with pytest.raises(ApbError, match="fewer values"):
        c.loadVtkParticleRecord(text.replace(b'NumberOfPoints="600"', b'NumberOfPoints="601"'))
    with pytest.raises(ApbError, match="not found"):
        c.loadVtkParticleRecord(text.replace(b'"forces"', b'"farces"'))
    with pytest.raises(ApbError, match="number of particles"):
        c.loadVtkParticleRecord(text.replace(b'NumberOfPoints="600"', b'NumberOfPoints="0"'))
    with pytest.raises(ApbError):
        c.loadVtkParticleRecord(b"hello")
    assert c.getNumberOfParticles("owned") == 0
    # values spread over lines differently, extra blanks, more values than needed: operator>> does not care
    loose = text.replace(b"\n        ", b" \t\n  ", 50)
    assert c.loadVtkParticleRecord(loose) == 600
    sid, _, _ = c.downloadIds()
    assert np.array_equal(sid, g["ref_ids"])
    c.close()


@pytest.mark.gpu
def test_gpu_vtk_round_trip_full_size():
    """1 M particles: write -> load into an empty container -> write gives the same bytes, and every loaded value is the
    correctly rounded double of its text (the oracle's float() of the same tokens)."""
    rng = np.random.default_rng(33)
    n, L = 1_000_000, 100.0
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    c.addParticles(rng.uniform(0, L, n), rng.uniform(0, L, n), rng.uniform(0, L, n), rng.permutation(n).astype(np.int64),
                   rng.integers(0, 3, n).astype(np.int32))
    for k in ("VX", "VY", "VZ", "FX", "FY", "FZ"):
        c.uploadColumn(k, rng.normal(size=n) * 10.0 ** rng.integers(-6, 7, n))
    piece = c.vtkParticleRecord()
    d = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    assert d.loadVtkParticleRecord(piece) == n
    assert np.array_equal(d.vtkParticleRecord(), piece)
    want = oracle.vtk_load_particle_record(piece)
    sid, stype, _ = d.downloadIds()
    col = lambda *names: np.stack([d.downloadColumn(k) for k in names], axis=1)  # noqa: E731
    assert np.array_equal(sid, want["ids"]) and np.array_equal(stype, want["types"])
    assert _same_bits(col("X", "Y", "Z"), want["r"]) and _same_bits(col("VX", "VY", "VZ"), want["v"]) and _same_bits(col("FX", "FY", "FZ"), want["f"])
    c.close()
    d.close()


if __name__ == "__main__" and "--chunked" in sys.argv:
    _chunked_record_equals_oracle()
