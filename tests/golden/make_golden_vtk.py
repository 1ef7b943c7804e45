"""Golden bytes of md-flexible's VTK checkpoint: seeded particles written by the UNMODIFIED reference writer
(oracle/_ref/vtk_ref_writer = examples/md-flexible/src/ParallelVtkWriter.cpp compiled where it lies, driven by
oracle/ref_driver_vtk.cpp on a stock AutoPas<MoleculeLJ>).
Run in the build container:  python tests/golden/make_golden_vtk.py  ->  tests/golden/fn_vtk.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

BOX_MIN = np.array([0.0, -5.0, 0.0])
BOX_MAX = np.array([10.0, 20.0, 7.5])


def particles(n=600, seed=77):
    rng = np.random.default_rng(seed)
    ids = rng.permutation(50 * n)[:n].astype(np.int64)
    ids[0] = 2 ** 40 + 5  # printed as unsigned long, whatever the attribute `type="Int32"` says
    r = BOX_MIN + rng.uniform(0, 1, (n, 3)) * (BOX_MAX - BOX_MIN)
    # writeWithDynamicPrecision: positions from 0.1 down to a few ulp below the upper corner, every dimension
    k = 0
    for d in range(3):
        for expo in np.linspace(1.0, 13.2, 40):
            r[k, d] = BOX_MAX[d] - 10.0 ** -expo
            k += 1
        r[k, d] = BOX_MAX[d] - 64 * np.spacing(BOX_MAX[d])  # still distinguishable with 15 digits (closer ones make the reference throw)
        k += 1
        r[k, d] = BOX_MAX[d] * (1 - 1e-7)
        k += 1
    r[k] = [9.95, 19.95, 7.45]  # inside the 0.1 band, no extra digits needed
    r[k + 1] = [9.9999995, 19.9999995, 7.4999995]  # rounds to the border with 6 and 7 digits
    v = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-9, 9, (n, 3))
    f = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-4, 13, (n, 3))
    v[0] = [0.0, -0.0, 1.0]
    v[1] = [1e-310, -2.5e-320, 1e300]  # denormals, a huge value
    v[2] = [0.5, 0.25, 0.125]
    v[3] = [1234565.0, 0.0001234565, 2.5e-5]  # decimal ties of "%.6g" that are exact in binary / not
    v[4] = [999999.5, 9999995.0, 0.00001]  # carries into the next decade, the %e / %f switch at 1e-5 and 1e6
    v[5] = [100000.0, 1000000.0, 123456.0]
    f[0] = [np.inf, -np.inf, np.nan]
    f[1] = [1e100, -1e-100, 1e22]
    f[2] = [2.0 ** 70, 2.0 ** -70, -(2.0 ** 53)]
    types = rng.integers(0, 4, n).astype(np.int64)
    return ids, r, v, f, types


if __name__ == "__main__":
    ids, r, v, f, types = particles()
    assert all(oracle.vtk_position_precision(r[i, d], BOX_MAX[d]) > 0 for i in range(len(r)) for d in range(3))
    piece, index = oracle.ref_vtk_records(ids, r, v, f, types, BOX_MIN, BOX_MAX, "fixture", 42, 6)
    order = oracle.vtk_parse_ids(piece)
    assert sorted(order) == sorted(ids)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fn_vtk.npz"), ids=ids, r=r, v=v, f=f, types=types,
                        box_min=BOX_MIN, box_max=BOX_MAX, ref_piece=piece, ref_pvtu=index, session=np.array("fixture"),
                        iteration=np.int64(42), digits=np.int64(6))
    print("fn_vtk.npz:", len(piece), "bytes in the piece,", len(index), "in the index")
    # The loader's fixture: the same particles without the inf / nan row (the reference's own loader cannot read those
    # back: operator>> fails on them), written by the reference writer and read by the unmodified reference loader.
    import tempfile
    f2 = f.copy()
    f2[0] = [1.5e300, -2.5e-300, 4.9e-324]
    piece2, index2 = oracle.ref_vtk_records(ids, r, v, f2, types, BOX_MIN, BOX_MAX, "fixture", 42, 6)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "fixture", "data"))
        piece2.tofile(os.path.join(d, "fixture", "data", "fixture_Particles_0_000042.vtu"))
        index2.tofile(os.path.join(d, "fixture", "fixture_Particles_000042.pvtu"))
        back = oracle.ref_vtk_load(os.path.join(d, "fixture", "fixture_Particles_000042.pvtu"))
    assert np.array_equal(back["ids"], oracle.vtk_parse_ids(piece2))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fn_vtk_load.npz"), piece=piece2, pvtu=index2,
                        box_min=BOX_MIN, box_max=BOX_MAX, ref_ids=back["ids"], ref_r=back["r"], ref_v=back["v"], ref_f=back["f"],
                        ref_types=back["types"])
    print("fn_vtk_load.npz:", len(piece2), "bytes,", len(back["ids"]), "particles from the reference loader")
