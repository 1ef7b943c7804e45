"""Golden bytes of md-flexible's MPI wire format: seeded particles serialised by the UNMODIFIED reference
(oracle/_ref/libautopas_ref_wire.so = examples/md-flexible/src/ParticleSerializationTools.cpp compiled where it lies).
Run in the build container:  python tests/golden/make_golden_wire.py  ->  tests/golden/fn_wire_format.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def particles(n=37, seed=2024):
    rng = np.random.default_rng(seed)
    ids = rng.permutation(10 * n)[:n].astype(np.int64)
    ids[0] = 2 ** 40 + 5  # ids beyond 32 bits survive
    r, v, f, oldf = (rng.normal(0, s, (n, 3)) for s in (10.0, 1.0, 100.0, 100.0))
    r[1] = [-0.0, 1e-310, 1e300]  # signed zero, a denormal, a huge value: the format is a memcpy
    types = rng.integers(0, 4, n).astype(np.int64)
    own = rng.integers(1, 3, n).astype(np.int64)  # owned / halo
    own[1] = 2  # (the particle with the huge coordinate is a halo copy: owned particles must lie inside a container box)
    return ids, r, v, f, oldf, types, own


if __name__ == "__main__":
    p = particles()
    data = oracle.ref_wire_serialize(*p)
    back = oracle.ref_wire_deserialize(data)
    assert np.array_equal(back["id"], p[0]) and np.array_equal(back["r"], p[1]) and np.array_equal(back["own"], p[6])
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fn_wire_format.npz"), ids=p[0], r=p[1], v=p[2], f=p[3],
             oldf=p[4], types=p[5], own=p[6], ref_bytes=data)
    print("fn_wire_format.npz:", len(data), "bytes")
