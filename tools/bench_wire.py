"""Throughput of apb_serialize_particles / apb_deserialize_particles (md-flexible's 120-byte MPI record) on one B200:
N particles device -> host buffer and back. The calls end in a PCIe transfer of N * 120 bytes, which bounds them; the
packing kernel itself moves N * (15 words read + 15 written) * 8 B through HBM. One JSON line.
usage: python tools/bench_wire.py [n=4000000]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autopas_b200 import GpuParticleContainer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rng = np.random.default_rng(0)
L = 200.0
r = rng.uniform(0, L, (n, 3))
c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
c.addParticles(r[:, 0], r[:, 1], r[:, 2], np.arange(n))
best_s, best_d = 1e9, 1e9
for _ in range(4):
    t0 = time.perf_counter()
    data = c.serializeParticles("owned")
    best_s = min(best_s, time.perf_counter() - t0)
    d = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    d.reserve(n, 0)
    t0 = time.perf_counter()
    d.deserializeParticles(data)
    best_d = min(best_d, time.perf_counter() - t0)
    d.close()
print(json.dumps({"config": "md-flexible MPI wire format, MoleculeLJ 120-byte records", "particles": n, "bytes": n * 120,
                  "serialize_ms": best_s * 1e3, "serialize_GBps": n * 120 / best_s * 1e-9,
                  "deserialize_ms": best_d * 1e3, "deserialize_GBps": n * 120 / best_d * 1e-9,
                  "note": "host-timed calls incl. the device<->host copy of the records (pageable numpy buffer)"}))
c.close()
