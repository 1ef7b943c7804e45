"""Kernel experiment helper: times repeated force evaluations (no integration) of the C2 / C3 workload through the C ABI.
usage: [APB_LIB_PATH=variant.so] python tools/force_only.py [cluster_size] [reps] [c2|c3] [n_per_dim] [traversal] [n3]"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"
npd = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else (100 if wl == "c2" else 126)
trav = sys.argv[5] if len(sys.argv) > 5 else "gpuvcl_pruned"
n3 = len(sys.argv) > 6 and sys.argv[6] == "n3"
skin = bench.C2["skin"] if wl == "c2" else bench.C3["skin"]
pos, vel, bmin, bmax, gmin, gmax = bench.make_workload(wl, npd, 0, [1, 1, 1])
n = len(pos)
c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, 2.5, skin, clusterSize=M)
c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
c.migrate()
c.exchangeHalos()
f = LJFunctor(2.5, applyShift=True, calculateGlobals=True, countFLOPs=True, virialTraceOnly=True)
f.setParticleProperties(24.0, 1.0)
t = GpuTraversal(trav, f, n3)
import torch
t0 = time.perf_counter()
c.rebuildNeighborLists(t)
tb = time.perf_counter() - t0
t0 = time.perf_counter()
c.rebuildNeighborLists(t)
tb = min(tb, time.perf_counter() - t0)
for k in range(3):
    f.initTraversal(); c.computeInteractions(t); f.endTraversal(n3)
best = 1e9
for k in range(reps):
    f.initTraversal()
    t0 = time.perf_counter()
    c.computeInteractions(t)
    best = min(best, time.perf_counter() - t0)
    f.endTraversal(n3)
r = f._raw
flops = 8 * r.num_dist_calls + 15 * r.num_kernel_calls_no_n3 + 18 * r.num_kernel_calls_n3 + 9 * r.num_global_calcs_no_n3 + 13 * r.num_global_calcs_n3
print(f"{wl} n={n} M={M} {trav}{' newton3' if n3 else ''}: force call best {best*1e3:.4f} ms ({flops / best / 1e12:.2f} TFLOP/s), rebuild call {tb*1e3:.3f} ms, "
      f"upot {f.getPotentialEnergy():.9e} dist {r.num_dist_calls} hits {r.num_kernel_calls_no_n3 + r.num_kernel_calls_n3}")
