"""Kernel experiment helper: times repeated force evaluations (no integration) of the C2 workload through the C ABI.
usage: [APB_LIB_PATH=variant.so] python tools/force_only.py [cluster_size] [reps]"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pos, vel, bmin, bmax, gmin, gmax = bench.make_workload(100, 0, [1, 1, 1])
n = len(pos)
halo = bench.periodic_images(pos, bmin, bmax, 2.8)
c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, 2.5, 0.3, clusterSize=M)
c.addParticles(pos[:, 0], pos[:, 1], pos[:, 2], np.arange(n))
c.addHaloParticles(halo[:, 0], halo[:, 1], halo[:, 2], np.arange(len(halo)) + n)
f = LJFunctor(2.5, applyShift=True, calculateGlobals=True, countFLOPs=True, virialTraceOnly=True)
f.setParticleProperties(24.0, 1.0)
t = GpuTraversal("gpuvcl_pruned", f, False)
c.rebuildNeighborLists(t)
for k in range(3):
    f.initTraversal(); c.computeInteractions(t); f.endTraversal(False)
best = 1e9
for k in range(reps):
    f.initTraversal()
    t0 = time.perf_counter()
    c.computeInteractions(t)
    best = min(best, time.perf_counter() - t0)
    f.endTraversal(False)
print(f"M={M} force call best {best*1e3:.4f} ms  upot {f.getPotentialEnergy():.6e} dist {f._raw.num_dist_calls} hits {f._raw.num_kernel_calls_no_n3}")
