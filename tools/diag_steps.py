import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from autopas_b200 import GpuParticleContainer, GpuTraversal, LJFunctor
pos, vel, bmin, bmax, gmin, gmax = bench.make_workload("c2", 100, 0, [1,1,1])
n = len(pos)
c = GpuParticleContainer("gpuVerletClusterLists", bmin, bmax, 2.5, 0.3, clusterSize=32)
c.addParticles(pos[:,0], pos[:,1], pos[:,2], np.arange(n))
for d, name in enumerate(("VX","VY","VZ")): c.uploadColumn(name, vel[:,d])
f = LJFunctor(2.5, applyShift=True, calculateGlobals=True, countFLOPs=True); f.setParticleProperties(24.0, 1.0)
t = GpuTraversal("gpuvcl_pruned", f, False)
c.enableLoopTiming(True)
it = 0
for blk in range(6):
    t0 = time.perf_counter()
    res = c.runSteps(t, 10, it, 0.002, [1.0], 10)
    wall = time.perf_counter() - t0
    it += 10
    tm = c.getLoopTiming()
    print(blk, 'wall ms/step', wall*100, {k: (round(v[0]/max(v[1],1),4), v[1]) for k, v in tm.items()}, 'upot', res[9].upot_sum/12, 'dist', res[9].num_dist_calls, flush=True)
