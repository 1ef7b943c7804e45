// Warp-cooperative walk over the linked-cells neighbourhood of one particle slot (gpuLinkedCells kernels).
//
// One warp owns slot i. The stencil cells are visited in (z, y, x) order; cells adjacent in x are adjacent slot ranges
// (slots are sorted by cell, x fastest), so they merge into runs that the 32 lanes read with coalesced loads. Every lane
// tests one candidate per round (`cand(j)`: the cheap distance test; returns whether the pair interacts), the hits
// are compacted by ballot into a 64-entry queue of the warp in shared memory, and the pair arithmetic (`heavy(j)`) runs
// on full rows of 32 hits - no lane idles through a miss, which is what the one-thread-per-particle kernels paid for at
// hit rates of 7-15 % (CellFunctor.h:173-184 visits every particle pair of a cell pair). The lanes keep partial sums;
// the caller reduces them with shuffles.
#pragma once
#include "internal.cuh"

struct LCWarpGeom {
  LCGeom g;
  const int *cellStart;
  const int *stencilSorted;  // 3 ints per entry, (z, y, x) lexicographic, self included
  int stencilN;
};

// HIGHER: only candidates in slots above i (newton3: each pair once, owned by its lower slot; a cell with a lower index
// holds lower slots only). `filterHaloPairs`: slot i lives in a halo cell, so partner cells that cannot hold owned
// particles are skipped (CellFunctor.h:173-184).
// Returns the number of hits (warp-uniform).
template <bool HIGHER, class Cand, class Heavy>
__device__ __forceinline__ int lcWarpWalk(const LCWarpGeom &w, int64_t i, int c, bool filterHaloPairs, int *queue,
                                           Cand &&cand, Heavy &&heavy) {
  const LCGeom &g = w.g;
  const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  const int cx = c % g.cellsPerDim[0], cy = (c / g.cellsPerDim[0]) % g.cellsPerDim[1],
            cz = c / (g.cellsPerDim[0] * g.cellsPerDim[1]);
  int qn = 0, r0 = 0, r1 = 0, total = 0;
  auto flushRun = [&]() {
    for (int base = r0; base < r1; base += 32) {
      const int j = base + static_cast<int>(lane);
      const bool hit = j < r1 && cand(j);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) queue[qn + __popc(m & below)] = j;
      qn += __popc(m);
      total += __popc(m);
      __syncwarp();
      if (qn >= 32) {
        const int jj = queue[lane];
        const int carry = static_cast<int>(lane) < qn - 32 ? queue[32 + lane] : 0;
        __syncwarp();
        if (static_cast<int>(lane) < qn - 32) queue[lane] = carry;
        qn -= 32;
        heavy(jj);
        __syncwarp();
      }
    }
  };
  for (int s = 0; s < w.stencilN; ++s) {
    const int ox = w.stencilSorted[3 * s], oy = w.stencilSorted[3 * s + 1], oz = w.stencilSorted[3 * s + 2];
    const int lin = (oz * g.cellsPerDim[1] + oy) * g.cellsPerDim[0] + ox;
    if (HIGHER && lin < 0) continue;
    const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
    if (nx < 0 || ny < 0 || nz < 0 || nx >= g.cellsPerDim[0] || ny >= g.cellsPerDim[1] || nz >= g.cellsPerDim[2]) continue;
    if (filterHaloPairs && !apbCellCanOwn(g, nx, ny, nz)) continue;
    const int c2 = c + lin;
    int j0 = w.cellStart[c2];
    const int j1 = w.cellStart[c2 + 1];
    if (HIGHER) j0 = max(j0, static_cast<int>(i) + 1);
    if (j0 >= j1) continue;
    if (j0 == r1) {
      r1 = j1;
    } else {
      flushRun();
      r0 = j0;
      r1 = j1;
    }
  }
  flushRun();
  if (static_cast<int>(lane) < qn) heavy(queue[lane]);
  __syncwarp();
  return total;
}

__device__ __forceinline__ double lcWarpSum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double lcWarpMax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#define LCW_WARPS 8  // warps (= particle slots in flight) per block
