"""Seeded particle scenarios shared by the CPU and GPU parity tests.

They follow the reference's own parity harness (tests/testAutopas/tests/containers/TraversalComparison.cpp:135-199:
uniformly random owned particles in the box plus random halo particles in the halo shell) and md-flexible's lattice
generators (tools/autopasTools/generators/GridGenerator.h:90-105) for the BASELINE configs.
"""
import numpy as np


def uniform_with_halo(n, nhalo, box_max, cutoff, seed=42, ntypes=1):
    rng = np.random.default_rng(seed)
    box_max = np.asarray(box_max, dtype=float)
    owned = rng.uniform(0.0, 1.0, (n, 3)) * box_max
    halo = []
    while len(halo) < nhalo:
        p = rng.uniform(-cutoff, 1.0, 3) * 1.0
        p = rng.uniform(-cutoff, box_max + cutoff)
        if np.any(p < 0) or np.any(p >= box_max):
            halo.append(p)
    halo = np.array(halo).reshape(-1, 3)
    pos = np.vstack([owned, halo])
    own = np.r_[np.ones(n), 2 * np.ones(nhalo)].astype(np.int64)
    types = rng.integers(0, ntypes, n + nhalo).astype(np.int64)
    return pos, own, types


def periodic_images(pos, box_min, box_max, width):
    """All periodic images of owned particles that lie within `width` outside the box (what md-flexible's halo exchange
    produces, RegularGridDecomposition.cpp:159-236). Returns (positions, source index)."""
    box_min, box_max = np.asarray(box_min, float), np.asarray(box_max, float)
    L = box_max - box_min
    out_pos, out_src = [], []
    shifts = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1) if (a, b, c) != (0, 0, 0)]
    idx = np.arange(len(pos))
    for s in shifts:
        p = pos + np.array(s) * L
        m = np.all((p >= box_min - width) & (p < box_max + width), axis=1)
        out_pos.append(p[m])
        out_src.append(idx[m])
    return np.vstack(out_pos), np.concatenate(out_src)


def grid_lattice(n_per_dim, spacing, jitter=0.0, seed=42):
    """GridGenerator::fillWithParticles: ids z-major, id = (z*ny+y)*nx+x; box padded by spacing/2 like md-flexible
    (MDFlexConfig.cpp:477-482)."""
    nx = ny = nz = n_per_dim
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1).astype(float) * spacing
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos = pos + rng.uniform(-jitter, jitter, pos.shape)
    box_min = np.full(3, -spacing / 2)
    box_max = np.full(3, (n_per_dim - 1) * spacing + spacing / 2)
    return pos, box_min, box_max
