"""Seeded inputs for the SPH / Axilrod-Teller / multi-site functor tests (shared by oracle, golden and GPU tests)."""
import numpy as np


def lattice_with_halo(n_per_dim, spacing, jitter, il, seed):
    """Jittered simple-cubic lattice over the box [0, L)^3 and its halo shell of width il. Returns pos, own, L."""
    rng = np.random.default_rng(seed)
    L = n_per_dim * spacing
    k = int(np.ceil(il / spacing)) + 1
    g = (np.arange(-k, n_per_dim + k) + 0.5) * spacing
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    pos = pos + rng.uniform(-jitter, jitter, pos.shape)
    inside = np.all((pos >= 0) & (pos < L), axis=1)
    in_halo = np.all((pos >= -il) & (pos < L + il), axis=1) & ~inside
    keep = inside | in_halo
    own = np.where(inside, 1, 2)[keep].astype(np.int64)
    pos = pos[keep]
    order = np.argsort(own, kind="stable")  # owned first
    return pos[order], own[order], L


def sph_scenario(seed=7, vary_h=True):
    cutoff, skin = 1.0, 0.1
    pos, own, L = lattice_with_halo(6, 0.35, 0.08, cutoff + skin, seed)
    rng = np.random.default_rng(seed + 1)
    n = len(pos)
    vel = rng.normal(0, 0.5, (n, 3))
    mass = rng.uniform(0.8, 1.2, n)
    smth = rng.uniform(0.30, 0.40, n) if vary_h else np.full(n, 0.36)  # support 2.5 h <= cutoff
    pressure = rng.uniform(0.5, 1.5, n)
    snd = rng.uniform(1.0, 1.4, n)
    return dict(pos=pos, vel=vel, mass=mass, smth=smth, pressure=pressure, snd=snd, own=own, box_min=np.zeros(3),
                box_max=np.full(3, L), cutoff=cutoff, skin=skin)


def atm_scenario(seed=11, ntypes=1):
    cutoff, skin = 2.0, 0.2
    pos, own, L = lattice_with_halo(5, 1.1, 0.15, cutoff + skin, seed)
    rng = np.random.default_rng(seed + 1)
    types = rng.integers(0, ntypes, len(pos)).astype(np.int64)
    return dict(pos=pos, types=types, own=own, box_min=np.zeros(3), box_max=np.full(3, L), cutoff=cutoff, skin=skin)


def multisite_scenario(seed=13):
    cutoff, skin = 2.5, 0.2
    pos, own, L = lattice_with_halo(4, 1.6, 0.1, cutoff + skin, seed)
    rng = np.random.default_rng(seed + 1)
    n = len(pos)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    mol_type = rng.integers(0, 2, n).astype(np.int64)
    eps, sigma = [1.0, 0.7], [1.0, 0.8]
    # molecule type 0: dumbbell of two type-0 sites; type 1: bent three-site molecule with mixed site types
    site_start = [0, 2, 5]
    site_pos = [[0.25, 0, 0], [-0.25, 0, 0], [0, 0, 0.1], [0.3, 0, -0.1], [-0.3, 0, -0.1]]
    site_type = [0, 0, 1, 0, 0]
    return dict(pos=pos, quat=q, mol_type=mol_type, own=own, box_min=np.zeros(3), box_max=np.full(3, L), cutoff=cutoff,
                skin=skin, eps=eps, sigma=sigma, site_start=site_start, site_pos=site_pos, site_type=site_type)
