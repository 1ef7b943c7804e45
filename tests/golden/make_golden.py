"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/libautopas_ref.so, built from
/root/reference by oracle/Makefile). Run in the development container:  python tests/golden/make_golden.py
Each .npz holds seeded inputs and the reference's outputs (forces by id, Upot, virial, getNumFLOPs, cell indices or the
cluster / cluster-pair structure). tests/test_oracle.py checks the C oracle against them; tests/test_gpu_parity.py the
CUDA path."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle  # noqa: E402
from scenarios import uniform_with_halo  # noqa: E402

CASES = [
    # name, container, n, nhalo, L, cutoff, skin, csf, M, shift, mixing, newton3
    ("lc_n3", "LinkedCells", 1200, 300, 8.0, 1.0, 0.1, 1.0, 0, True, False, True),
    ("lc_non3_mix", "LinkedCells", 1200, 300, 8.0, 1.0, 0.1, 1.0, 0, True, True, False),
    ("lc_csf05", "LinkedCells", 600, 150, 6.0, 1.0, 0.2, 0.5, 0, False, False, True),
    ("vcl4_non3", "VerletClusterLists", 1200, 300, 8.0, 1.0, 0.1, 1.0, 4, True, False, False),
    ("vcl4_n3_mix", "VerletClusterLists", 1200, 300, 8.0, 1.0, 0.1, 1.0, 4, True, True, True),
    ("vcl32_non3", "VerletClusterLists", 1500, 400, 8.0, 1.0, 0.1, 1.0, 32, False, False, False),
    ("vcl8_n3", "VerletClusterLists", 900, 200, 7.0, 1.0, 0.3, 1.0, 8, True, False, True),
]


def main():
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    for seed, (name, cont, n, nh, L, rc, skin, csf, M, shift, mixing, n3) in enumerate(CASES):
        ntypes = 2 if mixing else 1
        pos, own, types = uniform_with_halo(n, nh, [L, L, L], rc, seed=100 + seed, ntypes=ntypes)
        eps = np.array([1.0, 1.4][:ntypes])
        sigma = np.array([1.0, 0.85][:ntypes])
        bmin, bmax = np.zeros(3), np.full(3, L)
        out = dict(pos=pos, own=own, types=types, box_min=bmin, box_max=bmax, cutoff=rc, skin=skin, csf=csf,
                   cluster_size=M, shift=shift, mixing=mixing, newton3=n3, eps=eps, sigma=sigma, container=cont)
        if cont == "LinkedCells":
            r = oracle.ref_lj_linkedcells(pos[:, 0], pos[:, 1], pos[:, 2], types, own, bmin, bmax, rc, skin, csf,
                                          shift=shift, mixing=mixing, newton3=n3, soa=True, eps=eps, sigma=sigma)
            out.update(ref_cell=r["cell"])
        else:
            r = oracle.ref_lj_vcl(pos[:, 0], pos[:, 1], pos[:, 2], types, own, bmin, bmax, rc, skin, M, shift=shift,
                                  mixing=mixing, newton3=n3, soa=True,
                                  traversal="vcl_c06" if n3 else "vcl_cluster_iteration", eps=eps, sigma=sigma)
            out.update(ref_cluster_particles=r["cluster_particles"], ref_pairs=r["pairs"],
                       ref_towers_per_dim=np.array(r["towers_per_dim"]))
        out.update(ref_f=r["f"], ref_upot=r["upot"], ref_virial=r["virial"], ref_flops=np.uint64(r["flops"]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "upot", r["upot"], "flops", r["flops"])


def c1_full_size():
    """BASELINE configs[0] at full size (C1: 32^3 = 32 768 particles, spacing 1.1225, jittered, periodic images as halos,
    cutoff 2.5, skin 0.2): forces, Upot, virial and cell indices of the unmodified reference, LinkedCells / lc_c08 / SoA /
    newton3 with shift (the configuration the BASELINE names). Inputs are regenerated from the seed by the tests
    (scenarios.grid_lattice / periodic_images), so only the reference's outputs are stored."""
    from scenarios import grid_lattice, periodic_images
    rc, skin = 2.5, 0.2
    pos, bmin, bmax = grid_lattice(32, 1.1225, 0.1, 42)
    pos = bmin + np.mod(pos - bmin, bmax - bmin)
    hpos, _ = periodic_images(pos, bmin, bmax, rc + skin)
    allpos = np.vstack([pos, hpos])
    own = np.r_[np.ones(len(pos)), 2 * np.ones(len(hpos))].astype(np.int64)
    r = oracle.ref_lj_linkedcells(allpos[:, 0], allpos[:, 1], allpos[:, 2], None, own, bmin, bmax, rc, skin, 1.0, shift=True,
                                  newton3=True, soa=True)
    np.savez_compressed(os.path.join(HERE, "c1_full_size.npz"), n=len(pos), num_halo=len(hpos), cutoff=rc, skin=skin,
                        ref_f=r["f"][:len(pos)], ref_upot=r["upot"], ref_virial=r["virial"],
                        ref_cell=r["cell"].astype(np.int32), pos_checksum=float(allpos.sum()))
    print("c1_full_size upot", r["upot"], "virial", r["virial"])


if __name__ == "__main__":
    main()
    c1_full_size()
