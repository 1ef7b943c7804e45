"""Generates tests/golden/fn_*.npz from the UNMODIFIED reference (oracle/_ref/libautopas_ref.so and
libautopas_ref_ms.so, built by `make -C oracle ref` in the development container where /root/reference exists).
Run from the repo root:  python tests/golden/make_golden_functors.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from functor_scenarios import atm_scenario, multisite_scenario, sph_scenario  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    # SPH: density, then hydro force with the reference's own densities as input (sph-main.cpp:206-279 order)
    for n3 in (0, 1):
        s = sph_scenario(seed=7, vary_h=(n3 == 0))  # newton3 picks roles by traversal order: uniform h keeps it symmetric
        n = len(s["pos"])
        d = oracle.ref_sph(s["pos"], s["vel"], s["mass"], s["smth"], np.zeros(n), s["pressure"], s["snd"], s["own"],
                           s["box_min"], s["box_max"], s["cutoff"], s["skin"], 0, bool(n3))
        dens_in = np.where(s["own"] == 1, d["density"], 1.0) + 0.5  # positive everywhere (halo densities are partial)
        h = oracle.ref_sph(s["pos"], s["vel"], s["mass"], s["smth"], dens_in, s["pressure"], s["snd"], s["own"],
                           s["box_min"], s["box_max"], s["cutoff"], s["skin"], 1, bool(n3))
        np.savez_compressed(os.path.join(OUT, f"fn_sph_n3{n3}.npz"), ref_density=d["density"], density_in=dens_in,
                            ref_acc=h["acc"], ref_engdot=h["engdot"], ref_vsigmax=h["vsigmax"], **s)
    s = atm_scenario(seed=11)
    r = oracle.ref_atm(s["pos"], s["types"], s["own"], s["box_min"], s["box_max"], s["cutoff"], s["skin"], nu=0.073)
    np.savez_compressed(os.path.join(OUT, "fn_atm.npz"), nu=0.073, ref_f=r["f"], ref_upot=r["upot"],
                        ref_virial=r["virial"], ref_flops=r["flops"], **s)
    s = atm_scenario(seed=12, ntypes=2)
    nu_t = np.array([0.073, 0.11])
    r = oracle.ref_atm(s["pos"], s["types"], s["own"], s["box_min"], s["box_max"], s["cutoff"], s["skin"], nu_of_type=nu_t)
    np.savez_compressed(os.path.join(OUT, "fn_atm_mix.npz"), nu_of_type=nu_t, ref_f=r["f"], ref_upot=r["upot"],
                        ref_virial=r["virial"], ref_flops=r["flops"], **s)
    for n3 in (0, 1):
        s = multisite_scenario(seed=13)
        r = oracle.ref_multisite(s["pos"], s["quat"], s["mol_type"], s["own"], s["box_min"], s["box_max"], s["cutoff"],
                                 s["skin"], True, bool(n3), s["eps"], s["sigma"], s["site_start"], s["site_pos"],
                                 s["site_type"])
        np.savez_compressed(os.path.join(OUT, f"fn_multisite_n3{n3}.npz"), ref_f=r["f"], ref_torque=r["torque"],
                            ref_upot=r["upot"], ref_virial=r["virial"], **s)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("fn_")))


if __name__ == "__main__":
    main()
