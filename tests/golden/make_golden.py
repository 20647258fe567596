#!/usr/bin/env python3
"""Generate the golden vectors in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, built
from /root/reference by oracle/Makefile.ref) through oracle/ref_harness.c.  Run in the build container:

    python tests/golden/make_golden.py

Each .npz holds the inputs (canonical layout, see include/ludwig_b200.h) and the reference's outputs
after `nsteps` whole time steps, plus the parameters.  Interiors only are meaningful for outputs."""
import os
import zlib
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import refharness as rh  # noqa: E402

BINARY = dict(a=-0.00625, b=0.00625, kappa=0.004, mobility=1.25)

CASES = {
    # name: (kind, nvel, nlocal, nsteps, options)
    "binary_o1": ("binary", 19, (8, 6, 7), 5, dict(adv_order=1, nrelax=0, fbody=(0.0, 0.0, 0.0), gradmu=(0, 0, 0))),
    "binary_o3_force": ("binary", 19, (6, 8, 9), 5, dict(adv_order=3, nrelax=0, fbody=(1e-6, 2e-6, 3e-6), gradmu=(1e-5, 0, -1e-5))),
    "binary_o2_trt": ("binary", 19, (5, 5, 6), 4, dict(adv_order=2, nrelax=2, fbody=(0.0, 0.0, 0.0), gradmu=(0, 0, 0))),
    # advection order 4 (advection_le_4th) and cahn_hilliard_options_conserve 1 (phi_ch_update_conserve)
    "binary_o4_conserve": ("binary", 19, (7, 6, 8), 6, dict(adv_order=4, conserve=1, nrelax=0, fbody=(1e-6, 0.0, -1e-6), gradmu=(0, 0, 0))),
    "single_d3q19_m10": ("single", 19, (6, 5, 7), 6, dict(nrelax=0, reduced=0, fbody=(1e-6, 2e-6, 3e-6))),
    "single_d3q19_bgk_reduced": ("single", 19, (6, 5, 7), 6, dict(nrelax=1, reduced=1, fbody=(0.0, 0.0, 0.0))),
    "single_d3q15_trt": ("single", 15, (5, 6, 4), 6, dict(nrelax=2, reduced=0, fbody=(1e-6, 0.0, 0.0))),
    "single_d3q27_bgk": ("single", 27, (4, 5, 6), 6, dict(nrelax=1, reduced=0, fbody=(0.0, 2e-6, 0.0))),
    # free_energy symmetric_lb: two distributions, lb_collision_binary
    "symmlb_d3q19": ("symmlb", 19, (6, 7, 8), 5, dict(nrelax=0, reduced=0, fbody=(1e-6, -2e-6, 1e-6), mobility=3.75)),
    "symmlb_d3q19_trt_reduced": ("symmlb", 19, (5, 6, 7), 4, dict(nrelax=2, reduced=1, fbody=(0.0, 0.0, 0.0), mobility=0.45)),
    "symmlb_d3q15": ("symmlb", 15, (5, 5, 6), 4, dict(nrelax=0, reduced=0, fbody=(0.0, 1e-6, 0.0), mobility=3.75)),
}


def main():
    only = sys.argv[1:]            # python make_golden.py [case ...]: regenerate the named cases only
    for name, (kind, nvel, nlocal, nsteps, o) in CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        out = dict(kind=kind, nvel=nvel, nlocal=np.array(nlocal), nsteps=nsteps, nrelax=o["nrelax"],
                   fbody=np.array(o["fbody"], dtype=float))
        if kind == "binary":
            nhalo = 2
            with rh.RefSim(nlocal, nhalo=nhalo, have_phi=1, adv_order=o["adv_order"], conserve=o.get("conserve", 0), nrelax=o["nrelax"],
                           eta_shear=0.00625, fbody=o["fbody"], gradmu=o["gradmu"], **BINARY) as s:
                s.init_rest(1.0)
                s.init_spinodal(8361235, 0.0, 0.1)
                out["f0"], out["phi0"] = s.get(rh.REF_F), s.get(rh.REF_PHI)
                s.step(nsteps)
                for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U), ("rho", rh.REF_RHO),
                             ("force", rh.REF_FORCE), ("grad", rh.REF_GRAD), ("delsq", rh.REF_DELSQ)):
                    out[k] = s.get(w)
            out.update(nhalo=nhalo, adv_order=o["adv_order"], conserve=o.get("conserve", 0), gradmu=np.array(o["gradmu"], dtype=float),
                       eta=0.00625, **BINARY)
        elif kind == "symmlb":
            nhalo = 1
            par = dict(BINARY, mobility=o["mobility"])
            with rh.RefSim(nlocal, nhalo=nhalo, nvel=nvel, ndist=2, have_phi=1, nrelax=o["nrelax"], eta_shear=0.00625,
                           halo_reduced=o["reduced"], fbody=o["fbody"], **par) as s:
                s.init_rest(1.0)
                s.init_spinodal(8361235, 0.0, 0.1)
                s.op("phi_lb_from_field")
                out["f0"] = s.get(rh.REF_F)
                s.step(nsteps)
                for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U), ("grad", rh.REF_GRAD),
                             ("delsq", rh.REF_DELSQ)):
                    out[k] = s.get(w)
            out.update(nhalo=nhalo, reduced=o["reduced"], eta=0.00625, **par)
        else:
            nhalo = 1
            with rh.RefSim(nlocal, nhalo=nhalo, nrelax=o["nrelax"], halo_reduced=o["reduced"], eta_shear=0.05,
                           eta_bulk=0.08, fbody=o["fbody"], nvel=nvel) as s:
                s.init_uniform_u(1.0, (0.01, -0.02, 0.015))
                f0 = s.get(rh.REF_F)
                h = nhalo
                v = f0.reshape((nvel,) + tuple(n + 2 * h for n in nlocal))
                v[:, h:-h, h:-h, h:-h] *= 1.0 + 1e-2 * (rng.random((nvel,) + nlocal) - 0.5)
                s.set(rh.REF_F, f0)
                out["f0"] = f0
                s.step(nsteps)
                for k, w in (("f", rh.REF_F), ("u", rh.REF_U), ("rho", rh.REF_RHO)):
                    out[k] = s.get(w)
            out.update(nhalo=nhalo, reduced=o["reduced"], eta=0.05, eta_bulk=0.08)
        # halo values of the outputs are stale/undefined in the reference: blank them for compactness
        hh = out["nhalo"]
        nall = tuple(n + 2 * hh for n in nlocal)
        for k in ("f", "phi", "u", "rho", "force", "grad", "delsq"):
            if k in out:
                a = out[k].reshape((-1,) + nall).copy()
                keep = a[:, hh:-hh, hh:-hh, hh:-hh].copy()
                a[...] = 0.0
                a[:, hh:-hh, hh:-hh, hh:-hh] = keep
                out[k] = a.reshape(out[k].shape)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name, {k: getattr(v, "shape", v) for k, v in out.items() if k in ("f", "phi")})


if __name__ == "__main__":
    main()
