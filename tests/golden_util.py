"""Load a golden case (tests/golden/*.npz, produced by the reference: tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def case_names(prefix=None):
    """prefix None: the single-fluid / binary / symmetric_lb cases; "le_" / "lc_": Lees-Edwards / liquid crystal"""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))
    if prefix is None:
        return [n for n in names if not n.startswith(("le_", "lc_"))]
    return [n for n in names if n.startswith(prefix)]


def load(name):
    d = np.load(os.path.join(HERE, "golden", name + ".npz"))
    out = {k: d[k] for k in d.files}
    for k in ("kind",):
        out[k] = str(out[k])
    for k in ("nvel", "nsteps", "nrelax", "nhalo", "adv_order", "reduced", "nplanes"):
        if k in out:
            out[k] = int(out[k])
    for k in ("eta", "eta_bulk", "a", "b", "kappa", "mobility", "uy", "lc_a0", "lc_q0", "lc_gamma", "lc_kappa0", "lc_kappa1",
              "lc_xi", "lc_Gamma", "lc_epsilon"):
        if k in out:
            out[k] = float(out[k])
    out["nlocal"] = tuple(int(x) for x in out["nlocal"])
    return out
