"""Load a golden case (tests/golden/*.npz, produced by the reference: tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(name):
    d = np.load(os.path.join(HERE, "golden", name + ".npz"))
    out = {k: d[k] for k in d.files}
    for k in ("kind",):
        out[k] = str(out[k])
    for k in ("nvel", "nsteps", "nrelax", "nhalo", "adv_order", "reduced"):
        if k in out:
            out[k] = int(out[k])
    for k in ("eta", "eta_bulk", "a", "b", "kappa", "mobility"):
        if k in out:
            out[k] = float(out[k])
    out["nlocal"] = tuple(int(x) for x in out["nlocal"])
    return out
