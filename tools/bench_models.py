#!/usr/bin/env python3
"""Single-fluid collision + propagation (BASELINE config 1 family) for the three velocity sets and relaxation schemes:
MLUPS and fraction of the HBM roofline (algorithmic bytes 16 Q + 56 per site).  One GPU.
    python tools/bench_models.py [--size 256] [--steps 50]
Prints one JSON line per case."""
import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import ludwig_b200 as lb
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    n = args.size
    peak = 6543.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for nvel in (19, 15, 27):
        for name, nrelax in (("m10", lb.RELAX_M10), ("bgk", lb.RELAX_BGK), ("trt", lb.RELAX_TRT)):
            if nvel == 27 and nrelax == lb.RELAX_TRT:
                continue                      # not defined in the reference
            with lb.Lb200((n, n, n), nhalo=1, nvel=nvel) as sim:
                f = np.empty((nvel, sim.nsites))
                w = 1.0 / nvel
                f[...] = w
                sim.put(lb.F, f)
                cp = lb.CollideParam.make(nrelax, 1.0, 0.1, force=(1e-6, 0.0, 0.0))
                stream = torch.cuda.ExternalStream(sim.stream())
                sim.step(cp, None, 5)
                sim.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sim.step(cp, None, args.steps)
                e1.record(stream)
                sim.sync()
                ms = e0.elapsed_time(e1) / args.steps
            b = 16.0 * nvel + 56.0
            mlups = n ** 3 / (ms * 1e-3) / 1e6
            print(json.dumps({"workload": f"D3Q{nvel} single fluid {name}, {n}^3, pull-stream + collide (1 launch per step)",
                              "MLUPS": round(mlups, 1), "ms_per_step": round(ms, 4), "algorithmic_bytes_per_site": b,
                              "GBs": round(mlups * 1e6 * b / 1e9, 1), "frac_of_hbm_peak": round(mlups * 1e6 * b / 1e9 / peak, 4)}), flush=True)


def symmetric_lb(n, steps, peak):
    """`free_energy symmetric_lb` (two distributions, lb_collision_binary), D3Q19"""
    import numpy as np
    import torch
    import ludwig_b200 as lb
    with lb.Lb200((n, n, n), nhalo=1, nvel=19, ndist=2, have_phi=True) as sim:
        f = np.zeros((38, sim.nsites))
        f[:19] = 1.0 / 19
        rng = np.random.default_rng(1)
        f[19] = 0.05 * (rng.random(sim.nsites) - 0.5)
        sim.put(lb.F, f)
        del f
        cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, 0.00625)
        sp = lb.SymmParam.make(-0.00625, 0.00625, 0.004, 3.75)
        stream = torch.cuda.ExternalStream(sim.stream())
        sim.step(cp, sp, 3)
        sim.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.step(cp, sp, steps)
        e1.record(stream)
        sim.sync()
        ms = e0.elapsed_time(e1) / steps
        sim.profile(True); sim.step(cp, sp, 10); sim.sync()
        prof = {k: round(t / c, 4) for k, (t, c) in sim.profile_get().items() if c}
        sim.profile(False)
    # phi = sum g (152 + 8), gradient (8 + 32), two-distribution pull-collide (2*19*16 + 8 + 24 + 8 + 24 force-free)
    b = 160.0 + 40.0 + 672.0
    mlups = n ** 3 / (ms * 1e-3) / 1e6
    print(json.dumps({"workload": f"D3Q19 symmetric_lb (two distributions), {n}^3", "MLUPS": round(mlups, 1), "ms_per_step": round(ms, 4),
                      "algorithmic_bytes_per_site": b, "GBs": round(mlups * 1e6 * b / 1e9, 1),
                      "frac_of_hbm_peak": round(mlups * 1e6 * b / 1e9 / peak, 4), "ms_per_launch": prof}), flush=True)


if __name__ == "__main__":
    if "--symmetric-lb" in sys.argv:
        sys.argv.remove("--symmetric-lb")
        ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=256); ap.add_argument("--steps", type=int, default=30)
        a = ap.parse_args()
        symmetric_lb(a.size, a.steps, 6543.7)
        sys.exit(0)
    main()
