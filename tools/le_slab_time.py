"""python tools/le_slab_time.py [nx ny nz nplanes] -- device time per step of a sheared binary-fluid slab on one GPU (fast mode,
periodic in x on its own): what one rank of BASELINE config 5 (512 x 256 x 256 over 8 GPUs, one plane each) computes, without
the neighbour exchange."""
import os
import sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ludwig_b200 as lb
from ludwig_b200.initial import spinodal_phi

n = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 256, 256)
npl = int(sys.argv[4]) if len(sys.argv) >= 5 else 1
for planes in (npl, 0):
    with lb.Lb200(n, nhalo=2, have_phi=True, math=lb.MATH_FAST, le_nplanes=planes, le_uy=0.05) as sim:
        f = np.zeros((19, sim.nsites_lb if planes else sim.nsites))
        w = np.array([12.0] + [2.0 if sum(abs(c) for c in cv) == 1 else 1.0 for cv in lb.capi.CV19[1:]]) / 36.0 if hasattr(lb.capi, "CV19") else None
        f[...] = (1.0 / 19.0) if w is None else w[:, None]
        phi = np.zeros((1, sim.nsites))
        phi[:, :f.shape[1]] = spinodal_phi(n, 2, 13, 0.0, 0.1)
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, 0.1)
        sp = lb.SymmParam.make(-0.0625, 0.0625, 0.04, 0.15, adv_order=3)
        stream = torch.cuda.ExternalStream(sim.stream())
        sim.step(cp, sp, 10); sim.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); sim.step(cp, sp, 100); e1.record(stream); sim.sync()
        ms = e0.elapsed_time(e1) / 100
        print(f"{n} planes={planes}: {ms:.4f} ms/step, {n[0]*n[1]*n[2]/ms/1e3:.0f} MLUPS", flush=True)
