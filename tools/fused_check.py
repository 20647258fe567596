"""python tools/fused_check.py 40x26x64 -- a few whole binary-fluid steps through lb200_step on one GPU against the CPU oracle
(every field); LB200_FUSED_WS / LB200_PS_XC / LB200_MATH select the kernel variant.  Used under compute-sanitizer."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import ludwig_b200 as lb
from common import BINARY, ETA, close_fast, rel_err, seeded_state
from oracle import Oracle
nlocal = tuple(int(x) for x in sys.argv[1].split('x'))
orc = Oracle(nlocal, nhalo=2)
st0 = seeded_state(orc)
st = {k: v.copy() for k, v in st0.items()}
n = 6
orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=3, **BINARY), 1, n, st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
with lb.Lb200(nlocal, nhalo=2, have_phi=True, math=lb.MATH_STRICT if os.environ.get('LB200_MATH') == 'strict' else lb.MATH_FAST) as sim:
    sim.put(lb.F, st0["f"]); sim.put(lb.PHI, st0["phi"])
    sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), lb.SymmParam.make(adv_order=3, **BINARY), n)
    for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("force", lb.FORCE), ("rho", lb.RHO), ("grad", lb.GRAD)):
        g = sim.get(a)
        print(k, close_fast(orc.interior(g), orc.interior(st[k])), rel_err(orc.interior(g), orc.interior(st[k])))
