"""python tools/le_small.py -- quick check of the one-kernel step with Lees-Edwards planes against the oracle (one GPU)"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_le as t
from common import close_fast, rel_err
for n, npl, order, calls in (((16, 16, 40), 1, 3, 1), ((32, 26, 64), 2, 3, 2)):
    orc, got, want, prof = t._run_steps_profiled(n, npl, order, 6, calls)
    print(n, npl, {k: v[1] for k, v in prof.items() if v[1]})
    for k in want:
        g, w = orc.interior(got[k]), orc.interior(want[k])
        print("  ", k, close_fast(g, w), rel_err(g, w), flush=True)
