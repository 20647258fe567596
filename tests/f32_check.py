"""Measured errors of the FP32-storage mode (LB200_KNOB_F32) against the FP64 oracle (test infrastructure, like the
oracle itself): python tests/f32_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ludwig_b200 as lb                      # noqa: E402
from test_gpu_parity import f32_errors        # noqa: E402

for nlocal, nrelax, binary, u_amp, nsteps in (((24, 20, 16), lb.RELAX_M10, True, 0.01, 40), ((24, 20, 16), lb.RELAX_M10, True, 0.01, 400),
                                              ((16, 16, 40), lb.RELAX_TRT, True, 0.01, 40), ((20, 12, 24), lb.RELAX_BGK, False, 0.05, 40),
                                              ((32, 32, 32), lb.RELAX_M10, True, 0.01, 1000)):
    err, dev, scale = f32_errors(nlocal, nrelax, lb.MATH_FAST, nsteps, binary=binary, u_amp=u_amp)
    print(nlocal, "relax", nrelax, "binary", binary, "steps", nsteps, "| max|f - w| %.3e" % dev, "| bound N 2^-23 D %.3e" % (nsteps * 2.0 ** -23 * dev),
          "| abs err", {k: "%.3e" % v for k, v in err.items()}, "| scale", {k: "%.3e" % v for k, v in scale.items()}, flush=True)
