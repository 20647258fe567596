"""ludwig_b200 -- Blackwell-native (sm_100a) implementation of Ludwig's lattice-Boltzmann hot path.

The product is the C-ABI shared library `libludwig_b200.so` (include/ludwig_b200.h), built from
ludwig_b200/csrc (CUDA) and ludwig_b200/host (C host layer with the reference's function names).
This Python package is only a ctypes binding used by the tests and the benchmark; there is no
CPU fallback: importing works anywhere, creating a context without a CUDA device raises.
"""
from .capi import (Lb200, Lb200Error, slab_plan, SlabPlan, step_plan, StepPlan, STEP_PHI, STEP_UX, STEP_F, CollideParam, SymmParam, LcParam, Options, load_library, library_path,
                   F, PHI, U, RHO, FORCE, GRAD, DELSQ, MAP, GRAD_DELSQ, DELSQ_DELSQ, STR, Q, QGRAD, QDELSQ,
                   RELAX_M10, RELAX_BGK, RELAX_TRT, HALO_FULL, HALO_REDUCED, MATH_FAST, MATH_STRICT,
                   KNOB_WRAP, KNOB_PHI_SECTOR, KNOB_PEER, KNOB_PIPE, KNOB_PIPE_SMS, KNOB_F32, KNOB_GRAD_7PT, KNOB_FUSED, KNOB_QGRAD_2D5)

__all__ = ["Lb200", "Lb200Error", "slab_plan", "SlabPlan", "step_plan", "StepPlan", "STEP_PHI", "STEP_UX", "STEP_F", "CollideParam", "SymmParam", "LcParam", "Options", "load_library", "library_path",
           "F", "PHI", "U", "RHO", "FORCE", "GRAD", "DELSQ", "MAP", "GRAD_DELSQ", "DELSQ_DELSQ", "STR", "Q", "QGRAD", "QDELSQ",
           "RELAX_M10", "RELAX_BGK", "RELAX_TRT", "HALO_FULL", "HALO_REDUCED", "MATH_FAST", "MATH_STRICT",
           "KNOB_WRAP", "KNOB_PHI_SECTOR", "KNOB_PEER", "KNOB_PIPE", "KNOB_PIPE_SMS", "KNOB_F32", "KNOB_GRAD_7PT", "KNOB_FUSED", "KNOB_QGRAD_2D5"]
