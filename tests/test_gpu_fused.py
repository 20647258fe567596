"""The one-kernel binary-fluid step (LB200_KNOB_FUSED, ludwig_b200/csrc/lb200_fused.cuh): phi sector + pull-stream +
collision of the same plane in one sweep, the populations staged by TMA tensor copies, the y / z periodic images of f'
stored by the producing kernel.  Checked against the CPU oracle on identical inputs (every field the reference's step
leaves behind: f, phi, u, rho, force, grad, delsq), against the two-kernel step of the same library, and through the
transitions between the two (first step after an upload, single operators after a fused step, several calls).

Bar: fast mode within 1e-12 relative (absolute floor 1e-14 for velocities); strict mode (the same kernel compiled without FMA
contraction, its phi-sector warps in the reference's operation order) BIT-EXACT."""
import numpy as np
import pytest

import ludwig_b200 as lb
from common import BINARY, ETA, close_fast, rel_err, seeded_state
from oracle import Oracle

pytestmark = pytest.mark.gpu

FIELDS = (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("rho", lb.RHO), ("force", lb.FORCE), ("grad", lb.GRAD),
          ("delsq", lb.DELSQ))
FG = (1e-6, -2e-6, 5e-7)
GM = (1e-5, -2e-5, 3e-5)


def gpu_run(nlocal, st, order, nrelax, nsteps, calls=1, fused=1, math=lb.MATH_FAST):
    with lb.Lb200(nlocal, nhalo=2, have_phi=True, math=math) as sim:
        sim.set_knob(lb.KNOB_FUSED, fused)
        sim.put(lb.F, st["f"]); sim.put(lb.PHI, st["phi"])
        cp = lb.CollideParam.make(nrelax, 1.0, ETA, force=FG)
        sp = lb.SymmParam.make(adv_order=order, gradmu=GM, **BINARY)
        sim.profile(True)
        for c in range(calls):
            sim.step(cp, sp, nsteps // calls)
        sim.sync()
        prof = sim.profile_get()
        sim.profile(False)
        got = {k: sim.get(a) for k, a in FIELDS}
    return got, prof


def oracle_run(nlocal, order, nrelax, nsteps):
    orc = Oracle(nlocal, nhalo=2)
    st0 = seeded_state(orc)
    st = {k: v.copy() for k, v in st0.items()}
    orc.step(orc.collide_param(nrelax, 1.0, ETA, force=FG), orc.symm_param(adv_order=order, gradmu=GM, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    return orc, st0, st


# 40 x 26 x 64: 4 x 3 tiles of 8 x 30 columns (the last ones ragged), chunks of 8 .. 40 planes
@pytest.mark.parametrize("math_mode", [lb.MATH_FAST, lb.MATH_STRICT], ids=["fast", "strict"])
@pytest.mark.parametrize("xc", [0, 9, 20, 40])
@pytest.mark.parametrize("order,nrelax", [(1, lb.RELAX_M10), (2, lb.RELAX_TRT), (3, lb.RELAX_M10), (3, lb.RELAX_BGK)])
def test_fused_step_matches_the_oracle(order, nrelax, xc, math_mode, monkeypatch):
    if xc:
        monkeypatch.setenv("LB200_PS_XC", str(xc))
    nlocal = (40, 26, 64)
    orc, st0, st = oracle_run(nlocal, order, nrelax, 12)
    got, prof = gpu_run(nlocal, st0, order, nrelax, 12, math=math_mode)
    # the first step after the upload collides in place (two kernels), the other 11 are one kernel each
    assert prof["step_fused"][1] == 11 and prof["collide"][1] == 1 and prof["phi_sector"][1] == 1, prof
    for k in got:
        if math_mode == lb.MATH_STRICT:
            assert np.array_equal(orc.interior(got[k]), orc.interior(st[k])), k
        else:
            assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))


def test_every_warp_does_both_variant(monkeypatch):
    """LB200_FUSED_WS=0: the earlier form of the one-kernel step (no warp specialisation), kept for comparison runs"""
    monkeypatch.setenv("LB200_FUSED_WS", "0")
    nlocal = (40, 26, 64)
    orc, st0, st = oracle_run(nlocal, 3, lb.RELAX_M10, 8)
    got, prof = gpu_run(nlocal, st0, 3, lb.RELAX_M10, 8)
    assert prof["step_fused"][1] == 7
    for k in got:
        assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))


@pytest.mark.parametrize("math_mode", [lb.MATH_FAST, lb.MATH_STRICT], ids=["fast", "strict"])
@pytest.mark.parametrize("nlocal", [(16, 8, 8), (8, 30, 32), (24, 62, 34), (12, 4, 4)])
def test_fused_step_small_and_ragged_lattices(nlocal, math_mode):
    """one tile holding both periodic boundaries, tiles whose last row / column is cut, lattices as thin as the halo allows"""
    orc, st0, st = oracle_run(nlocal, 3, lb.RELAX_M10, 9)
    got, prof = gpu_run(nlocal, st0, 3, lb.RELAX_M10, 9, calls=3, math=math_mode)
    assert prof["step_fused"][1] == 8, prof
    for k in got:
        if math_mode == lb.MATH_STRICT:
            assert np.array_equal(orc.interior(got[k]), orc.interior(st[k])), k
        else:
            assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))


def test_fused_and_two_kernel_steps_agree():
    """same library, knob on / off: identical collision arithmetic, phi sector re-associated identically -> the two
    paths may differ only by x-chunk boundaries of the march (a few ulp), far inside the tolerance"""
    nlocal = (32, 32, 60)
    orc = Oracle(nlocal, nhalo=2)
    st0 = seeded_state(orc)
    a, pa = gpu_run(nlocal, st0, 3, lb.RELAX_M10, 20, fused=1)
    b, pb = gpu_run(nlocal, st0, 3, lb.RELAX_M10, 20, fused=0)
    assert pa["step_fused"][1] == 19 and pb["step_fused"][1] == 0
    for k in a:
        assert close_fast(orc.interior(a[k]), orc.interior(b[k])), (k, rel_err(orc.interior(a[k]), orc.interior(b[k])))


def test_odd_extent_takes_the_two_kernel_step():
    """the TMA boxes need 16-byte aligned rows (even Nz): odd extents keep the two-kernel step, in both arithmetic modes"""
    for nlocal, math in (((16, 16, 15), lb.MATH_FAST), ((16, 16, 15), lb.MATH_STRICT)):
        orc, st0, st = oracle_run(nlocal, 3, lb.RELAX_M10, 5)
        got, prof = gpu_run(nlocal, st0, 3, lb.RELAX_M10, 5, math=math)
        assert prof["step_fused"][1] == 0 and prof["collide"][1] == 5, prof
        for k in got:
            if math == lb.MATH_STRICT:
                assert np.array_equal(orc.interior(got[k]), orc.interior(st[k])), k
            else:
                assert close_fast(orc.interior(got[k]), orc.interior(st[k])), k


def test_single_operators_after_fused_steps():
    """lb_halo / lb_propagation / the single operators after fused steps see the state the reference would hold:
    fused steps, then one more step through the individual entry points, against the oracle's n + 1 steps"""
    nlocal = (24, 20, 32)
    orc, st0, st = oracle_run(nlocal, 3, lb.RELAX_M10, 7)
    with lb.Lb200(nlocal, nhalo=2, have_phi=True, math=lb.MATH_FAST) as sim:
        sim.put(lb.F, st0["f"]); sim.put(lb.PHI, st0["phi"])
        cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA, force=FG)
        sp = lb.SymmParam.make(adv_order=3, gradmu=GM, **BINARY)
        sim.step(cp, sp, 6)
        sim.step_api(cp, sp, 1)        # the reference driver's order (src/ludwig.c:528-860) through the single entry points
        got = {k: sim.get(a) for k, a in FIELDS}
    for k in got:
        assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))
