"""Parity of the code path the bench actually times: the one-sweep phi sector in its STEADY-STATE loop
(several (y,z) tiles per plane AND several x chunks long enough for the phase-unrolled march), followed by the
pull-stream collision, through lb200_step -- against the CPU oracle on identical inputs, and against the
statistics printed in the reference's own regression log of this configuration.

The x-chunk length is picked at launch from the SM count; LB200_PS_XC forces it, so that the chunk boundaries,
the pipeline fill / drain steps and the unrolled steady-state steps are all exercised at test sizes.

Bar: LB200_MATH_STRICT bit-exact; LB200_MATH_FAST within 1e-12 relative (absolute floor 1e-14 for velocities)
after N in {6, 20, 100} steps."""
import numpy as np
import pytest

import ludwig_b200 as lb
from common import BINARY, ETA, close_fast, rel_err, seeded_state
from ludwig_b200.initial import equilibrium_f, spinodal_phi
from oracle import Oracle, fed_density, stats_scalar

pytestmark = pytest.mark.gpu

FIELDS = (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("rho", lb.RHO), ("force", lb.FORCE), ("grad", lb.GRAD),
          ("delsq", lb.DELSQ))


def run_case(nlocal, order, math, nsteps, nrelax=lb.RELAX_M10, calls=1):
    orc = Oracle(nlocal, nhalo=2)
    st = seeded_state(orc)
    fg = (1e-6, -2e-6, 5e-7)
    gm = (1e-5, -2e-5, 3e-5)
    cpo = orc.collide_param(nrelax, 1.0, ETA, force=fg)
    spo = orc.symm_param(adv_order=order, gradmu=gm, **BINARY)
    with lb.Lb200(nlocal, nhalo=2, have_phi=True, math=math) as sim:
        sim.put(lb.F, st["f"]); sim.put(lb.PHI, st["phi"])
        cp = lb.CollideParam.make(nrelax, 1.0, ETA, force=fg)
        sp = lb.SymmParam.make(adv_order=order, gradmu=gm, **BINARY)
        for c in range(calls):
            sim.step(cp, sp, nsteps // calls)
        got = {k: sim.get(a) for k, a in FIELDS}
    orc.step(cpo, spo, 1, nsteps, st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    return orc, st, got


# 64 x 48 x 70: 4 x 3 tiles of 14 x 30 columns; chunks of 11 / 17 / 43 planes = 6 / 4 / 2 chunks, each with at
# least one pass of the phase-unrolled steady-state loop (needs >= 11 planes)
@pytest.mark.parametrize("xc", [11, 17, 43])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_steady_state_phi_sector_strict_bit_exact(order, xc, monkeypatch):
    monkeypatch.setenv("LB200_PS_XC", str(xc))
    orc, st, got = run_case((64, 48, 70), order, lb.MATH_STRICT, 6, calls=2)
    for k in got:
        assert np.array_equal(orc.interior(got[k]), orc.interior(st[k])), k


@pytest.mark.parametrize("xc", [11, 17, 43])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_steady_state_phi_sector_fast_tolerance(order, xc, monkeypatch):
    monkeypatch.setenv("LB200_PS_XC", str(xc))
    orc, st, got = run_case((64, 48, 70), order, lb.MATH_FAST, 20, nrelax=lb.RELAX_TRT if order == 2 else lb.RELAX_M10)
    for k in got:
        assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))


@pytest.mark.parametrize("nlocal", [(64, 48, 70), (48, 62, 64)])
def test_default_chunking_fast_tolerance(nlocal):
    """no LB200_PS_XC: the launch-time choice for this GPU (64 planes: one or two chunks with the steady-state loop)"""
    orc, st, got = run_case(nlocal, 3, lb.MATH_FAST, 10)
    for k in got:
        assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))


def test_hundred_steps_fast_tolerance(monkeypatch):
    """SURVEY 8(c) protocol, N = 100: the FMA / re-associated build stays within 1e-12 of the oracle on every field"""
    monkeypatch.setenv("LB200_PS_XC", "16")
    orc, st, got = run_case((32, 30, 62), 3, lb.MATH_FAST, 100, calls=4)
    for k in got:
        assert close_fast(orc.interior(got[k]), orc.interior(st[k])), (k, rel_err(orc.interior(got[k]), orc.interior(st[k])))


def test_hundred_steps_strict_bit_exact():
    orc, st, got = run_case((32, 30, 62), 3, lb.MATH_STRICT, 100, calls=4)
    for k in got:
        assert np.array_equal(orc.interior(got[k]), orc.interior(st[k])), k


# ---------------------------------------------------------------------------------------------------------------
# the reference's regression log of the headline configuration, reproduced by the CUDA path
# ---------------------------------------------------------------------------------------------------------------

def approx(v, digits):
    return pytest.approx(v, rel=0.5 * 10.0 ** (1 - digits), abs=1e-30)


@pytest.mark.parametrize("path", ["api", "fused"])
@pytest.mark.parametrize("math", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
def test_serial_spin_fd1_log_from_the_gpu(path, math):
    """tests/regression/d3q19-short/serial-spin-fd1.{inp,log} of the reference (64^3 spinodal binary fluid, A = -B =
    -0.00625, K = 0.004, M = 1.25, eta = 0.00625, 27pt gradient, advection order 1, seed 8361235): the statistics the
    reference prints at t = 10 (log lines 95-107), to the printed digits, from arrays computed by the CUDA library."""
    n = (64, 64, 64)
    orc = Oracle(n, nhalo=2)                      # geometry + the statistics routines only
    f = equilibrium_f(n, 2)
    phi = spinodal_phi(n, 2, 8361235, 0.0, 0.1)
    spo = orc.symm_param(adv_order=1, **BINARY)
    with lb.Lb200(n, nhalo=2, have_phi=True, math=math) as sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA)
        sp = lb.SymmParam.make(adv_order=1, **BINARY)
        if path == "api":
            sim.step_api(cp, sp, 10)
        else:
            sim.step(cp, sp, 10)
        # the reference prints [rho] and the momentum from the distributions after lb_propagation: the last operation
        # of step_api; pending after lb200_step and applied by the copy to the host
        gf, gphi, gu, ggrad = sim.get(lb.F), sim.get(lb.PHI), sim.get(lb.U), sim.get(lb.GRAD)

    s = stats_scalar(orc, gphi)
    assert s[0] == approx(3.1484764e+00, 8) and s[1] == approx(1.2010484e-05, 8)
    assert s[2] == approx(3.7820523e-04, 8)
    assert s[3] == approx(-4.7270149e-02, 8) and s[4] == approx(4.6821679e-02, 8)
    assert fed_density(orc, spo, gphi, ggrad) == approx(-9.7510518349e-07, 11)
    r = stats_scalar(orc, gf.sum(axis=0, keepdims=True))
    assert r[0] == approx(262144.00, 8)
    assert r[3] == approx(0.99998006808, 11) and r[4] == approx(1.00001625877, 11)
    ui = orc.interior(gu)
    for a, (lo, hi) in enumerate(((-1.3145696e-05, 1.2773457e-05), (-1.3301763e-05, 1.3768024e-05),
                                  (-1.2618505e-05, 1.2966490e-05))):
        assert ui[a].min() == approx(lo, 8) and ui[a].max() == approx(hi, 8)
    fi = orc.interior(gf)
    mom = [(fi * orc.cv[:, a, None, None, None]).sum() for a in range(3)]
    assert np.allclose(mom, 0.0, atol=1e-10)
