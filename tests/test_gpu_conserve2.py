"""cahn_hilliard_options_conserve 2 (PHI_CONSERVE_GLOBAL_SUBTRACT, src/phi_cahn_hilliard.c:1102-1169): after every forward
step (sum_fluid phi - sum0)/nfluid is subtracted everywhere; sum0 is the statistics code's initial sum
(cahn_hilliard_stats_time0).  The oracle equals the compiled reference bit for bit with one thread
(tests/test_oracle_vs_reference.py::test_conserve_2_global_subtraction_vs_reference); the reference's own result depends on its
thread count through the order of the plain sum, the library's order is fixed, so the comparison is to 1e-14 relative in
strict mode (a few ulp of the sum, divided by the number of sites) and to the fast-mode tolerance otherwise."""
import math

import numpy as np
import pytest

import ludwig_b200 as lb
from common import BINARY, ETA, close_fast, rel_err
from oracle import Oracle

pytestmark = pytest.mark.gpu


def start(orc, nlocal, seed=3):
    rng = np.random.default_rng(seed)
    f = np.zeros((19, orc.nsites)); orc.interior(f)[...] = orc.interior(orc.equilibrium(1.0, (0.0, 0.0, 0.0)))
    phi = np.zeros((1, orc.nsites)); orc.interior(phi)[...] = 0.05*(rng.random((1,) + nlocal) - 0.5) + 0.01
    return f, phi


@pytest.mark.parametrize("nlocal", [(8, 6, 10), (32, 30, 34)])
def test_initial_sum(nlocal):
    orc = Oracle(nlocal, nhalo=2)
    f, phi = start(orc, nlocal)
    exact = math.fsum(orc.interior(phi).ravel().tolist())
    with lb.Lb200(nlocal, nhalo=2, have_phi=True) as sim:
        sim.put(lb.PHI, phi)
        s = sim.phi_conserve_sum()
    assert abs(s - exact) <= 2*np.spacing(abs(exact))               # compensated: the correctly rounded sum or its neighbour
    assert abs(orc.phi_sum_time0(phi) - exact) <= 2*np.spacing(abs(exact))


@pytest.mark.parametrize("path", ["step", "api"])
@pytest.mark.parametrize("math_mode", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
@pytest.mark.parametrize("offset", [0.0, 1.0e-3])
def test_steps_with_global_subtraction(offset, math_mode, path):
    nlocal, nsteps = (16, 12, 20), 6
    orc = Oracle(nlocal, nhalo=2)
    f, phi = start(orc, nlocal)
    sum0 = orc.phi_sum_time0(phi) + offset
    st = dict(f=f.copy(), phi=phi.copy(), u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
              force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
    orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=3, conserve=2, phi_init_sum=sum0, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    with lb.Lb200(nlocal, nhalo=2, have_phi=True, math=math_mode) as sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA)
        sp = lb.SymmParam.make(adv_order=3, conserve=2, **BINARY)
        with pytest.raises(lb.Lb200Error):
            sim.step(cp, sp, 1)                                     # no initial sum yet
        if offset == 0.0:
            got0 = sim.phi_conserve_sum()                           # the library's own time-0 sum ...
            assert abs(got0 - sum0) <= 2*np.spacing(abs(sum0))
        sim.phi_init_sum_set(sum0)                                  # ... or the caller's (phi->field_init_sum)
        (sim.step if path == "step" else sim.step_api)(cp, sp, nsteps)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U))}
    for k in got:
        a, b = orc.interior(got[k]), orc.interior(st[k])
        if math_mode == lb.MATH_STRICT:
            assert np.abs(a - b).max() <= 1e-14*np.abs(b).max(), (k, rel_err(a, b))
        else:
            assert close_fast(a, b), (k, rel_err(a, b))
    if offset:
        assert abs(orc.interior(got["phi"]).sum() - sum0) < 1e-11
