"""Liquid crystal (Landau-de Gennes Q tensor + Beris-Edwards) on the GPU, through the C-ABI,
against the CPU oracle (oracle/lb_oracle_lc.c, pinned bit-for-bit to the compiled reference and to
pmpi08-chol-s01.log / serial-chol-fld.log in tests/test_lc_oracle.py) on identical inputs.

Bar: LB200_MATH_STRICT bit-exact for every operator and for whole time steps through every path;
LB200_MATH_FAST (FMA contraction) within 1e-12 relative (absolute floor 1e-14) after N steps."""
import numpy as np
import pytest

import ludwig_b200 as lb
from common import close_fast
from ludwig_b200.initial import equilibrium_f, lc_twist_q, lc_nematic_q
from oracle import Oracle

pytestmark = pytest.mark.gpu

CHOL = dict(a0=0.01, q0=0.19635, gamma=3.0, kappa0=0.000648456, kappa1=0.000648456, xi=0.7, Gamma=0.5)
FLD = dict(a0=0.084334998544, q0=0.05, gamma=3.085714285714, kappa0=0.01, kappa1=0.013, xi=0.7, Gamma=0.3,
           epsilon=41.4 * (1.0 / (12.0 * np.pi)), e0=(0.01, 0.0, 0.003))
ACT = dict(CHOL, zeta0=1.0 / 3.0, zeta1=0.005)       # lc_activity yes (serial-actv-s01.inp's constants)
RSH = dict(FLD, redshift=0.93)                       # lc_init_redshift != 1 (static)
ETA = 0.1


def state(orc, lc, seed=17, axis=2):
    rng = np.random.default_rng(seed)
    n = orc.nlocal
    q = lc_twist_q(n, orc.nhalo, lc["q0"], 1.0 / 3.0, axis)
    orc.interior(q)[...] += 0.02 * (rng.random(orc.interior(q).shape) - 0.5)
    u = np.zeros((3, orc.nsites))
    orc.interior(u)[...] = 0.02 * (rng.random((3,) + tuple(n)) - 0.5)
    f = equilibrium_f(n, orc.nhalo)
    return f, q, u


@pytest.mark.parametrize("n", [(8, 6, 10), (12, 12, 40)])
@pytest.mark.parametrize("lc", [CHOL, FLD, ACT, RSH], ids=["chol", "field", "active", "redshift"])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_lc_operators_strict_bit_exact(n, lc, order):
    orc = Oracle(n, nhalo=2)
    p = orc.lc_param(**lc)
    pg = lb.LcParam.make(adv_order=order, **lc)
    f, q, u = state(orc, lc)
    with lb.Lb200(n, nhalo=2, have_q=True, math=lb.MATH_STRICT) as sim:
        sim.put(lb.Q, q); sim.put(lb.U, u); sim.put(lb.F, f)
        # field_halo(q) + field_grad_compute
        sim.q_halo(); sim.q_grad_compute()
        qgrad = np.zeros((15, orc.nsites)); qdelsq = np.zeros((5, orc.nsites))
        orc.field_halo(q); orc.grad_7pt(q, qgrad, qdelsq)
        assert np.array_equal(sim.get(lb.Q), q)
        assert np.array_equal(orc.region(sim.get(lb.QGRAD), 1), orc.region(qgrad, 1))
        assert np.array_equal(orc.region(sim.get(lb.QDELSQ), 1), orc.region(qdelsq, 1))
        # pth_stress_compute (fe_lc_stress_v) and the force
        sim.lc_stress_compute(pg)
        s = np.zeros((9, orc.nsites))
        orc.lc_stress(p, q, qgrad, qdelsq, s)
        assert np.array_equal(orc.region(sim.get(lb.STR), 1), orc.region(s, 1))
        sim.hydro_f_zero(); sim.pth_force_fluid_driver()
        force = np.zeros((3, orc.nsites))
        orc.force_divergence(s, force)
        assert np.array_equal(orc.interior(sim.get(lb.FORCE)), orc.interior(force))
        # hydro_u_halo + beris_edw_update
        sim.hydro_u_halo(); sim.beris_edw_update(pg)
        orc.field_halo(u)
        h = np.zeros((5, orc.nsites)); flux = np.zeros((20, orc.nsites))
        orc.lc_mol_field(p, q, qgrad, qdelsq, h)
        orc.advection_nf(order, u, q, flux)
        orc.beris_edw_update(p, u, h, flux, q)
        assert np.array_equal(orc.interior(sim.get(lb.Q)), orc.interior(q))


def _run(n, lc, order, math, nsteps, path):
    orc = Oracle(n, nhalo=2)
    p = orc.lc_param(**lc)
    pg = lb.LcParam.make(adv_order=order, **lc)
    f, q, _ = state(orc, lc, seed=3, axis=0)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, qgrad, qdelsq = z(3), z(1), z(3), z(15), z(5)
    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA)
    with lb.Lb200(n, nhalo=2, have_q=True, math=math) as sim:
        sim.put(lb.F, f); sim.put(lb.Q, q)
        if path == "api":
            sim.step_lc_api(cp, pg, nsteps)
        elif path == "mixed":
            sim.step_lc(cp, pg, 2); sim.step_lc_api(cp, pg, 2); sim.step_lc(cp, pg, nsteps - 4)
        else:
            sim.set_knob(lb.KNOB_WRAP, 1 if path == "wrap" else 0)
            sim.step_lc(cp, pg, nsteps)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("q", lb.Q), ("u", lb.U), ("rho", lb.RHO), ("force", lb.FORCE))}
    orc.lc_step(orc.collide_param(0, 1.0, ETA), p, order, nsteps, f, q, u, rho, force, qgrad, qdelsq)
    return orc, got, dict(f=f, q=q, u=u, rho=rho, force=force)


@pytest.mark.parametrize("path", ["wrap", "halo", "api", "mixed"])
@pytest.mark.parametrize("n,lc,order", [((12, 10, 8), CHOL, 3), ((8, 8, 40), FLD, 1), ((10, 12, 6), FLD, 2), ((8, 10, 12), CHOL, 4),
                                        ((10, 8, 12), ACT, 3), ((8, 12, 10), RSH, 3)])
def test_lc_steps_strict_bit_exact(n, lc, order, path):
    orc, got, want = _run(n, lc, order, lb.MATH_STRICT, 10, path)
    for k in want:
        assert np.array_equal(orc.interior(got[k]), orc.interior(want[k])), k
    assert np.abs(orc.interior(want["u"])).max() > 1e-8


@pytest.mark.parametrize("path", ["wrap", "halo"])
@pytest.mark.parametrize("n,lc,order", [((12, 10, 8), CHOL, 3), ((8, 8, 40), FLD, 1), ((32, 32, 32), CHOL, 2), ((8, 10, 12), CHOL, 4),
                                        ((16, 16, 16), ACT, 3), ((16, 8, 16), RSH, 3)])
def test_lc_steps_fast_tolerance(n, lc, order, path):
    orc, got, want = _run(n, lc, order, lb.MATH_FAST, 20, path)
    for k in want:
        assert close_fast(orc.interior(got[k]), orc.interior(want[k])), (k, np.abs(orc.interior(got[k]) - orc.interior(want[k])).max())


def test_pmpi08_chol_s01_log_on_gpu():
    """the reference's regression answer tests/regression/d3q19/pmpi08-chol-s01.log (cholesteric twist, advection
    order 3, 10 steps; z-dependent state, so a 4 x 4 x 128 column reproduces the 128^3 statistics) from the CUDA path"""
    n = (4, 4, 128)
    orc = Oracle(n, nhalo=2)
    q = lc_twist_q(n, 2, CHOL["q0"], 0.333333333333333, 2)
    f = equilibrium_f(n, 2)
    approx = lambda v, d: pytest.approx(v, rel=0.5 * 10.0 ** (1 - d), abs=1e-30)
    with lb.Lb200(n, nhalo=2, have_q=True, math=lb.MATH_FAST) as sim:
        sim.put(lb.F, f); sim.put(lb.Q, q)
        sim.step_lc(lb.CollideParam.make(lb.RELAX_M10, 1.0, 1.0), lb.LcParam.make(adv_order=3, **CHOL), 10)
        qi = orc.interior(sim.get(lb.Q))
        uz = orc.interior(sim.get(lb.U))[2]
    v = qi[0].ravel()
    assert v.mean() == approx(8.3293376e-02, 8) and (v * v).mean() - v.mean() ** 2 == approx(3.1248723e-02, 8)
    assert v.min() == approx(-1.6670182e-01, 8) and v.max() == approx(3.3328742e-01, 8)
    assert qi[1].min() == approx(-2.4999462e-01, 8) and qi[1].max() == approx(2.4999462e-01, 8)
    assert qi[3].mean() == approx(8.3292223e-02, 8)
    assert uz.min() == approx(-4.2438997e-10, 6) and uz.max() == approx(4.2438990e-10, 6)


@pytest.mark.parametrize("math_mode", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
@pytest.mark.parametrize("path", ["step", "api"])
def test_lc_active_nematic_2d(path, math_mode):
    """an active nematic on a 2-d lattice (one plane in z, thinner than the halo) with fd_gradient_calculation 2d_5pt_fluid --
    the configuration of tests/regression/d3q19-short/serial-actv-s01.inp; the oracle is pinned to the compiled reference in
    tests/test_lc_oracle.py::test_lc_active_2d_steps_vs_reference"""
    lc = dict(a0=1.0, q0=0.0, gamma=3.0, kappa0=0.04, kappa1=0.04, xi=0.7, Gamma=0.3375, zeta0=1.0 / 3.0, zeta1=0.005)
    n, order, nsteps, eta = (32, 24, 1), 1, 12, 1.3333
    orc = Oracle(n, nhalo=2)
    p = orc.lc_param(grad_2d5=1, **lc)
    pg = lb.LcParam.make(adv_order=order, **lc)
    f, q, _ = state(orc, dict(lc, q0=0.19635), seed=5, axis=0)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, qgrad, qdelsq = z(3), z(1), z(3), z(15), z(5)
    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, eta)
    with lb.Lb200(n, nhalo=2, have_q=True, math=math_mode) as sim:
        sim.set_knob(lb.KNOB_QGRAD_2D5, 1)
        sim.put(lb.F, f); sim.put(lb.Q, q)
        (sim.step_lc if path == "step" else sim.step_lc_api)(cp, pg, nsteps)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("q", lb.Q), ("u", lb.U), ("force", lb.FORCE))}
    orc.lc_step(orc.collide_param(0, 1.0, eta), p, order, nsteps, f, q, u, rho, force, qgrad, qdelsq)
    want = dict(f=f, q=q, u=u, force=force)
    assert np.abs(orc.interior(u)).max() > 1e-8
    for k in want:
        a, b = orc.interior(got[k]), orc.interior(want[k])
        if math_mode == lb.MATH_STRICT:
            assert np.array_equal(a, b), k
        else:
            assert close_fast(a, b), (k, np.abs(a - b).max())
