"""Golden vectors of the Lees-Edwards and liquid-crystal steps produced by the unmodified reference
(tests/golden/make_golden_le_lc.py) vs (a) the oracle -- runs everywhere, keeps the oracle pinned where oracle/_ref
cannot be built -- and (b) the CUDA path through the C-ABI (gpu marker): strict bit-exact, fast within tolerance."""
import numpy as np
import pytest

import golden_util
from common import close_fast
from oracle import Oracle


def _lc_params(g):
    return dict(a0=g["lc_a0"], q0=g["lc_q0"], gamma=g["lc_gamma"], kappa0=g["lc_kappa0"], kappa1=g["lc_kappa1"],
                xi=g["lc_xi"], Gamma=g["lc_Gamma"], epsilon=g["lc_epsilon"], e0=tuple(g["lc_e0"]))


@pytest.mark.parametrize("name", golden_util.case_names("le_"))
def test_oracle_reproduces_le_golden(name):
    g = golden_util.load(name)
    orc = Oracle(g["nlocal"], nhalo=2, le_nplanes=g["nplanes"], le_uy=g["uy"])
    z = lambda k: np.zeros((k, orc.nsites))
    f, phi = g["f0"].copy(), g["phi0"].copy()
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    orc.le_step(orc.collide_param(0, 1.0, g["eta"]), orc.symm_param(g["a"], g["b"], g["kappa"], g["mobility"], adv_order=g["adv_order"]),
                0, g["nsteps"], f, phi, u, rho, force, grad, delsq)
    for k, a in (("f", f), ("phi", phi), ("u", u), ("rho", rho), ("force", force)):
        assert np.array_equal(orc.interior(a), g[k]), (name, k)


@pytest.mark.parametrize("name", golden_util.case_names("lc_"))
def test_oracle_reproduces_lc_golden(name):
    g = golden_util.load(name)
    orc = Oracle(g["nlocal"], nhalo=2)
    z = lambda k: np.zeros((k, orc.nsites))
    f, q = g["f0"].copy(), g["q0"].copy()
    u, rho, force, qgrad, qdelsq = z(3), z(1), z(3), z(15), z(5)
    orc.lc_step(orc.collide_param(0, 1.0, g["eta"]), orc.lc_param(**_lc_params(g)), g["adv_order"], g["nsteps"],
                f, q, u, rho, force, qgrad, qdelsq)
    for k, a in (("f", f), ("q", q), ("u", u), ("rho", rho), ("force", force)):
        assert np.array_equal(orc.interior(a), g[k]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", golden_util.case_names("le_"))
def test_cuda_reproduces_le_golden(name, strict):
    import ludwig_b200 as lb
    g = golden_util.load(name)
    with lb.Lb200(g["nlocal"], nhalo=2, have_phi=True, math=lb.MATH_STRICT if strict else lb.MATH_FAST,
                  le_nplanes=g["nplanes"], le_uy=g["uy"]) as sim:
        sim.put(lb.F, g["f0"]); sim.put(lb.PHI, g["phi0"])
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, g["eta"]),
                 lb.SymmParam.make(g["a"], g["b"], g["kappa"], g["mobility"], adv_order=g["adv_order"]), g["nsteps"])
        for k, arr in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("rho", lb.RHO), ("force", lb.FORCE)):
            got = sim.interior(sim.get(arr))
            assert (np.array_equal(got, g[k]) if strict else close_fast(got, g[k])), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", golden_util.case_names("lc_"))
def test_cuda_reproduces_lc_golden(name, strict):
    import ludwig_b200 as lb
    g = golden_util.load(name)
    with lb.Lb200(g["nlocal"], nhalo=2, have_q=True, math=lb.MATH_STRICT if strict else lb.MATH_FAST) as sim:
        sim.put(lb.F, g["f0"]); sim.put(lb.Q, g["q0"])
        sim.step_lc(lb.CollideParam.make(lb.RELAX_M10, 1.0, g["eta"]), lb.LcParam.make(adv_order=g["adv_order"], **_lc_params(g)),
                    g["nsteps"])
        for k, arr in (("f", lb.F), ("q", lb.Q), ("u", lb.U), ("rho", lb.RHO), ("force", lb.FORCE)):
            got = sim.interior(sim.get(arr))
            assert (np.array_equal(got, g[k]) if strict else close_fast(got, g[k])), (name, k)
