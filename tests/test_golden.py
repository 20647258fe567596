"""Golden vectors produced by the unmodified reference (tests/golden/make_golden.py) vs
(a) the oracle -- runs everywhere, keeps the oracle pinned where oracle/_ref cannot be built -- and
(b) the CUDA path through the C-ABI (gpu marker): strict mode bit-exact, fast mode within tolerance."""
import numpy as np
import pytest

import golden_util
from common import close_fast
from oracle import Oracle

CASES = golden_util.case_names()


def run_oracle(g):
    orc = Oracle(g["nlocal"], nhalo=g["nhalo"], nvel=g["nvel"])
    z = lambda k: np.zeros((k, orc.nsites))
    st = dict(f=g["f0"].copy(), u=z(3), rho=z(1), force=z(3))
    if g["kind"] == "symmlb":
        st = dict(f=g["f0"].copy(), phi=z(1), u=z(3), grad=z(3), delsq=z(1))
        cp = orc.collide_param(g["nrelax"], 1.0, g["eta"], force=tuple(g["fbody"]))
        sp = orc.symm_param(g["a"], g["b"], g["kappa"], g["mobility"])
        orc.step_lb2(cp, sp, g["nsteps"], st["f"], st["phi"], st["u"], z(3), st["grad"], st["delsq"],
                     halo_reduced=g["reduced"])
        return orc, st
    if g["kind"] == "binary":
        st.update(phi=g["phi0"].copy(), grad=z(3), delsq=z(1))
        cp = orc.collide_param(g["nrelax"], 1.0, g["eta"], force=tuple(g["fbody"]))
        sp = orc.symm_param(g["a"], g["b"], g["kappa"], g["mobility"], gradmu=tuple(g["gradmu"]), adv_order=g["adv_order"],
                            conserve=int(g.get("conserve", 0)))
        orc.step(cp, sp, 1, g["nsteps"], st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    else:
        cp = orc.collide_param(g["nrelax"], 1.0, g["eta"], eta_bulk=g["eta_bulk"], force=tuple(g["fbody"]))
        orc.step(cp, None, 0, g["nsteps"], st["f"], None, st["u"], st["rho"], st["force"], None, None,
                 halo_reduced=g["reduced"])
        st.pop("force")
    return orc, st


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(name):
    g = golden_util.load(name)
    orc, st = run_oracle(g)
    for k, a in st.items():
        assert np.array_equal(orc.interior(a), orc.interior(g[k])), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_reference_golden(name, strict):
    import ludwig_b200 as lb
    g = golden_util.load(name)
    binary = g["kind"] == "binary"
    orc = Oracle(g["nlocal"], nhalo=g["nhalo"], nvel=g["nvel"])
    if g["kind"] == "symmlb":
        with lb.Lb200(g["nlocal"], nhalo=g["nhalo"], nvel=g["nvel"], ndist=2, have_phi=True,
                      halo_scheme=lb.HALO_REDUCED if g["reduced"] else lb.HALO_FULL,
                      math=lb.MATH_STRICT if strict else lb.MATH_FAST) as sim:
            sim.put(lb.F, g["f0"])
            sim.step(lb.CollideParam.make(g["nrelax"], 1.0, g["eta"], force=tuple(g["fbody"])),
                     lb.SymmParam.make(g["a"], g["b"], g["kappa"], g["mobility"]), g["nsteps"])
            for k, arr in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("grad", lb.GRAD), ("delsq", lb.DELSQ)):
                got, ref = orc.interior(sim.get(arr)), orc.interior(g[k])
                if strict:
                    assert np.array_equal(got, ref), (name, k)
                else:
                    assert close_fast(got, ref), (name, k)
        return
    with lb.Lb200(g["nlocal"], nhalo=g["nhalo"], nvel=g["nvel"], have_phi=binary,
                  halo_scheme=lb.HALO_REDUCED if g.get("reduced", 0) else lb.HALO_FULL,
                  math=lb.MATH_STRICT if strict else lb.MATH_FAST) as sim:
        sim.put(lb.F, g["f0"])
        cp = lb.CollideParam.make(g["nrelax"], 1.0, g["eta"], eta_bulk=g.get("eta_bulk"), force=tuple(g["fbody"]))
        sp = None
        if binary:
            sim.put(lb.PHI, g["phi0"])
            sp = lb.SymmParam.make(g["a"], g["b"], g["kappa"], g["mobility"], gradmu=tuple(g["gradmu"]),
                                   adv_order=g["adv_order"], conserve=int(g.get("conserve", 0)))
        sim.step(cp, sp, g["nsteps"])
        keys = [("f", lb.F), ("u", lb.U), ("rho", lb.RHO)]
        if binary:
            keys += [("phi", lb.PHI), ("force", lb.FORCE), ("grad", lb.GRAD), ("delsq", lb.DELSQ)]
        for k, arr in keys:
            got, ref = orc.interior(sim.get(arr)), orc.interior(g[k])
            if strict:
                assert np.array_equal(got, ref), (name, k)
            else:
                assert close_fast(got, ref), (name, k)
