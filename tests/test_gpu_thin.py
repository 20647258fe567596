"""Lattices thinner than the halo (nlocal[d] < nhalo): the reference's 2-d runs, e.g. tests/regression/d3q19/pmpi08-le2d-fd1
(64 x 64 x 1 with nhalo 2).  The reference packs every send buffer before it unpacks any (src/field.c:1412-1531,
src/lb_data.c:1317-1477), so the outer halo layer receives the halo's content from BEFORE the swap, not the periodic image.
The oracle reproduces that bit for bit against the compiled reference (tests/test_oracle_vs_reference.py::test_*thin*,
tests/test_le_oracle.py::test_le2d_thin_lattice_long_run_vs_reference); here the CUDA library against the oracle."""
import numpy as np
import pytest

import ludwig_b200 as lb
from common import BINARY, ETA, close_fast, rel_err
from ludwig_b200.initial import equilibrium_f
from oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 1, 0)])
@pytest.mark.parametrize("nlocal", [(8, 8, 1), (1, 6, 5), (6, 1, 1), (1, 1, 1)])
def test_field_halo_thin_bit_exact(nlocal, periodic):
    orc = Oracle(nlocal, nhalo=2, periodic=periodic)
    rng = np.random.default_rng(31)
    phi = rng.random((1, orc.nsites))
    u = rng.random((3, orc.nsites))
    rphi, ru = phi.copy(), u.copy()
    orc.field_halo(rphi); orc.field_halo(ru)
    with lb.Lb200(nlocal, nhalo=2, periodic=periodic, have_phi=True) as sim:
        sim.put(lb.PHI, phi); sim.put(lb.U, u)
        sim.phi_halo(); sim.hydro_u_halo()
        assert np.array_equal(sim.get(lb.PHI), rphi)
        assert np.array_equal(sim.get(lb.U), ru)
        # a second swap sees the halo the first one left behind
        orc.field_halo(rphi)
        sim.phi_halo()
        assert np.array_equal(sim.get(lb.PHI), rphi)


@pytest.mark.parametrize("path", ["step", "api"])
@pytest.mark.parametrize("math", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
@pytest.mark.parametrize("nlocal,order", [((8, 8, 1), 3), ((8, 1, 6), 2), ((16, 12, 1), 1)])
def test_binary_steps_on_thin_lattices(nlocal, order, math, path):
    """whole binary-fluid steps on 2-d lattices (the gradient's z neighbours are the site itself, wz = 0 drops the z fluxes)"""
    nsteps = 6
    orc = Oracle(nlocal, nhalo=2)
    rng = np.random.default_rng(5)
    f = np.zeros((19, orc.nsites)); orc.interior(f)[...] = orc.interior(orc.equilibrium(1.0, (0.0, 0.0, 0.0)))
    phi = np.zeros((1, orc.nsites)); orc.interior(phi)[...] = 0.05*(rng.random((1,) + nlocal) - 0.5)
    st = dict(f=f.copy(), phi=phi.copy(), u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
              force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
    orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=order, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    with lb.Lb200(nlocal, nhalo=2, have_phi=True, math=math) as sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA)
        sp = lb.SymmParam.make(adv_order=order, **BINARY)
        (sim.step if path == "step" else sim.step_api)(cp, sp, nsteps)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("force", lb.FORCE), ("grad", lb.GRAD), ("delsq", lb.DELSQ))}
    for k in got:
        a, b = orc.interior(got[k]), orc.interior(st[k])
        if math == lb.MATH_STRICT:
            assert np.array_equal(a, b), k
        else:
            assert close_fast(a, b), (k, rel_err(a, b))


@pytest.mark.parametrize("math", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
def test_le2d_configuration(math):
    """the reference's pmpi08-le2d-fd1 configuration: 64 x 64 x 1, nhalo 2, 2 Lees-Edwards planes with speed 0.05, advection
    order 3; 100 of its steps (the oracle equals the compiled reference bit for bit over 650: tests/test_le_oracle.py)"""
    n = (64, 64, 1)
    fe = dict(a=-0.0625, b=0.0625, kappa=0.04, mobility=0.15)
    nsteps = 100
    orc = Oracle(n, nhalo=2, le_nplanes=2, le_uy=0.05)
    rng = np.random.default_rng(7361237)
    f = np.zeros((19, orc.nsites_lb)); orc.le_init_shear_profile(1.0, 0.1, f)
    phi = np.zeros((1, orc.nsites)); orc.interior(phi)[...] = 0.05*(rng.random((1,) + n) - 0.5)
    z = lambda k: np.zeros((k, orc.nsites))
    with lb.Lb200(n, nhalo=2, have_phi=True, math=math, le_nplanes=2, le_uy=0.05) as sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, 0.1), lb.SymmParam.make(adv_order=3, **fe), nsteps)
        gf, gphi, gu = sim.get(lb.F), sim.get(lb.PHI), sim.get(lb.U)
    u = z(3)
    orc.le_step(orc.collide_param(0, 1.0, 0.1), orc.symm_param(adv_order=3, **fe), 0, nsteps, f, phi, u, z(1), z(3), z(3), z(1))
    for name, a, b in (("f", gf, f), ("phi", gphi, phi), ("u", gu, u)):
        a, b = orc.interior(a), orc.interior(b)
        if math == lb.MATH_STRICT:
            assert np.array_equal(a, b), name
        else:
            # 100 steps of a sheared, phase-separating fluid amplify rounding differences: 1e-10 here (1e-12 holds for the
            # first 20 steps, tests/test_gpu_le.py)
            assert rel_err(a, b) <= 1e-10, (name, rel_err(a, b))
