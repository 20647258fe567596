"""Pins the oracle: oracle/lb_oracle.c (the CPU restatement the GPU tests check against) must agree
BIT FOR BIT with the UNMODIFIED reference (ludwig-cf/ludwig v0.23.0) compiled from /root/reference by
oracle/Makefile.ref into oracle/_ref/ (strict build: -O2 -ffp-contract=off, asserts on), driven through
the reference's own entry points by oracle/ref_harness.c.

Skipped when oracle/_ref has not been built (it needs /root/reference, i.e. this container); the
committed golden vectors in tests/golden/ (test_golden.py) carry the same pin everywhere else."""
import numpy as np
import pytest

import refharness as rh
from common import BINARY, ETA
from ludwig_b200.initial import equilibrium_f, spinodal_phi
from oracle import Oracle

pytestmark = pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built (needs /root/reference)")


def fill_random(sim, orc, what, rng, scale=1.0, shift=0.0):
    a = scale * (rng.random((sim.get(what).shape[0], orc.nsites)) + shift)
    sim.set(what, a)
    return a


@pytest.mark.parametrize("nvel", [19, 15, 27])
@pytest.mark.parametrize("nlocal,nhalo", [((6, 5, 7), 1), ((4, 6, 5), 2)])
def test_propagation(nvel, nlocal, nhalo):
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    orc = Oracle(nlocal, nhalo=nhalo, nvel=nvel)
    rng = np.random.default_rng(1)
    with rh.RefSim(nlocal, nhalo=nhalo, nvel=nvel) as s:
        f = fill_random(s, orc, rh.REF_F, rng)
        s.op("propagation")
        ref = s.get(rh.REF_F)
    fp = np.zeros_like(f)                 # reference fprime starts calloc'ed
    orc.propagation(f, fp)
    assert np.array_equal(fp, ref)


@pytest.mark.parametrize("reduced", [0, 1])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 0, 1), (0, 0, 0)])
@pytest.mark.parametrize("nvel,nlocal,nhalo", [(19, (6, 5, 7), 1), (19, (4, 6, 5), 2), (15, (4, 4, 4), 1), (27, (3, 4, 5), 1)])
def test_lb_halo(nvel, nlocal, nhalo, periodic, reduced):
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    orc = Oracle(nlocal, nhalo=nhalo, periodic=periodic, nvel=nvel)
    rng = np.random.default_rng(2)
    with rh.RefSim(nlocal, nhalo=nhalo, periodic=periodic, halo_reduced=reduced, nvel=nvel) as s:
        f = fill_random(s, orc, rh.REF_F, rng)
        s.op("lb_halo")
        ref = s.get(rh.REF_F)
    orc.lb_halo(f, reduced=reduced)
    assert np.array_equal(f, ref)


@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
@pytest.mark.parametrize("nlocal,nhalo", [((6, 5, 7), 2), ((4, 4, 4), 1), ((5, 3, 8), 3)])
def test_field_halo(nlocal, nhalo, periodic):
    orc = Oracle(nlocal, nhalo=nhalo, periodic=periodic)
    rng = np.random.default_rng(3)
    with rh.RefSim(nlocal, nhalo=nhalo, periodic=periodic, have_phi=int(nhalo >= 2), **(BINARY if nhalo >= 2 else {})) as s:
        u = fill_random(s, orc, rh.REF_U, rng)
        s.op("hydro_u_halo")
        ref_u = s.get(rh.REF_U)
        if nhalo >= 2:
            phi = fill_random(s, orc, rh.REF_PHI, rng)
            s.op("phi_halo")
            ref_phi = s.get(rh.REF_PHI)
    orc.field_halo(u)
    assert np.array_equal(u, ref_u)
    if nhalo >= 2:
        orc.field_halo(phi)
        assert np.array_equal(phi, ref_phi)


@pytest.mark.parametrize("nrelax", [0, 1, 2])
@pytest.mark.parametrize("nvel", [19, 15, 27])
def test_collide(nvel, nrelax):
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    if nvel == 27 and nrelax == 2:
        pytest.skip("TRT undefined for D3Q27 in the reference")
    nlocal = (5, 4, 6)
    orc = Oracle(nlocal, nhalo=1, nvel=nvel)
    rng = np.random.default_rng(4)
    fg = (1e-6, -2e-6, 3e-6)
    with rh.RefSim(nlocal, nhalo=1, nrelax=nrelax, eta_shear=0.02, eta_bulk=0.05, fbody=fg, nvel=nvel) as s:
        f = fill_random(s, orc, rh.REF_F, rng, scale=0.1, shift=0.5)
        force = fill_random(s, orc, rh.REF_FORCE, rng, scale=1e-4, shift=-0.5)
        s.op("collide")
        ref_f, ref_u, ref_rho = s.get(rh.REF_F), s.get(rh.REF_U), s.get(rh.REF_RHO)
    u = np.zeros((3, orc.nsites)); rho = np.zeros((1, orc.nsites))
    cp = orc.collide_param(nrelax, 1.0, 0.02, eta_bulk=0.05, force=fg)
    # the reference collides every y,z of the allocation for x in [1,N] (src/kernel_3d_v.c:50-71)
    orc.collide(cp, f, force, rho, u, include_halo=1)
    assert np.array_equal(f, ref_f)
    assert np.array_equal(u, ref_u)
    assert np.array_equal(rho, ref_rho)


def test_gradient_stress_force_fluxes():
    """Every stage of the phi sector on its own, including the stored stress and the four flux arrays."""
    nlocal = (6, 5, 7)
    orc = Oracle(nlocal, nhalo=2)
    rng = np.random.default_rng(5)
    gm = (1e-4, -2e-4, 3e-4)
    for order in (1, 2, 3, 4):
        with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=order, eta_shear=ETA, gradmu=gm, **BINARY) as s:
            phi = fill_random(s, orc, rh.REF_PHI, rng, scale=0.1, shift=-0.5)
            u = fill_random(s, orc, rh.REF_U, rng, scale=0.05, shift=-0.5)
            s.op("hydro_f_zero")
            s.op("phi_halo"); s.op("grad_compute")
            ref_phi_h, ref_grad, ref_delsq = s.get(rh.REF_PHI), s.get(rh.REF_GRAD), s.get(rh.REF_DELSQ)
            s.op("phi_force")
            ref_str, ref_force = s.get(rh.REF_STR), s.get(rh.REF_FORCE)
            s.op("cahn_hilliard")
            ref_flux, ref_phi = s.get(rh.REF_FLUX), s.get(rh.REF_PHI)
        sp = orc.symm_param(gradmu=gm, adv_order=order, **BINARY)
        grad = np.zeros((3, orc.nsites)); delsq = np.zeros((1, orc.nsites))
        strs = np.zeros((9, orc.nsites)); force = np.zeros((3, orc.nsites)); flux = np.zeros((4, orc.nsites))
        orc.field_halo(phi)
        assert np.array_equal(phi, ref_phi_h)
        orc.grad_27pt(phi, grad, delsq)
        assert np.array_equal(grad, ref_grad) and np.array_equal(delsq, ref_delsq)
        orc.stress_symm(sp, phi, grad, delsq, strs)
        # the reference's stress array is malloc'ed; it is defined for x in [0,N+1], all y,z
        xs_ = slice(orc.nhalo - 1, orc.nhalo + nlocal[0] + 1)
        assert np.array_equal(strs.reshape((9,) + orc.nall)[:, xs_], ref_str.reshape((9,) + orc.nall)[:, xs_])
        orc.force_divergence(strs, force)
        assert np.array_equal(force, ref_force)
        orc.field_halo(u)
        orc.advection(order, u, phi, flux)
        orc.flux_mu(sp, phi, delsq, flux)
        orc.flux_mu_ext(sp, flux)
        # flux arrays: compare where the reference defines them (x in [1,N], y,z in [0,N])
        h = orc.nhalo
        sl = (slice(None), slice(h, h + nlocal[0]), slice(h - 1, h + nlocal[1]), slice(h - 1, h + nlocal[2]))
        assert np.array_equal(flux.reshape((4,) + orc.nall)[sl], ref_flux.reshape((4,) + orc.nall)[sl])
        orc.phi_update(flux, phi)
        assert np.array_equal(orc.interior(phi), orc.interior(ref_phi))


def test_no_flux_mask_with_solid_sites():
    nlocal = (6, 5, 7)
    orc = Oracle(nlocal, nhalo=2)
    rng = np.random.default_rng(6)
    status = (rng.random(orc.nsites) < 0.15).astype(np.int8)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=3, eta_shear=ETA, **BINARY) as s:
        s.set(rh.REF_MAP, status.astype(np.float64))
        phi = fill_random(s, orc, rh.REF_PHI, rng, scale=0.1, shift=-0.5)
        u = fill_random(s, orc, rh.REF_U, rng, scale=0.05, shift=-0.5)
        delsq = fill_random(s, orc, rh.REF_DELSQ, rng, scale=0.1, shift=-0.5)
        s.op("cahn_hilliard")
        ref_phi = s.get(rh.REF_PHI)
    sp = orc.symm_param(adv_order=3, **BINARY)
    flux = np.zeros((4, orc.nsites))
    orc.field_halo(u)
    orc.advection(3, u, phi, flux)
    orc.flux_mu(sp, phi, delsq, flux)
    orc.flux_mu_ext(sp, flux)
    orc.no_flux(status, flux)
    orc.phi_update(flux, phi)
    assert np.array_equal(orc.interior(phi), orc.interior(ref_phi))


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("nlocal", [(12, 10, 8), (16, 16, 16)])
def test_binary_time_steps(order, nlocal):
    nsteps = 8
    orc = Oracle(nlocal, nhalo=2)
    fg = (1e-6, 2e-6, 3e-6)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=order, ghost_off=1, eta_shear=ETA, fbody=fg, **BINARY) as s:
        s.init_rest(1.0)
        s.init_spinodal(8361235, 0.0, 0.05)
        f, phi = s.get(rh.REF_F), s.get(rh.REF_PHI)
        assert np.array_equal(f, equilibrium_f(nlocal, 2))
        assert np.array_equal(phi, spinodal_phi(nlocal, 2, 8361235))
        s.step(nsteps)
        ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U), ("rho", rh.REF_RHO),
                                         ("force", rh.REF_FORCE), ("grad", rh.REF_GRAD), ("delsq", rh.REF_DELSQ))}
    st = dict(f=f, phi=phi, u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
              force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
    orc.step(orc.collide_param(0, 1.0, ETA, force=fg), orc.symm_param(adv_order=order, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    for k in ref:
        assert np.array_equal(orc.interior(st[k]), orc.interior(ref[k])), k


@pytest.mark.parametrize("gradmu", [(0.0, 0.0, 0.0), (1e-5, -2e-5, 3e-5)])
@pytest.mark.parametrize("order,nlocal", [(1, (12, 10, 8)), (3, (16, 16, 16))])
def test_binary_time_steps_force_method_phi_gradmu(order, nlocal, gradmu):
    """fe_force_method phi_gradmu (src/phi_force.c:110-121): force = -phi grad mu (phi_grad_mu_fluid) - phi grad_mu_ext
    (phi_grad_mu_external), instead of the stress divergence -- the configuration of serial-muex-st1.inp"""
    nsteps = 8
    orc = Oracle(nlocal, nhalo=2)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=order, ghost_off=1, eta_shear=ETA, gradmu=gradmu, force_gradmu=1,
                   **BINARY) as s:
        s.init_rest(1.0)
        s.init_spinodal(8361235, 0.0, 0.05)
        f, phi = s.get(rh.REF_F), s.get(rh.REF_PHI)
        s.step(nsteps)
        ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U), ("force", rh.REF_FORCE))}
    st = dict(f=f, phi=phi, u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
              force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
    orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=order, gradmu=gradmu, force_method=1, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    for k in ref:
        assert np.array_equal(orc.interior(st[k]), orc.interior(ref[k])), k


@pytest.mark.parametrize("order,nlocal", [(1, (12, 10, 8)), (3, (16, 16, 16)), (2, (6, 12, 10)), (4, (8, 8, 8))])
def test_binary_time_steps_compensated_sum(order, nlocal):
    """cahn_hilliard_options_conserve 1 (PHI_CONSERVE_COMPENSATED_SUM): phi_ch_update_conserve / phi_ch_csum_kernel
    (src/phi_cahn_hilliard.c:1059-1094, 1181-1215) with the per-site Kahan compensation carried from step to step;
    oracle == compiled reference bit for bit, and different from the plain forward step (the option does something)."""
    nsteps = 12
    orc = Oracle(nlocal, nhalo=2)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=order, conserve=1, ghost_off=1, eta_shear=ETA, **BINARY) as s:
        s.init_rest(1.0)
        s.init_spinodal(8361235, 0.0, 0.05)
        f, phi = s.get(rh.REF_F), s.get(rh.REF_PHI)
        s.step(nsteps)
        ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U))}
    plain = None
    for conserve in (1, 0):
        st = dict(f=f.copy(), phi=phi.copy(), u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
                  force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
        orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=order, conserve=conserve, **BINARY), 1, nsteps,
                 st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
        if conserve:
            for k in ref:
                assert np.array_equal(orc.interior(st[k]), orc.interior(ref[k])), k
        else:
            plain = st
    assert not np.array_equal(orc.interior(plain["phi"]), orc.interior(ref["phi"]))


@pytest.mark.parametrize("nrelax,reduced", [(0, 0), (1, 1), (2, 0)])
@pytest.mark.parametrize("nvel", [19, 15, 27])
def test_single_fluid_time_steps(nvel, nrelax, reduced):
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    if nvel == 27 and nrelax == 2:
        pytest.skip("TRT undefined for D3Q27 in the reference")
    nlocal, nsteps = (8, 6, 10), 6
    orc = Oracle(nlocal, nhalo=1, nvel=nvel)
    fg = (1e-6, 2e-6, 3e-6)
    rng = np.random.default_rng(7)
    with rh.RefSim(nlocal, nhalo=1, nrelax=nrelax, halo_reduced=reduced, eta_shear=0.1, fbody=fg, nvel=nvel) as s:
        s.init_uniform_u(1.0, (0.002, 0.003, 0.004))
        f = s.get(rh.REF_F)
        orc.interior(f)[...] *= 1.0 + 1e-3 * (rng.random((nvel,) + nlocal) - 0.5)
        s.set(rh.REF_F, f)
        s.step(nsteps)
        ref_f, ref_u, ref_rho = s.get(rh.REF_F), s.get(rh.REF_U), s.get(rh.REF_RHO)
    u = np.zeros((3, orc.nsites)); rho = np.zeros((1, orc.nsites)); force = np.zeros((3, orc.nsites))
    orc.step(orc.collide_param(nrelax, 1.0, 0.1, force=fg), None, 0, nsteps, f, None, u, rho, force, None, None,
             halo_reduced=reduced)
    assert np.array_equal(orc.interior(f), orc.interior(ref_f))
    assert np.array_equal(orc.interior(u), orc.interior(ref_u))
    assert np.array_equal(orc.interior(rho), orc.interior(ref_rho))


# ---- symmetric_lb: two distributions (src/collision.c:604-1013, src/phi_lb_coupler.c) --------------------

def lb2_state(orc, rng):
    """Interior state of a two-distribution binary fluid: f near equilibrium, g carrying phi plus a small flux."""
    nv, ns = orc.nvel, orc.nsites
    f = np.zeros((2 * nv, ns))
    fi = orc.interior(f)
    n = orc.nlocal
    rho = 1.0 + 0.01 * (rng.random(n) - 0.5)
    for p in range(nv):
        fi[p] = rho * orc.wv[p] * (1.0 + 0.05 * (rng.random(n) - 0.5))
        fi[nv + p] = orc.wv[p] * 0.02 * (rng.random(n) - 0.5)
    fi[nv] += 0.05 * (rng.random(n) - 0.5)
    return f


@pytest.mark.parametrize("nrelax", [0, 1, 2])
@pytest.mark.parametrize("nvel", [19, 15, 27])
def test_collide_binary(nvel, nrelax):
    """lb_collide with ndist = 2 -> lb_collision_binary, one call on random (f, g, phi, grad, delsq, force)."""
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    if nvel == 27 and nrelax == 2:
        pytest.skip("TRT is undefined for D3Q27 in the reference")
    nlocal, nhalo = (5, 4, 6), 1
    orc = Oracle(nlocal, nhalo=nhalo, nvel=nvel)
    rng = np.random.default_rng(40 + nvel + nrelax)
    fb = (1e-5, -2e-5, 3e-5)
    par = dict(a=-0.00625, b=0.00625, kappa=0.004, mobility=0.45)
    with rh.RefSim(nlocal, nhalo=nhalo, nvel=nvel, ndist=2, nrelax=nrelax, have_phi=1, eta_shear=0.02, eta_bulk=0.05,
                   fbody=fb, **par) as s:
        f = lb2_state(orc, rng)
        s.set(rh.REF_F, f)
        phi = fill_random(s, orc, rh.REF_PHI, rng, 0.1, -0.5)
        grad = fill_random(s, orc, rh.REF_GRAD, rng, 0.05, -0.5)
        delsq = fill_random(s, orc, rh.REF_DELSQ, rng, 0.1, -0.5)
        force = fill_random(s, orc, rh.REF_FORCE, rng, 1e-4, -0.5)
        u0 = fill_random(s, orc, rh.REF_U, rng)
        s.op("collide")
        ref_f, ref_u = s.get(rh.REF_F), s.get(rh.REF_U)
    u = u0.copy()
    orc.collide_binary(orc.collide_param(nrelax, 1.0, 0.02, eta_bulk=0.05, force=fb), orc.symm_param(**par),
                       f, force, phi, grad, delsq, u)
    # the reference's contiguous-range kernel also "collides" the y/z halo sites between the first and the
    # last interior site (no status test in lb_collision_mrt2); they hold zeros here (rho = 0 -> NaN) and are
    # overwritten by lb_halo before anything reads them: compare the interior
    assert np.array_equal(orc.interior(f), orc.interior(ref_f))
    assert np.array_equal(orc.interior(u), orc.interior(ref_u))


@pytest.mark.parametrize("nvel", [19, 15, 27])
def test_phi_lb_coupler(nvel):
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    nlocal, nhalo = (4, 5, 6), 1
    orc = Oracle(nlocal, nhalo=nhalo, nvel=nvel)
    rng = np.random.default_rng(50)
    with rh.RefSim(nlocal, nhalo=nhalo, nvel=nvel, ndist=2, have_phi=1, **BINARY) as s:
        f = fill_random(s, orc, rh.REF_F, rng)
        phi0 = fill_random(s, orc, rh.REF_PHI, rng)
        s.op("phi_lb_to_field")
        ref_phi = s.get(rh.REF_PHI)
        s.set(rh.REF_PHI, phi0)
        s.op("phi_lb_from_field")
        ref_f = s.get(rh.REF_F)
    phi = phi0.copy()
    orc.phi_lb_to_field(f, phi)
    assert np.array_equal(phi, ref_phi)
    orc.phi_lb_from_field(phi0, f)
    assert np.array_equal(f, ref_f)


@pytest.mark.parametrize("nvel,reduced", [(19, 0), (19, 1), (15, 0), (27, 0)])
def test_symmetric_lb_steps(nvel, reduced):
    """Whole symmetric_lb time steps (spinodal start through phi_lb_from_field), every field bit for bit."""
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    nlocal, nhalo, nsteps = (8, 6, 10), 1, 5
    orc = Oracle(nlocal, nhalo=nhalo, nvel=nvel)
    par = dict(a=-0.00625, b=0.00625, kappa=0.004, mobility=3.75)
    fb = (1e-6, 2e-6, -1e-6)
    with rh.RefSim(nlocal, nhalo=nhalo, nvel=nvel, ndist=2, have_phi=1, eta_shear=ETA, ghost_off=1, fbody=fb,
                   halo_reduced=reduced, **par) as s:
        s.init_rest(1.0)
        s.init_spinodal(8361235, 0.0, 0.05)
        s.op("phi_lb_from_field")
        f = s.get(rh.REF_F)
        phi = s.get(rh.REF_PHI)
        s.step(nsteps)
        ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U),
                                        ("grad", rh.REF_GRAD), ("delsq", rh.REF_DELSQ))}
    z3 = lambda: np.zeros((3, orc.nsites))
    u, force, grad, delsq = z3(), z3(), z3(), np.zeros((1, orc.nsites))
    orc.step_lb2(orc.collide_param(0, 1.0, ETA, force=fb), orc.symm_param(**par), nsteps, f, phi, u, force, grad, delsq,
                 halo_reduced=reduced)
    got = dict(f=f, phi=phi, u=u, grad=grad, delsq=delsq)
    for k in ref:
        assert np.array_equal(orc.interior(got[k]), orc.interior(ref[k])), k


def test_gradient_d4():
    """grad_3d_27pt_fluid_d4 (src/gradient_3d_27pt_fluid.c:112-134): the operator applied to delsq, nhalo = 3."""
    nlocal, nhalo = (5, 6, 7), 3
    orc = Oracle(nlocal, nhalo=nhalo)
    rng = np.random.default_rng(60)
    with rh.RefSim(nlocal, nhalo=nhalo, have_phi=1, grad_level=4, **BINARY) as s:
        fill_random(s, orc, rh.REF_PHI, rng, 1.0, -0.5)
        delsq = fill_random(s, orc, rh.REF_DELSQ, rng, 1.0, -0.5)
        s.op("grad_d4")
        ref_gd, ref_dd = s.get(rh.REF_GRAD_DELSQ), s.get(rh.REF_DELSQ_DELSQ)
    gd, dd = np.zeros((3, orc.nsites)), np.zeros((1, orc.nsites))
    orc.grad_27pt_d4(delsq, gd, dd)
    assert np.array_equal(orc.region(gd, 1), orc.region(ref_gd, 1))
    assert np.array_equal(orc.region(dd, 1), orc.region(ref_dd, 1))


def test_pth_stress_then_force_driver():
    """pth_stress_compute + pth_force_fluid_driver called separately == the oracle's two operators."""
    nlocal, nhalo = (6, 5, 7), 2
    orc = Oracle(nlocal, nhalo=nhalo)
    rng = np.random.default_rng(61)
    with rh.RefSim(nlocal, nhalo=nhalo, have_phi=1, **BINARY) as s:
        phi = fill_random(s, orc, rh.REF_PHI, rng, 0.1, -0.5)
        grad = fill_random(s, orc, rh.REF_GRAD, rng, 0.05, -0.5)
        delsq = fill_random(s, orc, rh.REF_DELSQ, rng, 0.1, -0.5)
        force = fill_random(s, orc, rh.REF_FORCE, rng, 1e-4, -0.5)
        s.op("pth_stress_compute")
        ref_str = s.get(rh.REF_STR)
        s.op("pth_force_fluid_driver")
        ref_force = s.get(rh.REF_FORCE)
    strs = np.zeros((9, orc.nsites))
    orc.stress_symm(orc.symm_param(**BINARY), phi, grad, delsq, strs)
    # stored for x in [0, N+1], every y, z of the allocation
    v = lambda a: a.reshape((-1,) + orc.nall)[:, nhalo - 1:-(nhalo - 1)]
    assert np.array_equal(v(strs), v(ref_str))
    orc.force_divergence(strs, force)
    assert np.array_equal(orc.interior(force), orc.interior(ref_force))


@pytest.mark.parametrize("order,nlocal,nvel", [(1, (12, 10, 8), 19), (3, (10, 12, 14), 19), (2, (8, 8, 8), 27)])
def test_binary_time_steps_7pt_gradient(order, nlocal, nvel):
    """fd_gradient_calculation 3d_7pt_fluid for the scalar order parameter (grad_3d_7pt_fluid_d2,
    src/gradient_3d_7pt_fluid.c:76-99, 231-300) in whole binary-fluid steps: oracle == compiled reference bit for bit"""
    if not rh.available(nvel=nvel):
        pytest.skip("reference build for this velocity set missing")
    nsteps = 8
    orc = Oracle(nlocal, nhalo=2, nvel=nvel)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=order, grad_7pt=1, eta_shear=ETA, nvel=nvel, **BINARY) as s:
        s.init_rest(1.0)
        s.init_spinodal(8361235, 0.0, 0.05)
        f, phi = s.get(rh.REF_F), s.get(rh.REF_PHI)
        s.step(nsteps)
        ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U), ("force", rh.REF_FORCE),
                                         ("grad", rh.REF_GRAD), ("delsq", rh.REF_DELSQ))}
    st = dict(f=f.copy(), phi=phi.copy(), u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
              force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
    orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=order, grad_7pt=1, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    for k in ref:
        assert np.array_equal(orc.interior(st[k]), orc.interior(ref[k])), k


@pytest.mark.parametrize("nlocal", [(8, 8, 1), (1, 6, 5), (6, 1, 1)])
def test_field_halo_on_lattices_thinner_than_the_halo(nlocal):
    """nlocal[d] < nhalo (the reference's own pmpi08-le2d-fd1 regression runs 64 x 64 x 1 with nhalo 2): every send buffer
    is packed before any is unpacked (src/field.c:1412-1531), so the outer halo layer receives the halo's content from
    BEFORE the swap; the oracle reproduces that bit for bit (and so does the CUDA library: tests/test_gpu_thin.py)."""
    orc = Oracle(nlocal, nhalo=2)
    rng = np.random.default_rng(31)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=2, eta_shear=ETA, **BINARY) as s:
        phi = rng.random((1, orc.nsites))
        s.set(rh.REF_PHI, phi)
        s.op("phi_halo")
        ref = s.get(rh.REF_PHI)
    orc.field_halo(phi)
    assert np.array_equal(phi, ref)


@pytest.mark.parametrize("nlocal,order", [((8, 8, 1), 3), ((8, 1, 6), 2)])
def test_binary_time_steps_on_thin_lattices(nlocal, order):
    nsteps = 6
    orc = Oracle(nlocal, nhalo=2)
    with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=order, ghost_off=1, eta_shear=ETA, **BINARY) as s:
        s.init_rest(1.0)
        s.init_spinodal(8361235, 0.0, 0.05)
        f, phi = s.get(rh.REF_F), s.get(rh.REF_PHI)
        s.step(nsteps)
        ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U), ("force", rh.REF_FORCE),
                                         ("grad", rh.REF_GRAD), ("delsq", rh.REF_DELSQ))}
    st = dict(f=f.copy(), phi=phi.copy(), u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
              force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
    orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=order, **BINARY), 1, nsteps,
             st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
    for k in ref:
        assert np.array_equal(orc.interior(st[k]), orc.interior(ref[k])), k


def test_conserve_2_global_subtraction_vs_reference():
    """cahn_hilliard_options_conserve 2 (phi_ch_subtract_sum_phi_after_forward_step, src/phi_cahn_hilliard.c:1102-1169): the
    initial sum of the statistics code (doubly compensated) and whole time steps with the correction, oracle == compiled
    reference bit for bit with one OpenMP thread (the reference's own summation order depends on its thread count).  With a
    deliberately offset initial sum the correction is far above rounding, so the test also shows that it is applied."""
    nlocal, nsteps = (8, 6, 10), 5
    orc = Oracle(nlocal, nhalo=2)
    before = rh.omp_threads(0)
    rh.omp_threads(1)
    try:
        for offset in (0.0, 1.0e-3):
            with rh.RefSim(nlocal, nhalo=2, have_phi=1, adv_order=3, conserve=2, ghost_off=1, eta_shear=ETA, **BINARY) as s:
                s.init_rest(1.0)
                s.init_spinodal(8361235, 0.0, 0.05)
                f, phi = s.get(rh.REF_F), s.get(rh.REF_PHI)
                sum0 = s.phi_stats_time0()
                assert sum0 == orc.phi_sum_time0(phi)
                s.phi_init_sum_set(sum0 + offset)
                s.step(nsteps)
                ref = {k: s.get(w) for k, w in (("f", rh.REF_F), ("phi", rh.REF_PHI), ("u", rh.REF_U))}
            st = dict(f=f.copy(), phi=phi.copy(), u=np.zeros((3, orc.nsites)), rho=np.zeros((1, orc.nsites)),
                      force=np.zeros((3, orc.nsites)), grad=np.zeros((3, orc.nsites)), delsq=np.zeros((1, orc.nsites)))
            orc.step(orc.collide_param(0, 1.0, ETA), orc.symm_param(adv_order=3, conserve=2, phi_init_sum=sum0 + offset, **BINARY), 1, nsteps,
                     st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
            for k in ref:
                assert np.array_equal(orc.interior(st[k]), orc.interior(ref[k])), (k, offset)
            if offset:
                # the sum has been pulled to the offset value
                assert abs(orc.interior(st["phi"]).sum() - (sum0 + offset)) < 1e-12
    finally:
        if before > 0:
            rh.omp_threads(before)
