"""Lees-Edwards planes on the GPU, through the C-ABI, against the CPU oracle
(oracle/lb_oracle_le.c, pinned bit-for-bit to the compiled reference and to serial-le3d-st5/6/7.log in
tests/test_le_oracle.py) on identical inputs.

Bar: LB200_MATH_STRICT bit-exact for every operator and for whole time steps; LB200_MATH_FAST (FMA contraction,
fused phi sector for the bulk + plane patches) within 1e-12 relative (absolute floor 1e-14) after N steps."""
import numpy as np
import pytest

import ludwig_b200 as lb
from common import close_fast
from ludwig_b200.initial import spinodal_phi
from oracle import Oracle, fed_density, stats_scalar

pytestmark = pytest.mark.gpu

FE = dict(a=-0.0625, b=0.0625, kappa=0.04, mobility=0.15)     # serial-le3d-st*.inp
ETA = 0.1
UY = 0.05


def le_state(orc, seed=11, amp=0.05):
    """shear-profile distributions with a non-equilibrium perturbation, noisy phi and u (interior; halos zero)"""
    rng = np.random.default_rng(seed)
    n = orc.nlocal
    f = np.zeros((orc.nvel, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f)
    orc.interior(f)[...] *= 1.0 + 1e-3 * (rng.random(orc.interior(f).shape) - 0.5)
    phi = np.zeros((1, orc.nsites))
    orc.interior(phi)[0] = amp * (rng.random(n) - 0.5)
    u = np.zeros((3, orc.nsites))
    orc.interior(u)[...] = 0.02 * (rng.random((3,) + tuple(n)) - 0.5)
    return f, phi, u


def make(n, nplanes, order, math, nvel=19):
    orc = Oracle(n, nhalo=2, le_nplanes=nplanes, le_uy=UY, nvel=nvel)
    sim = lb.Lb200(n, nhalo=2, nvel=nvel, have_phi=True, math=math, le_nplanes=nplanes, le_uy=UY)
    assert sim.nsites == orc.nsites and sim.nsites_lb == orc.nsites_lb
    sp_o = orc.symm_param(FE["a"], FE["b"], FE["kappa"], FE["mobility"], adv_order=order)
    sp_g = lb.SymmParam.make(FE["a"], FE["b"], FE["kappa"], FE["mobility"], adv_order=order)
    return orc, sim, sp_o, sp_g


def test_le_geometry_host():
    """lb200_le_plane_location / lb200_le_ic_to_buff (pure host arithmetic) against the oracle's (= the reference's)"""
    lib = lb.load_library()
    import ctypes as C
    for n, npl in (((32, 8, 8), 2), ((16, 4, 4), 1), ((48, 4, 4), 4)):
        orc = Oracle(n, nhalo=2, le_nplanes=npl)
        o = lb.capi.Options()
        o.nlocal[:] = n
        o.nhalo, o.cart_size, o.cart_rank, o.le_nplanes = 2, 1, 0, npl
        for p in range(npl):
            assert lib.lb200_le_plane_location(C.byref(o), p) == orc.le_plane_location(p)
        for ic in range(1, n[0] + 1):
            for di in (-2, -1, 1, 2):
                assert lib.lb200_le_ic_to_buff(C.byref(o), ic, di) == orc.le_ic_to_buff(ic, di), (ic, di)


@pytest.mark.parametrize("n,nplanes", [((16, 8, 6), 1), ((16, 12, 40), 2), ((24, 7, 5), 2)])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_le_operators_strict_bit_exact(n, nplanes, order):
    orc, sim, sp_o, sp_g = make(n, nplanes, order, lb.MATH_STRICT)
    with sim:
        f, phi, u = le_state(orc)
        tcur = 8                                   # time = 7: displacements 0.35 / 0.4, fractional
        time, tstep = tcur - 1.0, float(tcur)
        sim.physics_control_time_set(0, tcur)
        sim.put(lb.F, f); sim.put(lb.PHI, phi); sim.put(lb.U, u)

        # field_halo + field_grad_compute
        sim.phi_halo(); sim.phi_grad_compute()
        grad = np.zeros((3, orc.nsites)); delsq = np.zeros((1, orc.nsites))
        orc.field_halo(phi); orc.le_field(time, phi); orc.grad_27pt(phi, grad, delsq); orc.le_grad_buffer(phi, grad, delsq)
        assert np.array_equal(sim.get(lb.PHI), phi)                       # buffer planes included
        gg, gd = sim.get(lb.GRAD), sim.get(lb.DELSQ)
        assert np.array_equal(orc.region(gg, 1), orc.region(grad, 1)) and np.array_equal(orc.region(gd, 1), orc.region(delsq, 1))
        for p in range(nplanes):
            for x in (orc.le_ic_to_buff(orc.le_plane_location(p), 1), orc.le_ic_to_buff(orc.le_plane_location(p) + 1, -1)):
                sl = lambda a: a.reshape((a.shape[0], -1) + orc.nall[1:])[:, x + 1, 1:-1, 1:-1]
                assert np.array_equal(sl(gg), sl(grad)) and np.array_equal(sl(gd), sl(delsq))

        # phi_force_calculation: flux form + per-plane correction
        sim.hydro_f_zero(); sim.phi_force_calculation(sp_g)
        force = np.zeros((3, orc.nsites))
        orc.le_phi_force(sp_o, phi, grad, delsq, force)
        assert np.array_equal(orc.interior(sim.get(lb.FORCE)), orc.interior(force))

        # phi_cahn_hilliard: u halo, hydro_lees_edwards, fluxes, flux fix, update
        sim.phi_cahn_hilliard(sp_g)
        orc.field_halo(u); orc.le_hydro(time, u, nhcomm=2)
        assert np.array_equal(sim.get(lb.U), u)
        flux = np.zeros((4, orc.nsites))
        orc.advection(order, u, phi, flux); orc.flux_mu(sp_o, phi, delsq, flux); orc.flux_mu_ext(sp_o, flux)
        orc.le_fix_fluxes(time, flux); orc.phi_update(flux, phi)
        assert np.array_equal(orc.interior(sim.get(lb.PHI)), orc.interior(phi))

        # lb_data_apply_le_boundary_conditions
        sim.lb_le_apply_boundary_conditions()
        orc.le_lb_bc(tstep, f)
        assert np.array_equal(orc.interior(sim.get(lb.F)), orc.interior(f))


@pytest.mark.parametrize("nvel", [15, 27])
def test_le_lb_bc_other_models(nvel):
    """the plane-crossing re-projection for the other velocity sets (5 / 9 crossing populations per side)"""
    n, npl = (16, 9, 7), 2
    orc = Oracle(n, nhalo=1, le_nplanes=npl, le_uy=UY, nvel=nvel)
    rng = np.random.default_rng(3)
    f = np.zeros((nvel, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f)
    orc.interior(f)[...] *= 1.0 + 1e-3 * (rng.random(orc.interior(f).shape) - 0.5)
    with lb.Lb200(n, nhalo=1, nvel=nvel, math=lb.MATH_STRICT, le_nplanes=npl, le_uy=UY) as sim:
        sim.physics_control_time_set(0, 13)
        sim.put(lb.F, f)
        sim.lb_le_apply_boundary_conditions()
        got = sim.get(lb.F)
    orc.le_lb_bc(13.0, f)
    assert np.array_equal(orc.interior(got), orc.interior(f))


def _run_steps(n, nplanes, order, math, nsteps, seed=13, wrap=1):
    orc, sim, sp_o, sp_g = make(n, nplanes, order, math)
    f = np.zeros((19, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f)
    phi = np.zeros((1, orc.nsites))
    phi[:, :orc.nsites_lb] = spinodal_phi(n, 2, seed, 0.0, 0.1)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    with sim:
        sim.set_knob(lb.KNOB_WRAP, wrap)
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, nsteps)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("rho", lb.RHO),
                                          ("force", lb.FORCE), ("grad", lb.GRAD), ("delsq", lb.DELSQ))}
        assert sim.physics_control_timestep() == nsteps
    orc.le_step(orc.collide_param(0, 1.0, ETA), sp_o, 0, nsteps, f, phi, u, rho, force, grad, delsq)
    want = dict(f=f, phi=phi, u=u, rho=rho, force=force, grad=grad, delsq=delsq)
    return orc, sp_o, got, want


@pytest.mark.parametrize("n,nplanes,order", [((16, 12, 8), 2, 1), ((16, 8, 40), 1, 3), ((24, 8, 6), 2, 2), ((32, 32, 32), 2, 3), ((16, 10, 8), 2, 4)])
def test_le_steps_strict_bit_exact(n, nplanes, order):
    orc, sp, got, want = _run_steps(n, nplanes, order, lb.MATH_STRICT, 12)
    for k in want:
        assert np.array_equal(orc.interior(got[k]), orc.interior(want[k])), k


@pytest.mark.parametrize("wrap", [1, 0])
@pytest.mark.parametrize("n,nplanes,order", [((16, 12, 8), 2, 1), ((16, 16, 40), 1, 3), ((24, 16, 16), 2, 2), ((32, 32, 32), 2, 3),
                                             ((12, 16, 16), 2, 3), ((16, 16, 16), 2, 4)])
def test_le_steps_fast_tolerance(n, nplanes, order, wrap):
    """wrap = 1: the halo-free step (periodic images read in-kernel) with the plane patches; wrap = 0 (and the last
    shape, whose planes sit too close to the x boundary for the halo-free step): the reference's step structure with
    halo kernels, fused phi sector for the bulk + plane patches"""
    orc, sp, got, want = _run_steps(n, nplanes, order, lb.MATH_FAST, 20, wrap=wrap)
    for k in want:
        assert close_fast(orc.interior(got[k]), orc.interior(want[k])), (k, np.abs(orc.interior(got[k]) - orc.interior(want[k])).max())


@pytest.mark.parametrize("order,var,lo,hi,fed,uylo,uyhi", [
    (3, 3.0000123e-04, -4.4451160e-02, 4.6772004e-02, -7.4768699749e-06, -2.3465114e-02, 2.3466305e-02),      # serial-le3d-st7.log
    (4, 3.3084154e-04, -4.5460883e-02, 4.9576356e-02, -8.3701208477e-06, -2.3468862e-02, 2.3468484e-02)])     # serial-le3d-st8.log
def test_serial_le3d_logs_on_gpu(order, var, lo, hi, fed, uylo, uyhi):
    """the reference's own regression answers (tests/regression/d3q19-short/serial-le3d-st7.log, advection order 3, and
    serial-le3d-st8.log, order 4 -- a host loop in the reference) from the CUDA path in fast mode: printed statistics to
    the printed digits"""
    orc, sp, got, want = _run_steps((32, 32, 32), 2, order, lb.MATH_FAST, 10, seed=7361237)
    approx = lambda v, d: pytest.approx(v, rel=0.5 * 10.0 ** (1 - d), abs=1e-30)
    s = stats_scalar(orc, got["phi"])
    assert s[2] == approx(var, 8) and s[3] == approx(lo, 8) and s[4] == approx(hi, 8)
    assert fed_density(orc, sp, got["phi"], got["grad"]) == approx(fed, 10)
    ui = orc.interior(got["u"])
    assert ui[1].min() == approx(uylo, 8) and ui[1].max() == approx(uyhi, 8)


def _run_steps_profiled(n, nplanes, order, nsteps, calls, seed=13):
    orc, sim, sp_o, sp_g = make(n, nplanes, order, lb.MATH_FAST)
    f = np.zeros((19, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f)
    phi = np.zeros((1, orc.nsites))
    phi[:, :orc.nsites_lb] = spinodal_phi(n, 2, seed, 0.0, 0.1)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    with sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        sim.profile(True)
        for _ in range(calls):
            sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, nsteps // calls)
        sim.sync()
        prof = sim.profile_get()
        sim.profile(False)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("rho", lb.RHO),
                                          ("force", lb.FORCE), ("grad", lb.GRAD), ("delsq", lb.DELSQ))}
    orc.le_step(orc.collide_param(0, 1.0, ETA), sp_o, 0, nsteps, f, phi, u, rho, force, grad, delsq)
    want = dict(f=f, phi=phi, u=u, rho=rho, force=force, grad=grad, delsq=delsq)
    return orc, got, want, prof


@pytest.mark.parametrize("n,nplanes,order,calls", [((32, 26, 64), 2, 3, 1), ((16, 16, 40), 1, 3, 3), ((48, 12, 8), 3, 1, 1),
                                                   ((24, 16, 16), 2, 2, 2), ((64, 40, 36), 4, 3, 1)])
def test_le_one_kernel_step(n, nplanes, order, calls):
    """the one-kernel step with Lees-Edwards planes: the sweep over the whole lattice, then planes loc-1 .. loc+2 of every
    plane produced again through the buffer planes (gradient, flux-form force, Cahn-Hilliard, pull-stream + collision,
    plane-crossing populations, y / z images) -- 12 steps in `calls` calls against the oracle, every field"""
    orc, got, want, prof = _run_steps_profiled(n, nplanes, order, 12, calls)
    assert prof["step_fused"][1] == 11 and prof["phi_sector"][1] == 1, prof
    for k in want:
        assert close_fast(orc.interior(got[k]), orc.interior(want[k])), (k, np.abs(orc.interior(got[k]) - orc.interior(want[k])).max())


@pytest.mark.parametrize("mode", ["1", "2"])
def test_le_one_kernel_step_chain_modes(mode, monkeypatch):
    """LB200_FUSED_LE=2: the patch chain runs next to the sweep on its own stream on intermediate steps; 1 (default): after
    it.  Same kernels on the same data: identical results, and equal to the oracle within the fast bar after 30 steps"""
    monkeypatch.setenv("LB200_FUSED_LE", mode)
    n = (48, 26, 36)
    orc, got, want, prof = _run_steps_profiled(n, 2, 3, 30, 2)
    assert prof["step_fused"][1] == 29
    for k in want:
        assert close_fast(orc.interior(got[k]), orc.interior(want[k])), (k, np.abs(orc.interior(got[k]) - orc.interior(want[k])).max())


def test_le_chain_next_to_the_sweep_is_deterministic(monkeypatch):
    """the concurrent chain (not profiled: that is when it runs on its own stream) twice, and once serial: bit-identical"""
    def run(mode):
        monkeypatch.setenv("LB200_FUSED_LE", mode)
        orc, sim, sp_o, sp_g = make((64, 32, 32), 4, 3, lb.MATH_FAST)
        f = np.zeros((19, orc.nsites_lb))
        orc.le_init_shear_profile(1.0, ETA, f)
        phi = np.zeros((1, orc.nsites))
        phi[:, :orc.nsites_lb] = spinodal_phi((64, 32, 32), 2, 5, 0.0, 0.1)
        with sim:
            sim.put(lb.F, f); sim.put(lb.PHI, phi)
            for _ in range(3):
                sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, 15)
            return orc, {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("force", lb.FORCE), ("grad", lb.GRAD))}
    orc, a = run("2")
    orc, b = run("2")
    orc, s = run("1")
    for k in a:
        assert np.array_equal(orc.interior(a[k]), orc.interior(b[k])), k
        assert np.array_equal(orc.interior(a[k]), orc.interior(s[k])), k


def test_le_one_kernel_step_equals_two_kernel_step(monkeypatch):
    """LB200_FUSED_LE=0 (phi sector + collision + patches) and the one-kernel form give the same fields within the fast bar"""
    n = (32, 24, 32)
    orc, a, want, pa = _run_steps_profiled(n, 2, 3, 20, 1)
    monkeypatch.setenv("LB200_FUSED_LE", "0")
    orc, b, want, pb = _run_steps_profiled(n, 2, 3, 20, 1)
    assert pa["step_fused"][1] == 19 and pb["step_fused"][1] == 0
    for k in a:
        assert close_fast(orc.interior(a[k]), orc.interior(b[k])), k
        assert close_fast(orc.interior(a[k]), orc.interior(want[k])), k


@pytest.mark.parametrize("math_mode", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
@pytest.mark.parametrize("conserve", [1, 2])
def test_le_steps_with_conserve_options(conserve, math_mode):
    """cahn_hilliard_options_conserve 1 (compensated per-site sum: bit-exact in strict mode) and 2 (global subtraction after
    the forward step, to 1e-14: the device sum is compensated, the reference's a plain one) with Lees-Edwards planes; the
    oracle for both is pinned to the compiled reference in tests/test_le_oracle.py::test_le_steps_conserve_vs_reference"""
    n, nplanes, order, nsteps = (16, 12, 10), 2, 3, 8
    orc = Oracle(n, nhalo=2, le_nplanes=nplanes, le_uy=UY)
    f = np.zeros((19, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f)
    phi = np.zeros((1, orc.nsites))
    phi[:, :orc.nsites_lb] = spinodal_phi(n, 2, 13, 0.0, 0.1)
    sum0 = orc.phi_sum_time0(phi) + 1.0e-3 if conserve == 2 else 0.0
    with lb.Lb200(n, nhalo=2, have_phi=True, math=math_mode, le_nplanes=nplanes, le_uy=UY) as sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        if conserve == 2:
            sim.phi_init_sum_set(sum0)
        sp_g = lb.SymmParam.make(FE["a"], FE["b"], FE["kappa"], FE["mobility"], adv_order=order, conserve=conserve)
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, nsteps // 2)
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, nsteps - nsteps // 2)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("force", lb.FORCE))}
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    sp_o = orc.symm_param(FE["a"], FE["b"], FE["kappa"], FE["mobility"], adv_order=order, conserve=conserve, phi_init_sum=sum0)
    orc.le_step(orc.collide_param(0, 1.0, ETA), sp_o, 0, nsteps, f, phi, u, rho, force, grad, delsq)
    want = dict(f=f, phi=phi, u=u, force=force)
    for k in want:
        a, b = orc.interior(got[k]), orc.interior(want[k])
        if math_mode == lb.MATH_STRICT and conserve == 1:
            assert np.array_equal(a, b), k
        elif math_mode == lb.MATH_STRICT:
            assert np.abs(a - b).max() <= 1e-14*np.abs(b).max(), k
        else:
            assert close_fast(a, b), (k, np.abs(a - b).max())
    if conserve == 2:
        assert abs(orc.interior(got["phi"]).sum() - sum0) < 1e-11


LB2 = dict(a=-0.0625, b=0.0625, kappa=0.04, mobility=0.45)      # serial-le2d-lb1.inp


def _lb2_le_run(n, nplanes, math_mode, nsteps, seed=13, calls=1):
    orc = Oracle(n, nhalo=2, le_nplanes=nplanes, le_uy=UY)
    phi = np.zeros((1, orc.nsites))
    phi[:, :orc.nsites_lb] = spinodal_phi(n, 2, seed, 0.0, 0.1)
    f = np.zeros((38, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f[:19])
    orc.phi_lb_from_field(phi, f)
    with lb.Lb200(n, nhalo=2, ndist=2, have_phi=True, math=math_mode, le_nplanes=nplanes, le_uy=UY) as sim:
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        sp_g = lb.SymmParam.make(LB2["a"], LB2["b"], LB2["kappa"], LB2["mobility"])
        for _ in range(calls):
            sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, nsteps // calls)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("grad", lb.GRAD), ("delsq", lb.DELSQ))}
        assert sim.physics_control_timestep() == nsteps
    z = lambda k: np.zeros((k, orc.nsites))
    u, force, grad, delsq = z(3), z(3), z(3), z(1)
    orc.le_step_lb2(orc.collide_param(0, 1.0, ETA), orc.symm_param(LB2["a"], LB2["b"], LB2["kappa"], LB2["mobility"]),
                    0, nsteps, f, phi, u, force, grad, delsq)
    return orc, got, dict(f=f, phi=phi, u=u, grad=grad, delsq=delsq)


@pytest.mark.parametrize("n,nplanes", [((16, 12, 8), 2), ((16, 8, 10), 1), ((32, 16, 1), 2)])
def test_le_symmetric_lb_steps_strict_bit_exact(n, nplanes):
    """free_energy symmetric_lb (two distributions) with Lees-Edwards planes, 3-d and 2-d (a lattice thinner than the halo):
    the oracle is pinned to the compiled reference and to serial-le2d-lb1.log in tests/test_le_oracle.py"""
    orc, got, want = _lb2_le_run(n, nplanes, lb.MATH_STRICT, 10, calls=2)
    for k in want:
        assert np.array_equal(orc.interior(got[k]), orc.interior(want[k])), k


def test_serial_le2d_lb1_log_on_gpu():
    """tests/regression/d3q19-short/serial-le2d-lb1.log (64 x 64 x 1, 2 planes, 200 steps) from the CUDA path, fast mode:
    the printed statistics to the printed digits"""
    orc, got, want = _lb2_le_run((64, 64, 1), 2, lb.MATH_FAST, 200)
    approx = lambda v, d: pytest.approx(v, rel=0.5 * 10.0 ** (1 - d), abs=1e-30)
    phi = np.zeros((1, orc.nsites))
    orc.phi_lb_to_field(got["f"], phi)                      # the driver's statistics recompute phi first (src/ludwig.c:2415-2420)
    s = stats_scalar(orc, phi)
    assert s[2] == approx(4.5598599e-03, 7) and s[3] == approx(-2.1040506e-01, 7) and s[4] == approx(2.3299586e-01, 7)
    r = stats_scalar(orc, got["f"][:19].sum(axis=0, keepdims=True))
    assert r[3] == approx(0.99956275287, 10) and r[4] == approx(1.00179280622, 10)
    ui = orc.interior(got["u"])
    assert ui[1].min() == approx(-2.4492205e-02, 7) and ui[1].max() == approx(2.4612861e-02, 7)
    for k in want:
        assert close_fast(orc.interior(got[k]), orc.interior(want[k])), k


@pytest.mark.parametrize("math_mode", [lb.MATH_STRICT, lb.MATH_FAST], ids=["strict", "fast"])
@pytest.mark.parametrize("n,nplanes,order", [((16, 12, 8), 2, 3), ((16, 8, 10), 1, 2)])
def test_le_steps_7pt_gradient(n, nplanes, order, math_mode):
    """fd_gradient_calculation 3d_7pt_fluid with planes (grad_3d_7pt_fluid_le; the configuration of the reference's
    serial-le3d-st1..4): oracle pinned to the compiled reference in tests/test_le_oracle.py"""
    orc, sim, sp_o, sp_g = make(n, nplanes, order, math_mode)
    sp_o = orc.symm_param(FE["a"], FE["b"], FE["kappa"], FE["mobility"], adv_order=order, grad_7pt=1)
    f = np.zeros((19, orc.nsites_lb))
    orc.le_init_shear_profile(1.0, ETA, f)
    phi = np.zeros((1, orc.nsites))
    phi[:, :orc.nsites_lb] = spinodal_phi(n, 2, 13, 0.0, 0.1)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    with sim:
        sim.set_knob(lb.KNOB_GRAD_7PT, 1)
        sim.put(lb.F, f); sim.put(lb.PHI, phi)
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), sp_g, 10)
        got = {k: sim.get(a) for k, a in (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("force", lb.FORCE), ("grad", lb.GRAD),
                                          ("delsq", lb.DELSQ))}
    orc.le_step(orc.collide_param(0, 1.0, ETA), sp_o, 0, 10, f, phi, u, rho, force, grad, delsq)
    want = dict(f=f, phi=phi, u=u, force=force, grad=grad, delsq=delsq)
    for k in want:
        a, b = orc.interior(got[k]), orc.interior(want[k])
        if math_mode == lb.MATH_STRICT:
            assert np.array_equal(a, b), k
        else:
            assert close_fast(a, b), (k, np.abs(a - b).max())
