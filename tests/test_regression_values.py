"""Known answers printed by the reference's own regression logs (8-11 significant digits), reproduced by the
oracle from the reference's own initial conditions (restated in ludwig_b200/initial.py):

* tests/regression/d3q19-short/serial-spin-fd1.{inp,log}: 64^3 spinodal binary fluid, A = -B = -0.00625,
  K = 0.004, M = 1.25, eta = 0.00625, ghost modes off, 27pt gradient, advection order 1, seed 8361235.
  Log lines 82-84 (t = 0) and 95-107 (t = 10).
* tests/regression/d3q19-short/serial-dist-3du.log: uniform u = (0.002, 0.003, 0.004) on 32^3... momentum
  6.5536e+01 9.8304e+01 1.31072e+02 preserved exactly (lines 66, 75)."""
import numpy as np
import pytest

from common import BINARY, ETA
from ludwig_b200.initial import equilibrium_f, spinodal_phi
from oracle import Oracle, fed_density, stats_scalar


def approx(v, digits):
    return pytest.approx(v, rel=0.5 * 10.0 ** (1 - digits), abs=1e-30)


def test_serial_spin_fd1_log():
    n = (64, 64, 64)
    orc = Oracle(n, nhalo=2)
    f = equilibrium_f(n, 2)
    phi = spinodal_phi(n, 2, 8361235, 0.0, 0.1)       # default "noise" amplitude 0.1 (src/field_phi_init_rt.c:27)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    sp = orc.symm_param(adv_order=1, **BINARY)
    cp = orc.collide_param(0, 1.0, ETA)

    # t = 0 (log lines 82-86): statistics after the initial phi halo + gradient
    s = stats_scalar(orc, phi)
    assert s[0] == approx(3.1484764e+00, 8) and s[1] == approx(1.2010484e-05, 8)
    assert s[2] == approx(8.3289934e-04, 8)
    assert s[3] == approx(-4.9999916e-02, 8) and s[4] == approx(4.9999705e-02, 8)
    ph = phi.copy()
    orc.field_halo(ph)
    orc.grad_27pt(ph, grad, delsq)
    assert fed_density(orc, sp, ph, grad) == approx(-2.3227909424e-06, 11)

    orc.step(cp, sp, 1, 10, f, phi, u, rho, force, grad, delsq)

    # t = 10 (log lines 95-107)
    s = stats_scalar(orc, phi)
    assert s[0] == approx(3.1484764e+00, 8) and s[1] == approx(1.2010484e-05, 8)
    assert s[2] == approx(3.7820523e-04, 8)
    assert s[3] == approx(-4.7270149e-02, 8) and s[4] == approx(4.6821679e-02, 8)
    assert fed_density(orc, sp, phi, grad) == approx(-9.7510518349e-07, 11)
    # [rho] is printed from the distributions after propagation (src/stats_distribution.c:125-200)
    rho_f = f.sum(axis=0, keepdims=True)
    r = stats_scalar(orc, rho_f)
    assert r[0] == approx(262144.00, 8)
    # variance of rho ~ 1 is <q^2> - <q>^2 in double in the reference: only ~2 digits survive the cancellation
    assert r[2] == pytest.approx(1.5449642e-11, rel=5e-2)
    assert r[3] == approx(0.99998006808, 11) and r[4] == approx(1.00001625877, 11)
    ui = orc.interior(u)
    for a, (lo, hi) in enumerate(((-1.3145696e-05, 1.2773457e-05), (-1.3301763e-05, 1.3768024e-05),
                                  (-1.2618505e-05, 1.2966490e-05))):
        assert ui[a].min() == approx(lo, 8) and ui[a].max() == approx(hi, 8)
    # total momentum is a cancellation-dominated sum: the reference's tolerance is absolute 1e-12
    fi = orc.interior(f)
    mom = [(fi * orc.cv[:, a, None, None, None]).sum() for a in range(3)]
    assert np.allclose(mom, 0.0, atol=1e-10)


def test_serial_dist_3du_log():
    n = (64, 64, 64)                       # size 64_64_64 in serial-dist-3du.inp -> 262144 sites
    orc = Oracle(n, nhalo=1)
    f = equilibrium_f(n, 1, 1.0, (0.002, 0.003, 0.004))
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force = z(3), z(1), z(3)

    def momentum():
        fi = orc.interior(f).astype(np.longdouble)
        return [float((fi * orc.cv[:, a, None, None, None]).sum()) for a in range(3)]

    expect = [262144 * 0.002, 262144 * 0.003, 262144 * 0.004]     # 5.24288e+02 ... scaled from the log's 32^3
    assert momentum() == pytest.approx(expect, rel=1e-12)
    orc.step(orc.collide_param(0, 1.0, 0.1), None, 0, 10, f, None, u, rho, force, None, None)
    assert momentum() == pytest.approx(expect, rel=1e-12)
    assert np.allclose(orc.interior(u)[0], 0.002, rtol=1e-13) and np.allclose(orc.interior(u)[2], 0.004, rtol=1e-13)


def test_serial_dist_1dp_log():
    """tests/regression/d3q19-short/serial-dist-1dp.{inp,log}: single fluid, 32^3, viscosity 0.1, reduced distribution
    halo (lb_halo_openmp_reduced), 1-d Poiseuille profile u_x = umax x (L - x) 4 / L^2 at rho = 1 set through
    lb_1st_moment_equilib_set (src/distribution_rt.c:516-563).  Log lines 62-67 (t = 0) and 71-82 (t = 10)."""
    n = (32, 32, 32)
    orc = Oracle(n, nhalo=1)
    L, umax = 32.0, 0.001
    ux = np.zeros(orc.nall)
    for ic in range(1, 33):
        x = 1.0 * (0 + ic) - 0.5
        ux[ic, :, :] = umax * x * (L - x) * 4.0 / (L * L)
    f = orc.equilibrium(1.0, (ux.ravel(), 0.0, 0.0))
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force = z(3), z(1), z(3)

    def momentum():
        fi = orc.interior(f).astype(np.longdouble)
        return [float((fi * orc.cv[:, a, None, None, None]).sum()) for a in range(3)]

    assert momentum()[0] == approx(2.1856000e+01, 8)
    orc.step(orc.collide_param(0, 1.0, 0.1), None, 0, 10, f, None, u, rho, force, None, None, halo_reduced=1)
    assert momentum()[0] == approx(2.1856000e+01, 8) and abs(momentum()[1]) < 1e-10
    r = stats_scalar(orc, f.sum(axis=0, keepdims=True))
    # the reference forms the variance as sum(rho^2)/N - mean^2 in a sequential double sum (src/stats_distribution.c:90-110):
    # good to N eps ~ 4e-12 absolute, i.e. 5 digits of 1.9e-07; min / max / velocities are exact to the printed digits
    assert r[0] == approx(32768.00, 8) and r[2] == approx(1.9282755e-07, 5)
    assert r[3] == approx(0.99932343093, 11) and r[4] == approx(1.00067708627, 11)
    ui = orc.interior(u)
    assert ui[0].min() == approx(5.0228587e-04, 8) and ui[0].max() == approx(8.7868636e-04, 8)


def test_serial_spin_lb1_log():
    """tests/regression/d3q19-short/serial-spin-lb1.{inp,log}: `free_energy symmetric_lb` (two distributions,
    lb_collision_binary), 64^3, A = -B = -0.00625, K = 0.004, mobility 3.75, eta = 0.00625, nhalo = 1.
    Log lines 78-83 (t = 0) and 91-104 (t = 10)."""
    n = (64, 64, 64)
    orc = Oracle(n, nhalo=1)
    nv, ns = 19, orc.nsites
    f = np.zeros((2 * nv, ns))
    f[:nv] = equilibrium_f(n, 1)
    phi = spinodal_phi(n, 1, 8361235, 0.0, 0.1)
    orc.phi_lb_from_field(phi, f)                     # src/ludwig.c:402
    z = lambda k: np.zeros((k, ns))
    u, force, grad, delsq = z(3), z(3), z(3), z(1)
    par = dict(a=-0.00625, b=0.00625, kappa=0.004, mobility=3.75)
    sp = orc.symm_param(**par)
    cp = orc.collide_param(0, 1.0, ETA)

    orc.step_lb2(cp, sp, 10, f, phi, u, force, grad, delsq)

    # the printed phi is recomputed from the distribution after the last propagation (src/ludwig.c:2416-2419);
    # the free energy uses it with the gradients of the last step
    phi_end = phi.copy()
    orc.phi_lb_to_field(f, phi_end)
    s = stats_scalar(orc, phi_end)
    assert s[0] == approx(3.1484764e+00, 8) and s[1] == approx(1.2010484e-05, 8)
    assert s[2] == approx(6.9606559e-04, 8)
    assert s[3] == approx(-4.9172300e-02, 8) and s[4] == approx(4.9198035e-02, 8)
    assert fed_density(orc, sp, phi_end, grad) == approx(-1.9246336785e-06, 11)
    r = stats_scalar(orc, f[:nv].sum(axis=0, keepdims=True))
    assert r[0] == approx(262144.00, 8)
    assert r[3] == approx(0.99993867629, 11) and r[4] == approx(1.00004070448, 11)
    ui = orc.interior(u)
    for a, (lo, hi) in enumerate(((-1.1415683e-05, 1.1538576e-05), (-1.2562973e-05, 1.1995491e-05),
                                  (-1.3597212e-05, 1.1403472e-05))):
        assert ui[a].min() == approx(lo, 8) and ui[a].max() == approx(hi, 8)


def test_d3q27_serial_spin_n01_log():
    """tests/regression/d3q27/serial-spin-n01.{inp,log}: D3Q27, 16^3, symmetric free energy through the finite-difference
    route with fd_gradient_calculation 3d_7pt_fluid and advection order 2, viscosity 0.00625, mobility 1.25, seed 8361235:
    printed statistics at t = 0 and t = 10 to the printed digits."""
    n = (16, 16, 16)
    orc = Oracle(n, nhalo=2, nvel=27)
    f = orc.equilibrium(1.0, (0.0, 0.0, 0.0))
    phi = spinodal_phi(n, 2, 8361235, 0.0, 0.1)
    z = lambda k: np.zeros((k, orc.nsites))
    u, rho, force, grad, delsq = z(3), z(1), z(3), z(3), z(1)
    sp = orc.symm_param(adv_order=2, grad_7pt=1, **BINARY)
    cp = orc.collide_param(0, 1.0, ETA)

    s = stats_scalar(orc, phi)
    assert s[0] == approx(-2.2941536e+00, 8) and s[2] == approx(8.3652803e-04, 8)
    assert s[3] == approx(-4.9986867e-02, 8) and s[4] == approx(4.9997941e-02, 8)
    ph = phi.copy()
    orc.field_halo(ph)
    orc.grad_7pt(ph, grad, delsq)
    assert fed_density(orc, sp, ph, grad) == approx(-1.2851398593e-07, 11)

    orc.step(cp, sp, 1, 10, f, phi, u, rho, force, grad, delsq)

    s = stats_scalar(orc, phi)
    assert s[0] == approx(-2.2941536e+00, 8) and s[2] == approx(1.7010966e-04, 8)
    assert s[3] == approx(-4.2252684e-02, 8) and s[4] == approx(3.9099406e-02, 8)
    assert fed_density(orc, sp, phi, grad) == approx(-3.1541470368e-08, 11)
    r = stats_scalar(orc, f.sum(axis=0, keepdims=True))
    assert r[0] == approx(4096.00, 8) and r[3] == approx(0.99997295107, 11) and r[4] == approx(1.00002678655, 11)
    ui = orc.interior(u)
    for a, (lo, hi) in enumerate(((-1.3080740e-05, 1.4918725e-05), (-1.6836354e-05, 1.6014326e-05),
                                  (-1.8317661e-05, 1.4310906e-05))):
        assert ui[a].min() == approx(lo, 8) and ui[a].max() == approx(hi, 8)
