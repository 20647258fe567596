"""ctypes binding of oracle/liboracle.so (the plain-C CPU restatement, oracle/lb_oracle.c).
Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(HERE, "..", "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")


def build(force=False):
    src = os.path.join(ORACLE_DIR, "lb_oracle.c")
    src_le = os.path.join(ORACLE_DIR, "lb_oracle_le.c")
    src_lc = os.path.join(ORACLE_DIR, "lb_oracle_lc.c")
    deps = [src, src_le, src_lc, os.path.join(ORACLE_DIR, "lb_oracle.h"), os.path.join(ORACLE_DIR, "d3q19_tables.h")]
    if (not force and os.path.exists(ORACLE_SO)
            and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(d) for d in deps)):
        return ORACLE_SO
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall",
                           "-o", ORACLE_SO, src, src_le, src_lc, "-lm"])
    return ORACLE_SO


class Geom(C.Structure):
    _fields_ = [("nlocal", C.c_int * 3), ("nhalo", C.c_int), ("periodic", C.c_int * 3), ("le_nplanes", C.c_int)]


class LcParam(C.Structure):
    _fields_ = [("a0", C.c_double), ("q0", C.c_double), ("gamma", C.c_double), ("kappa0", C.c_double),
                ("kappa1", C.c_double), ("xi", C.c_double), ("Gamma", C.c_double), ("epsilon", C.c_double),
                ("e0", C.c_double * 3), ("is_active", C.c_int), ("zeta0", C.c_double), ("zeta1", C.c_double),
                ("zeta2", C.c_double), ("redshift", C.c_double), ("rredshift", C.c_double), ("grad_2d5", C.c_int)]


class LeParam(C.Structure):
    _fields_ = [("uy", C.c_double), ("time0", C.c_double)]


class Model(C.Structure):
    _fields_ = [("nvel", C.c_int), ("ndim", C.c_int), ("cv", (C.c_byte * 3) * 27),
                ("wv", C.c_double * 27), ("na", C.c_double * 27),
                ("ma", (C.c_double * 27) * 27), ("mi", (C.c_double * 27) * 27)]


class CollideParam(C.Structure):
    _fields_ = [("nrelax", C.c_int), ("rho0", C.c_double), ("eta_shear", C.c_double),
                ("eta_bulk", C.c_double), ("force_global", C.c_double * 3)]


class SymmParam(C.Structure):
    _fields_ = [("a", C.c_double), ("b", C.c_double), ("kappa", C.c_double),
                ("mobility", C.c_double), ("gradmu", C.c_double * 3), ("adv_order", C.c_int), ("conserve", C.c_int),
                ("grad_7pt", C.c_int), ("phi_init_sum", C.c_double), ("force_method", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(ORACLE_SO)
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Stateless operator set bound to one geometry + model."""

    def __init__(self, nlocal, nhalo=1, periodic=(1, 1, 1), nvel=19, le_nplanes=0, le_uy=0.0):
        self.lib = lib()
        self.g = Geom()
        self.g.nlocal[:] = nlocal
        self.g.nhalo = nhalo
        self.g.periodic[:] = periodic
        self.g.le_nplanes = le_nplanes
        self.nlocal = tuple(nlocal)
        self.nhalo = nhalo
        self.nall = tuple(n + 2 * nhalo for n in nlocal)
        self.nsites_lb = int(np.prod(self.nall))              # distributions, map (cs_nsites)
        # hydro / field arrays carry 2*nhalo buffer x-planes per Lees-Edwards plane (lees_edw_nsites)
        self.le_nplanes = le_nplanes
        self.nxbuffer = 2 * nhalo * le_nplanes
        self.nsites = (self.nall[0] + self.nxbuffer) * self.nall[1] * self.nall[2]
        self.le = LeParam(le_uy, 0.0)
        self.lib.orc_le_buffer_displacement.restype = C.c_double
        self.m = Model()
        rc = self.lib.orc_model_create(nvel, C.byref(self.m))
        assert rc == 0
        self.nvel = nvel
        self.cv = np.array([[self.m.cv[p][a] for a in range(3)] for p in range(nvel)], dtype=np.int64)
        self.wv = np.array(self.m.wv[:nvel])

    # ---- helpers -------------------------------------------------------------------------
    def interior(self, a):
        """View of the interior of a (ncomp, nsites) canonical array as (ncomp, Nx, Ny, Nz)."""
        h = self.nhalo
        v = a.reshape((a.shape[0], -1) + self.nall[1:])       # x extent: nall[0] (+ LE buffer planes)
        return v[:, h:h + self.nlocal[0], h:h + self.nlocal[1], h:h + self.nlocal[2]]

    def region(self, a, extra):
        """View of [1-extra, N+extra]^3."""
        h = self.nhalo - extra
        v = a.reshape((a.shape[0], -1) + self.nall[1:])
        return v[:, h:self.nall[0] - h, h:self.nall[1] - h, h:self.nall[2] - h]

    def buffer(self, a, extra=0):
        """View of the Lees-Edwards buffer planes of a field array, y/z in [1-extra, N+extra]."""
        h = self.nhalo - extra
        v = a.reshape((a.shape[0], -1) + self.nall[1:])
        return v[:, self.nall[0]:, h:self.nall[1] - h, h:self.nall[2] - h]

    def collide_param(self, nrelax=0, rho0=1.0, eta_shear=1.0 / 6.0, eta_bulk=None, force=(0, 0, 0)):
        cp = CollideParam()
        cp.nrelax, cp.rho0, cp.eta_shear = nrelax, rho0, eta_shear
        cp.eta_bulk = eta_shear if eta_bulk is None else eta_bulk
        cp.force_global[:] = force
        return cp

    def symm_param(self, a, b, kappa, mobility, gradmu=(0, 0, 0), adv_order=1, conserve=0, grad_7pt=0, phi_init_sum=0.0,
                   force_method=0):
        sp = SymmParam()
        sp.a, sp.b, sp.kappa, sp.mobility, sp.adv_order, sp.conserve = a, b, kappa, mobility, adv_order, conserve
        sp.grad_7pt = grad_7pt
        sp.phi_init_sum = phi_init_sum
        sp.force_method = force_method
        sp.gradmu[:] = gradmu
        return sp

    def equilibrium(self, rho, u):
        """f_p = rho w_p (1 + 3 u.c + 4.5 (cc - 1/3 I):uu), reference src/lb_data.c:809-834 (same op order)."""
        f = np.zeros((self.nvel, self.nsites_lb))
        cs2 = 1.0 / 3.0
        rcs2 = 1.0 / cs2
        for p in range(self.nvel):
            udotc = 0.0
            sdotq = 0.0
            for ia in range(3):
                udotc = udotc + u[ia] * float(self.cv[p, ia])
                for ib in range(3):
                    dab = 1.0 if ia == ib else 0.0
                    sdotq = sdotq + (float(self.cv[p, ia] * self.cv[p, ib]) - cs2 * dab) * u[ia] * u[ib]
            f[p] = rho * self.wv[p] * (1.0 + rcs2 * udotc + 0.5 * rcs2 * rcs2 * sdotq)
        return f

    # ---- operators (in place on canonical arrays) -------------------------------------------
    def propagation(self, f, fprime, ndist=1):
        self.lib.orc_propagation(C.byref(self.g), C.byref(self.m), ndist, _p(f), _p(fprime))

    def lb_halo(self, f, ndist=1, reduced=0):
        self.lib.orc_lb_halo(C.byref(self.g), C.byref(self.m), ndist, reduced, _p(f))

    def field_halo(self, a):
        self.lib.orc_field_halo(C.byref(self.g), a.shape[0], _p(a))

    def collide(self, cp, f, force, rho, u, status=None, include_halo=0):
        st = status.ctypes.data_as(C.c_void_p) if status is not None else None
        self.lib.orc_collide(C.byref(self.g), C.byref(self.m), C.byref(cp), st, include_halo,
                             _p(f), _p(force), _p(rho), _p(u))

    def grad_27pt(self, phi, grad, delsq):
        self.lib.orc_grad_27pt(C.byref(self.g), _p(phi), _p(grad), _p(delsq))

    def grad_27pt_d4(self, delsq, grad_delsq, delsq_delsq):
        """grad_3d_27pt_fluid_d4: the same operator on delsq, region nhalo - 2."""
        self.lib.orc_grad_27pt_ne(C.byref(self.g), self.nhalo - 2, _p(delsq), _p(grad_delsq), _p(delsq_delsq))

    def stress_symm(self, sp, phi, grad, delsq, str_):
        self.lib.orc_stress_symm(C.byref(self.g), C.byref(sp), _p(phi), _p(grad), _p(delsq), _p(str_))

    def force_divergence(self, str_, force):
        self.lib.orc_force_divergence(C.byref(self.g), _p(str_), _p(force))

    def advection(self, order, u, phi, flux):
        self.lib.orc_advection(C.byref(self.g), order, _p(u), _p(phi), _p(flux))

    def flux_mu(self, sp, phi, delsq, flux):
        self.lib.orc_flux_mu(C.byref(self.g), C.byref(sp), _p(phi), _p(delsq), _p(flux))

    def flux_mu_ext(self, sp, flux):
        self.lib.orc_flux_mu_ext(C.byref(self.g), C.byref(sp), _p(flux))

    def no_flux(self, status, flux):
        st = status.ctypes.data_as(C.c_void_p) if status is not None else None
        self.lib.orc_no_flux(C.byref(self.g), st, _p(flux))

    def phi_update(self, flux, phi):
        self.lib.orc_phi_update(C.byref(self.g), _p(flux), _p(phi))

    def phi_sum_time0(self, phi):
        self.lib.orc_phi_sum_time0.restype = C.c_double
        return self.lib.orc_phi_sum_time0(C.byref(self.g), _p(phi))

    def phi_subtract_sum(self, phi_init_sum, phi):
        self.lib.orc_phi_subtract_sum(C.byref(self.g), C.c_double(phi_init_sum), _p(phi))

    def phi_update_conserve(self, flux, csum, phi):
        self.lib.orc_phi_update_conserve(C.byref(self.g), _p(flux), _p(csum), _p(phi))

    # ---- Lees-Edwards (oracle/lb_oracle_le.c) ---------------------------------------------------
    def le_plane_location(self, p):
        return self.lib.orc_le_plane_location(C.byref(self.g), p)

    def le_ic_to_buff(self, ic, di):
        return self.lib.orc_le_ic_to_buff(C.byref(self.g), ic, di)

    def le_field(self, t, a):
        self.lib.orc_le_field(C.byref(self.g), C.byref(self.le), C.c_double(t), a.shape[0], _p(a))

    def le_hydro(self, t, u, nhcomm=1):
        self.lib.orc_le_hydro(C.byref(self.g), C.byref(self.le), C.c_double(t), nhcomm, _p(u))

    def le_grad_buffer(self, phi, grad, delsq):
        self.lib.orc_le_grad_buffer(C.byref(self.g), self.nhalo - 1, _p(phi), _p(grad), _p(delsq))

    def le_phi_force(self, sp, phi, grad, delsq, force):
        self.lib.orc_le_phi_force(C.byref(self.g), C.byref(sp), _p(phi), _p(grad), _p(delsq), _p(force))

    def le_fix_fluxes(self, t, flux):
        self.lib.orc_le_fix_fluxes(C.byref(self.g), C.byref(self.le), C.c_double(t), _p(flux))

    def le_lb_bc(self, tstep, f, ndist=1):
        self.lib.orc_le_lb_bc(C.byref(self.g), C.byref(self.m), C.byref(self.le), C.c_double(tstep), ndist, _p(f))

    def le_init_shear_profile(self, rho0, eta, f):
        self.lib.orc_le_init_shear_profile(C.byref(self.g), C.byref(self.m), C.byref(self.le), C.c_double(rho0),
                                           C.c_double(eta), _p(f))

    def le_step_lb2(self, cp, sp, tcurrent0, nsteps, f, phi, u, force, grad, delsq):
        self.lib.orc_le_step_lb2(C.byref(self.g), C.byref(self.m), C.byref(cp), C.byref(sp), C.byref(self.le),
                                 C.c_int(tcurrent0), C.c_int(nsteps), _p(f), _p(phi), _p(u), _p(force), _p(grad), _p(delsq))

    def le_step(self, cp, sp, tcurrent0, nsteps, f, phi, u, rho, force, grad, delsq):
        self.lib.orc_le_step(C.byref(self.g), C.byref(self.m), C.byref(cp), C.byref(sp), C.byref(self.le),
                             tcurrent0, nsteps, _p(f), _p(phi), _p(u), _p(rho), _p(force), _p(grad), _p(delsq))

    # ---- liquid crystal (oracle/lb_oracle_lc.c) ----------------------------------------------------
    def lc_param(self, a0, q0, gamma, kappa0, kappa1, xi, Gamma, epsilon=0.0, e0=(0.0, 0.0, 0.0), zeta0=None, zeta1=0.0, redshift=1.0, grad_2d5=0):
        p = LcParam()
        p.a0, p.q0, p.gamma, p.kappa0, p.kappa1, p.xi, p.Gamma, p.epsilon = a0, q0, gamma, kappa0, kappa1, xi, Gamma, epsilon
        p.e0[:] = e0
        p.is_active = int(zeta0 is not None)
        p.zeta0, p.zeta1, p.zeta2 = (zeta0 or 0.0), zeta1, 0.0
        p.redshift, p.rredshift = redshift, 1.0 / redshift
        p.grad_2d5 = grad_2d5
        return p

    def grad_2d_5pt(self, field, grad, delsq):
        self.lib.orc_grad_2d_5pt(C.byref(self.g), field.shape[0], _p(field), _p(grad), _p(delsq))

    def grad_7pt(self, field, grad, delsq):
        self.lib.orc_grad_7pt(C.byref(self.g), field.shape[0], _p(field), _p(grad), _p(delsq))

    def lc_stress(self, p, q, qgrad, qdelsq, str_):
        self.lib.orc_lc_stress(C.byref(self.g), C.byref(p), _p(q), _p(qgrad), _p(qdelsq), _p(str_))

    def lc_mol_field(self, p, q, qgrad, qdelsq, h):
        self.lib.orc_lc_mol_field(C.byref(self.g), C.byref(p), _p(q), _p(qgrad), _p(qdelsq), _p(h))

    def lc_fed_sum(self, p, q, qgrad):
        self.lib.orc_lc_fed_sum.restype = C.c_double
        return self.lib.orc_lc_fed_sum(C.byref(self.g), C.byref(p), _p(q), _p(qgrad))

    def advection_nf(self, order, u, field, flux):
        self.lib.orc_advection_nf(C.byref(self.g), order, field.shape[0], _p(u), _p(field), _p(flux))

    def beris_edw_update(self, p, u, h, flux, q):
        self.lib.orc_beris_edw_update(C.byref(self.g), C.c_double(p.xi), C.c_double(p.Gamma), _p(u), _p(h), _p(flux), _p(q))

    def lc_step(self, cp, p, adv_order, nsteps, f, q, u, rho, force, qgrad, qdelsq):
        self.lib.orc_lc_step(C.byref(self.g), C.byref(self.m), C.byref(cp), C.byref(p), adv_order, nsteps,
                             _p(f), _p(q), _p(u), _p(rho), _p(force), _p(qgrad), _p(qdelsq))

    # ---- symmetric_lb (two distributions: f is (2*nvel, nsites)) ------------------------------
    def phi_lb_to_field(self, f, phi):
        self.lib.orc_phi_lb_to_field(C.byref(self.g), C.byref(self.m), _p(f), _p(phi))

    def phi_lb_from_field(self, phi, f):
        self.lib.orc_phi_lb_from_field(C.byref(self.g), C.byref(self.m), _p(phi), _p(f))

    def collide_binary(self, cp, sp, f, force, phi, grad, delsq, u):
        self.lib.orc_collide_binary(C.byref(self.g), C.byref(self.m), C.byref(cp), C.byref(sp), _p(f), _p(force),
                                    _p(phi), _p(grad), _p(delsq), _p(u))

    def step_lb2(self, cp, sp, nsteps, f, phi, u, force, grad, delsq, halo_reduced=0):
        self.lib.orc_step_lb2(C.byref(self.g), C.byref(self.m), C.byref(cp), C.byref(sp), halo_reduced, nsteps,
                              _p(f), _p(phi), _p(u), _p(force), _p(grad), _p(delsq))

    def step(self, cp, sp, binary, nsteps, f, phi, u, rho, force, grad, delsq, halo_reduced=0):
        spp = C.byref(sp) if sp is not None else None
        args = [(_p(x) if x is not None else None) for x in (f, phi, u, rho, force, grad, delsq)]
        self.lib.orc_step(C.byref(self.g), C.byref(self.m), C.byref(cp), spp, int(binary),
                          halo_reduced, nsteps, *args)


# ---- observables the reference prints (numpy restatement) -----------------------------------

def stats_scalar(orc, a):
    """sum, mean, variance, min, max over the interior (reference src/phi_stats.c:91-145)."""
    v = orc.interior(a)[0].ravel()
    n = v.size
    s = float(np.sum(v, dtype=np.longdouble))
    mean = s / n
    var = float(np.sum(v.astype(np.longdouble) ** 2)) / n - mean * mean
    return s, mean, var, float(v.min()), float(v.max())


def fed_density(orc, sp, phi, grad):
    """(1/V) sum_interior [(a/2 + b/4 phi^2) phi^2 + kappa/2 |grad phi|^2]
    (reference src/symmetric.c:284-299, src/stats_free_energy.c:76-134)."""
    ph = orc.interior(phi)[0].astype(np.longdouble)
    g = orc.interior(grad).astype(np.longdouble)
    fed = (0.5 * sp.a + 0.25 * sp.b * ph * ph) * ph * ph + 0.5 * sp.kappa * (g * g).sum(axis=0)
    return float(fed.sum() / ph.size)
