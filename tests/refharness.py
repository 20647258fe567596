"""ctypes binding of oracle/_ref/libludwig_ref.so (the UNMODIFIED reference built by
oracle/Makefile.ref plus our harness oracle/ref_harness.c).  Test infrastructure only."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "..", "oracle", "_ref", "libludwig_ref.so")
REF_FAST_SO = os.path.join(HERE, "..", "oracle", "_ref", "libludwig_ref_fast.so")

REF_F, REF_PHI, REF_U, REF_RHO, REF_FORCE, REF_GRAD, REF_DELSQ, REF_STR, REF_FLUX, REF_MAP, \
    REF_GRAD_DELSQ, REF_DELSQ_DELSQ, REF_Q, REF_QGRAD, REF_QDELSQ, REF_H = range(16)
NCOMP = {REF_PHI: 1, REF_U: 3, REF_RHO: 1, REF_FORCE: 3, REF_GRAD: 3, REF_DELSQ: 1,
         REF_STR: 9, REF_FLUX: 4, REF_MAP: 1, REF_GRAD_DELSQ: 3, REF_DELSQ_DELSQ: 1,
         REF_Q: 5, REF_QGRAD: 15, REF_QDELSQ: 5, REF_H: 5}


class RefCfg(C.Structure):
    _fields_ = [("ntotal", C.c_int * 3), ("nhalo", C.c_int), ("periodic", C.c_int * 3),
                ("ndist", C.c_int), ("nrelax", C.c_int), ("ghost_off", C.c_int),
                ("halo_reduced", C.c_int), ("have_phi", C.c_int), ("adv_order", C.c_int),
                ("conserve", C.c_int),
                ("rho0", C.c_double), ("eta_shear", C.c_double), ("eta_bulk", C.c_double),
                ("fbody", C.c_double * 3), ("a", C.c_double), ("b", C.c_double),
                ("kappa", C.c_double), ("mobility", C.c_double), ("gradmu", C.c_double * 3),
                ("grad_level", C.c_int), ("le_nplanes", C.c_int), ("le_uy", C.c_double),
                ("have_q", C.c_int), ("lc_a0", C.c_double), ("lc_q0", C.c_double), ("lc_gamma", C.c_double),
                ("lc_kappa0", C.c_double), ("lc_kappa1", C.c_double), ("lc_xi", C.c_double), ("lc_Gamma", C.c_double),
                ("lc_epsilon", C.c_double), ("lc_e0", C.c_double * 3), ("grad_7pt", C.c_int),
                ("io_ascii", C.c_int), ("lc_active", C.c_int), ("lc_zeta0", C.c_double), ("lc_zeta1", C.c_double),
                ("lc_redshift", C.c_double), ("lc_grad_2d5", C.c_int), ("force_gradmu", C.c_int)]


def _so(fast=False, nvel=19):
    if nvel != 19:
        return os.path.join(HERE, "..", "oracle", "_ref", f"libludwig_ref_d3q{nvel}.so")
    return REF_FAST_SO if fast else REF_SO


def available(fast=False, nvel=19):
    return os.path.exists(_so(fast, nvel))


_libs = {}


def _lib(fast=False, nvel=19):
    key = (fast, nvel)
    if key not in _libs:
        lib = C.CDLL(_so(fast, nvel))
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = [C.POINTER(RefCfg)]
        lib.ref_time_steps.restype = C.c_double
        lib.ref_time_steps.argtypes = [C.c_void_p, C.c_int]
        for name in ("ref_free", "ref_nsites", "ref_hydro_f_zero", "ref_hydro_u_zero", "ref_hydro_u_halo",
                     "ref_phi_halo", "ref_grad_compute", "ref_phi_force", "ref_cahn_hilliard",
                     "ref_collide", "ref_lb_halo", "ref_propagation", "ref_phi_lb_to_field", "ref_phi_lb_from_field", "ref_grad_d4", "ref_pth_stress_compute",
                     "ref_pth_force_fluid_driver", "ref_nsites_le", "ref_le_field", "ref_le_hydro", "ref_le_lb_bc",
                     "ref_le_init_shear_profile", "ref_next_step", "ref_timestep", "ref_q_halo", "ref_q_grad_compute",
                     "ref_lc_stress_compute", "ref_beris_edw_update"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.ref_step.argtypes = [C.c_void_p, C.c_int]
        lib.ref_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_init_rest.argtypes = [C.c_void_p, C.c_double]
        lib.ref_init_uniform_u.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double * 3)]
        lib.ref_init_spinodal.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        for name in ("ref_lb_io_write", "ref_lb_io_read", "ref_field_io_write", "ref_field_io_read"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
        lib.ref_lc_twist_init.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib.ref_lc_o8m_init.argtypes = [C.c_void_p, C.c_double]
        lib.ref_lc_fed_sum.argtypes = [C.c_void_p]
        lib.ref_lc_fed_sum.restype = C.c_double
        lib.ref_nvel.restype = C.c_int
        lib.ref_phi_stats_time0.argtypes = [C.c_void_p]
        lib.ref_phi_stats_time0.restype = C.c_double
        lib.ref_phi_init_sum_set.argtypes = [C.c_void_p, C.c_double]
        if hasattr(lib, "ref_omp_threads"):
            lib.ref_omp_threads.argtypes = [C.c_int]
            lib.ref_omp_threads.restype = C.c_int
        assert lib.ref_nvel() == nvel
        _libs[key] = lib
    return _libs[key]


def omp_threads(n=0, fast=False, nvel=19):
    """Set (n > 0) and return the OpenMP team size the reference's kernels run with in this process."""
    lib = _lib(fast, nvel)
    return lib.ref_omp_threads(n) if hasattr(lib, "ref_omp_threads") else 0


class RefSim:
    """One reference simulation (the reference keeps a `physics` singleton: one at a time)."""

    def __init__(self, ntotal, nhalo=1, periodic=(1, 1, 1), ndist=1, nrelax=0, ghost_off=0,
                 halo_reduced=0, have_phi=0, adv_order=1, conserve=0, rho0=1.0, eta_shear=1.0 / 6.0,
                 eta_bulk=None, fbody=(0, 0, 0), a=0.0, b=0.0, kappa=0.0, mobility=0.0,
                 gradmu=(0, 0, 0), fast=False, nvel=19, grad_level=2, le_nplanes=0, le_uy=0.0, lc=None, grad_7pt=0, io_ascii=0, force_gradmu=0):
        self.lib = _lib(fast, nvel)
        cfg = RefCfg()
        cfg.ntotal[:] = ntotal
        cfg.nhalo = nhalo
        cfg.periodic[:] = periodic
        cfg.ndist, cfg.nrelax, cfg.ghost_off, cfg.halo_reduced = ndist, nrelax, ghost_off, halo_reduced
        cfg.have_phi, cfg.adv_order, cfg.conserve = have_phi, adv_order, conserve
        cfg.rho0, cfg.eta_shear = rho0, eta_shear
        cfg.eta_bulk = eta_shear if eta_bulk is None else eta_bulk
        cfg.fbody[:] = fbody
        cfg.a, cfg.b, cfg.kappa, cfg.mobility = a, b, kappa, mobility
        cfg.gradmu[:] = gradmu
        cfg.grad_level = grad_level
        cfg.grad_7pt = grad_7pt
        cfg.io_ascii = io_ascii
        cfg.force_gradmu = force_gradmu
        cfg.le_nplanes, cfg.le_uy = le_nplanes, le_uy
        if lc is not None:
            # liquid crystal: dict(a0, q0, gamma, kappa0, kappa1, xi, Gamma[, epsilon, e0])
            cfg.have_q = 1
            cfg.lc_a0, cfg.lc_q0, cfg.lc_gamma = lc["a0"], lc["q0"], lc["gamma"]
            cfg.lc_kappa0, cfg.lc_kappa1, cfg.lc_xi, cfg.lc_Gamma = lc["kappa0"], lc["kappa1"], lc["xi"], lc["Gamma"]
            cfg.lc_epsilon = lc.get("epsilon", 0.0)
            cfg.lc_e0[:] = lc.get("e0", (0.0, 0.0, 0.0))
            cfg.lc_active = int(lc.get("zeta0") is not None)
            cfg.lc_zeta0, cfg.lc_zeta1 = (lc.get("zeta0") or 0.0), lc.get("zeta1", 0.0)
            cfg.lc_redshift = lc.get("redshift", 1.0)
            cfg.lc_grad_2d5 = int(lc.get("grad_2d5", 0))
        self.cfg = cfg
        self.h = self.lib.ref_create(C.byref(cfg))
        self.nsites = self.lib.ref_nsites(self.h)
        self.nsites_le = self.lib.ref_nsites_le(self.h)      # hydro / field arrays carry the LE buffer planes
        self.ndist = ndist
        self.nvel = nvel

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def get(self, what):
        n = self.ndist * self.nvel if what == REF_F else NCOMP[what]
        ns = self.nsites if what in (REF_F, REF_MAP, REF_STR) else self.nsites_le
        out = np.empty((n, ns), dtype=np.float64)
        rc = self.lib.ref_get(self.h, what, out.ctypes.data)
        assert rc == 0
        return out

    def set(self, what, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        rc = self.lib.ref_set(self.h, what, arr.ctypes.data)
        assert rc == 0

    def init_rest(self, rho0=1.0):
        self.lib.ref_init_rest(self.h, rho0)

    def init_uniform_u(self, rho, u):
        self.lib.ref_init_uniform_u(self.h, rho, C.byref((C.c_double * 3)(*u)))

    def init_spinodal(self, seed, phi0, amp):
        self.lib.ref_init_spinodal(self.h, seed, phi0, amp)

    def lc_twist_init(self, axis, amplitude):
        self.lib.ref_lc_twist_init(self.h, axis, amplitude)

    def lc_o8m_init(self, amplitude):
        self.lib.ref_lc_o8m_init(self.h, amplitude)

    def lc_fed_sum(self):
        return self.lib.ref_lc_fed_sum(self.h)

    def io(self, name, timestep):
        """lb_io_write / lb_io_read / field_io_write / field_io_read in the current directory"""
        return getattr(self.lib, "ref_" + name)(self.h, timestep)

    def phi_stats_time0(self):
        """cahn_hilliard_stats_time0: sets and returns phi->field_init_sum (needed by conserve 2)"""
        return self.lib.ref_phi_stats_time0(self.h)

    def phi_init_sum_set(self, v):
        self.lib.ref_phi_init_sum_set(self.h, v)

    def op(self, name):
        return getattr(self.lib, "ref_" + name)(self.h)

    def step(self, n=1):
        self.lib.ref_step(self.h, n)

    def time_steps(self, n):
        return self.lib.ref_time_steps(self.h, n)
