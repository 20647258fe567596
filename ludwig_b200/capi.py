"""ctypes binding of libludwig_b200.so (include/ludwig_b200.h)."""
import ctypes as C
import os
import numpy as np

F, PHI, U, RHO, FORCE, GRAD, DELSQ, MAP, GRAD_DELSQ, DELSQ_DELSQ, STR, Q, QGRAD, QDELSQ = range(14)
RELAX_M10, RELAX_BGK, RELAX_TRT = 0, 1, 2
HALO_FULL, HALO_REDUCED = 1, 2
MATH_FAST, MATH_STRICT = 0, 1
H2D, D2H = 1, 2
KNOB_WRAP, KNOB_PHI_SECTOR, KNOB_PEER, KNOB_PIPE, KNOB_PIPE_SMS, KNOB_F32, KNOB_GRAD_7PT, KNOB_FUSED = 1, 2, 3, 4, 5, 6, 7, 8
KNOB_QGRAD_2D5 = 9

_NCOMP = {PHI: 1, U: 3, RHO: 1, FORCE: 3, GRAD: 3, DELSQ: 1, MAP: 1, GRAD_DELSQ: 3, DELSQ_DELSQ: 1, STR: 9,
          Q: 5, QGRAD: 15, QDELSQ: 5}


class Lb200Error(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [("nlocal", C.c_int * 3), ("nhalo", C.c_int), ("periodic", C.c_int * 3),
                ("nvel", C.c_int), ("ndist", C.c_int), ("have_phi", C.c_int),
                ("halo_scheme", C.c_int), ("math", C.c_int), ("device", C.c_int),
                ("cart_size", C.c_int), ("cart_rank", C.c_int),
                ("le_nplanes", C.c_int), ("le_uy", C.c_double), ("le_nt0", C.c_int), ("have_q", C.c_int)]


class CollideParam(C.Structure):
    _fields_ = [("nrelax", C.c_int), ("rho0", C.c_double), ("eta_shear", C.c_double),
                ("eta_bulk", C.c_double), ("force_global", C.c_double * 3)]

    @classmethod
    def make(cls, nrelax=RELAX_M10, rho0=1.0, eta_shear=1.0 / 6.0, eta_bulk=None, force=(0.0, 0.0, 0.0)):
        cp = cls()
        cp.nrelax, cp.rho0, cp.eta_shear = nrelax, rho0, eta_shear
        cp.eta_bulk = eta_shear if eta_bulk is None else eta_bulk
        cp.force_global[:] = force
        return cp


class SymmParam(C.Structure):
    _fields_ = [("a", C.c_double), ("b", C.c_double), ("kappa", C.c_double),
                ("mobility", C.c_double), ("gradmu", C.c_double * 3), ("adv_order", C.c_int), ("conserve", C.c_int),
                ("force_method", C.c_int)]

    @classmethod
    def make(cls, a, b, kappa, mobility, gradmu=(0.0, 0.0, 0.0), adv_order=1, conserve=0, force_method=0):
        sp = cls()
        sp.a, sp.b, sp.kappa, sp.mobility, sp.adv_order, sp.conserve = a, b, kappa, mobility, adv_order, conserve
        sp.force_method = force_method
        sp.gradmu[:] = gradmu
        return sp


class LcParam(C.Structure):
    _fields_ = [("a0", C.c_double), ("q0", C.c_double), ("gamma", C.c_double), ("kappa0", C.c_double),
                ("kappa1", C.c_double), ("xi", C.c_double), ("Gamma", C.c_double), ("epsilon", C.c_double),
                ("e0", C.c_double * 3), ("adv_order", C.c_int), ("is_active", C.c_int), ("zeta0", C.c_double),
                ("zeta1", C.c_double), ("zeta2", C.c_double), ("redshift", C.c_double)]

    @classmethod
    def make(cls, a0, q0, gamma, kappa0, kappa1, xi, Gamma, epsilon=0.0, e0=(0.0, 0.0, 0.0), adv_order=1,
             zeta0=None, zeta1=0.0, zeta2=0.0, redshift=1.0):
        p = cls()
        p.a0, p.q0, p.gamma, p.kappa0, p.kappa1, p.xi, p.Gamma, p.epsilon = a0, q0, gamma, kappa0, kappa1, xi, Gamma, epsilon
        p.e0[:] = e0
        p.adv_order = adv_order
        p.is_active = int(zeta0 is not None)      # lc_activity yes: zeta0 / zeta1 / zeta2 are read
        p.zeta0, p.zeta1, p.zeta2 = (zeta0 or 0.0), zeta1, zeta2
        p.redshift = redshift
        return p


class SlabPlan(C.Structure):
    _fields_ = [("left", C.c_int), ("right", C.c_int), ("has_lo", C.c_int), ("has_hi", C.c_int),
                ("nsites", C.c_longlong), ("chunk", C.c_longlong), ("off_lo", C.c_longlong),
                ("off_hi", C.c_longlong), ("halo_lo", C.c_longlong), ("halo_hi", C.c_longlong),
                ("count", C.c_longlong)]


class StepPlan(C.Structure):
    _fields_ = [("left", C.c_int), ("right", C.c_int), ("depth", C.c_int), ("ncomp_up", C.c_int), ("ncomp_down", C.c_int),
                ("comp_up", C.c_int * 27), ("comp_down", C.c_int * 27), ("nsites", C.c_longlong), ("chunk", C.c_longlong),
                ("src_up", C.c_longlong), ("dst_up", C.c_longlong), ("src_down", C.c_longlong),
                ("dst_down", C.c_longlong), ("peer_shift", C.c_longlong)]


STEP_PHI, STEP_UX, STEP_F = 0, 1, 2


def step_plan(nlocal, nhalo, cart_size, cart_rank, what, nvel=19):
    """lb200_step_plan: what lb200_step moves between x-slabs in halo-free mode (works without a GPU)."""
    lib = load_library()
    o = Options()
    o.nlocal[:] = nlocal
    o.nhalo = nhalo
    o.periodic[:] = (1, 1, 1)
    o.nvel, o.ndist, o.halo_scheme = nvel, 1, HALO_FULL
    o.cart_size, o.cart_rank = cart_size, cart_rank
    p = StepPlan()
    rc = lib.lb200_step_plan(C.byref(o), what, C.byref(p))
    if rc != 0:
        raise Lb200Error(lib.lb200_last_error().decode())
    return p


def slab_plan(nlocal, nhalo, periodic, cart_size, cart_rank, ncomp, depth):
    """lb200_slab_plan: the x-slab exchange plan (pure host arithmetic; works without a GPU)."""
    lib = load_library()
    o = Options()
    o.nlocal[:] = nlocal
    o.nhalo = nhalo
    o.periodic[:] = periodic
    o.nvel, o.ndist, o.halo_scheme = 19, 1, HALO_FULL
    o.cart_size, o.cart_rank = cart_size, cart_rank
    p = SlabPlan()
    rc = lib.lb200_slab_plan(C.byref(o), ncomp, depth, C.byref(p))
    if rc != 0:
        raise Lb200Error(lib.lb200_last_error().decode())
    return p


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libludwig_b200.so")


_lib = None


def load_library():
    """Load libludwig_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise Lb200Error(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                         " (make -C ludwig_b200/csrc)")
    lib = C.CDLL(path)
    lib.lb200_last_error.restype = C.c_char_p
    lib.lb200_create.argtypes = [C.POINTER(Options), C.POINTER(C.c_void_p)]
    lib.lb200_free.argtypes = [C.c_void_p]
    lib.lb200_nsites.argtypes = [C.c_void_p]
    lib.lb200_nsites_lb.argtypes = [C.c_void_p]
    lib.lb200_physics_control_time_set.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.lb200_physics_control_timestep.argtypes = [C.c_void_p]
    lib.lb200_le_plane_location.argtypes = [C.POINTER(Options), C.c_int]
    lib.lb200_le_ic_to_buff.argtypes = [C.POINTER(Options), C.c_int, C.c_int]
    lib.lb200_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.lb200_memcpy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.lb200_memcpy_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.lb200_pth_stress_compute.argtypes = [C.c_void_p, C.POINTER(SymmParam)]
    for name in ("lb200_sync", "lb200_hydro_f_zero", "lb200_hydro_u_zero", "lb200_hydro_u_halo",
                 "lb200_phi_halo", "lb200_phi_grad_compute", "lb200_lb_halo", "lb200_lb_propagation",
                 "lb200_phi_grad_compute_d4", "lb200_pth_force_fluid_driver", "lb200_field_leesedwards",
                 "lb200_hydro_lees_edwards", "lb200_lb_le_apply_boundary_conditions", "lb200_q_halo", "lb200_q_grad_compute"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.lb200_lc_stress_compute.argtypes = [C.c_void_p, C.POINTER(LcParam)]
    lib.lb200_lc_force_calculation.argtypes = [C.c_void_p, C.POINTER(LcParam)]
    lib.lb200_beris_edw_update.argtypes = [C.c_void_p, C.POINTER(LcParam)]
    lib.lb200_step_lc.argtypes = [C.c_void_p, C.POINTER(CollideParam), C.POINTER(LcParam), C.c_int]
    lib.lb200_phi_force_calculation.argtypes = [C.c_void_p, C.POINTER(SymmParam)]
    lib.lb200_phi_cahn_hilliard.argtypes = [C.c_void_p, C.POINTER(SymmParam)]
    lib.lb200_phi_conserve_sum.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.lb200_phi_init_sum_set.argtypes = [C.c_void_p, C.c_double]
    lib.lb200_lb_collide.argtypes = [C.c_void_p, C.POINTER(CollideParam)]
    lib.lb200_step.argtypes = [C.c_void_p, C.POINTER(CollideParam), C.POINTER(SymmParam), C.c_int]
    lib.lb200_lb_collision_binary.argtypes = [C.c_void_p, C.POINTER(CollideParam), C.POINTER(SymmParam)]
    lib.lb200_phi_lb_to_field.argtypes = [C.c_void_p]
    lib.lb200_phi_lb_from_field.argtypes = [C.c_void_p]
    lib.lb200_launch_count.argtypes = [C.c_void_p]
    lib.lb200_set_knob.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.lb200_exchange_mode.argtypes = [C.c_void_p]
    lib.lb200_pipe_state.argtypes = [C.c_void_p, C.POINTER(C.c_int * 2)]
    lib.lb200_launch_count.restype = C.c_longlong
    lib.lb200_stream.argtypes = [C.c_void_p]
    lib.lb200_stream.restype = C.c_void_p
    lib.lb200_profile.argtypes = [C.c_void_p, C.c_int]
    lib.lb200_profile_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.lb200_slab_plan.argtypes = [C.POINTER(Options), C.c_int, C.c_int, C.POINTER(SlabPlan)]
    lib.lb200_step_plan.argtypes = [C.POINTER(Options), C.c_int, C.POINTER(StepPlan)]
    lib.lb200_attach_nccl.argtypes = [C.c_void_p, C.c_void_p]
    lib.lb200_nccl_unique_id.argtypes = [C.c_void_p]
    lib.lb200_nccl_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.lb200_nccl_comm_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


class Lb200:
    """One device-resident lattice (the lb_t / hydro_t / field_t / field_grad_t / map_t device sides)."""

    def __init__(self, nlocal, nhalo=1, periodic=(1, 1, 1), nvel=19, ndist=1, have_phi=False,
                 halo_scheme=HALO_FULL, math=MATH_FAST, device=-1, cart_size=1, cart_rank=0,
                 le_nplanes=0, le_uy=0.0, le_nt0=0, have_q=False):
        self.lib = load_library()
        o = Options()
        o.nlocal[:] = nlocal
        o.nhalo = nhalo
        o.periodic[:] = periodic
        o.nvel, o.ndist, o.have_phi = nvel, ndist, int(bool(have_phi))
        o.halo_scheme, o.math, o.device = halo_scheme, math, device
        o.cart_size, o.cart_rank = cart_size, cart_rank
        o.le_nplanes, o.le_uy, o.le_nt0 = le_nplanes, le_uy, le_nt0
        o.have_q = int(bool(have_q))
        self.options = o
        self.h = C.c_void_p()
        self._check(self.lib.lb200_create(C.byref(o), C.byref(self.h)))
        self.nlocal = tuple(nlocal)
        self.nhalo = nhalo
        self.nall = tuple(n + 2 * nhalo for n in nlocal)
        self.nsites = self.lib.lb200_nsites(self.h)          # hydro / field arrays (with Lees-Edwards buffer planes)
        self.nsites_lb = self.lib.lb200_nsites_lb(self.h)    # F, MAP
        self.nvel, self.ndist = nvel, ndist
        self._nccl = None

    def _check(self, rc):
        if rc != 0:
            raise Lb200Error(f"lb200 error {rc}: {self.lib.lb200_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.lb200_free(self.h)
            self.h = None
        if getattr(self, "_nccl", None):
            self.lib.lb200_nccl_comm_destroy(self._nccl)
            self._nccl = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data movement -------------------------------------------------------------------
    def ncomp(self, array):
        return self.nvel * self.ndist if array == F else _NCOMP[array]

    def host_nsites(self, array):
        return self.nsites_lb if array in (F, MAP) else self.nsites

    def put(self, array, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        ns = self.host_nsites(array)
        assert host.size == self.ncomp(array) * ns, (host.shape, self.ncomp(array), ns)
        self._check(self.lib.lb200_memcpy(self.h, array, host.ctypes.data, H2D))

    def get(self, array):
        out = np.empty((self.ncomp(array), self.host_nsites(array)), dtype=np.float64)
        self._check(self.lib.lb200_memcpy(self.h, array, out.ctypes.data, D2H))
        return out

    def memcpy_async(self, array, host_ptr, kind):
        self._check(self.lib.lb200_memcpy_async(self.h, array, host_ptr, kind))

    def device_ptr(self, array):
        p = C.c_void_p()
        self._check(self.lib.lb200_device_ptr(self.h, array, C.byref(p)))
        return p.value

    def interior(self, a):
        h = self.nhalo
        v = a.reshape((a.shape[0], -1) + self.nall[1:])        # x extent: nall[0] (+ Lees-Edwards buffer planes)
        return v[:, h:h + self.nlocal[0], h:h + self.nlocal[1], h:h + self.nlocal[2]]

    # ---- operators (names follow the reference entry points) ------------------------------------
    def sync(self):
        self._check(self.lib.lb200_sync(self.h))

    def hydro_f_zero(self):
        self._check(self.lib.lb200_hydro_f_zero(self.h))

    def hydro_u_zero(self):
        self._check(self.lib.lb200_hydro_u_zero(self.h))

    def hydro_u_halo(self):
        self._check(self.lib.lb200_hydro_u_halo(self.h))

    def phi_halo(self):
        self._check(self.lib.lb200_phi_halo(self.h))

    def phi_grad_compute(self):
        self._check(self.lib.lb200_phi_grad_compute(self.h))

    # liquid crystal
    def q_halo(self):
        self._check(self.lib.lb200_q_halo(self.h))

    def q_grad_compute(self):
        self._check(self.lib.lb200_q_grad_compute(self.h))

    def lc_stress_compute(self, lc):
        self._check(self.lib.lb200_lc_stress_compute(self.h, C.byref(lc)))

    def lc_force_calculation(self, lc):
        self._check(self.lib.lb200_lc_force_calculation(self.h, C.byref(lc)))

    def beris_edw_update(self, lc):
        self._check(self.lib.lb200_beris_edw_update(self.h, C.byref(lc)))

    def step_lc(self, cp, lc, nsteps=1):
        self._check(self.lib.lb200_step_lc(self.h, C.byref(cp), C.byref(lc), nsteps))

    def step_lc_api(self, cp, lc, nsteps=1):
        """The liquid-crystal step through the individual entry points, reference driver order (src/ludwig.c:528-860)."""
        for _ in range(nsteps):
            self.hydro_f_zero()
            self.q_halo()
            self.q_grad_compute()
            self.lc_stress_compute(lc)
            self.pth_force_fluid_driver()
            self.hydro_u_halo()
            self.beris_edw_update(lc)
            self.hydro_u_zero()
            self.lb_collide(cp)
            self.lb_halo()
            self.lb_propagation()

    # Lees-Edwards planes
    def physics_control_time_set(self, t_start, t_current):
        self._check(self.lib.lb200_physics_control_time_set(self.h, t_start, t_current))

    def physics_control_timestep(self):
        return self.lib.lb200_physics_control_timestep(self.h)

    def field_leesedwards(self):
        self._check(self.lib.lb200_field_leesedwards(self.h))

    def hydro_lees_edwards(self):
        self._check(self.lib.lb200_hydro_lees_edwards(self.h))

    def lb_le_apply_boundary_conditions(self):
        self._check(self.lib.lb200_lb_le_apply_boundary_conditions(self.h))

    def phi_grad_compute_d4(self):
        self._check(self.lib.lb200_phi_grad_compute_d4(self.h))

    def pth_stress_compute(self, sp):
        self._check(self.lib.lb200_pth_stress_compute(self.h, C.byref(sp)))

    def pth_force_fluid_driver(self):
        self._check(self.lib.lb200_pth_force_fluid_driver(self.h))

    def phi_force_calculation(self, sp):
        self._check(self.lib.lb200_phi_force_calculation(self.h, C.byref(sp)))

    def phi_conserve_sum(self):
        """lb200_phi_conserve_sum: the compensated sum of phi over the fluid sites of the whole lattice (collective over
        the slabs); kept as the value cahn_hilliard_options_conserve 2 restores."""
        v = C.c_double()
        self._check(self.lib.lb200_phi_conserve_sum(self.h, C.byref(v)))
        return v.value

    def phi_init_sum_set(self, v):
        self._check(self.lib.lb200_phi_init_sum_set(self.h, C.c_double(v)))

    def phi_cahn_hilliard(self, sp):
        self._check(self.lib.lb200_phi_cahn_hilliard(self.h, C.byref(sp)))

    def lb_collide(self, cp):
        self._check(self.lib.lb200_lb_collide(self.h, C.byref(cp)))

    def lb_halo(self):
        self._check(self.lib.lb200_lb_halo(self.h))

    def lb_propagation(self):
        self._check(self.lib.lb200_lb_propagation(self.h))

    def step(self, cp, sp=None, nsteps=1):
        self._check(self.lib.lb200_step(self.h, C.byref(cp), C.byref(sp) if sp is not None else None, nsteps))

    def set_knob(self, knob, value):
        """lb200_set_knob: KNOB_WRAP (halo-free steps) / KNOB_PHI_SECTOR (one-sweep phi sector)."""
        self._check(self.lib.lb200_set_knob(self.h, knob, int(value)))

    def lb_collision_binary(self, cp, sp):
        self._check(self.lib.lb200_lb_collision_binary(self.h, C.byref(cp), C.byref(sp)))

    def phi_lb_to_field(self):
        self._check(self.lib.lb200_phi_lb_to_field(self.h))

    def phi_lb_from_field(self):
        self._check(self.lib.lb200_phi_lb_from_field(self.h))

    def step_api(self, cp, sp=None, nsteps=1):
        """The same time step through the individual reference-named entry points, in the
        reference driver's order (src/ludwig.c:528-860)."""
        le = self.options.le_nplanes > 0
        for _ in range(nsteps):
            if le:
                # physics_control_next_step: the plane displacement is a function of the step counter
                self.physics_control_time_set(0, self.physics_control_timestep() + 1)
            self.hydro_f_zero()
            if self.ndist == 2:
                self.phi_lb_to_field()
                self.phi_halo()
                self.phi_grad_compute()
                self.hydro_u_zero()
                self.lb_collision_binary(cp, sp)
                self.lb_halo()
                self.lb_propagation()
                continue
            if sp is not None:
                self.phi_halo()
                self.phi_grad_compute()
                self.phi_force_calculation(sp)
                self.phi_cahn_hilliard(sp)
            self.hydro_u_zero()
            self.lb_collide(cp)
            if le:
                self.lb_le_apply_boundary_conditions()      # src/ludwig.c:817-819
            self.lb_halo()
            self.lb_propagation()

    KCLASSES = ("collide", "propagate", "halo", "grad", "force_ch", "phi_sector", "le", "lc_stress", "lc_be", "step_fused")

    def profile(self, on=True):
        self._check(self.lib.lb200_profile(self.h, int(on)))

    def profile_get(self):
        """{class: (total_ms, launches)} accumulated since profile(True)."""
        out = {}
        for i, name in enumerate(self.KCLASSES):
            t, n = C.c_double(), C.c_int()
            self._check(self.lib.lb200_profile_get(self.h, i, C.byref(t), C.byref(n)))
            out[name] = (t.value, n.value)
        return out

    def exchange_mode(self):
        """0 single GPU, 1 NCCL send/recv, 2 NVLink peer stores from inside the kernels."""
        return int(self.lib.lb200_exchange_mode(self.h))

    def pipe_state(self):
        """(state, (phi-sector SMs, collision SMs)): 0 not used, 1 green-context SM partitions, 2 priority streams."""
        sms = (C.c_int * 2)()
        st = int(self.lib.lb200_pipe_state(self.h, C.byref(sms)))
        return st, (int(sms[0]), int(sms[1]))

    def launch_count(self):
        return int(self.lib.lb200_launch_count(self.h))

    def stream(self):
        return self.lib.lb200_stream(self.h)

    # ---- multi-GPU ------------------------------------------------------------------------------
    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        self._check(self.lib.lb200_nccl_unique_id(buf))
        return bytes(buf)

    def nccl_init(self, unique_id, nranks, rank):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        comm = C.c_void_p()
        self._check(self.lib.lb200_nccl_comm_create(buf, nranks, rank, C.byref(comm)))
        self._nccl = comm
        self._check(self.lib.lb200_attach_nccl(self.h, comm))
