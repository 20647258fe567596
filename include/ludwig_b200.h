/*
 * ludwig_b200.h -- C-ABI of libludwig_b200.so: a Blackwell (sm_100a) implementation of the
 * per-timestep lattice-Boltzmann hot path of Ludwig (ludwig-cf/ludwig v0.23.0).
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes.  Each entry point replaces the
 * device side of one reference host function (cited as reference file:line, paths relative to
 * the reference root).  The reference-side binding (what a Ludwig maintainer adds to collision.c,
 * propagation.c, ... to route the hot path through this library) is shown in INTEGRATION.md, and
 * ludwig_b200/host/ holds a C host layer with the reference's own function names on top of this ABI.
 *
 * Host arrays crossing this boundary use the "canonical" structure-of-arrays layout on the
 * reference's allocated lattice (identical to the reference's -DADDR_SOA host layout,
 * src/memory.h:182-195):
 *     nall[a] = nlocal[a] + 2*nhalo;  nsites = nall[X]*nall[Y]*nall[Z]
 *     index(ic,jc,kc) = ((ic+nhalo-1)*nall[Y] + (jc+nhalo-1))*nall[Z] + (kc+nhalo-1)   src/coords.c:617-631
 *     scalar a[index];  vector a[ia*nsites + index];  f[(n*nvel + p)*nsites + index]
 *
 * All functions return 0 on success and a negative LB200_E* code on failure (and record a message
 * retrievable with lb200_last_error()); like the reference they are synchronous unless stated.
 * There is NO CPU fallback: without a CUDA device lb200_create() fails with LB200_ENODEVICE.
 */
#ifndef LUDWIG_B200_H
#define LUDWIG_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lb200_s lb200_t;          /* opaque: geometry + every device-resident lattice array */

enum lb200_error {
  LB200_OK = 0,
  LB200_EINVAL = -1,        /* bad argument / unsupported option */
  LB200_ENODEVICE = -2,     /* no CUDA device (there is no CPU fallback) */
  LB200_ECUDA = -3,         /* CUDA runtime error (message in lb200_last_error) */
  LB200_ENOMEM = -4,
  LB200_ESTATE = -5,        /* operation needs an array this context does not hold (e.g. no phi) */
  LB200_ECOMM = -6          /* NCCL / peer communication error */
};

/* lb_relaxation_enum_t: src/lb_data_options.h:21-24 */
enum lb200_relaxation {LB200_RELAXATION_M10 = 0, LB200_RELAXATION_BGK = 1, LB200_RELAXATION_TRT = 2};
/* lb_halo_enum_t: src/lb_data_options.h:26-29 */
enum lb200_halo_scheme {LB200_HALO_FULL = 1, LB200_HALO_REDUCED = 2};
/* tdpMemcpyKind: target/target.h */
enum lb200_memcpy_kind {LB200_HOST_TO_DEVICE = 1, LB200_DEVICE_TO_HOST = 2};
/* arithmetic mode of the kernels */
enum lb200_math {
  LB200_MATH_FAST = 0,      /* FMA contraction allowed (default): within 1e-12 relative of the reference */
  LB200_MATH_STRICT = 1     /* no contraction, reference operation order: bit-identical to the reference
                               built -ffp-contract=off */
};

/* lattice arrays held by a context */
enum lb200_array {
  LB200_F = 0,              /* lb->f              ndist*nvel x nsites   src/lb_data.h:123 */
  LB200_PHI = 1,            /* field "phi"        1 x nsites            src/field.h:68-85 */
  LB200_U = 2,              /* hydro->u           3 x nsites            src/hydro.h:32-47 */
  LB200_RHO = 3,            /* hydro->rho         1 x nsites */
  LB200_FORCE = 4,          /* hydro->force       3 x nsites */
  LB200_GRAD = 5,           /* field_grad->grad   3 x nsites            src/field_grad.h:24-42 */
  LB200_DELSQ = 6,          /* field_grad->delsq  1 x nsites */
  LB200_MAP = 7,            /* map->status        1 x nsites, passed as double, 0 = MAP_FLUID  src/map.h:26-48 */
  LB200_GRAD_DELSQ = 8,     /* field_grad->grad_delsq  3 x nsites  (after lb200_phi_grad_compute_d4) */
  LB200_DELSQ_DELSQ = 9,    /* field_grad->delsq_delsq 1 x nsites */
  LB200_STR = 10,           /* pth->str           9 x nsites, component ia*3 + ib (after lb200_pth_stress_compute) */
  LB200_Q = 11,             /* field "q"          5 x nsites: Q_xx, Q_xy, Q_xz, Q_yy, Q_yz   src/field.h, NQAB */
  LB200_QGRAD = 12,         /* field_grad(q)->grad   15 x nsites, component n*3 + ia (after lb200_q_grad_compute) */
  LB200_QDELSQ = 13         /* field_grad(q)->delsq   5 x nsites */
};

typedef struct lb200_options_s {
  int nlocal[3];            /* local lattice extent (cs_nlocal), src/coords.c:211-215 */
  int nhalo;                /* cs_nhalo: 1 (single fluid) or 2 (binary fluid FD route, src/ludwig.c:1198) */
  int periodic[3];          /* cs periodicity of the GLOBAL system */
  int nvel;                 /* 15, 19 or 27 (-D_D3Q15_/-D_D3Q19_/-D_D3Q27_, src/lb_data.h:30-42) */
  int ndist;                /* 1, or 2 = symmetric_lb (needs have_phi; nhalo >= 1) */
  int have_phi;             /* allocate phi, grad, delsq (free_energy symmetric) */
  int halo_scheme;          /* enum lb200_halo_scheme */
  int math;                 /* enum lb200_math */
  int device;               /* CUDA device ordinal, -1 = current */
  /* x-slab domain decomposition across the GPUs of one box (cs "grid Px_1_1", src/coords.c:151-201) */
  int cart_size;            /* number of slabs (1 = single GPU) */
  int cart_rank;            /* this slab */
  /* Lees-Edwards sliding periodic planes, lees_edw_options_t (src/lees_edwards_options.h:31-38): steady shear.
   * N_LE_plane planes in the GLOBAL system, equally spaced in x at (p + 1/2) Lx / nplanes (src/leesedwards.c:
   * 245-257, 615-634); must divide evenly over the slabs and lie further than nhalo from a slab boundary
   * (lees_edw_checks, :433-470).  With planes, every array except LB200_F and LB200_MAP carries
   * 2*nhalo*(planes in this slab) buffer x-planes after the high x halo, as the reference's
   * (lees_edw_nsites, :485-495): lb200_nsites() counts them, lb200_nsites_lb() does not. */
  int le_nplanes;           /* 0 = none */
  double le_uy;             /* LE_plane_vel */
  int le_nt0;               /* reference time step (lees_edw_options_t.nt0, usually 0) */
  /* free_energy lc_blue_phase (src/ludwig.c:1598-1666): the tensor order parameter q (5 components, nhalo >= 2).
   * Exclusive with have_phi; no Lees-Edwards planes in this round.  x-slabs (cart_size > 1) exchange the x-planes
   * of q (depth nhalo), u and f over NCCL. */
  int have_q;
} lb200_options_t;

/* lb_collide_param_t / collide_param_t as seen by the collision: src/lb_data.h:59-76,
 * src/collision.c:60-75, 1906-1958 (re-read from `physics` on every lb_collide call) */
typedef struct lb200_collide_param_s {
  int nrelax;               /* enum lb200_relaxation */
  double rho0;              /* physics rho0 */
  double eta_shear;         /* physics eta_shear */
  double eta_bulk;          /* physics eta_bulk */
  double force_global[3];   /* constant + pulsatile body force, src/collision.c:1942-1945 */
} lb200_collide_param_t;

/* fe_symm_param_t + Cahn-Hilliard controls: src/symmetric.h:41-47, src/phi_cahn_hilliard.c:298-404,
 * 1329-1397, src/advection.c:74 */
typedef struct lb200_symm_param_s {
  double a, b, kappa;       /* symmetric free energy A, B, kappa */
  double mobility;          /* physics mobility */
  double gradmu[3];         /* physics grad_mu (external chemical potential gradient) */
  int adv_order;            /* fd_advection_scheme_order 1-4 (src/advection.c:453-468; order 5 needs a 3-deep halo: not built) */
  int conserve;             /* cahn_hilliard_options_conserve (phi_ch_info_t.conserve, src/phi_cahn_hilliard.h:31-36):
                             * 0 = plain forward step; 1 = compensated (Kahan) sum per site, phi_ch_update_conserve,
                             * src/phi_cahn_hilliard.c:1059-1094, 1181-1215 -- the compensation field lives in the context
                             * like pch->csum; 2 = global subtraction after the forward step (lb200_phi_conserve_sum below) */
  int force_method;         /* fe_force_method (src/phi_force.c:99-133): 0 = stress_divergence (default), 1 = phi_gradmu
                             * (force = -phi grad mu - phi grad_mu_ext, src/phi_grad_mu.c).  With Lees-Edwards planes the
                             * reference uses its flux method whatever this says, and so does the library */
} lb200_symm_param_t;

/* fe_lc_param_t + beris_edw_param_t as the liquid-crystal kernels see them (src/blue_phase.h:52-75,
 * src/blue_phase_beris_edwards.h:30-37); static redshift, no noise */
typedef struct lb200_lc_param_s {
  double a0, q0, gamma;     /* lc_a0, lc_q0, lc_gamma */
  double kappa0, kappa1;    /* lc_kappa0, lc_kappa1 (as the reference, its vectorised molecular field and free-energy
                             * density use kappa0 for both: src/blue_phase.c:1934, 2119-2120) */
  double xi;                /* lc_xi: flow-aligning parameter */
  double Gamma;             /* lc_Gamma: rotational diffusion constant */
  double epsilon;           /* lc_dielectric_anisotropy / (12 pi), as stored by fe_lc_param_set (src/blue_phase.c:249-252) */
  double e0[3];             /* electric_e0 */
  int adv_order;            /* fd_advection_scheme_order 1-4 (src/advection.c:453-468; order 5 needs a 3-deep halo: not built) */
  int is_active;            /* lc_activity: the active stress zeta0 d_ab - zeta1 Q_ab is added to the stress
                             * (fe_lc_compute_stress_active, src/blue_phase.c:934-972; fe_lc_stress_v :1825-1845) */
  double zeta0, zeta1;      /* lc_active_zeta0, lc_active_zeta1 */
  double zeta2;             /* lc_active_zeta2: the polarisation-gradient term; must be 0 (LB200_EINVAL otherwise) */
  double redshift;          /* lc_init_redshift (fe_lc_param_t.redshift; 0 is read as 1): q0 / redshift, kappa redshift^2 in the
                             * molecular field, free-energy density and stress (src/blue_phase.c:2117-2120, 1933-1934, 2300-2302).
                             * Static: lc_redshift_update (fe_lc_redshift_compute) is not built */
} lb200_lc_param_t;

const char * lb200_last_error(void);
int lb200_version(void);

/* lb_data_create + hydro_create + field_create + field_grad_create + map_create device sides:
 * src/lb_data.c:67-241, src/hydro.c:52-121, src/field.c:59-136, src/field_grad.c:35-164 */
int lb200_create(const lb200_options_t * options, lb200_t ** ctx);
/* lb_free / hydro_free / field_free: src/lb_data.c:251-296 */
int lb200_free(lb200_t * ctx);

int lb200_nsites(const lb200_t * ctx);        /* hydro / field arrays (with the Lees-Edwards buffer planes, if any) */
int lb200_nsites_lb(const lb200_t * ctx);     /* LB200_F, LB200_MAP: cs_nsites */
/* device pointer of an array in its current (device) layout; for zero-copy interop */
int lb200_device_ptr(lb200_t * ctx, int array, void ** ptr);

/* lb_memcpy / field_memcpy / hydro_memcpy / map_memcpy: src/lb_data.c:529-583, src/field.c:224-288.
 * `host` is a canonical array (see top); synchronous. */
int lb200_memcpy(lb200_t * ctx, int array, double * host, int kind);
/* same, but asynchronous on the context's stream; `host` should be pinned */
int lb200_memcpy_async(lb200_t * ctx, int array, double * host, int kind);
int lb200_sync(lb200_t * ctx);

/* hydro_f_zero / hydro_u_zero: src/hydro.c:217-263 (the value is 0 in every caller) */
int lb200_hydro_f_zero(lb200_t * ctx);
int lb200_hydro_u_zero(lb200_t * ctx);
/* hydro_u_halo: src/hydro.c:185-191; field_halo(phi): src/field.c:371-404 */
int lb200_hydro_u_halo(lb200_t * ctx);
int lb200_phi_halo(lb200_t * ctx);
/* field_grad_compute with d2 = grad_3d_27pt_fluid_d2: src/field_grad.c:319-340,
 * src/gradient_3d_27pt_fluid.c:76-99, 219-363 */
int lb200_phi_grad_compute(lb200_t * ctx);
/* field_grad level 4, d4 = grad_3d_27pt_fluid_d4 (src/gradient_3d_27pt_fluid.c:112-134): the same operator
 * applied to delsq on [1-(nhalo-2), N+(nhalo-2)]^3 -> LB200_GRAD_DELSQ, LB200_DELSQ_DELSQ */
int lb200_phi_grad_compute_d4(lb200_t * ctx);
/* the two halves of phi_force_calculation as separate operators: pth_stress_compute
 * (src/phi_force_stress.c:171-284, P stored in LB200_STR) and pth_force_fluid_driver
 * (src/phi_force_colloid.c:274-465, force -= div P from the stored stress) */
int lb200_pth_stress_compute(lb200_t * ctx, const lb200_symm_param_t * sp);
int lb200_pth_force_fluid_driver(lb200_t * ctx);
/* phi_force_calculation, stress-divergence method, fluid only: src/phi_force.c:74-137,
 * src/phi_force_stress.c:171-284, src/phi_force_colloid.c:274-465 */
int lb200_phi_force_calculation(lb200_t * ctx, const lb200_symm_param_t * sp);
/* phi_cahn_hilliard (no noise; conserve = 0 or 1): src/phi_cahn_hilliard.c:213-288 */
int lb200_phi_cahn_hilliard(lb200_t * ctx, const lb200_symm_param_t * sp);
/* cahn_hilliard_options_conserve 2 (PHI_CONSERVE_GLOBAL_SUBTRACT, src/phi_cahn_hilliard.c:1102-1169): after every forward step
 * (sum_fluid phi - sum0)/nfluid is subtracted at every fluid site -- the one place of the path with a global reduction
 * (x-slabs: NCCL all-gather of one pair per GPU, added in rank order in compensated arithmetic, so the result does not
 * depend on the GPU count beyond the last bit of the sum).  sum0 is phi->field_init_sum of the reference, which its
 * statistics code computes at time 0 (cahn_hilliard_stats_time0, src/cahn_hilliard_stats.c:58-76):
 *   lb200_phi_conserve_sum   computes the compensated sum of the CURRENT phi over the fluid sites of the whole lattice
 *                            (collective over the slabs), returns it and keeps it as sum0;
 *   lb200_phi_init_sum_set   sets sum0 to a value the caller has (the binding passes phi->field_init_sum). */
int lb200_phi_conserve_sum(lb200_t * ctx, double * sum);
int lb200_phi_init_sum_set(lb200_t * ctx, double sum0);
/* lb_collide (ndist = 1): src/collision.c:143-232, 253-593 */
int lb200_lb_collide(lb200_t * ctx, const lb200_collide_param_t * cp);
/* symmetric_lb (`free_energy symmetric_lb`, options.ndist = 2: array LB200_F holds [density | order parameter]
 * distributions).  lb_collide with ndist == 2 -> lb_collision_binary: src/collision.c:157-159, 604-1013,
 * 2856-3135; phi_lb_to_field / phi_lb_from_field: src/phi_lb_coupler.c:39-137.  lb200_step on such a context
 * runs the symmetric_lb time step (src/ludwig.c:551-571, 683-685, 802-860). */
int lb200_lb_collision_binary(lb200_t * ctx, const lb200_collide_param_t * cp, const lb200_symm_param_t * sp);
int lb200_phi_lb_to_field(lb200_t * ctx);
int lb200_phi_lb_from_field(lb200_t * ctx);

/* ---- Lees-Edwards planes (options.le_nplanes > 0) ----------------------------------------------------------
 * The plane displacement is uy * time; like the reference (physics_control_next_step / _timestep / _time,
 * src/physics.c:600-647) the context keeps t_start and the step counter t_current: lb200_step advances
 * t_current by one at the start of every step; a host driving the individual entry points sets it.
 * With planes present
 *   lb200_phi_grad_compute       = field_leesedwards + d2 + grad_3d_27pt_fluid_le   (src/field_grad.c:319-340)
 *   lb200_phi_force_calculation  = phi_force_flux: flux form + per-plane correction  (src/phi_force.c:91-97, 289-673)
 *   lb200_phi_cahn_hilliard      = ... hydro_lees_edwards ... phi_ch_le_fix_fluxes    (src/phi_cahn_hilliard.c:213-288)
 *   lb200_step                   = the reference step with lb_data_apply_le_boundary_conditions after the collision
 * all on the device (the reference's GPU build runs them on the host with whole-array copies). */
int lb200_physics_control_time_set(lb200_t * ctx, int t_start, int t_current);
int lb200_physics_control_timestep(const lb200_t * ctx);
/* field_leesedwards(phi): src/field.c:418-510;  hydro_lees_edwards: src/hydro.c:350-440 */
int lb200_field_leesedwards(lb200_t * ctx);
int lb200_hydro_lees_edwards(lb200_t * ctx);
/* lb_data_apply_le_boundary_conditions: src/model_le.c:78-180 */
int lb200_lb_le_apply_boundary_conditions(lb200_t * ctx);
/* lees_edw_plane_location (local x of plane np of this slab) and lees_edw_ic_to_buff: src/leesedwards.c:615-634,
 * 1030-1065; pure host arithmetic on the options */
int lb200_le_plane_location(const lb200_options_t * options, int np);
int lb200_le_ic_to_buff(const lb200_options_t * options, int ic, int di);

/* ---- liquid crystal (options.have_q) ------------------------------------------------------------------------
 * field_halo(q): src/field.c:371-404;  field_grad_compute(q_grad) with d2 = grad_3d_7pt_fluid_d2
 * (src/gradient_3d_7pt_fluid.c:76-99, 231-300) -> LB200_QGRAD, LB200_QDELSQ (the stress and the Beris-Edwards kernels
 * below do not need these arrays: they rebuild the gradients from Q in registers) */
int lb200_q_halo(lb200_t * ctx);
int lb200_q_grad_compute(lb200_t * ctx);
/* pth_stress_compute with fe_lc_stress_v (src/phi_force_stress.c:171-284, src/blue_phase.c:1737-1800, 2279-2775)
 * -> LB200_STR; follow with lb200_pth_force_fluid_driver.  lb200_lc_force_calculation = phi_force_calculation
 * for this free energy (both, src/phi_force.c:100-110) */
int lb200_lc_stress_compute(lb200_t * ctx, const lb200_lc_param_t * lc);
int lb200_lc_force_calculation(lb200_t * ctx, const lb200_lc_param_t * lc);
/* beris_edw_update (hydro present, no colloids, no noise): advection_x + beris_edw_h_driver + beris_edw_update_driver,
 * src/blue_phase_beris_edwards.c:266-296, 538-850, 942-985.  The caller has done hydro_u_halo (src/ludwig.c:771-773). */
int lb200_beris_edw_update(lb200_t * ctx, const lb200_lc_param_t * lc);
/* nsteps whole liquid-crystal time steps (src/ludwig.c:528-860 with ludwig->q): hydro_f_zero; field_halo(q);
 * field_grad_compute; phi_force_calculation; hydro_u_halo; beris_edw_update; hydro_u_zero; lb_collide; lb_halo;
 * lb_propagation -- three sweeps per step (stress; force + Beris-Edwards; pull-stream + collide), halo-free on
 * periodic lattices.  Asynchronous like lb200_step. */
int lb200_step_lc(lb200_t * ctx, const lb200_collide_param_t * cp, const lb200_lc_param_t * lc, int nsteps);

/* lb_halo: src/lb_data.c:754-762, 1124-1477 */
int lb200_lb_halo(lb200_t * ctx);
/* lb_propagation: src/propagation.c:43-95, 153-240 */
int lb200_lb_propagation(lb200_t * ctx);

/* nsteps whole time steps in the reference driver's order (src/ludwig.c:528-860), fused:
 * pull-stream + collide in one sweep, stress + force + Cahn-Hilliard in one sweep, zeroing folded
 * into the producers.  sp == NULL (or no phi) = single fluid.  Asynchronous on the context stream;
 * call lb200_sync() (or any lb200_memcpy) to wait. */
int lb200_step(lb200_t * ctx, const lb200_collide_param_t * cp, const lb200_symm_param_t * sp,
	       int nsteps);

/* Execution knobs of lb200_step (results do not depend on them; the parity tests run both settings).
 * Defaults come from the environment (LB200_WRAP, LB200_PHI_SECTOR), else 1. */
enum lb200_knob {
  LB200_KNOB_WRAP = 1,        /* 1: halo-free time steps on periodic lattices (periodic images are read from
                               *    the interior; only the planes the kernels read cross NVLink)
                               * 0: the reference's step structure with three halo swaps (phi, u, f) */
  LB200_KNOB_PHI_SECTOR = 2,  /* 1: gradient + force + Cahn-Hilliard in one sweep (all-fluid lattices) */
  LB200_KNOB_PEER = 3,        /* 1: x-slab neighbours exchange planes by NVLink peer stores from inside the kernels
                               *    (cudaIpc-mapped arrays + one flag per kernel); 0: NCCL send/recv.  Must be set
                               *    identically on every rank.  Default LB200_PEER, else 1 */
  LB200_KNOB_PIPE = 4,        /* S >= 2: slab pipeline of the single-GPU binary-fluid step -- the lattice is cut into S
                               *    x-slabs and the two kernels of a step run CONCURRENTLY on disjoint SM partitions
                               *    (CUDA green contexts), the collision of slab s next to the phi sector of the slabs
                               *    after it (same operations on the same data: results unchanged).  0: off.
                               *    Default LB200_PIPE, else 0 */
  LB200_KNOB_PIPE_SMS = 5,    /* SMs provisioned for the phi-sector partition (rounded up to the device's partition
                               *    granularity, 8 on sm_100); the collision gets the rest.  Default LB200_PIPE_SMS, else 56.
                               *    Must be set before the first pipelined step */
  LB200_KNOB_F32 = 6,         /* 1: FP32 STORAGE of the D3Q19 distributions inside lb200_step (single GPU, halo-free path,
                               *    no planes): the two distribution arrays hold float(f_p - w_p) for the steps of one
                               *    call, arithmetic stays FP64, LB200_F is converted on entry and back on exit.  Not
                               *    the reference's arithmetic: each population is rounded once per step with relative
                               *    error <= 2^-24 of its deviation from the rest weight w_p (bound and measured
                               *    errors: DESIGN.md, tests/test_gpu_parity.py::test_f32_storage_error_bound).
                               *    208 instead of 360 bytes per site and step.  Default LB200_F32, else 0 */
  LB200_KNOB_FUSED = 8,       /* 1: the whole binary-fluid time step in ONE kernel (fast arithmetic mode, D3Q19, all-fluid,
                               *    halo-free path, advection order 1-3, no planes): the phi sector and the pull-stream +
                               *    collision of the same plane share a sweep, so the body force never goes through
                               *    memory (368 instead of 416 bytes per site and step).  Where it does not apply the
                               *    step runs the phi-sector kernel followed by the collision kernel.
                               *    Default LB200_FUSED, else 1 */
  LB200_KNOB_QGRAD_2D5 = 9,   /* 1: fd_gradient_calculation 2d_5pt_fluid for the Q tensor (grad_2d_5pt_fluid_d2,
                               *    src/gradient_2d_5pt_fluid.c:53-72): lattices with nlocal[Z] = 1.  0 (default): 3d_7pt_fluid */
  LB200_KNOB_GRAD_7PT = 7     /* NOT an execution knob: selects the finite-difference scheme of the scalar order parameter,
                               *    `fd_gradient_calculation`: 0 = 3d_27pt_fluid (default), 1 = 3d_7pt_fluid
                               *    (grad_3d_7pt_fluid_d2, src/gradient_3d_7pt_fluid.c:76-99, 231-300) for
                               *    lb200_phi_grad_compute and lb200_step (which then runs the gradient and the
                               *    force / Cahn-Hilliard kernels separately; no Lees-Edwards planes).  Default LB200_GRAD_7PT, else 0 */
};
int lb200_set_knob(lb200_t * ctx, int knob, int value);
/* slab pipeline of this context: 0 = not used yet, 1 = green contexts (sms[0] / sms[1] = SMs of the phi-sector /
 * collision partition), 2 = two priority streams sharing every SM (no partition API), -1 = unavailable */
int lb200_pipe_state(const lb200_t * ctx, int sms[2]);
/* how lb200_step exchanges x-planes on this context: 0 = single GPU, 1 = NCCL send/recv, 2 = peer stores
 * (meaningful after the first lb200_step, which sets the mapping up collectively) */
int lb200_exchange_mode(const lb200_t * ctx);

/* number of kernels this library has launched on this context since creation */
long long lb200_launch_count(const lb200_t * ctx);
/* the context's CUDA stream (cudaStream_t) for event timing by the caller */
void * lb200_stream(lb200_t * ctx);

/* per-kernel-class device timing (CUDA events on the context stream) for the roofline report */
enum lb200_kernel_class {
  LB200_K_COLLIDE = 0,      /* (pull-stream +) collision sweep */
  LB200_K_PROPAGATE = 1,    /* stand-alone propagation sweep */
  LB200_K_HALO = 2,         /* halo shells incl. x-plane exchange */
  LB200_K_GRAD = 3,         /* 27-point gradient */
  LB200_K_FORCE_CH = 4,     /* stress-divergence force and/or Cahn-Hilliard update */
  LB200_K_PHI_SECTOR = 5,   /* gradient + force + Cahn-Hilliard in one sweep (lb200_step, all-fluid) */
  LB200_K_LE = 6,           /* Lees-Edwards: buffer interpolation, plane patches, plane-crossing populations */
  LB200_K_LC_STRESS = 7,    /* liquid crystal: gradients + molecular field + stress */
  LB200_K_LC_BE = 8,        /* liquid crystal: force divergence + Beris-Edwards update */
  LB200_K_STEP_FUSED = 9,   /* phi sector + pull-stream + collision in one sweep (lb200_step, LB200_KNOB_FUSED) */
  LB200_KCLASS_MAX = 10
};
int lb200_profile(lb200_t * ctx, int on);      /* clears accumulated timings */
int lb200_profile_get(lb200_t * ctx, int kernel_class, double * total_ms, int * count);

/* The x-slab exchange plan of one halo swap (pure host arithmetic, usable without a device): which
 * ranks are the neighbours, which contiguous element ranges of each component leave this slab, and how
 * the received planes are laid out in the staging areas the halo-shell kernel reads.  All offsets and
 * counts are in doubles.  Replaces the send/recv region set-up of lb_halo_create / field_halo_create
 * (src/lb_data.c:1183-1210, src/field.c:1329-1355) for the x direction. */
typedef struct lb200_slab_plan_s {
  int left, right;          /* neighbour ranks (periodic wrap) */
  int has_lo, has_hi;       /* 0 at a non-periodic global boundary: nothing is exchanged there */
  long long nsites;         /* component stride in the lattice array (with the Lees-Edwards buffer planes) */
  long long chunk;          /* depth * nall[Y] * nall[Z]: one component's boundary planes, contiguous */
  long long off_lo;         /* first element of my planes i in [1, depth]        (sent to `left`) */
  long long off_hi;         /* first element of my planes i in [N-depth+1, N]    (sent to `right`) */
  long long halo_lo;        /* first element of my halo planes i in [1-depth, 0]     (filled from xlo) */
  long long halo_hi;        /* first element of my halo planes i in [N+1, N+depth]   (filled from xhi) */
  long long count;          /* ncomp * chunk: message size; staging layout is [comp][depth][y][z] */
} lb200_slab_plan_t;
int lb200_slab_plan(const lb200_options_t * options, int ncomp, int depth, lb200_slab_plan_t * plan);

/* What lb200_step moves between x-slabs in halo-free mode (pure host arithmetic): per array only the components
 * and planes the NEXT kernel reads across the slab boundary -- y/z images are read in-kernel from the interior.
 * `up` goes to the high neighbour (my top planes -> its low halo planes), `down` to the low neighbour.
 * With the peer-store exchange a kernel adds `peer_shift` (down) or subtracts it (up) to the index of a boundary
 * site to get the index of its image in the neighbour's array. */
enum lb200_step_array {LB200_STEP_PHI = 0, LB200_STEP_UX = 1, LB200_STEP_F = 2};
typedef struct lb200_step_plan_s {
  int left, right;          /* neighbour ranks */
  int depth;                /* planes per direction: nhalo for phi, 1 for u_x and f */
  int ncomp_up, ncomp_down; /* components travelling up / down */
  int comp_up[27];          /* f: populations with c_x = +1; phi, u_x: component 0 */
  int comp_down[27];        /* f: populations with c_x = -1 */
  long long nsites;         /* component stride */
  long long chunk;          /* depth * nall[Y] * nall[Z] */
  long long src_up;         /* first element of my planes [N-depth+1, N] */
  long long dst_up;         /* first element of the receiver's halo planes [1-depth, 0] */
  long long src_down;       /* first element of my planes [1, depth] */
  long long dst_down;       /* first element of the receiver's halo planes [N+1, N+depth] */
  long long peer_shift;     /* nlocal[X] * nall[Y] * nall[Z] */
} lb200_step_plan_t;
int lb200_step_plan(const lb200_options_t * options, int step_array, lb200_step_plan_t * plan);

/* multi-GPU: attach an NCCL communicator (ncclComm_t, one rank per slab, rank == cart_rank) used
 * for the x-direction halo planes; replaces MPI_Isend/Irecv of lb_halo_post/field_halo_post
 * (src/lb_data.c:1317-1420, src/field.c:1412-1531). */
int lb200_attach_nccl(lb200_t * ctx, void * nccl_comm);
/* helpers so that a host without nccl.h can create the communicator: unique id is 128 bytes */
int lb200_nccl_unique_id(void * id128);
int lb200_nccl_comm_create(const void * id128, int nranks, int rank, void ** nccl_comm);
int lb200_nccl_comm_destroy(void * nccl_comm);

#ifdef __cplusplus
}
#endif
#endif
