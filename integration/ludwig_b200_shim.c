/* ludwig_b200_shim.c -- the reference-side binding of INTEGRATION.md as real code.
 *
 * Compiled against the reference's OWN headers (lb_t, field_t, hydro_t, fe_t ... are the reference's structs, nothing
 * is re-declared here) and linked into the reference's OWN objects (src/ludwig.c, the run-time set-up, statistics, I/O,
 * the unit tests) with GNU ld's --wrap: every call those objects make to a hot-path entry point listed in
 * integration/Makefile arrives at the __wrap_ function below, which forwards it to the C-ABI of libludwig_b200.so
 * (include/ludwig_b200.h).  The reference's own definition stays reachable as __real_ and is used while an object
 * is not device backed (before its first host-to-device copy, or for things this library does not hold).
 *
 * Coherence protocol = the reference's own GPU-build protocol: host arrays are authoritative until
 * X_memcpy(obj, tdpMemcpyHostToDevice), the device until X_memcpy(obj, tdpMemcpyDeviceToHost); src/ludwig.c does the
 * former once before the time-step loop (:501-506) and the latter before every statistics / output call
 * (:871, 985, 2410-2442), tests/unit/*.c around every operation they check.  The reference is built with
 * -DADDR_SOA, the layout of all its GPU configurations, which is the layout of the C-ABI's host arrays.
 *
 * One device lattice (lb200_t) per coordinate system, created at the first host-to-device copy on it and freed
 * with the last object created on it.
 */

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pe.h"
#include "coords.h"
#include "physics.h"
#include "lb_data.h"
#include "collision.h"
#include "propagation.h"
#include "model_le.h"
#include "field.h"
#include "field_grad.h"
#include "gradient_3d_27pt_fluid.h"
#include "gradient_3d_7pt_fluid.h"
#include "gradient_2d_5pt_fluid.h"
#include "hydro.h"
#include "free_energy.h"
#include "symmetric.h"
#include "phi_force.h"
#include "phi_force_stress.h"
#include "phi_cahn_hilliard.h"
#include "advection.h"
#include "leesedwards.h"
#include "noise.h"
#include "wall.h"
#include "colloids.h"
#include "blue_phase.h"
#include "blue_phase_beris_edwards.h"
#include "runtime.h"
#include "wall_rt.h"
#include "colloids_rt.h"
#include "map.h"

#include "ludwig_b200.h"

/* LB200_SHIM_TRACE=1: one line per re-routed call on stderr */
static int trace_on(void) {
  static int on = -1;
  if (on < 0) on = (getenv("LB200_SHIM_TRACE") != NULL);
  return on;
}
#define TRACE(...) do { if (trace_on()) { fprintf(stderr, "[shim] " __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); } } while (0)

/* ---- one device lattice per coordinate system ------------------------------------------------------------- */

typedef struct {
  cs_t * cs;
  lb200_t * ctx;
  lees_edw_t * le;
  lb_t * lb;
  field_t * phi;          /* scalar order parameter (nf = 1, not hydro->rho) */
  field_t * q;            /* tensor order parameter (nf = 5) */
  field_t * u, * rho, * force;
  int nobj;               /* objects created on this cs and not yet freed */
  int has_phi, has_q;     /* what the device lattice was created with */
} slot_t;

#define NSLOT 64
static slot_t slots_[NSLOT];

static slot_t * slot_find(cs_t * cs, int create) {
  slot_t * empty = NULL;
  for (int i = 0; i < NSLOT; i++) {
    if (slots_[i].cs == cs) return &slots_[i];
    if (slots_[i].cs == NULL && empty == NULL) empty = &slots_[i];
  }
  if (!create || empty == NULL) return NULL;
  memset(empty, 0, sizeof(*empty));
  empty->cs = cs;
  return empty;
}

static void slot_release(slot_t * s) {
  s->nobj -= 1;
  if (s->nobj <= 0) {
    TRACE("last object on this cs freed%s", s->ctx ? ": lb200_free" : "");
    if (s->ctx) lb200_free(s->ctx);
    memset(s, 0, sizeof(*s));
  }
}

static void b200_check(pe_t * pe, int rc, const char * what) {
  TRACE("%s -> %d", what, rc);
  if (rc != 0) pe_fatal(pe, "libludwig_b200: %s: %s\n", what, lb200_last_error());
}

static int b200_kind(tdpMemcpyKind flag) {
  return (flag == tdpMemcpyHostToDevice) ? LB200_HOST_TO_DEVICE : LB200_DEVICE_TO_HOST;
}

/* the device lattice of this cs; created on demand (create != 0: a host-to-device copy is about to fill it) */
static lb200_t * b200_context(pe_t * pe, cs_t * cs, int create) {
  slot_t * s = slot_find(cs, 0);
  if (s == NULL) return NULL;
  if (s->ctx == NULL && create) {
    lb200_options_t o;
    /* velocity sets without device kernels (D2Q9) stay on the reference's own code: host == target, as in its CPU build */
    if (s->lb && s->lb->nvel != 15 && s->lb->nvel != 19 && s->lb->nvel != 27) return NULL;
    int cartsz[3], coords[3];
    const char * math = getenv("LB200_MATH");
    memset(&o, 0, sizeof(o));
    cs_nlocal(cs, o.nlocal);
    cs_nhalo(cs, &o.nhalo);
    cs_periodic(cs, o.periodic);
    cs_cartsz(cs, cartsz);
    cs_cart_coords(cs, coords);
    if (cartsz[Y] != 1 || cartsz[Z] != 1) pe_fatal(pe, "libludwig_b200: decomposition %d_%d_%d: x-slabs (grid P_1_1) only\n", cartsz[X], cartsz[Y], cartsz[Z]);
    o.nvel = s->lb ? s->lb->nvel : NVEL;
    o.ndist = s->lb ? s->lb->ndist : 1;
    o.have_phi = (o.ndist == 2) || (s->phi != NULL && o.nhalo >= 2);     /* symmetric_lb carries phi in the second distribution */
    o.have_q = (s->q != NULL && !o.have_phi && o.nhalo >= 2);
    o.halo_scheme = (s->lb && s->lb->haloscheme == LB_HALO_REDUCED) ? LB200_HALO_REDUCED : LB200_HALO_FULL;
    o.math = (math && strcmp(math, "strict") == 0) ? LB200_MATH_STRICT : LB200_MATH_FAST;
    o.device = -1;
    o.cart_size = cartsz[X];
    o.cart_rank = coords[X];
    if (s->le && lees_edw_nplane_total(s->le) > 0) {
      o.le_nplanes = lees_edw_nplane_total(s->le);
      lees_edw_plane_uy(s->le, &o.le_uy);
      o.le_nt0 = 0;
    }
    TRACE("lb200_create %d x %d x %d nhalo %d ndist %d phi %d q %d le %d", o.nlocal[0], o.nlocal[1], o.nlocal[2], o.nhalo, o.ndist,
	  o.have_phi, o.have_q, o.le_nplanes);
    b200_check(pe, lb200_create(&o, &s->ctx), "lb200_create");
    s->has_phi = o.have_phi; s->has_q = o.have_q;
  }
  return s->ctx;
}

/* the plane displacement is a function of the step counter: the device context keeps its own copy (src/physics.c:600-647) */
static void b200_time_sync(pe_t * pe, slot_t * s) {
  physics_t * phys = NULL;
  if (s->ctx == NULL || s->le == NULL || lees_edw_nplane_total(s->le) == 0) return;
  physics_ref(&phys);
  b200_check(pe, lb200_physics_control_time_set(s->ctx, 0, physics_control_timestep(phys)), "physics_control_time");
}

/* which array of the device lattice a field_t is (-1: none -- it stays on the reference's own code path) */
static int b200_field_array(slot_t * s, field_t * f) {
  if (s == NULL || s->ctx == NULL) return -1;
  if (f == s->phi && s->phi != NULL) return s->has_phi ? LB200_PHI : -1;
  if (f == s->q && s->q != NULL) return s->has_q ? LB200_Q : -1;
  if (f == s->u) return LB200_U;
  if (f == s->rho) return LB200_RHO;
  if (f == s->force) return LB200_FORCE;
  return -1;
}

static void b200_symm_param(fe_t * fe, phi_ch_t * pch, lb200_symm_param_t * sp) {
  physics_t * phys = NULL;
  fe_symm_param_t p;
  memset(sp, 0, sizeof(*sp));
  fe_symm_param((fe_symm_t *) fe, &p);
  physics_ref(&phys);
  sp->a = p.a; sp->b = p.b; sp->kappa = p.kappa;
  physics_mobility(phys, &sp->mobility);
  physics_grad_mu(phys, sp->gradmu);
  advection_order(&sp->adv_order);
  sp->conserve = pch ? pch->info.conserve : 0;
}

/* ---- object life cycle: who lives on which cs ------------------------------------------------------------ */

int __real_lb_data_create(pe_t * pe, cs_t * cs, const lb_data_options_t * opts, lb_t ** lb);
int __wrap_lb_data_create(pe_t * pe, cs_t * cs, const lb_data_options_t * opts, lb_t ** lb) {
  int rc = __real_lb_data_create(pe, cs, opts, lb);
  slot_t * s = slot_find(cs, 1);
  if (rc == 0 && s) { s->lb = *lb; s->nobj += 1; }
  return rc;
}

int __real_lb_free(lb_t * lb);
int __wrap_lb_free(lb_t * lb) {
  slot_t * s = slot_find(lb->cs, 0);
  if (s && s->lb == lb) { s->lb = NULL; slot_release(s); }
  return __real_lb_free(lb);
}

int __real_field_create(pe_t * pe, cs_t * cs, lees_edw_t * le, const char * name, const field_options_t * opts, field_t ** pobj);
int __wrap_field_create(pe_t * pe, cs_t * cs, lees_edw_t * le, const char * name, const field_options_t * opts, field_t ** pobj) {
  int rc = __real_field_create(pe, cs, le, name, opts, pobj);
  slot_t * s = slot_find(cs, 1);
  if (rc == 0 && s) {
    field_t * f = *pobj;
    s->nobj += 1;
    if (le) s->le = le;
    /* hydro_create names its fields "rho", "vel", "force" (src/hydro.c:75-110); the order parameters are the driver's */
    if (strcmp(name, "rho") == 0 && f->nf == 1) s->rho = f;
    else if (strcmp(name, "vel") == 0 && f->nf == 3) s->u = f;
    else if (strcmp(name, "force") == 0 && f->nf == 3) s->force = f;
    else if (strcmp(name, "eta") == 0) { /* viscosity field: host only */ }
    else if (f->nf == 1 && s->phi == NULL) s->phi = f;
    else if (f->nf == 5 && s->q == NULL) s->q = f;
    else if (f->nf == 3 && s->u == NULL) s->u = f;        /* a bare vector field (tests/unit/test_field.c) */
  }
  return rc;
}

int __real_field_free(field_t * obj);
int __wrap_field_free(field_t * obj) {
  slot_t * s = slot_find(obj->cs, 0);
  if (s) {
    if (s->phi == obj) s->phi = NULL;
    if (s->q == obj) s->q = NULL;
    if (s->u == obj) s->u = NULL;
    if (s->rho == obj) s->rho = NULL;
    if (s->force == obj) s->force = NULL;
    slot_release(s);
  }
  return __real_field_free(obj);
}

/* ---- copies: where authority changes hands -------------------------------------------------------------------- */

int __wrap_lb_memcpy(lb_t * lb, tdpMemcpyKind flag) {            /* src/lb_data.c:529-583 */
  lb200_t * ctx = b200_context(lb->pe, lb->cs, flag == tdpMemcpyHostToDevice);
  if (ctx == NULL || flag == tdpMemcpyDeviceToDevice) return 0;
  b200_check(lb->pe, lb200_memcpy(ctx, LB200_F, lb->f, b200_kind(flag)), "lb_memcpy");
  return 0;
}

int __wrap_field_memcpy(field_t * obj, tdpMemcpyKind flag) {     /* src/field.c:224-288 */
  slot_t * s = slot_find(obj->cs, 0);
  if (s && flag == tdpMemcpyHostToDevice) b200_context(obj->pe, obj->cs, 1);
  int a = b200_field_array(s, obj);
  if (a < 0) return 0;                                          /* host == target for anything else, as in a CPU build */
  b200_check(obj->pe, lb200_memcpy(s->ctx, a, obj->data, b200_kind(flag)), "field_memcpy");
  return 0;
}

int __wrap_hydro_memcpy(hydro_t * obj, tdpMemcpyKind flag) {     /* src/hydro.c:123-160: rho, u, force */
  __wrap_field_memcpy(obj->rho, flag);
  __wrap_field_memcpy(obj->u, flag);
  __wrap_field_memcpy(obj->force, flag);
  return 0;
}

int __wrap_field_grad_memcpy(field_grad_t * obj, tdpMemcpyKind flag) {    /* src/field_grad.c:236-262 */
  slot_t * s = slot_find(obj->field->cs, 0);
  int a = b200_field_array(s, obj->field);
  if (a == LB200_PHI) {
    b200_check(obj->pe, lb200_memcpy(s->ctx, LB200_GRAD, obj->grad, b200_kind(flag)), "field_grad_memcpy");
    b200_check(obj->pe, lb200_memcpy(s->ctx, LB200_DELSQ, obj->delsq, b200_kind(flag)), "field_grad_memcpy");
  }
  else if (a == LB200_Q) {
    b200_check(obj->pe, lb200_memcpy(s->ctx, LB200_QGRAD, obj->grad, b200_kind(flag)), "field_grad_memcpy");
    b200_check(obj->pe, lb200_memcpy(s->ctx, LB200_QDELSQ, obj->delsq, b200_kind(flag)), "field_grad_memcpy");
  }
  return 0;
}

/* ---- hydro_t ----------------------------------------------------------------------------------------------- */

int __real_hydro_f_zero(hydro_t * obj, const double fzero[3]);
int __wrap_hydro_f_zero(hydro_t * obj, const double fzero[3]) {   /* src/hydro.c:240-263 */
  slot_t * s = slot_find(obj->cs, 0);
  if (b200_field_array(s, obj->force) < 0) return __real_hydro_f_zero(obj, fzero);
  if (fzero[X] != 0.0 || fzero[Y] != 0.0 || fzero[Z] != 0.0) pe_fatal(obj->pe, "libludwig_b200: hydro_f_zero with a non-zero value\n");
  b200_check(obj->pe, lb200_hydro_f_zero(s->ctx), "hydro_f_zero");
  return 0;
}

int __real_hydro_u_zero(hydro_t * obj, const double uzero[3]);
int __wrap_hydro_u_zero(hydro_t * obj, const double uzero[3]) {   /* src/hydro.c:217-238 */
  slot_t * s = slot_find(obj->cs, 0);
  if (b200_field_array(s, obj->u) < 0) return __real_hydro_u_zero(obj, uzero);
  if (uzero[X] != 0.0 || uzero[Y] != 0.0 || uzero[Z] != 0.0) pe_fatal(obj->pe, "libludwig_b200: hydro_u_zero with a non-zero value\n");
  b200_check(obj->pe, lb200_hydro_u_zero(s->ctx), "hydro_u_zero");
  return 0;
}

int __real_hydro_u_halo(hydro_t * obj);
int __wrap_hydro_u_halo(hydro_t * obj) {                          /* src/hydro.c:185-191 */
  slot_t * s = slot_find(obj->cs, 0);
  if (b200_field_array(s, obj->u) < 0) return __real_hydro_u_halo(obj);
  b200_check(obj->pe, lb200_hydro_u_halo(s->ctx), "hydro_u_halo");
  return 0;
}

/* ---- field_t ------------------------------------------------------------------------------------------------ */

static int b200_field_halo(slot_t * s, field_t * obj, int a) {
  if (a == LB200_PHI) b200_check(obj->pe, lb200_phi_halo(s->ctx), "field_halo");
  else if (a == LB200_U) b200_check(obj->pe, lb200_hydro_u_halo(s->ctx), "field_halo");
  else if (a == LB200_Q) b200_check(obj->pe, lb200_q_halo(s->ctx), "field_halo");
  else return -1;
  return 0;
}

int __real_field_halo(field_t * obj);
int __wrap_field_halo(field_t * obj) {                            /* src/field.c:371-404 */
  slot_t * s = slot_find(obj->cs, 0);
  int a = b200_field_array(s, obj);
  if (a < 0 || b200_field_halo(s, obj, a) != 0) return __real_field_halo(obj);
  return 0;
}

int __real_field_halo_swap(field_t * obj, field_halo_enum_t flag);
int __wrap_field_halo_swap(field_t * obj, field_halo_enum_t flag) {   /* src/field.c:1541-1560 */
  slot_t * s = slot_find(obj->cs, 0);
  int a = b200_field_array(s, obj);
  if (a < 0 || flag == FIELD_HALO_HOST || b200_field_halo(s, obj, a) != 0) return __real_field_halo_swap(obj, flag);
  return 0;
}

int __real_field_grad_compute(field_grad_t * obj);
int __wrap_field_grad_compute(field_grad_t * obj) {               /* src/field_grad.c:319-340 */
  slot_t * s = slot_find(obj->field->cs, 0);
  int a = b200_field_array(s, obj->field);
  if (a == LB200_PHI && (obj->d2 == grad_3d_27pt_fluid_d2 || obj->d2 == grad_3d_7pt_fluid_d2)) {
    /* with planes: field_leesedwards + d2 + the buffer-region gradients, all on the device */
    b200_time_sync(obj->pe, s);
    b200_check(obj->pe, lb200_set_knob(s->ctx, LB200_KNOB_GRAD_7PT, obj->d2 == grad_3d_7pt_fluid_d2), "field_grad_compute");
    b200_check(obj->pe, lb200_phi_grad_compute(s->ctx), "field_grad_compute");
    if (obj->level >= 4) {
      if (obj->d4 != grad_3d_27pt_fluid_d4) pe_fatal(obj->pe, "libludwig_b200: fourth-order gradients: 3d_27pt_fluid only\n");
      b200_check(obj->pe, lb200_phi_grad_compute_d4(s->ctx), "field_grad_compute");
    }
    return 0;
  }
  if (a == LB200_Q && (obj->d2 == grad_3d_7pt_fluid_d2 || obj->d2 == grad_2d_5pt_fluid_d2)) {
    b200_check(obj->pe, lb200_set_knob(s->ctx, LB200_KNOB_QGRAD_2D5, obj->d2 == grad_2d_5pt_fluid_d2), "field_grad_compute");
    b200_check(obj->pe, lb200_q_grad_compute(s->ctx), "field_grad_compute");
    return 0;
  }
  if (a >= 0) pe_fatal(obj->pe, "libludwig_b200: fd_gradient_calculation of this run has no device kernel (3d_27pt_fluid, 3d_7pt_fluid; Q tensor: 3d_7pt_fluid, 2d_5pt_fluid)\n");
  return __real_field_grad_compute(obj);
}

/* ---- liquid crystal: parameters ------------------------------------------------------------------------------ */

/* beris_edw_t is opaque outside its own file: the rotational diffusion constant is taken where the driver sets it */
static beris_edw_param_t be_param_;
static int be_param_known_ = 0;

int __real_beris_edw_param_set(beris_edw_t * be, beris_edw_param_t * values);
int __wrap_beris_edw_param_set(beris_edw_t * be, beris_edw_param_t * values) {   /* src/blue_phase_beris_edwards.c:215-225 */
  be_param_ = *values;
  be_param_known_ = 1;
  return __real_beris_edw_param_set(be, values);
}

static void b200_lc_param(pe_t * pe, fe_t * fe, lb200_lc_param_t * lc) {
  const fe_lc_param_t * p = ((fe_lc_t *) fe)->param;
  memset(lc, 0, sizeof(*lc));
  /* the reference refreshes the phase of the electric field (param->coswt) whenever a caller asks for the device copy of the
   * free energy before a kernel (fe_lc_target -> fe_lc_param_commit, src/blue_phase.c:191-225); those callers are replaced here */
  fe_lc_param_commit((fe_lc_t *) fe);
  if (p->is_active && p->zeta2 != 0.0) pe_fatal(pe, "libludwig_b200: lc_active_zeta2 != 0 is outside this library\n");
  if (p->is_redshift_updated) pe_fatal(pe, "libludwig_b200: lc_redshift_update is outside this library\n");
  lc->a0 = p->a0; lc->q0 = p->q0; lc->gamma = p->gamma; lc->kappa0 = p->kappa0; lc->kappa1 = p->kappa1; lc->xi = p->xi;
  lc->epsilon = p->epsilon;
  for (int a = 0; a < 3; a++) lc->e0[a] = p->e0[a]*p->coswt;
  lc->Gamma = be_param_known_ ? be_param_.gamma : 0.0;
  advection_order(&lc->adv_order);
  lc->is_active = p->is_active; lc->zeta0 = p->zeta0; lc->zeta1 = p->zeta1; lc->zeta2 = p->zeta2;
  lc->redshift = p->redshift;
}

int __real_beris_edw_update(beris_edw_t * be, fe_t * fe, field_t * fq, field_grad_t * fq_grad, hydro_t * hydro,
			    colloids_info_t * cinfo, map_t * map, noise_t * noise);
int __wrap_beris_edw_update(beris_edw_t * be, fe_t * fe, field_t * fq, field_grad_t * fq_grad, hydro_t * hydro,
			    colloids_info_t * cinfo, map_t * map, noise_t * noise) {   /* src/blue_phase_beris_edwards.c:266-296 */
  slot_t * s = slot_find(fq->cs, 0);
  lb200_lc_param_t lc;
  int ncolloid = 0;
  if (s == NULL || s->ctx == NULL || b200_field_array(s, fq) != LB200_Q) return __real_beris_edw_update(be, fe, fq, fq_grad, hydro, cinfo, map, noise);
  /* hydro == NULL: relaxational dynamics only -- the reference leaves the velocity gradient and the advective fluxes at zero
   * (src/blue_phase_beris_edwards.c:278-283, 605); the same update with u = 0 */
  if (hydro == NULL) b200_check(fq->pe, lb200_hydro_u_zero(s->ctx), "beris_edw_update (no hydrodynamics)");
  if (cinfo) colloids_info_ntotal(cinfo, &ncolloid);
  if (ncolloid > 0) pe_fatal(fq->pe, "libludwig_b200: colloids are outside this library\n");
  if (!be_param_known_ || be_param_.noise) pe_fatal(fq->pe, "libludwig_b200: order-parameter noise is outside this library\n");
  if (fe == NULL || fe->id != FE_LC) pe_fatal(fq->pe, "libludwig_b200: beris_edw_update: free_energy lc_blue_phase only\n");
  b200_lc_param(fq->pe, fe, &lc);
  b200_check(fq->pe, lb200_beris_edw_update(s->ctx, &lc), "beris_edw_update");
  return 0;
}

/* ---- order-parameter sector ---------------------------------------------------------------------------------- */

int __real_phi_force_calculation(pe_t * pe, cs_t * cs, lees_edw_t * le, wall_t * wall, pth_t * pth, fe_t * fe, map_t * map,
				 field_t * phi, hydro_t * hydro);
int __wrap_phi_force_calculation(pe_t * pe, cs_t * cs, lees_edw_t * le, wall_t * wall, pth_t * pth, fe_t * fe, map_t * map,
				 field_t * phi, hydro_t * hydro) {   /* src/phi_force.c:74-137 */
  slot_t * s = slot_find(cs, 0);
  lb200_symm_param_t sp;
  if (hydro == NULL) return 0;
  if (pth->method == FE_FORCE_METHOD_NO_FORCE) return 0;
  if (s != NULL && s->ctx != NULL && s->has_q && fe != NULL && fe->id == FE_LC) {
    /* liquid crystal: pth_stress_compute (fe_lc_stress_v) + pth_force_fluid_driver, src/phi_force.c:100-110 */
    lb200_lc_param_t lc;
    if (pth->method != FE_FORCE_METHOD_STRESS_DIVERGENCE) pe_fatal(pe, "libludwig_b200: fe_force_method stress_divergence only\n");
    if (wall_present(wall)) pe_fatal(pe, "libludwig_b200: walls are outside this library\n");
    b200_lc_param(pe, fe, &lc);
    b200_check(pe, lb200_lc_force_calculation(s->ctx, &lc), "phi_force_calculation (liquid crystal)");
    return 0;
  }
  if (s == NULL || s->ctx == NULL || b200_field_array(s, phi) != LB200_PHI) {
    return __real_phi_force_calculation(pe, cs, le, wall, pth, fe, map, phi, hydro);
  }
  if (pth->method != FE_FORCE_METHOD_STRESS_DIVERGENCE && pth->method != FE_FORCE_METHOD_PHI_GRADMU) {
    pe_fatal(pe, "libludwig_b200: fe_force_method stress_divergence / phi_gradmu only\n");
  }
  if (wall_present(wall)) pe_fatal(pe, "libludwig_b200: walls are outside this library\n");
  if (fe == NULL || fe->id != FE_SYMMETRIC) pe_fatal(pe, "libludwig_b200: phi_force_calculation: free_energy symmetric / lc_blue_phase only\n");
  b200_symm_param(fe, NULL, &sp);
  sp.force_method = (pth->method == FE_FORCE_METHOD_PHI_GRADMU);
  b200_time_sync(pe, s);
  b200_check(pe, lb200_phi_force_calculation(s->ctx, &sp), "phi_force_calculation");
  return 0;
}

int __real_phi_cahn_hilliard(phi_ch_t * pch, fe_t * fe, field_t * phi, hydro_t * hydro, map_t * map, noise_t * noise);
int __wrap_phi_cahn_hilliard(phi_ch_t * pch, fe_t * fe, field_t * phi, hydro_t * hydro, map_t * map, noise_t * noise) {
  /* src/phi_cahn_hilliard.c:213-288 (hydro_u_halo, advection, fluxes, update: one device call) */
  slot_t * s = slot_find(phi->cs, 0);
  lb200_symm_param_t sp;
  if (s == NULL || s->ctx == NULL || b200_field_array(s, phi) != LB200_PHI) return __real_phi_cahn_hilliard(pch, fe, phi, hydro, map, noise);
  if (fe == NULL || fe->id != FE_SYMMETRIC) pe_fatal(pch->pe, "libludwig_b200: phi_cahn_hilliard: free_energy symmetric only\n");
  if (pch->info.noise) pe_fatal(pch->pe, "libludwig_b200: order-parameter noise is outside this library\n");
  b200_symm_param(fe, pch, &sp);
  b200_time_sync(pch->pe, s);
  /* cahn_hilliard_options_conserve 2 restores the sum the driver's statistics took at time 0 (src/cahn_hilliard_stats.c:58-76) */
  if (sp.conserve == 2) b200_check(pch->pe, lb200_phi_init_sum_set(s->ctx, phi->field_init_sum), "phi_init_sum");
  b200_check(pch->pe, lb200_phi_cahn_hilliard(s->ctx, &sp), "phi_cahn_hilliard");
  return 0;
}

/* ---- what this library covers, checked every time step ------------------------------------------------------------
 * fluid-only lattices (no walls, no colloids, no porous-media map), free energy none / symmetric / symmetric_lb /
 * lc_blue_phase.  Anything else would run partly on the device and partly on the host with different views of the data:
 * the run is refused, loudly, rather than finished with wrong numbers.  The driver's wall and colloid objects are seen where it
 * creates them (wall_rt_init, colloids_init_rt: src/ludwig.c:274-277). */

static wall_t * wall_seen_ = NULL;
static colloids_info_t * cinfo_seen_ = NULL;

int __real_wall_rt_init(pe_t * pe, cs_t * cs, rt_t * rt, lb_t * lb, map_t * map, wall_t ** wall);
int __wrap_wall_rt_init(pe_t * pe, cs_t * cs, rt_t * rt, lb_t * lb, map_t * map, wall_t ** wall) {
  int rc = __real_wall_rt_init(pe, cs, rt, lb, map, wall);
  wall_seen_ = wall ? *wall : NULL;
  return rc;
}

int __real_colloids_init_rt(pe_t * pe, rt_t * rt, cs_t * cs, colloids_info_t ** pinfo, colloid_io_t ** pcio,
			    interact_t ** interact, wall_t * wall, map_t * map, const lb_model_t * model);
int __wrap_colloids_init_rt(pe_t * pe, rt_t * rt, cs_t * cs, colloids_info_t ** pinfo, colloid_io_t ** pcio,
			    interact_t ** interact, wall_t * wall, map_t * map, const lb_model_t * model) {
  int rc = __real_colloids_init_rt(pe, rt, cs, pinfo, pcio, interact, wall, map, model);
  cinfo_seen_ = pinfo ? *pinfo : NULL;
  return rc;
}

static void b200_scope_check(pe_t * pe, map_t * map, fe_t * fe) {
  int n = 0;
  if (wall_seen_ && wall_present(wall_seen_)) pe_fatal(pe, "libludwig_b200: walls are outside this library\n");
  if (cinfo_seen_) colloids_info_ntotal(cinfo_seen_, &n);
  if (n > 0) pe_fatal(pe, "libludwig_b200: colloids are outside this library\n");
  if (map) map_pm(map, &n);
  if (map && n) pe_fatal(pe, "libludwig_b200: porous media are outside this library\n");
  if (fe && fe->id != FE_SYMMETRIC && fe->id != FE_LC) {
    pe_fatal(pe, "libludwig_b200: this free energy is outside this library (none, symmetric, symmetric_lb, lc_blue_phase)\n");
  }
}

/* ---- distributions -------------------------------------------------------------------------------------------- */

int __real_lb_collide(lb_t * lb, hydro_t * hydro, map_t * map, noise_t * noise, fe_t * fe, visc_t * visc);
int __wrap_lb_collide(lb_t * lb, hydro_t * hydro, map_t * map, noise_t * noise, fe_t * fe, visc_t * visc) {
  /* src/collision.c:143-162; parameters re-read where the reference re-reads them (:1163-1246, 1906-1958) */
  slot_t * s = slot_find(lb->cs, 0);
  physics_t * phys = NULL;
  lb200_collide_param_t cp;
  double fpulse[3], freq;
  int t;
  if (hydro == NULL) return 0;                                    /* :147 */
  if (s == NULL || s->ctx == NULL) return __real_lb_collide(lb, hydro, map, noise, fe, visc);
  b200_scope_check(lb->pe, map, fe);
  if (visc != NULL) pe_fatal(lb->pe, "libludwig_b200: viscosity models are outside this library\n");
  if (lb->param->noise) pe_fatal(lb->pe, "libludwig_b200: lb_fluctuations are outside this library\n");
  memset(&cp, 0, sizeof(cp));
  physics_ref(&phys);
  cp.nrelax = (int) lb->nrelax;                                   /* LB_RELAXATION_M10 / BGK / TRT = 0 / 1 / 2 */
  physics_rho0(phys, &cp.rho0);
  physics_eta_shear(phys, &cp.eta_shear);
  physics_eta_bulk(phys, &cp.eta_bulk);
  physics_fbody(phys, cp.force_global);
  physics_fpulse(phys, fpulse);
  physics_fpulse_frequency(phys, &freq);
  t = physics_control_timestep(phys);
  for (int ia = 0; ia < 3; ia++) cp.force_global[ia] += fpulse[ia]*sin(2.0*4.0*atan(1.0)*freq*t);
  if (lb->ndist == 2) {
    lb200_symm_param_t sp;
    b200_symm_param(fe, NULL, &sp);
    b200_check(lb->pe, lb200_lb_collision_binary(s->ctx, &cp, &sp), "lb_collide (lb_collision_binary)");
  }
  else {
    b200_check(lb->pe, lb200_lb_collide(s->ctx, &cp), "lb_collide");
  }
  return 0;
}

int __real_lb_halo(lb_t * lb);
int __wrap_lb_halo(lb_t * lb) {                                   /* src/lb_data.c:754-762 */
  slot_t * s = slot_find(lb->cs, 0);
  if (s == NULL || s->ctx == NULL) return __real_lb_halo(lb);
  b200_check(lb->pe, lb200_lb_halo(s->ctx), "lb_halo");
  return 0;
}

int __real_lb_propagation(lb_t * lb);
int __wrap_lb_propagation(lb_t * lb) {                            /* src/propagation.c:43-50 */
  slot_t * s = slot_find(lb->cs, 0);
  if (s == NULL || s->ctx == NULL) return __real_lb_propagation(lb);
  b200_check(lb->pe, lb200_lb_propagation(s->ctx), "lb_propagation");
  return 0;
}

int __real_lb_data_apply_le_boundary_conditions(lb_t * lb, lees_edw_t * le);
int __wrap_lb_data_apply_le_boundary_conditions(lb_t * lb, lees_edw_t * le) {   /* src/model_le.c:78-180 */
  slot_t * s = slot_find(lb->cs, 0);
  if (s == NULL || s->ctx == NULL) return __real_lb_data_apply_le_boundary_conditions(lb, le);
  if (lees_edw_nplane_total(le) == 0) return 0;
  b200_time_sync(lb->pe, s);
  b200_check(lb->pe, lb200_lb_le_apply_boundary_conditions(s->ctx), "lb_data_apply_le_boundary_conditions");
  return 0;
}

int __real_phi_lb_to_field(field_t * phi, lb_t * lb);
int __wrap_phi_lb_to_field(field_t * phi, lb_t * lb) {            /* src/phi_lb_coupler.c:39-67 */
  slot_t * s = slot_find(lb->cs, 0);
  if (s == NULL || s->ctx == NULL) return __real_phi_lb_to_field(phi, lb);
  b200_check(lb->pe, lb200_phi_lb_to_field(s->ctx), "phi_lb_to_field");
  return 0;
}
