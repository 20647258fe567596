/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin driver, written for this repository, that is compiled TOGETHER WITH the unmodified
 * reference sources (see oracle/Makefile.ref) into oracle/_ref/libludwig_ref.so.  It builds
 * the reference's own objects (pe/cs/physics/lb/hydro/map/field/field_grad/fe_symm/pth/phi_ch)
 * exactly as the reference driver does for `free_energy none` and `free_energy symmetric`
 * (/root/reference/src/ludwig.c:1147-1260, 202-429), calls the reference's own hot-path entry
 * points in the reference's own order (/root/reference/src/ludwig.c:528-860), and copies raw
 * arrays in/out in a layout-neutral "canonical" form so that tests can compare
 * (a) our C restatement in oracle/lb_oracle.c and (b) the CUDA product against the reference
 * itself on identical inputs.
 *
 * Canonical array layout used at this interface (and by oracle/ and the product's C-ABI):
 *   site index  = reference cs_index(): ((ic+nhalo-1)*nall[Y] + (jc+nhalo-1))*nall[Z] + (kc+nhalo-1)
 *   scalar a    : a[index]
 *   rank-1 (nf) : a[n*nsites + index]            (structure of arrays on the allocated lattice)
 *   f           : f[(n*nvel + p)*nsites + index]
 */

#include <assert.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "pe.h"
#include "coords.h"
#include "physics.h"
#include "leesedwards.h"
#include "lb_data.h"
#include "collision.h"
#include "propagation.h"
#include "hydro.h"
#include "map.h"
#include "field.h"
#include "field_grad.h"
#include "field_phi_init.h"
#include "gradient_3d_27pt_fluid.h"
#include "symmetric.h"
#include "phi_force.h"
#include "phi_force_stress.h"
#include "phi_force_colloid.h"
#include "phi_cahn_hilliard.h"
#include "advection.h"
#include "advection_s.h"
#include "wall.h"
#include "fe_force_method.h"
#include "noise.h"
#include "phi_lb_coupler.h"
#include "model_le.h"
#include "blue_phase.h"
#include "blue_phase_init.h"
#include "blue_phase_beris_edwards.h"
#include "gradient_3d_7pt_fluid.h"
#include "gradient_2d_5pt_fluid.h"
#include "colloids.h"
#include "io_event.h"
#include "cahn_hilliard_stats.h"

typedef struct ref_cfg_s {
  int ntotal[3];
  int nhalo;
  int periodic[3];
  int ndist;
  int nrelax;          /* 0 m10, 1 bgk, 2 trt */
  int ghost_off;       /* 1 = "ghost_modes off" */
  int halo_reduced;    /* 1 = lb_halo_reduced */
  int have_phi;        /* 1 = free_energy symmetric (FD route) */
  int adv_order;       /* 1,2,3 */
  int conserve;        /* cahn_hilliard_options_conserve */
  double rho0;
  double eta_shear;
  double eta_bulk;
  double fbody[3];
  double a, b, kappa;
  double mobility;
  double gradmu[3];
  int grad_level;      /* 0/2: field_grad level 2; 4: also grad_delsq, delsq_delsq (grad_3d_27pt_fluid_d4) */
  int le_nplanes;      /* N_LE_plane: number of Lees-Edwards planes (0 = none) */
  double le_uy;        /* LE_plane_vel (steady shear) */
  /* free_energy lc_blue_phase (have_phi = 0): Landau-de Gennes Q tensor + Beris-Edwards, 7-point gradient */
  int have_q;
  double lc_a0, lc_q0, lc_gamma, lc_kappa0, lc_kappa1, lc_xi;
  double lc_Gamma;     /* rotational diffusion constant (beris_edw_param_t.gamma) */
  double lc_epsilon;   /* dielectric anisotropy as stored in fe_lc_param_t (includes the 1/12pi) */
  double lc_e0[3];     /* external electric field */
  int grad_7pt;        /* fd_gradient_calculation 3d_7pt_fluid for the scalar order parameter (default 3d_27pt_fluid) */
  int io_ascii;        /* 1: default_io_format ascii (distributions and order-parameter field, input and output) */
  int lc_active;       /* lc_activity yes */
  double lc_zeta0, lc_zeta1;   /* lc_active_zeta0, lc_active_zeta1 (zeta2 = 0) */
  double lc_redshift;  /* lc_init_redshift (0: 1.0); no dynamic update */
  int lc_grad_2d5;     /* fd_gradient_calculation 2d_5pt_fluid for the Q tensor */
  int force_gradmu;    /* fe_force_method phi_gradmu (binary fluid, FD route) */
} ref_cfg_t;

typedef struct ref_sim_s {
  ref_cfg_t cfg;
  pe_t * pe;
  cs_t * cs;
  physics_t * phys;
  lees_edw_t * le;
  lb_t * lb;
  hydro_t * hydro;
  map_t * map;
  wall_t * wall;
  field_t * phi;
  field_grad_t * phi_grad;
  fe_symm_t * fe;
  pth_t * pth;
  phi_ch_t * pch;
  field_t * q;
  field_grad_t * q_grad;
  fe_lc_t * fe_lc;
  beris_edw_t * be;
  colloids_info_t * cinfo;
  int nsites;          /* cs_nsites: lb->f, map */
  int nsites_le;       /* lees_edw_nsites: hydro, fields, gradients, fluxes (== nsites without planes) */
} ref_sim_t;

enum {REF_F = 0, REF_PHI = 1, REF_U = 2, REF_RHO = 3, REF_FORCE = 4,
      REF_GRAD = 5, REF_DELSQ = 6, REF_STR = 7, REF_FLUX = 8, REF_MAP = 9,
      REF_GRAD_DELSQ = 10, REF_DELSQ_DELSQ = 11,
      REF_Q = 12, REF_QGRAD = 13, REF_QDELSQ = 14, REF_H = 15};

static int mpi_up = 0;

ref_sim_t * ref_create(const ref_cfg_t * cfg) {

  ref_sim_t * s = (ref_sim_t *) calloc(1, sizeof(ref_sim_t));
  s->cfg = *cfg;

  if (!mpi_up) {
    static char arg0[] = "ref_harness";
    static char * argv_[] = {arg0, NULL};
    int argc = 1;
    char ** argv = argv_;
    MPI_Init(&argc, &argv);
    mpi_up = 1;
  }

  pe_create(MPI_COMM_WORLD, PE_QUIET, &s->pe);
  cs_create(s->pe, &s->cs);
  cs_nhalo_set(s->cs, cfg->nhalo);
  cs_ntotal_set(s->cs, cfg->ntotal);
  cs_periodicity_set(s->cs, cfg->periodic);
  cs_init(s->cs);
  cs_nsites(s->cs, &s->nsites);

  physics_create(s->pe, &s->phys);
  physics_rho0_set(s->phys, cfg->rho0);
  physics_eta_shear_set(s->phys, cfg->eta_shear);
  physics_eta_bulk_set(s->phys, cfg->eta_bulk);
  { double fb[3] = {cfg->fbody[0], cfg->fbody[1], cfg->fbody[2]};
    physics_fbody_set(s->phys, fb); }
  if (cfg->mobility != 0.0) physics_mobility_set(s->phys, cfg->mobility);
  { double gm[3] = {cfg->gradmu[0], cfg->gradmu[1], cfg->gradmu[2]};
    physics_grad_mu_set(s->phys, gm); }

  { lees_edw_options_t opts = {0};
    if (cfg->le_nplanes > 0) {
      /* /root/reference/src/leesedwards_rt.c: N_LE_plane, LE_plane_vel, steady shear, nt0 = 0 */
      opts.nplanes = cfg->le_nplanes;
      opts.type = LE_SHEAR_TYPE_STEADY;
      opts.uy = cfg->le_uy;
      opts.nt0 = 0;
      physics_control_init_time(s->phys, 0, 1000000);
    }
    lees_edw_create(s->pe, s->cs, &opts, &s->le);
    lees_edw_nsites(s->le, &s->nsites_le); }

  { lb_data_options_t opts = lb_data_options_ndim_nvel_ndist(NDIM, NVEL, cfg->ndist);
    opts.nrelax = (lb_relaxation_enum_t) cfg->nrelax;
    opts.halo = cfg->halo_reduced ? LB_HALO_REDUCED : LB_HALO_FULL;
    if (cfg->io_ascii) opts.iodata.input.iorformat = opts.iodata.output.iorformat = IO_RECORD_ASCII;
    lb_data_create(s->pe, s->cs, &opts, &s->lb); }
  if (cfg->ghost_off) lb_collision_ghost_modes_off(s->lb);

  { hydro_options_t opts = hydro_options_default();
    hydro_create(s->pe, s->cs, s->le, &opts, &s->hydro); }

  { map_options_t opts = map_options_default();
    map_create(s->pe, s->cs, &opts, &s->map); }

  wall_create(s->pe, s->cs, s->map, s->lb, &s->wall);

  if (cfg->have_phi) {
    field_options_t opts = field_options_ndata_nhalo(1, cfg->nhalo);
    phi_ch_info_t ch = {0};
    fe_symm_param_t p = {0};
    if (cfg->io_ascii) opts.iodata.input.iorformat = opts.iodata.output.iorformat = IO_RECORD_ASCII;
    field_create(s->pe, s->cs, s->le, "phi", &opts, &s->phi);
    field_grad_create(s->pe, s->phi, cfg->grad_level == 4 ? 4 : 2, &s->phi_grad);
    if (cfg->grad_7pt) field_grad_set(s->phi_grad, grad_3d_7pt_fluid_d2, grad_3d_7pt_fluid_d4);
    else               field_grad_set(s->phi_grad, grad_3d_27pt_fluid_d2, grad_3d_27pt_fluid_d4);
    fe_symm_create(s->pe, s->cs, s->phi, s->phi_grad, &s->fe);
    p.a = cfg->a; p.b = cfg->b; p.kappa = cfg->kappa;
    fe_symm_param_set(s->fe, p);
    if (cfg->ndist == 2) {
      /* free_energy symmetric_lb (/root/reference/src/ludwig.c:1340-1383): order parameter carried by the
       * second distribution, dynamics in lb_collision_binary, no explicit force */
      pth_create(s->pe, s->cs, FE_FORCE_METHOD_NO_FORCE, &s->pth);
    }
    else {
      ch.conserve = cfg->conserve;
      phi_ch_create(s->pe, s->cs, s->le, &ch, &s->pch);
      pth_create(s->pe, s->cs, cfg->force_gradmu ? FE_FORCE_METHOD_PHI_GRADMU : FE_FORCE_METHOD_STRESS_DIVERGENCE, &s->pth);
      advection_order_set(cfg->adv_order);
    }
  }
  else if (cfg->have_q) {
    /* /root/reference/src/ludwig.c:1598-1666, src/blue_phase_rt.c:50-410 */
    field_options_t opts = field_options_ndata_nhalo(NQAB, cfg->nhalo);
    fe_lc_param_t p = {0};
    beris_edw_param_t bp = {0};
    int ncell[3] = {2, 2, 2};
    if (cfg->io_ascii) opts.iodata.input.iorformat = opts.iodata.output.iorformat = IO_RECORD_ASCII;
    field_create(s->pe, s->cs, s->le, "q", &opts, &s->q);
    field_grad_create(s->pe, s->q, 2, &s->q_grad);
    if (cfg->lc_grad_2d5) field_grad_set(s->q_grad, grad_2d_5pt_fluid_d2, grad_2d_5pt_fluid_d4);
    else                  field_grad_set(s->q_grad, grad_3d_7pt_fluid_d2, grad_3d_7pt_fluid_d4);
    fe_lc_create(s->pe, s->cs, s->le, s->q, s->q_grad, &s->fe_lc);
    p.a0 = cfg->lc_a0; p.q0 = cfg->lc_q0; p.gamma = cfg->lc_gamma;
    p.kappa0 = cfg->lc_kappa0; p.kappa1 = cfg->lc_kappa1; p.xi = cfg->lc_xi;
    p.redshift = (cfg->lc_redshift != 0.0) ? cfg->lc_redshift : 1.0; p.rredshift = 1.0/p.redshift;
    p.epsilon = cfg->lc_epsilon;
    /* /root/reference/src/blue_phase_rt.c:157-170 */
    p.is_active = cfg->lc_active; p.zeta0 = cfg->lc_zeta0; p.zeta1 = cfg->lc_zeta1; p.zeta2 = 0.0;
    p.e0[0] = cfg->lc_e0[0]; p.e0[1] = cfg->lc_e0[1]; p.e0[2] = cfg->lc_e0[2];
    p.coswt = 1.0;
    fe_lc_param_set(s->fe_lc, &p);
    beris_edw_create(s->pe, s->cs, s->le, &s->be);
    bp.xi = cfg->lc_xi; bp.gamma = cfg->lc_Gamma;
    beris_edw_param_set(s->be, &bp);
    pth_create(s->pe, s->cs, FE_FORCE_METHOD_STRESS_DIVERGENCE, &s->pth);
    advection_order_set(cfg->adv_order);
    colloids_info_create(s->pe, s->cs, ncell, &s->cinfo);
  }
  else {
    pth_create(s->pe, s->cs, FE_FORCE_METHOD_NO_FORCE, &s->pth);
  }

  return s;
}

void ref_free(ref_sim_t * s) {
  if (s == NULL) return;
  if (s->cinfo) colloids_info_free(s->cinfo);
  if (s->be) beris_edw_free(s->be);
  if (s->fe_lc) fe_lc_free(s->fe_lc);
  if (s->q_grad) field_grad_free(s->q_grad);
  if (s->q) field_free(s->q);
  if (s->pch) phi_ch_free(s->pch);
  if (s->pth) pth_free(s->pth);
  if (s->fe) fe_symm_free(s->fe);
  if (s->phi_grad) field_grad_free(s->phi_grad);
  if (s->phi) field_free(s->phi);
  wall_free(s->wall);
  map_free(&s->map);
  hydro_free(s->hydro);
  lb_free(s->lb);
  lees_edw_free(s->le);
  physics_free(s->phys);
  cs_free(s->cs);
  pe_free(s->pe);
  free(s);
}

int ref_nsites(ref_sim_t * s) { return s->nsites; }
int ref_nsites_le(ref_sim_t * s) { return s->nsites_le; }
int ref_nvel(void) { return NVEL; }

/* --- initial conditions through the reference's own routines --------------------------- */

int ref_init_rest(ref_sim_t * s, double rho0) { return lb_init_rest_f(s->lb, rho0); }

int ref_init_uniform_u(ref_sim_t * s, double rho, const double u[3]) {
  int nlocal[3];
  double uu[3] = {u[0], u[1], u[2]};
  cs_nlocal(s->cs, nlocal);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	lb_1st_moment_equilib_set(s->lb, cs_index(s->cs, ic, jc, kc), rho, uu);
  return 0;
}

int ref_init_spinodal(ref_sim_t * s, int seed, double phi0, double amp) {
  assert(s->phi);
  return field_phi_init_spinodal(s->phi, seed, phi0, amp);
}

/* --- array access ------------------------------------------------------------------------ */

static int ref_copy(ref_sim_t * s, int what, double * buf, int put) {

  const int ns = (what == REF_F || what == REF_MAP || what == REF_STR) ? s->nsites : s->nsites_le;

#define XFER(hostexpr, k) do { if (put) (hostexpr) = buf[(k)]; else buf[(k)] = (hostexpr); } while (0)

  switch (what) {
  case REF_F:
    for (int n = 0; n < s->lb->ndist; n++)
      for (int p = 0; p < NVEL; p++)
	for (int i = 0; i < ns; i++)
	  XFER(s->lb->f[LB_ADDR(ns, s->lb->ndist, NVEL, i, n, p)], (size_t) (n*NVEL + p)*ns + i);
    break;
  case REF_PHI:
    for (int i = 0; i < ns; i++) XFER(s->phi->data[addr_rank1(ns, 1, i, 0)], i);
    break;
  case REF_U:
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < ns; i++) XFER(s->hydro->u->data[addr_rank1(ns, 3, i, a)], (size_t) a*ns + i);
    break;
  case REF_RHO:
    for (int i = 0; i < ns; i++) XFER(s->hydro->rho->data[addr_rank0(ns, i)], i);
    break;
  case REF_FORCE:
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < ns; i++) XFER(s->hydro->force->data[addr_rank1(ns, 3, i, a)], (size_t) a*ns + i);
    break;
  case REF_GRAD:
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < ns; i++) XFER(s->phi_grad->grad[addr_rank2(ns, 1, 3, i, 0, a)], (size_t) a*ns + i);
    break;
  case REF_DELSQ:
    for (int i = 0; i < ns; i++) XFER(s->phi_grad->delsq[addr_rank1(ns, 1, i, 0)], i);
    break;
  case REF_GRAD_DELSQ:
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < ns; i++) XFER(s->phi_grad->grad_delsq[addr_rank2(ns, 1, 3, i, 0, a)], (size_t) a*ns + i);
    break;
  case REF_DELSQ_DELSQ:
    for (int i = 0; i < ns; i++) XFER(s->phi_grad->delsq_delsq[addr_rank1(ns, 1, i, 0)], i);
    break;
  case REF_STR:
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
	for (int i = 0; i < ns; i++) XFER(s->pth->str[addr_rank2(ns, 3, 3, i, a, b)], (size_t) (a*3 + b)*ns + i);
    break;
  case REF_FLUX:
    for (int i = 0; i < ns; i++) {
      XFER(s->pch->flux->fw[addr_rank0(ns, i)], (size_t) 0*ns + i);
      XFER(s->pch->flux->fe[addr_rank0(ns, i)], (size_t) 1*ns + i);
      XFER(s->pch->flux->fy[addr_rank0(ns, i)], (size_t) 2*ns + i);
      XFER(s->pch->flux->fz[addr_rank0(ns, i)], (size_t) 3*ns + i);
    }
    break;
  case REF_Q:
    for (int n = 0; n < NQAB; n++)
      for (int i = 0; i < ns; i++) XFER(s->q->data[addr_rank1(ns, NQAB, i, n)], (size_t) n*ns + i);
    break;
  case REF_QGRAD:
    for (int n = 0; n < NQAB; n++)
      for (int a = 0; a < 3; a++)
	for (int i = 0; i < ns; i++) XFER(s->q_grad->grad[addr_rank2(ns, NQAB, 3, i, n, a)], (size_t) (n*3 + a)*ns + i);
    break;
  case REF_QDELSQ:
    for (int n = 0; n < NQAB; n++)
      for (int i = 0; i < ns; i++) XFER(s->q_grad->delsq[addr_rank1(ns, NQAB, i, n)], (size_t) n*ns + i);
    break;
  case REF_H:
    /* the molecular field through the reference's public per-site function (the stored copy is private to
     * blue_phase_beris_edwards.c); read only */
    if (put) return -1;
    for (int i = 0; i < ns; i++) {
      int nall[3], nlocal[3], nh;
      cs_nlocal(s->cs, nlocal); cs_nhalo(s->cs, &nh);
      for (int a2 = 0; a2 < 3; a2++) nall[a2] = nlocal[a2] + 2*nh;
      int kc = i % nall[Z] - nh + 1, jc = (i/nall[Z]) % nall[Y] - nh + 1, ic = i/(nall[Y]*nall[Z]) - nh + 1;
      double h[3][3] = {{0.0}};
      if (ic >= 1 && ic <= nlocal[X] && jc >= 1 && jc <= nlocal[Y] && kc >= 1 && kc <= nlocal[Z]) fe_lc_mol_field(s->fe_lc, i, h);
      buf[(size_t) 0*ns + i] = h[X][X]; buf[(size_t) 1*ns + i] = h[X][Y]; buf[(size_t) 2*ns + i] = h[X][Z];
      buf[(size_t) 3*ns + i] = h[Y][Y]; buf[(size_t) 4*ns + i] = h[Y][Z];
    }
    break;
  case REF_MAP:
    for (int i = 0; i < ns; i++) {
      if (put) s->map->status[i] = (char) buf[i]; else buf[i] = (double) s->map->status[i];
    }
    break;
  default:
    return -1;
  }
#undef XFER
  return 0;
}

int ref_get(ref_sim_t * s, int what, double * out) { return ref_copy(s, what, out, 0); }
int ref_set(ref_sim_t * s, int what, const double * in) { return ref_copy(s, what, (double *) in, 1); }

/* --- individual hot-path operators, reference entry points unchanged ---------------------- */

int ref_hydro_f_zero(ref_sim_t * s) { double z[3] = {0.0, 0.0, 0.0}; return hydro_f_zero(s->hydro, z); }
int ref_hydro_u_zero(ref_sim_t * s) { double z[3] = {0.0, 0.0, 0.0}; return hydro_u_zero(s->hydro, z); }
int ref_hydro_u_halo(ref_sim_t * s) { return hydro_u_halo(s->hydro); }
int ref_phi_halo(ref_sim_t * s) { return field_halo(s->phi); }
int ref_grad_compute(ref_sim_t * s) { return field_grad_compute(s->phi_grad); }
int ref_grad_d4(ref_sim_t * s) { return grad_3d_27pt_fluid_d4(s->phi_grad); }
int ref_pth_stress_compute(ref_sim_t * s) { return pth_stress_compute(s->pth, (fe_t *) s->fe); }
int ref_pth_force_fluid_driver(ref_sim_t * s) { return pth_force_fluid_driver(s->pth, s->hydro); }
int ref_phi_force(ref_sim_t * s) {
  fe_t * fe = s->fe_lc ? (fe_t *) s->fe_lc : (fe_t *) s->fe;
  return phi_force_calculation(s->pe, s->cs, s->le, s->wall, s->pth, fe, s->map,
			       s->phi, s->hydro);
}
/* liquid crystal: /root/reference/src/ludwig.c:579-586, 769-779 */
int ref_q_halo(ref_sim_t * s) { return field_halo(s->q); }
int ref_q_grad_compute(ref_sim_t * s) { return field_grad_compute(s->q_grad); }
int ref_lc_stress_compute(ref_sim_t * s) { return pth_stress_compute(s->pth, (fe_t *) s->fe_lc); }
int ref_beris_edw_update(ref_sim_t * s) {
  return beris_edw_update(s->be, (fe_t *) s->fe_lc, s->q, s->q_grad, s->hydro, s->cinfo, s->map, NULL);
}
int ref_lc_twist_init(ref_sim_t * s, int helical_axis, double amplitude) {
  fe_lc_param_t p;
  fe_lc_param(s->fe_lc, &p);
  p.amplitude0 = amplitude;
  return blue_phase_twist_init(s->cs, &p, s->q, helical_axis);
}
int ref_lc_o8m_init(ref_sim_t * s, double amplitude) {
  fe_lc_param_t p;
  double angles[3] = {0.0, 0.0, 0.0};
  fe_lc_param(s->fe_lc, &p);
  p.amplitude0 = amplitude;
  return blue_phase_O8M_init(s->cs, &p, s->q, angles);
}
/* total free energy density sum over the interior (stats_free_energy_density, src/stats_free_energy.c:76-134) */
double ref_lc_fed_sum(ref_sim_t * s) {
  int nlocal[3];
  double sum = 0.0;
  cs_nlocal(s->cs, nlocal);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	double fed;
	fe_lc_fed(s->fe_lc, cs_index(s->cs, ic, jc, kc), &fed);
	sum += fed;
      }
  return sum;
}
/* cahn_hilliard_options_conserve 2: the initial sum of the driver's statistics code (src/ludwig.c calls it once before the loop) */
double ref_phi_stats_time0(ref_sim_t * s) {
  cahn_hilliard_stats_time0(s->pch, s->phi, s->map);
  return s->phi->field_init_sum;
}
int ref_phi_init_sum_set(ref_sim_t * s, double v) { s->phi->field_init_sum = v; return 0; }

int ref_cahn_hilliard(ref_sim_t * s) {
  return phi_cahn_hilliard(s->pch, (fe_t *) s->fe, s->phi, s->hydro, s->map, NULL);
}
int ref_collide(ref_sim_t * s) {
  return lb_collide(s->lb, s->hydro, s->map, NULL, (fe_t *) s->fe, NULL);
}
int ref_lb_halo(ref_sim_t * s) { return lb_halo(s->lb); }
int ref_phi_lb_to_field(ref_sim_t * s) { return phi_lb_to_field(s->phi, s->lb); }
int ref_phi_lb_from_field(ref_sim_t * s) { return phi_lb_from_field(s->phi, s->lb); }
int ref_propagation(ref_sim_t * s) { return lb_propagation(s->lb); }

/* Lees-Edwards: the pieces of field_grad_compute / phi_cahn_hilliard / the LB boundary condition that the
 * planes add (/root/reference/src/field.c:418-510, src/hydro.c:350-440, src/model_le.c:78-180) */
int ref_le_field(ref_sim_t * s) { return field_leesedwards(s->phi); }
int ref_le_hydro(ref_sim_t * s) { return hydro_lees_edwards(s->hydro); }
int ref_le_lb_bc(ref_sim_t * s) { return lb_data_apply_le_boundary_conditions(s->lb, s->le); }
int ref_le_init_shear_profile(ref_sim_t * s) { return lb_le_init_shear_profile(s->lb, s->le); }
int ref_next_step(ref_sim_t * s) { return physics_control_next_step(s->phys); }
int ref_timestep(ref_sim_t * s) { return physics_control_timestep(s->phys); }

/* on-disk formats: the reference's own writers / readers, files land in the current directory
 * as dist-%9.9d.001-001, phi-%9.9d.001-001, q-%9.9d.001-001 (+ their .meta JSON): src/lb_data.c:1716-1830,
 * src/field.c:1633-1740, src/io_subfile.c:186-204 */
int ref_lb_io_write(ref_sim_t * s, int timestep) { io_event_t ev = {0}; return lb_io_write(s->lb, timestep, &ev); }
int ref_lb_io_read(ref_sim_t * s, int timestep) { io_event_t ev = {0}; return lb_io_read(s->lb, timestep, &ev); }
int ref_field_io_write(ref_sim_t * s, int timestep) {
  io_event_t ev = {0};
  return field_io_write(s->q ? s->q : s->phi, timestep, &ev);
}
int ref_field_io_read(ref_sim_t * s, int timestep) {
  io_event_t ev = {0};
  return field_io_read(s->q ? s->q : s->phi, timestep, &ev);
}

/* One full time step in the reference driver's order (/root/reference/src/ludwig.c:528-860) */

int ref_step(ref_sim_t * s, int nsteps) {
  for (int n = 0; n < nsteps; n++) {
    if (s->cfg.le_nplanes > 0) physics_control_next_step(s->phys);
    ref_hydro_f_zero(s);
    if (s->lb->ndist == 2) {
      /* symmetric_lb: /root/reference/src/ludwig.c:551-571, 683-685 */
      ref_phi_lb_to_field(s);
      ref_phi_halo(s);
      ref_grad_compute(s);
    }
    else if (s->phi) {
      ref_phi_halo(s);
      ref_grad_compute(s);
      ref_phi_force(s);
      ref_cahn_hilliard(s);
    }
    else if (s->q) {
      ref_q_halo(s);
      ref_q_grad_compute(s);
      ref_phi_force(s);
      ref_hydro_u_halo(s);
      ref_beris_edw_update(s);
    }
    ref_hydro_u_zero(s);
    ref_collide(s);
    if (s->cfg.le_nplanes > 0) ref_le_lb_bc(s);       /* /root/reference/src/ludwig.c:817-819 */
    ref_lb_halo(s);
    ref_propagation(s);
  }
  return 0;
}

/* OpenMP team size of the reference's kernels (target/target_x86.c runs every kernel in a `#pragma omp parallel`):
 * n > 0 sets it for this process whatever OMP_NUM_THREADS said at start-up (torchrun exports OMP_NUM_THREADS=1);
 * returns the team size the next kernel will use, which is what the bench reports as "cores". */
#ifdef _OPENMP
#include <omp.h>
int ref_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int ref_omp_threads(int n) { (void) n; return 1; }
#endif

/* Wall-clock of nsteps full steps, for the CPU baseline (bench.py --impl reference) */

double ref_time_steps(ref_sim_t * s, int nsteps) {
  double t0 = MPI_Wtime();
  ref_step(s, nsteps);
  return MPI_Wtime() - t0;
}
