/*
 * lb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded-semantics (OpenMP over independent sites only) CPU restatement of
 * the reference's (ludwig-cf/ludwig v0.23.0) per-timestep lattice-Boltzmann hot path.
 * It is the checker the CUDA product is compared with; the product never calls it
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it).
 *
 * Parity status: PINNED.  Every function below is compared bit-for-bit with the unmodified
 * reference compiled from /root/reference (oracle/_ref/libludwig_ref.so, oracle/Makefile.ref)
 * in tests/test_oracle_vs_reference.py, with golden vectors produced by that reference in
 * tests/golden/ (script tests/golden/make_golden.py), and with the printed statistics of the
 * reference's own regression logs (tests/regression/d3q19-short/serial-spin-fd1.log,
 * serial-dist-3du.log).
 *
 * Array layout ("canonical", also used by the product's C-ABI for host arrays):
 *   nall[a] = nlocal[a] + 2*nhalo,  nsites = nall[X]*nall[Y]*nall[Z]
 *   index(ic,jc,kc) = ((ic+nhalo-1)*nall[Y] + (jc+nhalo-1))*nall[Z] + (kc+nhalo-1)   (z fastest;
 *                      reference src/coords.c:617-631), ic in [1-nhalo, nlocal+nhalo]
 *   scalar: a[index]; nf-vector: a[n*nsites + index]; distributions: f[(n*nvel + p)*nsites + index]
 */
#ifndef LB_ORACLE_H
#define LB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_geom_s {
  int nlocal[3];
  int nhalo;
  int periodic[3];
  int le_nplanes;        /* Lees-Edwards planes (0 = none): hydro / field arrays then carry
			  * nxbuffer = 2*nhalo*le_nplanes extra x-planes after the high x halo
			  * (lees_edw_nsites, src/leesedwards.c:485-495); lb->f never does (src/lb_data.c:102-115) */
} orc_geom_t;

typedef struct orc_model_s {
  int nvel;
  int ndim;
  signed char cv[27][3];
  double wv[27];
  double na[27];
  double ma[27][27];     /* ma[m][p] */
  double mi[27][27];     /* mi[p][m] = wv[p]*na[m]*ma[m][p] */
} orc_model_t;

enum {ORC_RELAX_M10 = 0, ORC_RELAX_BGK = 1, ORC_RELAX_TRT = 2};
enum {ORC_MAP_FLUID = 0};

typedef struct orc_collide_param_s {
  int nrelax;            /* ORC_RELAX_* */
  double rho0;
  double eta_shear;
  double eta_bulk;
  double force_global[3];
} orc_collide_param_t;

typedef struct orc_symm_param_s {
  double a, b, kappa;
  double mobility;
  double gradmu[3];
  int adv_order;         /* 1 .. 4 */
  int conserve;          /* cahn_hilliard_options_conserve: 0, 1 = compensated sum, 2 = global subtraction after the forward step */
  int grad_7pt;          /* fd_gradient_calculation: 0 = 3d_27pt_fluid, 1 = 3d_7pt_fluid (whole steps, orc_step) */
  double phi_init_sum;   /* conserve 2: phi->field_init_sum, the sum the global correction restores (orc_phi_sum_time0) */
  int force_method;      /* fe_force_method: 0 = stress_divergence, 1 = phi_gradmu (src/phi_force.c:99-121) */
} orc_symm_param_t;

int orc_nsites(const orc_geom_t * g);          /* hydro, fields, gradients, fluxes: with the LE buffer planes */
int orc_nsites_lb(const orc_geom_t * g);       /* distributions, map: cs_nsites */
int orc_index(const orc_geom_t * g, int ic, int jc, int kc);

int orc_model_create(int nvel, orc_model_t * model);

/* cahn_hilliard_options_conserve 2 */
double orc_phi_sum_time0(const orc_geom_t * g, const double * phi);
void orc_phi_subtract_sum(const orc_geom_t * g, double phi_init_sum, double * phi);

/* lb_propagation: fprime <- pull(f) on the interior, self copy on y/z halo of x in [1,N] */
void orc_propagation(const orc_geom_t * g, const orc_model_t * m, int ndist,
		     const double * f, double * fprime);
/* lb_halo, full or reduced */
void orc_lb_halo(const orc_geom_t * g, const orc_model_t * m, int ndist, int reduced, double * f);
/* field_halo: nhalo deep, nf components */
void orc_field_halo(const orc_geom_t * g, int nf, double * data);

/* lb_collide (ndist = 1): all sites x in [1,N], all y,z incl. halo when include_halo != 0
 * (as the reference), interior only otherwise. status may be NULL (all fluid). */
void orc_collide(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		 const char * status, int include_halo,
		 double * f, const double * force, double * rho, double * u);

void orc_grad_27pt(const orc_geom_t * g, const double * phi, double * grad, double * delsq);
void orc_grad_27pt_ne(const orc_geom_t * g, int nextra, const double * field, double * grad, double * delsq);
void orc_stress_symm(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi,
		     const double * grad, const double * delsq, double * str);
void orc_force_divergence(const orc_geom_t * g, const double * str, double * force);
void orc_advection(const orc_geom_t * g, int order, const double * u, const double * phi,
		   double * flux);
void orc_flux_mu(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi,
		 const double * delsq, double * flux);
void orc_flux_mu_ext(const orc_geom_t * g, const orc_symm_param_t * sp, double * flux);
void orc_no_flux(const orc_geom_t * g, const char * status, double * flux);
void orc_phi_update(const orc_geom_t * g, const double * flux, double * phi);
/* fe_force_method phi_gradmu: force += -phi grad mu (phi_grad_mu_fluid), then += -phi grad_mu_ext (phi_grad_mu_external) */
void orc_phi_force_gradmu(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi, const double * delsq,
			  double * force);
void orc_phi_update_conserve(const orc_geom_t * g, const double * flux, double * csum, double * phi);

void orc_field_set(const orc_geom_t * g, int nf, double * data, const double * values);

/* Whole step(s), reference order (reference src/ludwig.c:528-860). binary != 0: symmetric FD route.
 * Work arrays are allocated inside.  f has ndist = 1. */
void orc_step(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
	      const orc_symm_param_t * sp, int binary, int halo_reduced, int nsteps,
	      double * f, double * phi, double * u, double * rho, double * force,
	      double * grad, double * delsq);

/* symmetric_lb (ndist = 2): f holds [LB_RHO dist | LB_PHI dist].  phi_lb_to_field / phi_lb_from_field
 * (src/phi_lb_coupler.c:39-137), lb_collision_binary (src/collision.c:604-1013), whole steps. */
void orc_phi_lb_to_field(const orc_geom_t * g, const orc_model_t * m, const double * f, double * phi);
void orc_phi_lb_from_field(const orc_geom_t * g, const orc_model_t * m, const double * phi, double * f);
void orc_collide_binary(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
			const orc_symm_param_t * sp, double * f, const double * force,
			const double * phi, const double * grad, const double * delsq, double * u);
void orc_step_lb2(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		  const orc_symm_param_t * sp, int halo_reduced, int nsteps,
		  double * f, double * phi, double * u, double * force, double * grad, double * delsq);

/* ---- Lees-Edwards sliding periodic boundaries, oracle/lb_oracle_le.c -----------------
 * Steady shear, single domain in y (as the x-slab decomposition of the product).  `t` arguments are the
 * reference's physics_control_time() = t_start + t_current - 1 (src/physics.c:622-630) unless stated. */
typedef struct orc_le_s {
  double uy;             /* plane speed (lees_edw_options_t.uy) */
  double time0;          /* reference time nt0 (src/leesedwards.c:262) */
} orc_le_t;

int orc_le_nxbuffer(const orc_geom_t * g);
int orc_le_plane_location(const orc_geom_t * g, int np);
int orc_le_ic_to_buff(const orc_geom_t * g, int ic, int di);
int orc_le_ibuff_to_real(const orc_geom_t * g, int ib);
double orc_le_buffer_displacement(const orc_geom_t * g, const orc_le_t * le, int ib, double t);
/* field_leesedwards (src/field.c:418-510): buffer planes <- 4-point Lagrange interpolation; uses time t */
void orc_le_field(const orc_geom_t * g, const orc_le_t * le, double t, int nf, double * data);
/* hydro_lees_edwards (src/hydro.c:350-440): linear interpolation + velocity jump; the reference calls the
 * displacement with t + 1 (src/hydro.c:401): pass t, the + 1 is applied inside.  nhcomm = z extent */
void orc_le_hydro(const orc_geom_t * g, const orc_le_t * le, double t, int nhcomm, double * u);
/* grad_3d_27pt_fluid_le (src/gradient_3d_27pt_fluid.c:375-651): gradients in the buffer region */
void orc_le_grad_buffer(const orc_geom_t * g, int nextra, const double * field, double * grad, double * delsq);
/* phi_force_flux (src/phi_force.c:289-345, 360-473, 595-673): flux form of the force with the per-plane fix */
void orc_le_phi_force(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi,
		      const double * grad, const double * delsq, double * force);
/* phi_ch_le_fix_fluxes (src/phi_cahn_hilliard.c:613-745) */
void orc_le_fix_fluxes(const orc_geom_t * g, const orc_le_t * le, double t, double * flux);
/* lb_data_apply_le_boundary_conditions (src/model_le.c:78-180, 264-345, 358-400, 584-640); tstep = the integer
 * time step physics_control_timestep() */
void orc_le_lb_bc(const orc_geom_t * g, const orc_model_t * m, const orc_le_t * le, double tstep,
		  int ndist, double * f);
/* lb_le_init_shear_profile (src/model_le.c:652-714) */
void orc_le_init_shear_profile(const orc_geom_t * g, const orc_model_t * m, const orc_le_t * le,
			       double rho0, double eta, double * f);
/* whole steps with planes (src/ludwig.c:528-860); tcurrent0 = physics t_current before the first step
 * (t_start = 0); on return the caller's clock is tcurrent0 + nsteps */
void orc_le_grad7_buffer(const orc_geom_t * g, int nextra, const double * field, double * grad, double * delsq);
void orc_le_step_lb2(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		     const orc_symm_param_t * sp, const orc_le_t * le, int tcurrent0, int nsteps,
		     double * f, double * phi, double * u, double * force, double * grad, double * delsq);
void orc_le_step(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		 const orc_symm_param_t * sp, const orc_le_t * le, int tcurrent0, int nsteps,
		 double * f, double * phi, double * u, double * rho, double * force,
		 double * grad, double * delsq);

/* ---- liquid crystal: Landau-de Gennes Q tensor + Beris-Edwards, oracle/lb_oracle_lc.c ----
 * q[n*ns + i], n = XX, XY, XZ, YY, YZ; qgrad[(n*3 + a)*ns + i]; qdelsq[n*ns + i]; str[(a*3 + b)*ns + i];
 * flux[(face*5 + n)*ns + i], face = w, e, y, z.  Static redshift, no noise / colloids; activity: zeta0, zeta1 (zeta2 = 0). */
typedef struct orc_lc_param_s {
  double a0, q0, gamma, kappa0, kappa1, xi;   /* fe_lc_param_t, src/blue_phase.h:52-75 */
  double Gamma;                               /* beris_edw_param_t.gamma: rotational diffusion constant */
  double epsilon;                             /* dielectric anisotropy (as stored: includes 1/12pi) */
  double e0[3];                               /* external electric field */
  int is_active;                              /* lc_activity */
  double zeta0, zeta1, zeta2;                 /* lc_active_zeta0/1/2 (zeta2: the dp field is zero unless fe_lc_active_stress ran) */
  double redshift, rredshift;                 /* lc_init_redshift and its reciprocal (fe_lc_redshift_set, src/blue_phase.c:1357-1366); static */
  int grad_2d5;                               /* fd_gradient_calculation 2d_5pt_fluid (orc_lc_step; lattices with nlocal[Z] = 1) */
} orc_lc_param_t;

void orc_grad_7pt(const orc_geom_t * g, int nf, const double * field, double * grad, double * delsq);
void orc_grad_2d_5pt(const orc_geom_t * g, int nf, const double * field, double * grad, double * delsq);
void orc_lc_compute_h(const orc_lc_param_t * p, double q[3][3], double dq[3][3][3], double dsq[3][3], double h[3][3]);
double orc_lc_compute_fed(const orc_lc_param_t * p, double q[3][3], double dq[3][3][3]);
void orc_lc_compute_stress(const orc_lc_param_t * p, double q[3][3], double dq[3][3][3], double h[3][3], double s[3][3]);
void orc_lc_stress(const orc_geom_t * g, const orc_lc_param_t * p, const double * q, const double * grad,
		   const double * delsq, double * str);
void orc_lc_mol_field(const orc_geom_t * g, const orc_lc_param_t * p, const double * q, const double * grad,
		      const double * delsq, double * h);
double orc_lc_fed_sum(const orc_geom_t * g, const orc_lc_param_t * p, const double * q, const double * grad);
void orc_advection_nf(const orc_geom_t * g, int order, int nf, const double * u, const double * field, double * flux);
void orc_beris_edw_update(const orc_geom_t * g, double xi, double Gamma, const double * u, const double * h,
			  const double * flux, double * q);
void orc_lc_step(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		 const orc_lc_param_t * p, int adv_order, int nsteps,
		 double * f, double * q, double * u, double * rho, double * force,
		 double * qgrad, double * qdelsq);

#ifdef __cplusplus
}
#endif
#endif
