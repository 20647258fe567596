/*
 * lb_oracle_lc.c -- TEST INFRASTRUCTURE ONLY (see lb_oracle.h for status and layout).
 *
 * CPU restatement of the liquid-crystal (Landau-de Gennes Q tensor, `free_energy lc_blue_phase`) additions to
 * the reference's time step: the 7-point gradient of the five independent Q components, the
 * molecular field, the free-energy density and the (non-symmetric) stress in the reference's *vectorised* forms
 * (the ones its kernels call), and the Beris-Edwards update with advective fluxes.  Redshift 1, no activity,
 * no noise, no colloids / walls.  Written in the reference's order of floating-point operations: compile with
 * -ffp-contract=off.  Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Layout: q[n*ns + index], n = XX, XY, XZ, YY, YZ; qgrad[(n*3 + ia)*ns + index]; qdelsq[n*ns + index];
 * str[(ia*3 + ib)*ns + index]; flux[(face*5 + n)*ns + index], face = w, e, y, z.
 *
 * Parity status: PINNED bit-for-bit to the unmodified reference compiled here (tests/test_lc_oracle.py).
 */

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lb_oracle.h"

enum {X = 0, Y = 1, Z = 2};
enum {XX = 0, XY = 1, XZ = 2, YY = 3, YZ = 4, NQAB = 5};

static const signed char d_[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};                 /* KRONECKER_DELTA_CHAR */
static const signed char e_[3][3][3] = {{{0, 0, 0}, {0, 0, 1}, {0, -1, 0}},             /* LEVI_CIVITA_CHAR */
					{{0, 0, -1}, {0, 0, 0}, {1, 0, 0}},
					{{0, 1, 0}, {-1, 0, 0}, {0, 0, 0}}};

/* ---- 7-point gradient, nf components: src/gradient_3d_7pt_fluid.c:76-99 (extent nhalo - 1), :231-300 ------- */

void orc_grad_7pt(const orc_geom_t * g, int nf, const double * field, double * grad, double * delsq) {

  const size_t ns = (size_t) orc_nsites(g);
  const int nextra = g->nhalo - 1;
  const int ys = g->nlocal[Z] + 2*g->nhalo;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1 - nextra; ic <= g->nlocal[X] + nextra; ic++) {
    for (int jc = 1 - nextra; jc <= g->nlocal[Y] + nextra; jc++) {
      for (int kc = 1 - nextra; kc <= g->nlocal[Z] + nextra; kc++) {
	const int index = orc_index(g, ic, jc, kc);
	/* x-neighbours through the Lees-Edwards buffer planes next to a plane (src/gradient_3d_7pt_fluid.c:231-246) */
	const int indexm1 = orc_index(g, orc_le_ic_to_buff(g, ic, -1), jc, kc);
	const int indexp1 = orc_index(g, orc_le_ic_to_buff(g, ic, +1), jc, kc);
	for (int n = 0; n < nf; n++) {
	  const double * f = field + (size_t) n*ns;
	  grad[(size_t) (n*3 + X)*ns + index] = 0.5*(f[indexp1] - f[indexm1]);
	  grad[(size_t) (n*3 + Y)*ns + index] = 0.5*(f[index + ys] - f[index - ys]);
	  grad[(size_t) (n*3 + Z)*ns + index] = 0.5*(f[index + 1] - f[index - 1]);
	  delsq[(size_t) n*ns + index] = f[indexp1] + f[indexm1] + f[index + ys] + f[index - ys]
	    + f[index + 1] + f[index - 1] - 6.0*f[index];
	}
      }
    }
  }
}

/* ---- grad_2d_5pt_fluid_d2 -> grad_2d_5pt_fluid_operator: src/gradient_2d_5pt_fluid.c:53-72, 108-172.  A host loop in the reference
 * over (ic, jc) at kc = 1 ONLY: the z component of the gradient is zero, and the arrays keep whatever they held at kc != 1 (zero
 * from their allocation) -- which the stress on the z halo sites then reads (pth_stress_compute sweeps [0, N + 1]^3). ---- */

void orc_grad_2d_5pt(const orc_geom_t * g, int nf, const double * field, double * grad, double * delsq) {
  const size_t ns = (size_t) orc_nsites(g);
  const int nextra = g->nhalo - 1;
  const int ys = g->nlocal[Z] + 2*g->nhalo;
  for (int ic = 1 - nextra; ic <= g->nlocal[X] + nextra; ic++) {
    for (int jc = 1 - nextra; jc <= g->nlocal[Y] + nextra; jc++) {
      const int index = orc_index(g, ic, jc, 1);
      const int indexm1 = orc_index(g, orc_le_ic_to_buff(g, ic, -1), jc, 1);
      const int indexp1 = orc_index(g, orc_le_ic_to_buff(g, ic, +1), jc, 1);
      for (int n = 0; n < nf; n++) {
	const double * f = field + (size_t) n*ns;
	grad[(size_t) (n*3 + X)*ns + index] = 0.5*(f[indexp1] - f[indexm1]);
	grad[(size_t) (n*3 + Y)*ns + index] = 0.5*(f[index + ys] - f[index - ys]);
	grad[(size_t) (n*3 + Z)*ns + index] = 0.0;
	delsq[(size_t) n*ns + index] = f[indexp1] + f[indexm1] + f[index + ys] + f[index - ys] - 4.0*f[index];
      }
    }
  }
}

/* ---- expansion of the compressed tensors at a site: src/blue_phase.c:1689-1722 ------------------------------ */

static void lc_expand(const double * q_, const double * grad, const double * delsq, size_t ns, int index,
		      double q[3][3], double dq[3][3][3], double dsq[3][3]) {
  q[X][X] = q_[XX*ns + index]; q[X][Y] = q_[XY*ns + index]; q[X][Z] = q_[XZ*ns + index];
  q[Y][X] = q[X][Y]; q[Y][Y] = q_[YY*ns + index]; q[Y][Z] = q_[YZ*ns + index];
  q[Z][X] = q[X][Z]; q[Z][Y] = q[Y][Z]; q[Z][Z] = 0.0 - q[X][X] - q[Y][Y];
  for (int ia = 0; ia < 3; ia++) {
    dq[ia][X][X] = grad[(size_t) (XX*3 + ia)*ns + index];
    dq[ia][X][Y] = grad[(size_t) (XY*3 + ia)*ns + index];
    dq[ia][X][Z] = grad[(size_t) (XZ*3 + ia)*ns + index];
    dq[ia][Y][X] = dq[ia][X][Y];
    dq[ia][Y][Y] = grad[(size_t) (YY*3 + ia)*ns + index];
    dq[ia][Y][Z] = grad[(size_t) (YZ*3 + ia)*ns + index];
    dq[ia][Z][X] = dq[ia][X][Z];
    dq[ia][Z][Y] = dq[ia][Y][Z];
    dq[ia][Z][Z] = 0.0 - dq[ia][X][X] - dq[ia][Y][Y];
  }
  if (delsq) {
    dsq[X][X] = delsq[XX*ns + index]; dsq[X][Y] = delsq[XY*ns + index]; dsq[X][Z] = delsq[XZ*ns + index];
    dsq[Y][X] = dsq[X][Y]; dsq[Y][Y] = delsq[YY*ns + index]; dsq[Y][Z] = delsq[YZ*ns + index];
    dsq[Z][X] = dsq[X][Z]; dsq[Z][Y] = dsq[Y][Z]; dsq[Z][Z] = 0.0 - dsq[X][X] - dsq[Y][Y];
  }
}

/* ---- fe_lc_compute_h_v, src/blue_phase.c:2094-2270 (redshift = 1; note kappa1 = kappa0 there, :2119-2120) --- */

void orc_lc_compute_h(const orc_lc_param_t * p, double q[3][3], double dq[3][3][3], double dsq[3][3],
		      double h[3][3]) {
  const double r3 = (1.0/3.0);
  const double q0 = p->rredshift*p->q0;
  const double kappa0 = p->redshift*p->redshift*p->kappa0;
  const double kappa1 = kappa0;
  const double gamma = p->gamma;
  double q2 = 0.0, edq = 0.0, e2 = 0.0, sum;

  for (int ia = 0; ia < 3; ia++)
    for (int ib = 0; ib < 3; ib++) q2 += q[ia][ib]*q[ia][ib];

  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      sum = 0.0;
      for (int ic = 0; ic < 3; ic++) sum += q[ia][ic]*q[ib][ic];
      h[ia][ib] = - p->a0*(1.0 - r3*gamma)*q[ia][ib] + p->a0*gamma*(sum - r3*q2*d_[ia][ib]) - p->a0*gamma*q2*q[ia][ib];
    }
  }

  for (int ib = 0; ib < 3; ib++)
    for (int ic = 0; ic < 3; ic++)
      for (int ia = 0; ia < 3; ia++) edq += e_[ib][ic][ia]*dq[ib][ic][ia];

  /* The unrolled contraction e_acd d_c Q_bd + e_bcd d_c Q_ad (:2150-2242): the non-zero terms in (ic, id) order,
   * the two of a diagonal element added together before they join the sum */
  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      sum = 0.0;
      for (int ic = 0; ic < 3; ic++) {
	for (int id = 0; id < 3; id++) {
	  const int ea = e_[ia][ic][id], eb = e_[ib][ic][id];
	  if (ea != 0 && eb != 0) sum += ea*dq[ic][ib][id] + eb*dq[ic][ia][id];
	  else if (ea != 0) sum += ea*dq[ic][ib][id];
	  else if (eb != 0) sum += eb*dq[ic][ia][id];
	}
      }
      if (ia == ib) {
	h[ia][ib] += kappa0*dsq[ia][ib] - 2.0*kappa1*q0*sum + 4.0*r3*kappa1*q0*edq - 4.0*kappa1*q0*q0*q[ia][ib];
      }
      else {
	h[ia][ib] += kappa0*dsq[ia][ib] - 2.0*kappa1*q0*sum - 4.0*kappa1*q0*q0*q[ia][ib];
      }
    }
  }

  /* electric field (:2246-2263), coswt = 1 */
  for (int ia = 0; ia < 3; ia++) { double ea = p->e0[ia]*1.0; e2 += ea*ea; }
  for (int ia = 0; ia < 3; ia++) {
    double ea = p->e0[ia]*1.0;
    for (int ib = 0; ib < 3; ib++) {
      double eb = p->e0[ib]*1.0;
      h[ia][ib] += p->epsilon*(ea*eb - r3*d_[ia][ib]*e2);
    }
  }
}

/* ---- fe_lc_compute_fed_v, src/blue_phase.c:1908-2075 (kappa1 = kappa0, :1934) ------------------------------- */

double orc_lc_compute_fed(const orc_lc_param_t * p, double q[3][3], double dq[3][3][3]) {
  const double r3 = 1.0/3.0;
  const double q0 = p->rredshift*p->q0;
  const double kappa0 = p->redshift*p->redshift*p->kappa0;
  const double kappa1 = kappa0;
  double q2 = 0.0, q3 = 0.0, dq0 = 0.0, dq1 = 0.0, efield = 0.0, sum;

  for (int ia = 0; ia < 3; ia++)
    for (int ib = 0; ib < 3; ib++) q2 += q[ia][ib]*q[ia][ib];
  for (int ia = 0; ia < 3; ia++)
    for (int ib = 0; ib < 3; ib++)
      for (int ic = 0; ic < 3; ic++) q3 += q[ia][ib]*q[ib][ic]*q[ia][ic];

  for (int ia = 0; ia < 3; ia++) {
    sum = 0.0;
    for (int ib = 0; ib < 3; ib++) sum += dq[ib][ia][ib];
    dq0 += sum*sum;
  }

  /* (e_acd d_c Q_bd + 2 q0 Q_ab)^2, unrolled there over the non-zero e_acd in (ic, id) order (:1975-2048) */
  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      sum = 0.0;
      for (int ic = 0; ic < 3; ic++) {
	for (int id = 0; id < 3; id++) {
	  if (e_[ia][ic][id] > 0) sum += dq[ic][ib][id];
	  if (e_[ia][ic][id] < 0) sum -= dq[ic][ib][id];
	}
      }
      sum += 2.0*q0*q[ia][ib];
      dq1 += sum*sum;
    }
  }

  for (int ia = 0; ia < 3; ia++) {
    double ea = p->e0[ia]*1.0;
    for (int ib = 0; ib < 3; ib++) {
      double eb = p->e0[ib]*1.0;
      efield += ea*q[ia][ib]*eb;
    }
  }

  return 0.5*p->a0*(1.0 - r3*p->gamma)*q2 - r3*p->a0*p->gamma*q3 + 0.25*p->a0*p->gamma*q2*q2
    + 0.5*kappa0*dq0 + 0.5*kappa1*dq1 - p->epsilon*efield;
}

/* ---- fe_lc_compute_stress_v, src/blue_phase.c:2279-2775: the automatically unrolled form of
 * fe_lc_compute_stress (:827-925) -- per element: the isotropic + 2 xi (Q + 1/3) Q:H start, three xi terms, nine
 * gradient statements each followed by its single non-zero Levi-Civita term, three antisymmetric terms, sign -- */

void orc_lc_compute_stress(const orc_lc_param_t * p, double q[3][3], double dq[3][3][3], double h[3][3],
			   double s[3][3]) {
  const double r3 = (1.0/3.0);
  const double q0 = p->q0*p->rredshift;
  const double kappa0 = p->kappa0*p->redshift*p->redshift;
  const double kappa1 = p->kappa1*p->redshift*p->redshift;
  const double xi = p->xi;
  double qh = 0.0;
  double p0 = orc_lc_compute_fed(p, q, dq);
  p0 = 0.0 - p0;

  for (int ia = 0; ia < 3; ia++)
    for (int ib = 0; ib < 3; ib++) qh += q[ia][ib]*h[ia][ib];

  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      double sth;
      if (ia == ib) sth = 2.0*xi*(q[ia][ib] + r3)*qh - p0;
      else          sth = 2.0*xi*(q[ia][ib])*qh;
      for (int ic = 0; ic < 3; ic++) {
	const double qb = (ib == ic) ? (q[ib][ic] + r3) : (q[ib][ic]);
	const double qa = (ia == ic) ? (q[ia][ic] + r3) : (q[ia][ic]);
	sth += -xi*h[ia][ic]*qb - xi*qa*h[ib][ic];
      }
      for (int ic = 0; ic < 3; ic++) {
	for (int id = 0; id < 3; id++) {
	  sth += - kappa0*dq[ia][ib][ic]*dq[id][ic][id] - kappa1*dq[ia][ic][id]*dq[ib][ic][id]
	    + kappa1*dq[ia][ic][id]*dq[ic][ib][id];
	  if (ib != ic) {
	    const int ie = 3 - ib - ic;
	    if (e_[ib][ic][ie] > 0) sth -= 2.0*kappa1*q0*dq[ia][ic][id]*q[id][ie];
	    else                    sth += 2.0*kappa1*q0*dq[ia][ic][id]*q[id][ie];
	  }
	}
      }
      for (int ic = 0; ic < 3; ic++) sth += q[ia][ic]*h[ib][ic] - h[ia][ic]*q[ib][ic];
      s[ia][ib] = -sth;
    }
  }
}

/* ---- pth_stress_compute with fe_lc_stress_v: src/phi_force_stress.c:171-284, src/blue_phase.c:1745-1800.
 * Computed on [0, N+1]^3 (what the divergence reads); the reference's flat kernel also sweeps the rest of the
 * y, z halo, whose values nothing reads ------------------------------------------------------------------------ */

void orc_lc_stress(const orc_geom_t * g, const orc_lc_param_t * p, const double * q_, const double * grad,
		   const double * delsq, double * str) {
  const size_t ns = (size_t) orc_nsites(g);
  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 0; ic <= g->nlocal[X] + 1; ic++) {
    for (int jc = 0; jc <= g->nlocal[Y] + 1; jc++) {
      for (int kc = 0; kc <= g->nlocal[Z] + 1; kc++) {
	const int index = orc_index(g, ic, jc, kc);
	double q[3][3], dq[3][3][3], dsq[3][3], h[3][3], s[3][3];
	lc_expand(q_, grad, delsq, ns, index, q, dq, dsq);
	orc_lc_compute_h(p, q, dq, dsq, h);
	orc_lc_compute_stress(p, q, dq, h, s);
	if (p->is_active) {
	  /* fe_lc_stress_v, src/blue_phase.c:1825-1845, with fe_lc_compute_stress_active, :934-972 (the documented form);
	   * dp = grad of the polarisation field, identically zero here (zeta2 term of fe_lc_active_stress not built) */
	  const double dp[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
	  double sa[3][3];
	  for (int ia = 0; ia < 3; ia++)
	    for (int ib = 0; ib < 3; ib++)
	      sa[ia][ib] = p->zeta0*(ia == ib) - p->zeta1*q[ia][ib] - p->zeta2*(dp[ia][ib] + dp[ib][ia]);
	  for (int ia = 0; ia < 3; ia++)
	    for (int ib = 0; ib < 3; ib++) sa[ia][ib] = -sa[ia][ib];
	  for (int ia = 0; ia < 3; ia++)
	    for (int ib = 0; ib < 3; ib++) s[ia][ib] += sa[ia][ib];
	}
	for (int ia = 0; ia < 3; ia++)
	  for (int ib = 0; ib < 3; ib++) str[(size_t) (ia*3 + ib)*ns + index] = s[ia][ib];
      }
    }
  }
}

/* ---- beris_edw_h_kernel_v: src/blue_phase_beris_edwards.c:942-985 (interior) -------------------------------- */

void orc_lc_mol_field(const orc_geom_t * g, const orc_lc_param_t * p, const double * q_, const double * grad,
		      const double * delsq, double * hq) {
  const size_t ns = (size_t) orc_nsites(g);
  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	const int index = orc_index(g, ic, jc, kc);
	double q[3][3], dq[3][3][3], dsq[3][3], h[3][3];
	lc_expand(q_, grad, delsq, ns, index, q, dq, dsq);
	orc_lc_compute_h(p, q, dq, dsq, h);
	hq[XX*ns + index] = h[X][X]; hq[XY*ns + index] = h[X][Y]; hq[XZ*ns + index] = h[X][Z];
	hq[YY*ns + index] = h[Y][Y]; hq[YZ*ns + index] = h[Y][Z];
      }
    }
  }
}

double orc_lc_fed_sum(const orc_geom_t * g, const orc_lc_param_t * p, const double * q_, const double * grad) {
  const size_t ns = (size_t) orc_nsites(g);
  double sum = 0.0;
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	double q[3][3], dq[3][3][3], dsq[3][3];
	lc_expand(q_, grad, NULL, ns, orc_index(g, ic, jc, kc), q, dq, dsq);
	sum += orc_lc_compute_fed(p, q, dq);
      }
  return sum;
}

/* ---- advective fluxes of nf components: src/advection.c:538-629 (order 1), 770-893 (2), 946-1141 (3);
 * extent x in [1, N], y, z in [0, N] ---------------------------------------------------------------------------- */

static double adv3(double u, double fd1, double fd2, double fd3) {
  const double a1 = -0.213933;
  const double a2 =  0.927865;
  const double a3 =  0.286067;
  return u*(a1*fd1 + a2*fd2 + a3*fd3);
}

void orc_advection_nf(const orc_geom_t * g, int order, int nf, const double * u, const double * field, double * flux) {
  const size_t ns = (size_t) orc_nsites(g);
  const int ys = g->nlocal[Z] + 2*g->nhalo;
  const int xs = ys*(g->nlocal[Y] + 2*g->nhalo);

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 0; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 0; kc <= g->nlocal[Z]; kc++) {
	const int i0 = orc_index(g, ic, jc, kc);
	const double u0[3] = {u[0*ns + i0], u[1*ns + i0], u[2*ns + i0]};
	const int off[4] = {-xs, +xs, +ys, +1};
	const int comp[4] = {X, X, Y, Z};
	for (int face = 0; face < 4; face++) {
	  const int o = off[face];
	  const int i1 = i0 + o;
	  double uf;
	  for (int n = 0; n < nf; n++) {
	    const double * f = field + (size_t) n*ns;
	    double * fl = flux + (size_t) (face*nf + n)*ns;
	    if (order == 1) {
	      int index = i0;
	      uf = 0.5*(u0[comp[face]] + u[comp[face]*ns + i1]);
	      if (face == 0) { if (uf > 0.0) index = i1; }
	      else           { if (uf < 0.0) index = i1; }
	      fl[i0] = uf*f[index];
	    }
	    else if (order == 2) {
	      if (face == 0) fl[i0] = 0.5*(u0[X] + u[0*ns + i1])*1*0.5*(f[i1] + f[i0]);
	      else           fl[i0] = 0.5*(u0[comp[face]] + u[comp[face]*ns + i1])*1*0.5*(f[i0] + f[i1]);
	    }
	    else if (order == 4) {
	      /* advection_le_4th, src/advection.c:1153-1262 */
	      const double a1 = (1.0/16.0), a2 = (9.0/16.0);
	      uf = 0.5*(u0[comp[face]] + u[comp[face]*ns + i1]);
	      if (face == 0) fl[i0] = uf*(- a1*f[i0 + 2*o] + a2*f[i0 + o] + a2*f[i0] - a1*f[i0 - o]);
	      else           fl[i0] = uf*(- a1*f[i0 - o] + a2*f[i0] + a2*f[i0 + o] - a1*f[i0 + 2*o]);
	    }
	    else {
	      uf = 0.5*1*(u0[comp[face]] + u[comp[face]*ns + i1]);
	      if (face == 0) {
		if (uf > 0.0) fl[i0] = adv3(uf, f[i0 + 2*o], f[i0 + o], f[i0]);
		else          fl[i0] = adv3(uf, f[i0 - o], f[i0], f[i0 + o]);
	      }
	      else {
		if (uf < 0.0) fl[i0] = adv3(uf, f[i0 + 2*o], f[i0 + o], f[i0]);
		else          fl[i0] = adv3(uf, f[i0 - o], f[i0], f[i0 + o]);
	      }
	    }
	  }
	}
      }
    }
  }
}

/* ---- beris_edw_kernel_v: src/blue_phase_beris_edwards.c:538-850 (no noise; all fluid) ------------------------ */

void orc_beris_edw_update(const orc_geom_t * g, double xi, double Gamma, const double * u, const double * hq,
			  const double * flux, double * q_) {
  const size_t ns = (size_t) orc_nsites(g);
  const int ys = g->nlocal[Z] + 2*g->nhalo;
  const int xs = ys*(g->nlocal[Y] + 2*g->nhalo);
  const double dt = 1.0;
  const double r3 = (1.0/3.0);

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	const int index = orc_index(g, ic, jc, kc);
	double q[3][3], w[3][3], d[3][3], omega[3][3], s[3][3];
	double trace_qw, tr;
	const int off[3] = {xs, ys, 1};

	q[X][X] = q_[XX*ns + index]; q[X][Y] = q_[XY*ns + index]; q[X][Z] = q_[XZ*ns + index];
	q[Y][X] = q[X][Y]; q[Y][Y] = q_[YY*ns + index]; q[Y][Z] = q_[YZ*ns + index];
	q[Z][X] = q[X][Z]; q[Z][Y] = q[Y][Z]; q[Z][Z] = 0.0 - q[X][X] - q[Y][Y];

	/* w[a][b] = d_b u_a (:600-680) */
	for (int ib = 0; ib < 3; ib++)
	  for (int ia = 0; ia < 3; ia++)
	    w[ia][ib] = 0.5*(u[(size_t) ia*ns + index + off[ib]] - u[(size_t) ia*ns + index - off[ib]]);

	tr = r3*(w[X][X] + w[Y][Y] + w[Z][Z]);
	w[X][X] -= tr; w[Y][Y] -= tr; w[Z][Z] -= tr;

	trace_qw = 0.0;
	for (int ia = 0; ia < 3; ia++) {
	  for (int ib = 0; ib < 3; ib++) {
	    trace_qw += q[ia][ib]*w[ib][ia];
	    d[ia][ib] = 0.5*(w[ia][ib] + w[ib][ia]);
	    omega[ia][ib] = 0.5*(w[ia][ib] - w[ib][ia]);
	  }
	}
	for (int ia = 0; ia < 3; ia++) {
	  for (int ib = 0; ib < 3; ib++) {
	    s[ia][ib] = -2.0*xi*(q[ia][ib] + r3*d_[ia][ib])*trace_qw;
	    for (int id = 0; id < 3; id++) {
	      s[ia][ib] += (xi*d[ia][id] + omega[ia][id])*(q[id][ib] + r3*d_[id][ib])
		+ (q[ia][id] + r3*d_[ia][id])*(xi*d[id][ib] - omega[id][ib]);
	    }
	  }
	}

	{
	  const int a[NQAB] = {X, X, X, Y, Y}, b[NQAB] = {X, Y, Z, Y, Z};
	  const double * fw = flux + (size_t) 0*NQAB*ns, * fe = flux + (size_t) 1*NQAB*ns;
	  const double * fy = flux + (size_t) 2*NQAB*ns, * fz = flux + (size_t) 3*NQAB*ns;
	  for (int n = 0; n < NQAB; n++) {
	    double qn = q[a[n]][b[n]];
	    qn += dt*(s[a[n]][b[n]] + 0.0 + Gamma*hq[(size_t) n*ns + index]
		      - fe[(size_t) n*ns + index] + fw[(size_t) n*ns + index]
		      - fy[(size_t) n*ns + index] + fy[(size_t) n*ns + index - ys]
		      - fz[(size_t) n*ns + index] + fz[(size_t) n*ns + index - 1]);
	    q_[(size_t) n*ns + index] = qn;
	  }
	}
      }
    }
  }
}

/* ---- one liquid-crystal time step, reference driver order src/ludwig.c:528-860 ------------------------------
 * hydro_f_zero; field_halo(q); field_grad_compute(q_grad); phi_force_calculation (pth_stress_compute with
 * fe_lc_stress_v + pth_force_fluid_driver); hydro_u_halo; beris_edw_update (advection_x, no-normal-flux masks
 * all 1, beris_edw_h_driver, beris_edw_update_driver); hydro_u_zero; lb_collide; lb_halo; lb_propagation. */

void orc_lc_step(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		 const orc_lc_param_t * p, int adv_order, int nsteps,
		 double * f, double * q, double * u, double * rho, double * force,
		 double * qgrad, double * qdelsq) {

  const size_t ns = (size_t) orc_nsites(g);
  const double zero[3] = {0.0, 0.0, 0.0};
  double * fprime = (double *) calloc(ns*m->nvel, sizeof(double));
  double * str = (double *) calloc(ns*9, sizeof(double));
  double * flux = (double *) calloc(ns*4*NQAB, sizeof(double));
  double * hq = (double *) calloc(ns*NQAB, sizeof(double));
  assert(fprime && str && flux && hq);
  memcpy(fprime, f, ns*m->nvel*sizeof(double));

  for (int n = 0; n < nsteps; n++) {
    orc_field_set(g, 3, force, zero);
    orc_field_halo(g, NQAB, q);
    if (p->grad_2d5) orc_grad_2d_5pt(g, NQAB, q, qgrad, qdelsq);       /* fd_gradient_calculation 2d_5pt_fluid */
    else             orc_grad_7pt(g, NQAB, q, qgrad, qdelsq);
    orc_lc_stress(g, p, q, qgrad, qdelsq, str);
    orc_force_divergence(g, str, force);
    orc_field_halo(g, 3, u);
    orc_advection_nf(g, adv_order, NQAB, u, q, flux);
    orc_lc_mol_field(g, p, q, qgrad, qdelsq, hq);
    orc_beris_edw_update(g, p->xi, p->Gamma, u, hq, flux, q);
    orc_field_set(g, 3, u, zero);
    orc_collide(g, m, cp, NULL, 0, f, force, rho, u);
    orc_lb_halo(g, m, 1, 0, f);
    orc_propagation(g, m, 1, f, fprime);
    memcpy(f, fprime, ns*m->nvel*sizeof(double));
  }
  free(fprime); free(str); free(flux); free(hq);
}
