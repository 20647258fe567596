/*
 * lb_oracle_le.c -- TEST INFRASTRUCTURE ONLY (see lb_oracle.h for status and layout).
 *
 * CPU restatement of what Lees-Edwards sliding periodic planes add to the reference's time step
 * (SURVEY 8f row f1): the plane / buffer geometry of src/leesedwards.c, the interpolation of phi and u into
 * the buffer planes, the gradients in the buffer region, the flux form of the thermodynamic force with its
 * per-plane correction, the Cahn-Hilliard flux fix, and the re-projection / displacement / interpolation of
 * the plane-crossing distributions.  Steady shear, no decomposition in y (cartsz[Y] == 1).
 * Written in the reference's order of floating-point operations: must be compiled -ffp-contract=off.
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity status: PINNED bit-for-bit to the unmodified reference compiled here (tests/test_le_oracle.py) and to
 * the printed statistics of the reference's regression logs serial-le3d-st5/6/7.log.
 */

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lb_oracle.h"

enum {X = 0, Y = 1, Z = 2};

static int imin_(int a, int b) { return a < b ? a : b; }
static int imax_(int a, int b) { return a > b ? a : b; }

/* ---- plane and buffer geometry ----------------------------------------------------------------------- */

/* lees_edw_nxbuffer: src/leesedwards.c:405 */
int orc_le_nxbuffer(const orc_geom_t * g) { return 2*g->nhalo*g->le_nplanes; }

/* lees_edw_plane_location, src/leesedwards.c:615-634, with dx_sep = ntotal[X]/nplanes, dx_min = dx_sep/2
 * (:256-257), one domain (offset 0) */
int orc_le_plane_location(const orc_geom_t * g, int np) {
  const double dx_sep = 1.0*g->nlocal[X]/g->le_nplanes;
  const double dx_min = 0.5*dx_sep;
  int ix = dx_min + np*dx_sep - 0;
  return ix;
}

/* lees_edw_ibuff_to_real, src/leesedwards.c:1008-1022 */
int orc_le_ibuff_to_real(const orc_geom_t * g, int ib) {
  const int nh = g->nhalo;
  int p = ib/(2*nh);
  int ic = orc_le_plane_location(g, p) - (nh - 1);
  return ic + ib % (2*nh);
}

/* lees_edw_ic_to_buff, src/leesedwards.c:1030-1065 */
int orc_le_ic_to_buff(const orc_geom_t * g, int ic, int di) {
  int ib = ic + di;
  if (g->le_nplanes > 0) {
    const int nh = g->nhalo;
    int p = ic/(g->nlocal[X]/g->le_nplanes);
    int ip;
    p = imax_(0, imin_(p, g->le_nplanes - 1));
    ip = orc_le_plane_location(g, p) - (nh - 1);
    if (di > 0 && (ic >= ip && ic < ip + nh) && (ic + di >= ip + nh)) {
      return g->nlocal[X] + (1 + 2*p)*nh + (ic - ip + 1) + di;
    }
    ip = orc_le_plane_location(g, p) + 1;
    if (di < 0 && (ic >= ip && ic < ip + nh) && (ic + di < ip)) {
      return g->nlocal[X] + (2 + 2*p)*nh + (ic - ip + 1) + di;
    }
  }
  return ib;
}

/* lees_edw_buffer_duy, src/leesedwards.c:1082-1095 */
static int le_buffer_duy(const orc_geom_t * g, int ib) {
  return (ib % (2*g->nhalo) < g->nhalo) ? -1 : +1;
}

/* lees_edw_buffer_displacement (steady shear), src/leesedwards.c:649-673 */
double orc_le_buffer_displacement(const orc_geom_t * g, const orc_le_t * le, int ib, double t) {
  double tle;
  if (t < 0.0) t = 0.0;
  tle = t - le->time0;
  return tle*le->uy*le_buffer_duy(g, ib);
}

/* ---- field_leesedwards, src/field.c:418-510 (serial branch :460-505) ----------------------------------- */

void orc_le_field(const orc_geom_t * g, const orc_le_t * le, double t, int nf, double * data) {

  const size_t ns = (size_t) orc_nsites(g);
  const int nh = g->nhalo;
  const int * nl = g->nlocal;
  const int nxb = orc_le_nxbuffer(g);
  const int ib0 = nl[X] + nh + 1;
  const double r6 = (1.0/6.0);
  const double ltot_y = 1.0*nl[Y];

  for (int ib = 0; ib < nxb; ib++) {
    int ic = orc_le_ibuff_to_real(g, ib);
    double dy = orc_le_buffer_displacement(g, le, ib, t + 0.0);    /* lees_edw_buffer_dy(le, ib, 0.0, &dy) */
    int jdy;
    double fr;
    dy = fmod(dy, ltot_y);
    jdy = floor(dy);
    fr = 1.0 - (dy - jdy);

    for (int jc = 1 - nh; jc <= nl[Y] + nh; jc++) {
      int j0 = 1 + (jc - jdy - 3 + 2*nl[Y]) % nl[Y];
      int j1 = 1 + j0 % nl[Y];
      int j2 = 1 + j1 % nl[Y];
      int j3 = 1 + j2 % nl[Y];
      for (int kc = 1 - nh; kc <= nl[Z] + nh; kc++) {
	int index  = orc_index(g, ib0 + ib, jc, kc);
	int index0 = orc_index(g, ic, j0, kc);
	int index1 = orc_index(g, ic, j1, kc);
	int index2 = orc_index(g, ic, j2, kc);
	int index3 = orc_index(g, ic, j3, kc);
	for (int n = 0; n < nf; n++) {
	  double * d = data + (size_t) n*ns;
	  d[index] =
	    -  r6*fr*(fr-1.0)*(fr-2.0)*d[index0]
	    + 0.5*(fr*fr-1.0)*(fr-2.0)*d[index1]
	    - 0.5*fr*(fr+1.0)*(fr-2.0)*d[index2]
	    +        r6*fr*(fr*fr-1.0)*d[index3];
	}
      }
    }
  }
}

/* ---- hydro_lees_edwards, src/hydro.c:350-440 (serial branch :386-430) ------------------------------------ */

void orc_le_hydro(const orc_geom_t * g, const orc_le_t * le, double t, int nhcomm, double * u) {

  const size_t ns = (size_t) orc_nsites(g);
  const int nh = g->nhalo;
  const int * nl = g->nlocal;
  const int nxb = orc_le_nxbuffer(g);
  const int ib0 = nl[X] + nh + 1;
  const double ltot_y = 1.0*nl[Y];

  for (int ib = 0; ib < nxb; ib++) {
    int ic = orc_le_ibuff_to_real(g, ib);
    double ule[3] = {0.0, le->uy*le_buffer_duy(g, ib), 0.0};       /* lees_edw_buffer_du, src/leesedwards.c:985-1000 */
    double dy = orc_le_buffer_displacement(g, le, ib, t + 1.0);    /* lees_edw_buffer_dy(le, ib, 1.0, &dy) */
    int jdy;
    double fr;
    dy = fmod(dy, ltot_y);
    jdy = floor(dy);
    fr = dy - jdy;

    for (int jc = 1 - nh; jc <= nl[Y] + nh; jc++) {
      int j1 = 1 + (jc - jdy - 2 + 2*nl[Y]) % nl[Y];
      int j2 = 1 + j1 % nl[Y];
      for (int kc = 1 - nhcomm; kc <= nl[Z] + nhcomm; kc++) {
	int index0 = orc_index(g, ib0 + ib, jc, kc);
	int index1 = orc_index(g, ic, j1, kc);
	int index2 = orc_index(g, ic, j2, kc);
	for (int ia = 0; ia < 3; ia++) {
	  u[(size_t) ia*ns + index0] = ule[ia] + u[(size_t) ia*ns + index1]*fr + u[(size_t) ia*ns + index2]*(1.0 - fr);
	}
      }
    }
  }
}

/* ---- grad_3d_27pt_fluid_le, src/gradient_3d_27pt_fluid.c:375-651 ------------------------------------------
 * The same 27-point stencil at the nextra buffer planes either side of each plane. */

static void grad27_at(const double * field, size_t ns, int ys, int indexm1, int index, int indexp1,
		      double * grad, double * delsq) {
  const double r9 = (1.0/9.0);
  grad[0*ns + index] = 0.5*r9*
    (+ field[indexp1-ys-1] - field[indexm1-ys-1]
     + field[indexp1-ys  ] - field[indexm1-ys  ]
     + field[indexp1-ys+1] - field[indexm1-ys+1]
     + field[indexp1   -1] - field[indexm1   -1]
     + field[indexp1     ] - field[indexm1     ]
     + field[indexp1   +1] - field[indexm1   +1]
     + field[indexp1+ys-1] - field[indexm1+ys-1]
     + field[indexp1+ys  ] - field[indexm1+ys  ]
     + field[indexp1+ys+1] - field[indexm1+ys+1]);
  grad[1*ns + index] = 0.5*r9*
    (+ field[indexm1+ys-1] - field[indexm1-ys-1]
     + field[indexm1+ys  ] - field[indexm1-ys  ]
     + field[indexm1+ys+1] - field[indexm1-ys+1]
     + field[index  +ys-1] - field[index  -ys-1]
     + field[index  +ys  ] - field[index  -ys  ]
     + field[index  +ys+1] - field[index  -ys+1]
     + field[indexp1+ys-1] - field[indexp1-ys-1]
     + field[indexp1+ys  ] - field[indexp1-ys  ]
     + field[indexp1+ys+1] - field[indexp1-ys+1]);
  grad[2*ns + index] = 0.5*r9*
    (+ field[indexm1-ys+1] - field[indexm1-ys-1]
     + field[indexm1   +1] - field[indexm1   -1]
     + field[indexm1+ys+1] - field[indexm1+ys-1]
     + field[index  -ys+1] - field[index  -ys-1]
     + field[index     +1] - field[index     -1]
     + field[index  +ys+1] - field[index  +ys-1]
     + field[indexp1-ys+1] - field[indexp1-ys-1]
     + field[indexp1   +1] - field[indexp1   -1]
     + field[indexp1+ys+1] - field[indexp1+ys-1]);
  delsq[index] = r9*
    (+ field[indexm1-ys-1] + field[indexm1-ys  ] + field[indexm1-ys+1]
     + field[indexm1   -1] + field[indexm1     ] + field[indexm1   +1]
     + field[indexm1+ys-1] + field[indexm1+ys  ] + field[indexm1+ys+1]
     + field[index  -ys-1] + field[index  -ys  ] + field[index  -ys+1]
     + field[index     -1]                       + field[index     +1]
     + field[index  +ys-1] + field[index  +ys  ] + field[index  +ys+1]
     + field[indexp1-ys-1] + field[indexp1-ys  ] + field[indexp1-ys+1]
     + field[indexp1   -1] + field[indexp1     ] + field[indexp1   +1]
     + field[indexp1+ys-1] + field[indexp1+ys  ] + field[indexp1+ys+1]
     - 26.0*field[index]);
}

/* grad_3d_7pt_fluid_le, src/gradient_3d_7pt_fluid.c:317-440: the same buffer-plane triples with the 7-point formulae */
static void grad7_at(const double * field, size_t ns, int ys, int indexm1, int index, int indexp1, double * grad, double * delsq) {
  grad[0*ns + index] = 0.5*(field[indexp1] - field[indexm1]);
  grad[1*ns + index] = 0.5*(field[index + ys] - field[index - ys]);
  grad[2*ns + index] = 0.5*(field[index + 1] - field[index - 1]);
  delsq[index] = field[indexp1] + field[indexm1] + field[index + ys] + field[index - ys] + field[index + 1] + field[index - 1]
    - 6.0*field[index];
}

void orc_le_grad7_buffer(const orc_geom_t * g, int nextra, const double * field, double * grad, double * delsq) {
  const size_t ns = (size_t) orc_nsites(g);
  const int * nl = g->nlocal;
  const int ys = nl[Z] + 2*g->nhalo;
  for (int np = 0; np < g->le_nplanes; np++) {
    int ic = orc_le_plane_location(g, np);
    for (int nh = 1; nh <= nextra; nh++) {
      int ic0 = orc_le_ic_to_buff(g, ic, nh - 1), ic1 = orc_le_ic_to_buff(g, ic, nh), ic2 = orc_le_ic_to_buff(g, ic, nh + 1);
      for (int jc = 1 - nextra; jc <= nl[Y] + nextra; jc++)
	for (int kc = 1 - nextra; kc <= nl[Z] + nextra; kc++)
	  grad7_at(field, ns, ys, orc_index(g, ic0, jc, kc), orc_index(g, ic1, jc, kc), orc_index(g, ic2, jc, kc), grad, delsq);
    }
    ic += 1;
    for (int nh = 1; nh <= nextra; nh++) {
      int ic2 = orc_le_ic_to_buff(g, ic, -nh + 1), ic1 = orc_le_ic_to_buff(g, ic, -nh), ic0 = orc_le_ic_to_buff(g, ic, -nh - 1);
      for (int jc = 1 - nextra; jc <= nl[Y] + nextra; jc++)
	for (int kc = 1 - nextra; kc <= nl[Z] + nextra; kc++)
	  grad7_at(field, ns, ys, orc_index(g, ic0, jc, kc), orc_index(g, ic1, jc, kc), orc_index(g, ic2, jc, kc), grad, delsq);
    }
  }
}

void orc_le_grad_buffer(const orc_geom_t * g, int nextra, const double * field, double * grad, double * delsq) {

  const size_t ns = (size_t) orc_nsites(g);
  const int * nl = g->nlocal;
  const int ys = nl[Z] + 2*g->nhalo;

  for (int np = 0; np < g->le_nplanes; np++) {
    int ic = orc_le_plane_location(g, np);

    /* looking across in the +ve x-direction (:427-533) */
    for (int nh = 1; nh <= nextra; nh++) {
      int ic0 = orc_le_ic_to_buff(g, ic, nh - 1);
      int ic1 = orc_le_ic_to_buff(g, ic, nh);
      int ic2 = orc_le_ic_to_buff(g, ic, nh + 1);
      for (int jc = 1 - nextra; jc <= nl[Y] + nextra; jc++) {
	for (int kc = 1 - nextra; kc <= nl[Z] + nextra; kc++) {
	  grad27_at(field, ns, ys, orc_index(g, ic0, jc, kc), orc_index(g, ic1, jc, kc), orc_index(g, ic2, jc, kc),
		    grad, delsq);
	}
      }
    }

    /* looking across the plane in the -ve x-direction (:535-641) */
    ic += 1;
    for (int nh = 1; nh <= nextra; nh++) {
      int ic2 = orc_le_ic_to_buff(g, ic, -nh + 1);
      int ic1 = orc_le_ic_to_buff(g, ic, -nh);
      int ic0 = orc_le_ic_to_buff(g, ic, -nh - 1);
      for (int jc = 1 - nextra; jc <= nl[Y] + nextra; jc++) {
	for (int kc = 1 - nextra; kc <= nl[Z] + nextra; kc++) {
	  grad27_at(field, ns, ys, orc_index(g, ic0, jc, kc), orc_index(g, ic1, jc, kc), orc_index(g, ic2, jc, kc),
		    grad, delsq);
	}
      }
    }
  }
}

/* ---- phi_force_flux, src/phi_force.c:289-345: fluxes :360-440, per-plane fix :595-673, divergence :452-495 -- */

static void symm_stress(const orc_symm_param_t * sp, const double * phi, const double * grad,
			const double * delsq_, size_t ns, int index, double s[3][3]) {
  /* fe_symm_str, src/symmetric.c:333-361 */
  const double kappa = sp->kappa;
  double ph = phi[index];
  double delsq = delsq_[index];
  double gr[3] = {grad[0*ns + index], grad[1*ns + index], grad[2*ns + index]};
  double p0 = 0.5*sp->a*ph*ph + 0.75*sp->b*ph*ph*ph*ph - kappa*ph*delsq
    - 0.5*kappa*(gr[X]*gr[X] + gr[Y]*gr[Y] + gr[Z]*gr[Z]);
  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      double d = (ia == ib);
      s[ia][ib] = p0*d + kappa*gr[ia]*gr[ib];
    }
  }
}

void orc_le_phi_force(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi,
		      const double * grad, const double * delsq, double * force) {

  const size_t ns = (size_t) orc_nsites(g);            /* phi, grad, delsq, force: LE size */
  const size_t nf = (size_t) orc_nsites_lb(g);         /* flux work arrays: cs_nsites (:304, 383) */
  const int * nl = g->nlocal;
  double * fluxe = (double *) calloc(3*nf, sizeof(double));
  double * fluxw = (double *) calloc(3*nf, sizeof(double));
  double * fluxy = (double *) calloc(3*nf, sizeof(double));
  double * fluxz = (double *) calloc(3*nf, sizeof(double));
  assert(fluxe && fluxw && fluxy && fluxz);

  /* phi_force_compute_fluxes :360-440 */
  #pragma omp parallel for schedule(static)
  for (int ic = 1; ic <= nl[X]; ic++) {
    int icm1 = orc_le_ic_to_buff(g, ic, -1);
    int icp1 = orc_le_ic_to_buff(g, ic, +1);
    for (int jc = 0; jc <= nl[Y]; jc++) {
      for (int kc = 0; kc <= nl[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double pth0[3][3], pth1[3][3];
	symm_stress(sp, phi, grad, delsq, ns, index, pth0);
	symm_stress(sp, phi, grad, delsq, ns, orc_index(g, icm1, jc, kc), pth1);
	for (int ia = 0; ia < 3; ia++) fluxw[ia*nf + index] = 0.5*(pth1[ia][X] + pth0[ia][X]);
	symm_stress(sp, phi, grad, delsq, ns, orc_index(g, icp1, jc, kc), pth1);
	for (int ia = 0; ia < 3; ia++) fluxe[ia*nf + index] = 0.5*(pth1[ia][X] + pth0[ia][X]);
	symm_stress(sp, phi, grad, delsq, ns, orc_index(g, ic, jc + 1, kc), pth1);
	for (int ia = 0; ia < 3; ia++) fluxy[ia*nf + index] = 0.5*(pth1[ia][Y] + pth0[ia][Y]);
	symm_stress(sp, phi, grad, delsq, ns, orc_index(g, ic, jc, kc + 1), pth1);
	for (int ia = 0; ia < 3; ia++) fluxz[ia*nf + index] = 0.5*(pth1[ia][Z] + pth0[ia][Z]);
      }
    }
  }

  /* phi_force_flux_fix_local :595-673 (sums in the reference's jc, kc order) */
  {
    const double ra = 0.5/((1.0*nl[Y])*(1.0*nl[Z]));
    for (int ip = 0; ip < g->le_nplanes; ip++) {
      int ic = orc_le_plane_location(g, ip);
      double fbar[3] = {0.0, 0.0, 0.0};
      for (int jc = 1; jc <= nl[Y]; jc++) {
	for (int kc = 1; kc <= nl[Z]; kc++) {
	  int index = orc_index(g, ic, jc, kc);
	  int index1 = orc_index(g, ic + 1, jc, kc);
	  for (int ia = 0; ia < 3; ia++) fbar[ia] += - fluxe[ia*nf + index] + fluxw[ia*nf + index1];
	}
      }
      for (int jc = 1; jc <= nl[Y]; jc++) {
	for (int kc = 1; kc <= nl[Z]; kc++) {
	  int index = orc_index(g, ic, jc, kc);
	  int index1 = orc_index(g, ic + 1, jc, kc);
	  for (int ia = 0; ia < 3; ia++) {
	    fluxe[ia*nf + index] += ra*fbar[ia];
	    fluxw[ia*nf + index1] -= ra*fbar[ia];
	  }
	}
      }
    }
  }

  /* phi_force_flux_divergence :452-495, hydro_f_local_add */
  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= nl[X]; ic++) {
    for (int jc = 1; jc <= nl[Y]; jc++) {
      for (int kc = 1; kc <= nl[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	int indexj = orc_index(g, ic, jc - 1, kc);
	int indexk = orc_index(g, ic, jc, kc - 1);
	for (int ia = 0; ia < 3; ia++) {
	  double fo = -(+ fluxe[ia*nf + index] - fluxw[ia*nf + index]
			+ fluxy[ia*nf + index] - fluxy[ia*nf + indexj]
			+ fluxz[ia*nf + index] - fluxz[ia*nf + indexk]);
	  force[(size_t) ia*ns + index] += fo;
	}
      }
    }
  }

  free(fluxz); free(fluxy); free(fluxw); free(fluxe);
}

/* ---- phi_ch_le_fix_fluxes, src/phi_cahn_hilliard.c:613-745 (serial branch :648-735); the displacement is
 * lees_edw_plane_dy = t*uy (src/leesedwards.c:940-953) ------------------------------------------------------ */

void orc_le_fix_fluxes(const orc_geom_t * g, const orc_le_t * le, double t, double * flux) {

  const size_t ns = (size_t) orc_nsites(g);
  const int * nl = g->nlocal;
  double * fw = flux + 0*ns;
  double * fe = flux + 1*ns;
  const double ltot_y = 1.0*nl[Y];
  double * bufferw = (double *) malloc((size_t) nl[Y]*nl[Z]*sizeof(double));
  double * buffere = (double *) malloc((size_t) nl[Y]*nl[Z]*sizeof(double));
  assert(bufferw && buffere);

  for (int ip = 0; ip < g->le_nplanes; ip++) {
    int ic = orc_le_plane_location(g, ip);
    double dy, fr;
    int jdy;

    /* looking up */
    dy = t*le->uy;
    dy = fmod(+dy, ltot_y);
    jdy = floor(dy);
    fr = dy - jdy;
    for (int jc = 1; jc <= nl[Y]; jc++) {
      int j1 = 1 + (jc - jdy - 2 + 2*nl[Y]) % nl[Y];
      int j2 = 1 + j1 % nl[Y];
      for (int kc = 1; kc <= nl[Z]; kc++) {
	bufferw[nl[Z]*(jc - 1) + (kc - 1)] =
	  fw[orc_index(g, ic + 1, j1, kc)]*fr + fw[orc_index(g, ic + 1, j2, kc)]*(1.0 - fr);
      }
    }

    /* looking down */
    dy = t*le->uy;
    dy = fmod(-dy, ltot_y);
    jdy = floor(dy);
    fr = dy - jdy;
    for (int jc = 1; jc <= nl[Y]; jc++) {
      int j1 = 1 + (jc - jdy - 2 + 2*nl[Y]) % nl[Y];
      int j2 = 1 + j1 % nl[Y];
      for (int kc = 1; kc <= nl[Z]; kc++) {
	buffere[nl[Z]*(jc - 1) + (kc - 1)] =
	  fe[orc_index(g, ic, j1, kc)]*fr + fe[orc_index(g, ic, j2, kc)]*(1.0 - fr);
      }
    }

    /* average */
    for (int jc = 1; jc <= nl[Y]; jc++) {
      for (int kc = 1; kc <= nl[Z]; kc++) {
	int i1 = nl[Z]*(jc - 1) + (kc - 1);
	int index = orc_index(g, ic, jc, kc);
	fe[index] = 0.5*(fe[index] + bufferw[i1]);
	index = orc_index(g, ic + 1, jc, kc);
	fw[index] = 0.5*(fw[index] + buffere[i1]);
      }
    }
  }
  free(bufferw);
  free(buffere);
}

/* ---- lb_data_apply_le_boundary_conditions, src/model_le.c:78-180 ------------------------------------------
 * reproject :264-345, displace (serial kernel) :358-400, interpolate :584-640; buffer order le_ibuf :232-252 */

void orc_le_lb_bc(const orc_geom_t * g, const orc_model_t * m, const orc_le_t * le, double t,
		  int ndist, double * f) {

  const size_t ns = (size_t) orc_nsites_lb(g);
  const int * nl = g->nlocal;
  const int nplane = g->le_nplanes;
  int prop[2][9];
  int nprop = 0;
  if (nplane == 0) return;

  { int ip = 0;
    for (int p = 1; p < m->nvel; p++) if (m->cv[p][X] == +1) prop[0][ip++] = p;
    ip = 0;
    for (int p = 1; p < m->nvel; p++) if (m->cv[p][X] == -1) prop[1][ip++] = p;
    nprop = ip; }

  const size_t nxdist = (size_t) ndist*nprop*(nl[Y] + 1)*nl[Z];
  const size_t nxbuff = 2*nplane*nxdist;
  double * sbuff = (double *) calloc(nxbuff, sizeof(double));
  double * rbuff = (double *) calloc(nxbuff, sizeof(double));
  assert(sbuff && rbuff);

#define IBUF(jc, kc, iplane, iside, n, p) \
  ((size_t) (iside)*nxdist*nplane + (size_t) ((p) + nprop*((n) + ndist*((iplane) + nplane*((kc) - 1 + nl[Z]*((jc) - 1))))))

  const double cs2 = (1.0/3.0);
  const double rcs2 = 1.0/cs2;
  const double ltot_y = 1.0*nl[Y];
  /* lees_edw_plane_uy_now (steady): uy; lees_edw_buffer_displacement(le, nhalo, t): ib = nhalo has duy = +1 */
  const double dy_le = orc_le_buffer_displacement(g, le, g->nhalo, t);

  /* reproject */
  for (int ix = 0; ix < 2*nplane; ix++) {
    int iplane = ix/2;
    int iside = ix % 2;
    int cx = 1 - 2*iside;
    int ic = iside + orc_le_plane_location(g, iplane);
    double du[3] = {0.0, 0.0, 0.0};
    du[Y] = le->uy;
    du[Y] = -1.0*cx*du[Y];
    for (int jc = 1; jc <= nl[Y]; jc++) {
      for (int kc = 1; kc <= nl[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	for (int n = 0; n < ndist; n++) {
	  double rho = 0.0;
	  double gv[3] = {0.0, 0.0, 0.0};
	  double ds[3][3];
	  /* lb_0th_moment / lb_1st_moment: src/lb_data.c:1589-1608, 1647-1672 (sums in p order) */
	  for (int p = 0; p < m->nvel; p++) rho += f[(size_t) (n*m->nvel + p)*ns + index];
	  for (int p = 0; p < m->nvel; p++) {
	    for (int ia = 0; ia < 3; ia++) gv[ia] += m->cv[p][ia]*f[(size_t) (n*m->nvel + p)*ns + index];
	  }
	  for (int ia = 0; ia < 3; ia++)
	    for (int ib = 0; ib < 3; ib++)
	      ds[ia][ib] = (gv[ia]*du[ib] + du[ia]*gv[ib] + rho*du[ia]*du[ib]);
	  for (int ip = 0; ip < nprop; ip++) {
	    int p = prop[iside][ip];
	    double udotc = du[Y]*m->cv[p][Y];
	    double sdotq = 0.0;
	    for (int ia = 0; ia < 3; ia++) {
	      for (int ib = 0; ib < 3; ib++) {
		double dab = cs2*(ia == ib);
		double q = (m->cv[p][ia]*m->cv[p][ib] - dab);
		sdotq += ds[ia][ib]*q;
	      }
	    }
	    {
	      double fp = f[(size_t) (n*m->nvel + p)*ns + index];
	      fp += m->wv[p]*(rho*udotc*rcs2 + 0.5*sdotq*rcs2*rcs2);
	      sbuff[IBUF(jc, kc, iplane, iside, n, ip)] = fp;
	    }
	  }
	}
      }
    }
  }

  /* displace (integer part) */
  for (int ix = 0; ix < 2*nplane; ix++) {
    int iplane = ix/2;
    int iside = ix % 2;
    int cx = 1 - 2*iside;
    int dj = floor(fmod(dy_le*cx, ltot_y));
    for (int jc = 1; jc <= nl[Y] + 1; jc++) {
      int js = 1 + (jc + dj - 1 + 2*nl[Y]) % nl[Y];
      for (int kc = 1; kc <= nl[Z]; kc++)
	for (int n = 0; n < ndist; n++)
	  for (int ip = 0; ip < nprop; ip++)
	    rbuff[IBUF(jc, kc, iplane, iside, n, ip)] = sbuff[IBUF(js, kc, iplane, iside, n, ip)];
    }
  }

  /* interpolate (fractional part) */
  for (int ix = 0; ix < 2*nplane; ix++) {
    int iplane = ix/2;
    int iside = ix % 2;
    int cx = 1 - 2*iside;
    int ic = iside + orc_le_plane_location(g, iplane);
    double dy = fmod(dy_le*cx, ltot_y);
    int jdy = floor(dy);
    double fr = dy - jdy;
    for (int jc = 1; jc <= nl[Y]; jc++) {
      for (int kc = 1; kc <= nl[Z]; kc++) {
	int index0 = orc_index(g, ic, jc, kc);
	for (int n = 0; n < ndist; n++) {
	  for (int ip = 0; ip < nprop; ip++) {
	    int p = prop[iside][ip];
	    f[(size_t) (n*m->nvel + p)*ns + index0] =
	      (1.0 - fr)*rbuff[IBUF(jc, kc, iplane, iside, n, ip)] + fr*rbuff[IBUF(jc + 1, kc, iplane, iside, n, ip)];
	  }
	}
      }
    }
  }
#undef IBUF
  free(sbuff);
  free(rbuff);
}

/* ---- lb_le_init_shear_profile, src/model_le.c:652-714; lees_edw_steady_uy src/leesedwards.c:508-533;
 * lees_edw_shear_rate :759-769 ---------------------------------------------------------------------------- */

void orc_le_init_shear_profile(const orc_geom_t * g, const orc_model_t * m, const orc_le_t * le,
			       double rho0, double eta, double * f) {

  const size_t ns = (size_t) orc_nsites_lb(g);
  const int * nl = g->nlocal;
  const double dx_sep = 1.0*nl[X]/g->le_nplanes;
  const double dx_min = 0.5*dx_sep;
  double u[3] = {0.0, 0.0, 0.0};
  double gradu[3][3] = {{0.0}};
  const double cs2 = (1.0/3.0);
  const double rcs2 = 1.0/cs2;

  gradu[X][Y] = le->uy*g->le_nplanes/(1.0*nl[X]);       /* gammadot = uy*nplanetotal/ltot[X] */

  for (int ic = 1; ic <= nl[X]; ic++) {
    double xglobal = 0 + (double) ic - 0.5;
    int nplane = (int) ((dx_min + xglobal)/dx_sep);
    u[Y] = xglobal*gradu[X][Y] - le->uy*nplane;
    for (int jc = 1; jc <= nl[Y]; jc++) {
      for (int kc = 1; kc <= nl[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	for (int p = 0; p < m->nvel; p++) {
	  double cdotu = 0.0;
	  double sdotq = 0.0;
	  for (int i = 0; i < 3; i++) {
	    cdotu += m->cv[p][i]*u[i];
	    for (int j = 0; j < 3; j++) {
	      double dij = (i == j);
	      double qij = m->cv[p][i]*m->cv[p][j] - cs2*dij;
	      sdotq += (rho0*u[i]*u[j] - eta*gradu[i][j])*qij;
	    }
	  }
	  f[(size_t) p*ns + index] = m->wv[p]*(rho0 + rcs2*rho0*cdotu + 0.5*rcs2*rcs2*sdotq);
	}
      }
    }
  }
}

/* ---- one time step with planes, reference driver order src/ludwig.c:528-860 --------------------------------- */

/* symmetric_lb (two distributions) with planes, src/ludwig.c:528-860: phi_lb_to_field; field_halo(phi); field_grad_compute
 * (field_leesedwards, gradient, buffer-region gradients); lb_collide (lb_collision_binary); lb_data_apply_le_boundary_conditions
 * on both distributions; lb_halo; lb_propagation.  (Regression case: tests/regression/d3q19-short/serial-le2d-lb1.inp) */
void orc_le_step_lb2(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		     const orc_symm_param_t * sp, const orc_le_t * le, int tcurrent0, int nsteps,
		     double * f, double * phi, double * u, double * force, double * grad, double * delsq) {
  const size_t nsf = (size_t) orc_nsites_lb(g);
  const size_t nf = nsf*2*m->nvel;
  const double zero[3] = {0.0, 0.0, 0.0};
  double * fprime = (double *) calloc(nf, sizeof(double));
  assert(fprime);
  memcpy(fprime, f, nf*sizeof(double));

  for (int n = 0; n < nsteps; n++) {
    const int tcurrent = tcurrent0 + n + 1;
    const double tstep = 1.0*tcurrent;
    const double time = 1.0*(0 + tcurrent - 1.0);
    orc_field_set(g, 3, force, zero);
    orc_phi_lb_to_field(g, m, f, phi);
    orc_field_halo(g, 1, phi);
    orc_le_field(g, le, time, 1, phi);
    orc_grad_27pt(g, phi, grad, delsq);
    orc_le_grad_buffer(g, g->nhalo - 1, phi, grad, delsq);
    orc_field_set(g, 3, u, zero);
    orc_collide_binary(g, m, cp, sp, f, force, phi, grad, delsq, u);
    orc_le_lb_bc(g, m, le, tstep, 2, f);
    orc_lb_halo(g, m, 2, 0, f);
    orc_propagation(g, m, 2, f, fprime);
    memcpy(f, fprime, nf*sizeof(double));
  }
  free(fprime);
}

void orc_le_step(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		 const orc_symm_param_t * sp, const orc_le_t * le, int tcurrent0, int nsteps,
		 double * f, double * phi, double * u, double * rho, double * force,
		 double * grad, double * delsq) {

  const size_t ns = (size_t) orc_nsites(g);
  const size_t nsf = (size_t) orc_nsites_lb(g);
  const double zero[3] = {0.0, 0.0, 0.0};
  double * fprime = (double *) calloc(nsf*m->nvel, sizeof(double));
  double * flux = (double *) calloc(ns*4, sizeof(double));
  double * csum = NULL;     /* cahn_hilliard_options_conserve 1: pch->csum, zero when the phi_ch_t is created */
  assert(fprime && flux);
  memcpy(fprime, f, nsf*m->nvel*sizeof(double));
  if (sp->conserve == 1) { csum = (double *) calloc(ns, sizeof(double)); assert(csum); }

  for (int n = 0; n < nsteps; n++) {
    const int tcurrent = tcurrent0 + n + 1;              /* physics_control_next_step */
    const double tstep = 1.0*tcurrent;                   /* physics_control_timestep */
    const double time = 1.0*(0 + tcurrent - 1.0);        /* physics_control_time, t_start = 0 */

    orc_field_set(g, 3, force, zero);
    orc_field_halo(g, 1, phi);
    orc_le_field(g, le, time, 1, phi);                   /* field_grad_compute: src/field_grad.c:324 */
    if (sp->grad_7pt) {                                  /* fd_gradient_calculation 3d_7pt_fluid */
      orc_grad_7pt(g, 1, phi, grad, delsq);
      orc_le_grad7_buffer(g, g->nhalo - 1, phi, grad, delsq);
    }
    else {
      orc_grad_27pt(g, phi, grad, delsq);
      orc_le_grad_buffer(g, g->nhalo - 1, phi, grad, delsq);
    }
    orc_le_phi_force(g, sp, phi, grad, delsq, force);
    orc_field_halo(g, 3, u);
    orc_le_hydro(g, le, time, 1, u);
    orc_advection(g, sp->adv_order, u, phi, flux);
    orc_flux_mu(g, sp, phi, delsq, flux);
    orc_flux_mu_ext(g, sp, flux);
    orc_no_flux(g, NULL, flux);
    orc_le_fix_fluxes(g, le, time, flux);
    /* src/phi_cahn_hilliard.c:276-285: the same choice of update with or without planes */
    if (csum) orc_phi_update_conserve(g, flux, csum, phi);
    else      orc_phi_update(g, flux, phi);
    if (sp->conserve == 2) orc_phi_subtract_sum(g, sp->phi_init_sum, phi);

    orc_field_set(g, 3, u, zero);
    orc_collide(g, m, cp, NULL, 0, f, force, rho, u);
    orc_le_lb_bc(g, m, le, tstep, 1, f);
    orc_lb_halo(g, m, 1, 0, f);
    orc_propagation(g, m, 1, f, fprime);
    memcpy(f, fprime, nsf*m->nvel*sizeof(double));
  }

  free(fprime);
  free(flux);
  free(csum);
}
