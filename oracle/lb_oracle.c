/*
 * lb_oracle.c -- TEST INFRASTRUCTURE ONLY (see lb_oracle.h for status and layout).
 *
 * CPU restatement of the reference's LB hot path, written from the reference's behaviour, with
 * the reference's order of floating-point operations so that results are bit-identical with the
 * reference built -O2 -ffp-contract=off (this file must be compiled with -ffp-contract=off).
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lb_oracle.h"
#include "d3q19_tables.h"

enum {X = 0, Y = 1, Z = 2};

/* ---- geometry: src/coords.c:211-215, 617-631 ------------------------------------------- */

static void orc_nall(const orc_geom_t * g, int nall[3]) {
  for (int a = 0; a < 3; a++) nall[a] = g->nlocal[a] + 2*g->nhalo;
}

int orc_nsites_lb(const orc_geom_t * g) {
  int nall[3];
  orc_nall(g, nall);
  return nall[X]*nall[Y]*nall[Z];
}

/* lees_edw_nsites, src/leesedwards.c:485-495 (== cs_nsites without planes) */
int orc_nsites(const orc_geom_t * g) {
  int nall[3];
  orc_nall(g, nall);
  return (nall[X] + 2*g->nhalo*g->le_nplanes)*nall[Y]*nall[Z];
}

/* index offset of the x-neighbour di planes away, through the LE buffer region if a plane is crossed
 * (lees_edw_ic_to_buff, src/leesedwards.c:1030-1065); di*xs without planes */
static int xoff(const orc_geom_t * g, int ic, int di, int xs) {
  if (g->le_nplanes == 0) return di*xs;
  return (orc_le_ic_to_buff(g, ic, di) - ic)*xs;
}

int orc_index(const orc_geom_t * g, int ic, int jc, int kc) {
  int nall[3];
  orc_nall(g, nall);
  return ((ic + g->nhalo - 1)*nall[Y] + (jc + g->nhalo - 1))*nall[Z] + (kc + g->nhalo - 1);
}

/* ---- velocity sets and mode matrices ----------------------------------------------------
 * D3Q19: src/lb_d3q19.h:26-39, src/lb_d3q19.c:108-150
 * D3Q15: src/lb_d3q15.h:22-32, src/lb_d3q15.c:150-178
 * D3Q27: src/lb_d3q27.h:26-40, src/lb_d3q27.c:155-195
 * normalisers na: src/lb_d3q19.c:72-78; inverse mi: src/lb_data.c:640-646               */

static const signed char cv19[19][3] = {
  { 0,  0,  0},
  { 1,  1,  0}, { 1,  0,  1}, { 1,  0,  0}, { 1,  0, -1}, { 1, -1,  0}, { 0,  1,  1},
  { 0,  1,  0}, { 0,  1, -1}, { 0,  0,  1}, { 0,  0, -1}, { 0, -1,  1}, { 0, -1,  0},
  { 0, -1, -1}, {-1,  1,  0}, {-1,  0,  1}, {-1,  0,  0}, {-1,  0, -1}, {-1, -1,  0}};

static const signed char cv15[15][3] = {
  { 0,  0,  0},
  { 1,  1,  1}, { 1,  1, -1}, { 1,  0,  0}, { 1, -1,  1}, { 1, -1, -1}, { 0,  1,  0},
  { 0,  0,  1}, { 0,  0, -1}, { 0, -1,  0}, {-1,  1,  1}, {-1,  1, -1}, {-1,  0,  0},
  {-1, -1,  1}, {-1, -1, -1}};

int orc_model_create(int nvel, orc_model_t * model) {

  const double cs2 = (1.0/3.0);

  memset(model, 0, sizeof(*model));
  model->nvel = nvel;
  model->ndim = 3;

  if (nvel == 19) {
    for (int p = 0; p < 19; p++) {
      int c1 = 0;
      for (int a = 0; a < 3; a++) { model->cv[p][a] = cv19[p][a]; c1 += abs(cv19[p][a]); }
      model->wv[p] = (c1 == 0) ? 12.0/36.0 : (c1 == 1) ? 2.0/36.0 : 1.0/36.0;
    }
  }
  else if (nvel == 15) {
    for (int p = 0; p < 15; p++) {
      int c1 = 0;
      for (int a = 0; a < 3; a++) { model->cv[p][a] = cv15[p][a]; c1 += abs(cv15[p][a]); }
      model->wv[p] = (c1 == 0) ? 16.0/72.0 : (c1 == 1) ? 8.0/72.0 : 1.0/72.0;
    }
  }
  else if (nvel == 27) {
    /* p = 0 rest; then lexicographic (x slowest) over {-1,0,1}^3 skipping the rest vector */
    int p = 1;
    model->wv[0] = 64.0/216.0;
    for (int i = -1; i <= 1; i++)
      for (int j = -1; j <= 1; j++)
	for (int k = -1; k <= 1; k++) {
	  int c1 = abs(i) + abs(j) + abs(k);
	  if (c1 == 0) continue;
	  model->cv[p][X] = i; model->cv[p][Y] = j; model->cv[p][Z] = k;
	  model->wv[p] = (c1 == 1) ? 16.0/216.0 : (c1 == 2) ? 4.0/216.0 : 1.0/216.0;
	  p++;
	}
  }
  else {
    return -1;
  }

  for (int p = 0; p < nvel; p++) {
    double rho = 1.0;
    double cx = rho*model->cv[p][X];
    double cy = rho*model->cv[p][Y];
    double cz = rho*model->cv[p][Z];
    double (*ma)[27] = model->ma;

    ma[0][p] = rho;
    ma[1][p] = cx;
    ma[2][p] = cy;
    ma[3][p] = cz;
    ma[4][p] = cx*cx - cs2;
    ma[5][p] = cx*cy;
    ma[6][p] = cx*cz;
    ma[7][p] = cy*cy - cs2;
    ma[8][p] = cy*cz;
    ma[9][p] = cz*cz - cs2;

    if (nvel == 19) {
      double c2   = cx*cx + cy*cy + cz*cz;
      double chi1 = (2.0*c2 - 3.0)*(3.0*cz*cz - c2);
      double chi2 = (2.0*c2 - 3.0)*(cy*cy - cx*cx);
      double chi3 = 3.0*c2*c2 - 6.0*c2 + 1;
      ma[10][p] = chi1;
      ma[11][p] = chi1*cx;
      ma[12][p] = chi1*cy;
      ma[13][p] = chi1*cz;
      ma[14][p] = chi2;
      ma[15][p] = chi2*cx;
      ma[16][p] = chi2*cy;
      ma[17][p] = chi2*cz;
      ma[18][p] = chi3;
    }
    if (nvel == 15) {
      ma[10][p] = cx*cy*cz;
      ma[11][p] = 3.0*(cz*cz - cs2)*cx;
      ma[12][p] = 3.0*(cx*cx - cs2)*cy;
      ma[13][p] = 3.0*(cy*cy - cs2)*cz;
      ma[14][p] = 9.0*(cx*cx - cs2)*(cy*cy - cs2) - 3.0*(cz*cz - cs2);
    }
    if (nvel == 27) {
      ma[10][p] = 3.0*(cx*cx - cs2)*cy;
      ma[11][p] = 3.0*(cx*cx - cs2)*cz;
      ma[12][p] = 3.0*(cy*cy - cs2)*cz;
      ma[13][p] = 3.0*(cy*cy - cs2)*cx;
      ma[14][p] = 3.0*(cz*cz - cs2)*cx;
      ma[15][p] = 3.0*(cz*cz - cs2)*cy;
      ma[16][p] = cx*cy*cz;
      ma[17][p] = 9.0*(cx*cx - cs2)*(cy*cy - cs2);
      ma[18][p] = 9.0*(cy*cy - cs2)*(cz*cz - cs2);
      ma[19][p] = 9.0*(cz*cz - cs2)*(cx*cx - cs2);
      ma[20][p] = 9.0*(cx*cx - cs2)*cy*cz;
      ma[21][p] = 9.0*(cy*cy - cs2)*cz*cx;
      ma[22][p] = 9.0*(cz*cz - cs2)*cx*cy;
      ma[23][p] = 9.0*(cx*cx - cs2)*(cy*cy - cs2)*cz;
      ma[24][p] = 9.0*(cy*cy - cs2)*(cz*cz - cs2)*cx;
      ma[25][p] = 9.0*(cz*cz - cs2)*(cx*cx - cs2)*cy;
      ma[26][p] = 27.0*(cx*cx - cs2)*(cy*cy - cs2)*(cz*cz - cs2);
    }
  }

  for (int m = 0; m < nvel; m++) {
    double wip = 0.0;
    for (int p = 0; p < nvel; p++) wip += model->wv[p]*model->ma[m][p]*model->ma[m][p];
    model->na[m] = 1.0/wip;
  }

  for (int p = 0; p < nvel; p++) {
    for (int m = 0; m < nvel; m++) {
      double maba = model->ma[m][p];
      model->mi[p][m] = model->wv[p]*model->na[m]*maba;
    }
  }

  return 0;
}

/* ---- lb_propagation: src/propagation.c:153-200 (kernel), limits src/propagation.c:58-66 ---- */

void orc_propagation(const orc_geom_t * g, const orc_model_t * m, int ndist,
		     const double * f, double * fprime) {
  int nall[3];
  const int nh = g->nhalo;
  const size_t ns = (size_t) orc_nsites_lb(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const int xs = nall[Y]*nall[Z];

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1 - nh; jc <= g->nlocal[Y] + nh; jc++) {
      for (int kc = 1 - nh; kc <= g->nlocal[Z] + nh; kc++) {
	int index = orc_index(g, ic, jc, kc);
	int mask = (jc >= 1 && jc <= g->nlocal[Y] && kc >= 1 && kc <= g->nlocal[Z]);
	for (int n = 0; n < ndist; n++) {
	  for (int p = 0; p < m->nvel; p++) {
	    int indexp = index - mask*(m->cv[p][X]*xs + m->cv[p][Y]*ys + m->cv[p][Z]);
	    fprime[(size_t) (n*m->nvel + p)*ns + index] = f[(size_t) (n*m->nvel + p)*ns + indexp];
	  }
	}
      }
    }
  }
}

/* ---- halo exchange on one periodic rank --------------------------------------------------
 * lb_halo regions: src/lb_data.c:1183-1210; reduced set: src/lb_data.c:1224-1239;
 * self-message short circuit: src/lb_data.c:1009-1011; no neighbour across a non-periodic
 * boundary (src/lb_data.c:1160-1172): zeros are unpacked there.
 * field_halo regions: src/field.c:1329-1355.                                                */

typedef struct { int imin, imax, jmin, jmax, kmin, kmax; } lim_t;

static void halo_limits(const orc_geom_t * g, int depth, const int c[3], lim_t * s, lim_t * r) {
  const int * nl = g->nlocal;
  lim_t send = {1, nl[X], 1, nl[Y], 1, nl[Z]};
  lim_t recv = {1, nl[X], 1, nl[Y], 1, nl[Z]};

  if (c[X] == -1) send.imax = depth;
  if (c[X] == +1) send.imin = send.imax - (depth - 1);
  if (c[Y] == -1) send.jmax = depth;
  if (c[Y] == +1) send.jmin = send.jmax - (depth - 1);
  if (c[Z] == -1) send.kmax = depth;
  if (c[Z] == +1) send.kmin = send.kmax - (depth - 1);

  if (c[X] == +1) { recv.imin = 1 - depth;     recv.imax = 0; }
  if (c[X] == -1) { recv.imin = recv.imax + 1; recv.imax = recv.imax + depth; }
  if (c[Y] == +1) { recv.jmin = 1 - depth;     recv.jmax = 0; }
  if (c[Y] == -1) { recv.jmin = recv.jmax + 1; recv.jmax = recv.jmax + depth; }
  if (c[Z] == +1) { recv.kmin = 1 - depth;     recv.kmax = 0; }
  if (c[Z] == -1) { recv.kmin = recv.kmax + 1; recv.kmax = recv.kmax + depth; }
  *s = send;
  *r = recv;
}

void orc_lb_halo(const orc_geom_t * g, const orc_model_t * m, int ndist, int reduced, double * f) {

  const size_t ns = (size_t) orc_nsites_lb(g);

  for (int cx = -1; cx <= 1; cx++) {
    for (int cy = -1; cy <= 1; cy++) {
      for (int cz = -1; cz <= 1; cz++) {
	int c[3] = {cx, cy, cz};
	int mm = cx*cx + cy*cy + cz*cz;
	lim_t s, r;
	int absent;
	if (mm == 0) continue;
	/* No neighbour across a non-periodic boundary: nothing is received, but the reference still
	 * unpacks its (calloc'ed, never written) receive buffer, i.e. zeros arrive
	 * (src/lb_data.c:1010-1011 vs 1262-1266). */
	absent = ((cx && !g->periodic[X]) || (cy && !g->periodic[Y]) || (cz && !g->periodic[Z]));
	halo_limits(g, 1, c, &s, &r);

	for (int q = 0; q < m->nvel; q++) {
	  int dot = cx*m->cv[q][X] + cy*m->cv[q][Y] + cz*m->cv[q][Z];
	  if (reduced && dot != mm) continue;
	  for (int n = 0; n < ndist; n++) {
	    double * fq = f + (size_t) (n*m->nvel + q)*ns;
	    for (int i = 0; i <= s.imax - s.imin; i++)
	      for (int j = 0; j <= s.jmax - s.jmin; j++)
		for (int k = 0; k <= s.kmax - s.kmin; k++) {
		  fq[orc_index(g, r.imin + i, r.jmin + j, r.kmin + k)]
		    = absent ? 0.0 : fq[orc_index(g, s.imin + i, s.jmin + j, s.kmin + k)];
		}
	  }
	}
      }
    }
  }
}

void orc_field_halo(const orc_geom_t * g, int nf, double * data) {

  const size_t ns = (size_t) orc_nsites(g);
  /* field_halo_post packs every send buffer before field_halo_wait unpacks any (src/field.c:1412-1531).  On a lattice
   * thinner than the halo (nlocal[d] < nhalo, e.g. 64 x 64 x 1 with nhalo 2: tests/regression/d3q19/pmpi08-le2d-fd1)
   * the send regions reach into the halo, so what travels is the halo's content BEFORE this swap: copy from a snapshot. */
  double * snap = NULL;
  if (g->nlocal[X] < g->nhalo || g->nlocal[Y] < g->nhalo || g->nlocal[Z] < g->nhalo) {
    snap = (double *) malloc((size_t) nf*ns*sizeof(double));
    assert(snap);
    memcpy(snap, data, (size_t) nf*ns*sizeof(double));
  }

  for (int cx = -1; cx <= 1; cx++) {
    for (int cy = -1; cy <= 1; cy++) {
      for (int cz = -1; cz <= 1; cz++) {
	int c[3] = {cx, cy, cz};
	lim_t s, r;
	int absent;
	if (cx == 0 && cy == 0 && cz == 0) continue;
	/* as for lb_halo: zeros arrive from an absent neighbour (src/field.c:1178-1187, 1364) */
	absent = ((cx && !g->periodic[X]) || (cy && !g->periodic[Y]) || (cz && !g->periodic[Z]));
	halo_limits(g, g->nhalo, c, &s, &r);
	for (int n = 0; n < nf; n++) {
	  double * d = data + (size_t) n*ns;
	  const double * src = snap ? snap + (size_t) n*ns : d;
	  for (int i = 0; i <= s.imax - s.imin; i++)
	    for (int j = 0; j <= s.jmax - s.jmin; j++)
	      for (int k = 0; k <= s.kmax - s.kmin; k++) {
		d[orc_index(g, r.imin + i, r.jmin + j, r.kmin + k)]
		  = absent ? 0.0 : src[orc_index(g, s.imin + i, s.jmin + j, s.kmin + k)];
	      }
	}
      }
    }
  }
  free(snap);
}

/* ---- lb_collide, single distribution: src/collision.c:253-593 ---------------------------
 * relaxation rates: src/collision.c:1269-1285 (shear), :1330-1360 (bulk), :1420-1520 (ghosts)
 * D3Q19 unrolled projections: src/collision.c:1990-2845 (tables in d3q19_tables.h)          */

static void collide_site(const orc_model_t * m, const orc_collide_param_t * cp,
			 double * fs, const double hforce[3], double * rho_out, double u_out[3]) {

  const int nvel = m->nvel;
  const int nhydro = 10;
  const double cs2 = (1.0/3.0);
  const double rdim = (1.0/3);
  double mode[27];
  double rho, rrho;
  double u[3];
  double s[3][3];
  double seq[3][3];
  double force[3];
  double rtau, rtau_bulk;
  double rtau_ghost[27];
  double tr_s, tr_seq;

  for (int ia = 0; ia < 3; ia++) force[ia] = cp->force_global[ia] + hforce[ia];

  if (nvel == 19) {
    for (int mm = 0; mm < 19; mm++) {
      mode[mm] = 0.0;
      for (int p = 0; p < 19; p++) {
	if (d3q19_fwd[mm][p] != 0.0) mode[mm] += fs[p]*d3q19_fwd[mm][p];
      }
    }
  }
  else {
    for (int mm = 0; mm < nvel; mm++) {
      mode[mm] = 0.0;
      for (int p = 0; p < nvel; p++) mode[mm] += fs[p]*m->ma[mm][p];
    }
  }

  rho = mode[0];
  for (int ia = 0; ia < 3; ia++) u[ia] = mode[1 + ia];
  {
    int k = 0;
    for (int ia = 0; ia < 3; ia++)
      for (int ib = ia; ib < 3; ib++) { s[ia][ib] = mode[4 + k]; k++; }
  }

  rrho = 1.0/rho;
  for (int ia = 0; ia < 3; ia++) u[ia] = rrho*(u[ia] + 0.5*force[ia]);

  rtau = 1.0/(0.5 + cp->eta_shear/(cp->rho0*cs2));
  if (cp->nrelax == ORC_RELAX_BGK) {
    rtau_bulk = 1.0/(0.5 + cp->eta_shear/(cp->rho0*cs2));
  }
  else {
    rtau_bulk = 1.0/(0.5 + cp->eta_bulk/(cp->rho0*cs2));
  }

  for (int mm = nhydro; mm < nvel; mm++) rtau_ghost[mm] = 1.0;  /* M10 */
  if (cp->nrelax == ORC_RELAX_BGK) {
    for (int mm = nhydro; mm < nvel; mm++) rtau_ghost[mm] = rtau;
  }
  if (cp->nrelax == ORC_RELAX_TRT) {
    double tau = cp->eta_shear/(cp->rho0*cs2);
    double rg = 0.5 + 2.0*tau/(tau + 3.0/8.0);
    if (rg > 2.0) rg = 2.0;
    if (nvel == 15) {
      rtau_ghost[10] = rtau;
      rtau_ghost[11] = rg; rtau_ghost[12] = rg; rtau_ghost[13] = rg;
      rtau_ghost[14] = rtau;
    }
    if (nvel == 19) {
      rtau_ghost[10] = rtau; rtau_ghost[14] = rtau; rtau_ghost[18] = rtau;
      rtau_ghost[11] = rg; rtau_ghost[12] = rg; rtau_ghost[13] = rg;
      rtau_ghost[15] = rg; rtau_ghost[16] = rg; rtau_ghost[17] = rg;
    }
  }

  tr_s = 0.0;
  tr_seq = 0.0;
  for (int ia = 0; ia < 3; ia++) {
    for (int ib = ia; ib < 3; ib++) seq[ia][ib] = rho*u[ia]*u[ib];
    tr_s   += s[ia][ia];
    tr_seq += seq[ia][ia];
  }
  for (int ia = 0; ia < 3; ia++) {
    s[ia][ia]   -= rdim*tr_s;
    seq[ia][ia] -= rdim*tr_seq;
  }
  tr_s = tr_s - rtau_bulk*(tr_s - tr_seq);

  for (int ia = 0; ia < 3; ia++) {
    for (int ib = ia; ib < 3; ib++) {
      s[ia][ib] -= rtau*(s[ia][ib] - seq[ia][ib]);
      if (ia == ib) s[ia][ib] += rdim*tr_s;
      s[ia][ib] += (2.0 - rtau)*(u[ia]*force[ib] + force[ia]*u[ib]);
    }
  }

  for (int ia = 0; ia < 3; ia++) mode[1 + ia] += force[ia];
  {
    int k = 0;
    for (int ia = 0; ia < 3; ia++)
      for (int ib = ia; ib < 3; ib++) { mode[4 + k] = s[ia][ib] + 0.0; k++; }
  }
  for (int mm = nhydro; mm < nvel; mm++) {
    mode[mm] = mode[mm] - rtau_ghost[mm]*(mode[mm] - 0.0) + 0.0;
  }

  if (nvel == 19) {
    for (int p = 0; p < 19; p++) {
      double ftmp = 0.0;
      for (int mm = 0; mm < 19; mm++) {
	if (d3q19_bwd[p][mm] != 0.0) ftmp += d3q19_bwd[p][mm]*mode[mm];
      }
      fs[p] = ftmp;
    }
  }
  else {
    for (int p = 0; p < nvel; p++) {
      double ftmp = 0.0;
      for (int mm = 0; mm < nvel; mm++) ftmp += m->mi[p][mm]*mode[mm];
      fs[p] = ftmp;
    }
  }

  *rho_out = rho;
  for (int ia = 0; ia < 3; ia++) u_out[ia] = u[ia];
}

void orc_collide(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		 const char * status, int include_halo,
		 double * f, const double * force, double * rho, double * u) {

  const size_t ns = (size_t) orc_nsites(g);         /* hydro arrays */
  const size_t nsf = (size_t) orc_nsites_lb(g);     /* distributions */
  const int nh = include_halo ? g->nhalo : 0;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1 - nh; jc <= g->nlocal[Y] + nh; jc++) {
      for (int kc = 1 - nh; kc <= g->nlocal[Z] + nh; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double fs[27];
	double hf[3];
	double r, uu[3];
	if (status && status[index] != ORC_MAP_FLUID) continue;
	for (int p = 0; p < m->nvel; p++) fs[p] = f[(size_t) p*nsf + index];
	for (int ia = 0; ia < 3; ia++) hf[ia] = force[(size_t) ia*ns + index];
	collide_site(m, cp, fs, hf, &r, uu);
	for (int p = 0; p < m->nvel; p++) f[(size_t) p*nsf + index] = fs[p];
	rho[index] = r;
	for (int ia = 0; ia < 3; ia++) u[(size_t) ia*ns + index] = uu[ia];
      }
    }
  }
}

/* ---- 27-point gradient: src/gradient_3d_27pt_fluid.c:76-99 (extent), :219-363 (stencil) --- */

void orc_grad_27pt(const orc_geom_t * g, const double * phi, double * grad, double * delsq) {
  orc_grad_27pt_ne(g, g->nhalo - 1, phi, grad, delsq);
}

/* the operator on [1-nextra, N+nextra]^3: nextra = nhalo - 1 for d2 (:91-95); grad_3d_27pt_fluid_d4 (:112-134)
 * applies it to delsq with nextra = nhalo - 2 to produce grad_delsq, delsq_delsq */
void orc_grad_27pt_ne(const orc_geom_t * g, int nextra, const double * phi, double * grad, double * delsq) {

  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  const double r9 = (1.0/9.0);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const int xs = nall[Y]*nall[Z];
  const double * field = phi;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1 - nextra; ic <= g->nlocal[X] + nextra; ic++) {
    for (int jc = 1 - nextra; jc <= g->nlocal[Y] + nextra; jc++) {
      for (int kc = 1 - nextra; kc <= g->nlocal[Z] + nextra; kc++) {
	int index = orc_index(g, ic, jc, kc);
	int indexm1 = index + xoff(g, ic, -1, xs);        /* lees_edw_ic_to_buff, :250-253 */
	int indexp1 = index + xoff(g, ic, +1, xs);

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
    }
  }
}

/* ---- chemical stress (symmetric): src/symmetric.c:371-416, extent src/phi_force_stress.c:181-192
 * (x in [0,N+1], ALL y,z including the halo: contiguous-range kernel, no mask) -------------- */

void orc_stress_symm(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi,
		     const double * grad, const double * delsq_, double * str) {

  const size_t ns = (size_t) orc_nsites(g);
  const int nh = g->nhalo;
  const double a = sp->a, b = sp->b, kappa = sp->kappa;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 0; ic <= g->nlocal[X] + 1; ic++) {
    for (int jc = 1 - nh; jc <= g->nlocal[Y] + nh; jc++) {
      for (int kc = 1 - nh; kc <= g->nlocal[Z] + nh; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double gr[3] = {grad[0*ns + index], grad[1*ns + index], grad[2*ns + index]};
	double ph = phi[index];
	double delsq = delsq_[index];
	double p0 = 0.5*a*ph*ph + 0.75*b*ph*ph*ph*ph - kappa*ph*delsq
	  - 0.5*kappa*(gr[X]*gr[X] + gr[Y]*gr[Y] + gr[Z]*gr[Z]);
	for (int ia = 0; ia < 3; ia++) {
	  for (int ib = 0; ib < 3; ib++) {
	    double d = (ia == ib);
	    str[(size_t) (ia*3 + ib)*ns + index] = p0*d + kappa*gr[ia]*gr[ib];
	  }
	}
      }
    }
  }
}

/* ---- force = - divergence of stress: src/phi_force_colloid.c:315-465 ---------------------- */

void orc_force_divergence(const orc_geom_t * g, const double * str, double * force) {

  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const int xs = nall[Y]*nall[Z];

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	const int off[6] = {+xs, -xs, +ys, -ys, +1, -1};
	double fo[3];
	for (int d = 0; d < 6; d++) {
	  int ib = d/2;
	  int index1 = index + off[d];
	  for (int ia = 0; ia < 3; ia++) {
	    double p1 = str[(size_t) (ia*3 + ib)*ns + index1];
	    double p0 = str[(size_t) (ia*3 + ib)*ns + index];
	    if (d == 0)          fo[ia]  = -0.5*(p1 + p0);
	    else if (d % 2 == 1) fo[ia] += 0.5*(p1 + p0);
	    else                 fo[ia] -= 0.5*(p1 + p0);
	  }
	}
	for (int ia = 0; ia < 3; ia++) force[(size_t) ia*ns + index] += fo[ia]*1;
      }
    }
  }
}

/* ---- advective fluxes: order 1 src/advection.c:538-629; order 2 :770-893; order 3 :946-1141.
 * flux layout: flux[0] = fw, flux[1] = fe, flux[2] = fy, flux[3] = fz (each nsites).
 * extent x in [1,N], y,z in [0,N] (src/advection.c:507, 653, 917) ------------------------- */

static double adv3(double u, double fd1, double fd2, double fd3) {
  const double a1 = -0.213933;
  const double a2 =  0.927865;
  const double a3 =  0.286067;
  return u*(a1*fd1 + a2*fd2 + a3*fd3);
}

void orc_advection(const orc_geom_t * g, int order, const double * u, const double * phi,
		   double * flux) {
  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const int xs = nall[Y]*nall[Z];
  double * fw = flux + 0*ns;
  double * fe = flux + 1*ns;
  double * fy = flux + 2*ns;
  double * fz = flux + 3*ns;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 0; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 0; kc <= g->nlocal[Z]; kc++) {
	int index0 = orc_index(g, ic, jc, kc);
	double u0[3] = {u[0*ns + index0], u[1*ns + index0], u[2*ns + index0]};
	double uf;
	/* x-neighbours through lees_edw_ic_to_buff (src/advection.c:559-560, 792-795, 988-991) */
	const int xm1 = xoff(g, ic, -1, xs), xp1 = xoff(g, ic, +1, xs);
	const int xm2 = xoff(g, ic, -2, xs), xp2 = xoff(g, ic, +2, xs);

	if (order == 1) {
	  int index1, index;
	  index1 = index0 + xm1;
	  uf = 0.5*(u0[X] + u[0*ns + index1]);
	  index = index0; if (uf > 0.0) index = index1;
	  fw[index0] = uf*phi[index];
	  index1 = index0 + xp1;
	  uf = 0.5*(u0[X] + u[0*ns + index1]);
	  index = index0; if (uf < 0.0) index = index1;
	  fe[index0] = uf*phi[index];
	  index1 = index0 + ys;
	  uf = 0.5*(u0[Y] + u[1*ns + index1]);
	  index = index0; if (uf < 0.0) index = index1;
	  fy[index0] = uf*phi[index];
	  index1 = index0 + 1;
	  uf = 0.5*(u0[Z] + u[2*ns + index1]);
	  index = index0; if (uf < 0.0) index = index1;
	  fz[index0] = uf*phi[index];
	}
	else if (order == 2) {
	  int index1;
	  index1 = index0 + xm1;
	  fw[index0] = 0.5*(u0[X] + u[0*ns + index1])*1*0.5*(phi[index1] + phi[index0]);
	  index1 = index0 + xp1;
	  fe[index0] = 0.5*(u0[X] + u[0*ns + index1])*1*0.5*(phi[index0] + phi[index1]);
	  index1 = index0 + ys;
	  fy[index0] = 0.5*(u0[Y] + u[1*ns + index1])*1*0.5*(phi[index0] + phi[index1]);
	  index1 = index0 + 1;
	  fz[index0] = 0.5*(u0[Z] + u[2*ns + index1])*1*0.5*(phi[index0] + phi[index1]);
	}
	else if (order == 4) {
	  /* advection_le_4th, src/advection.c:1153-1262: four-point central interpolation */
	  const double a1 = (1.0/16.0);
	  const double a2 = (9.0/16.0);
	  uf = 0.5*(u0[X] + u[0*ns + index0 + xm1]);
	  fw[index0] = uf*(- a1*phi[index0 + xm2] + a2*phi[index0 + xm1] + a2*phi[index0] - a1*phi[index0 + xp1]);
	  uf = 0.5*(u0[X] + u[0*ns + index0 + xp1]);
	  fe[index0] = uf*(- a1*phi[index0 + xm1] + a2*phi[index0] + a2*phi[index0 + xp1] - a1*phi[index0 + xp2]);
	  uf = 0.5*(u0[Y] + u[1*ns + index0 + ys]);
	  fy[index0] = uf*(- a1*phi[index0 - ys] + a2*phi[index0] + a2*phi[index0 + ys] - a1*phi[index0 + 2*ys]);
	  uf = 0.5*(u0[Z] + u[2*ns + index0 + 1]);
	  fz[index0] = uf*(- a1*phi[index0 - 1] + a2*phi[index0] + a2*phi[index0 + 1] - a1*phi[index0 + 2]);
	}
	else {
	  /* west: index2 = -2, index1 = -1, index3 = +1 */
	  uf = 0.5*1*(u0[X] + u[0*ns + index0 + xm1]);
	  if (uf > 0.0) fw[index0] = adv3(uf, phi[index0 + xm2], phi[index0 + xm1], phi[index0]);
	  else          fw[index0] = adv3(uf, phi[index0 + xp1], phi[index0], phi[index0 + xm1]);
	  /* east */
	  uf = 0.5*1*(u0[X] + u[0*ns + index0 + xp1]);
	  if (uf < 0.0) fe[index0] = adv3(uf, phi[index0 + xp2], phi[index0 + xp1], phi[index0]);
	  else          fe[index0] = adv3(uf, phi[index0 + xm1], phi[index0], phi[index0 + xp1]);
	  /* y */
	  uf = 0.5*1*(u0[Y] + u[1*ns + index0 + ys]);
	  if (uf < 0.0) fy[index0] = adv3(uf, phi[index0 + 2*ys], phi[index0 + ys], phi[index0]);
	  else          fy[index0] = adv3(uf, phi[index0 - ys], phi[index0], phi[index0 + ys]);
	  /* z */
	  uf = 0.5*1*(u0[Z] + u[2*ns + index0 + 1]);
	  if (uf < 0.0) fz[index0] = adv3(uf, phi[index0 + 2], phi[index0 + 1], phi[index0]);
	  else          fz[index0] = adv3(uf, phi[index0 - 1], phi[index0], phi[index0 + 1]);
	}
      }
    }
  }
}

/* ---- diffusive fluxes - M (mu1 - mu0): src/phi_cahn_hilliard.c:350-404, mu src/symmetric.c:307-319 */

static double symm_mu(const orc_symm_param_t * sp, double phi, double delsq) {
  return sp->a*phi + sp->b*phi*phi*phi - sp->kappa*delsq;
}

/* ---- fe_force_method phi_gradmu: phi_grad_mu_fluid + phi_grad_mu_external, src/phi_grad_mu.c:52-82, 124-176, kernels :272-340,
 * 352-384.  Each adds to hydro->force (hydro_f_local_add); the external part runs only when grad_mu != 0. ---- */

void orc_phi_force_gradmu(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi, const double * delsq,
			  double * force) {
  const size_t ns = (size_t) orc_nsites(g);
  const int ys = g->nlocal[Z] + 2*g->nhalo;
  const int xs = ys*(g->nlocal[Y] + 2*g->nhalo);
  const int is_grad_mu = (sp->gradmu[X] != 0.0 || sp->gradmu[Y] != 0.0 || sp->gradmu[Z] != 0.0);
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	const int index = orc_index(g, ic, jc, kc);
	const int off[3] = {xs, ys, 1};
	const double phi0 = phi[index];
	for (int ia = 0; ia < 3; ia++) {
	  const double mum1 = symm_mu(sp, phi[index - off[ia]], delsq[index - off[ia]]);
	  const double mup1 = symm_mu(sp, phi[index + off[ia]], delsq[index + off[ia]]);
	  double f = 0.0;
	  f += -phi0*0.5*(mup1 - mum1);
	  force[(size_t) ia*ns + index] += f;
	}
      }
  if (is_grad_mu) {
    for (int ic = 1; ic <= g->nlocal[X]; ic++)
      for (int jc = 1; jc <= g->nlocal[Y]; jc++)
	for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	  const int index = orc_index(g, ic, jc, kc);
	  const double phi0 = phi[index];
	  for (int ia = 0; ia < 3; ia++) force[(size_t) ia*ns + index] += -phi0*sp->gradmu[ia];
	}
  }
}

void orc_flux_mu(const orc_geom_t * g, const orc_symm_param_t * sp, const double * phi,
		 const double * delsq, double * flux) {
  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const int xs = nall[Y]*nall[Z];
  const double mobility = sp->mobility;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 0; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 0; kc <= g->nlocal[Z]; kc++) {
	int index0 = orc_index(g, ic, jc, kc);
	double mu0 = symm_mu(sp, phi[index0], delsq[index0]);
	double mu1;
	mu1 = symm_mu(sp, phi[index0 + xoff(g, ic, -1, xs)], delsq[index0 + xoff(g, ic, -1, xs)]);   /* :369-370 */
	flux[0*ns + index0] -= mobility*(mu0 - mu1);
	mu1 = symm_mu(sp, phi[index0 + xoff(g, ic, +1, xs)], delsq[index0 + xoff(g, ic, +1, xs)]);
	flux[1*ns + index0] -= mobility*(mu1 - mu0);
	mu1 = symm_mu(sp, phi[index0 + ys], delsq[index0 + ys]);
	flux[2*ns + index0] -= mobility*(mu1 - mu0);
	mu1 = symm_mu(sp, phi[index0 + 1], delsq[index0 + 1]);
	flux[3*ns + index0] -= mobility*(mu1 - mu0);
      }
    }
  }
}

/* ---- external chemical potential gradient: src/phi_cahn_hilliard.c:1373-1397 ------------- */

void orc_flux_mu_ext(const orc_geom_t * g, const orc_symm_param_t * sp, double * flux) {
  const size_t ns = (size_t) orc_nsites(g);
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 0; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 0; kc <= g->nlocal[Z]; kc++) {
	int index0 = orc_index(g, ic, jc, kc);
	flux[0*ns + index0] -= sp->mobility*sp->gradmu[X];
	flux[1*ns + index0] -= sp->mobility*sp->gradmu[X];
	flux[2*ns + index0] -= sp->mobility*sp->gradmu[Y];
	flux[3*ns + index0] -= sp->mobility*sp->gradmu[Z];
      }
    }
  }
}

/* ---- no normal flux at solid/fluid faces: src/advection_bcs.c:80-130 ---------------------- */

void orc_no_flux(const orc_geom_t * g, const char * status, double * flux) {
  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const int xs = nall[Y]*nall[Z];
  if (status == NULL) return;    /* all fluid: every mask is 1.0 and x *= 1.0 is the identity */
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 0; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 0; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double mask  = (status[index] == ORC_MAP_FLUID);
	double maskw = (status[index - xs] == ORC_MAP_FLUID);
	double maske = (status[index + xs] == ORC_MAP_FLUID);
	double masky = (status[index + ys] == ORC_MAP_FLUID);
	double maskz = (status[index + 1] == ORC_MAP_FLUID);
	flux[0*ns + index] *= mask*maskw;
	flux[1*ns + index] *= mask*maske;
	flux[2*ns + index] *= mask*masky;
	flux[3*ns + index] *= mask*maskz;
      }
    }
  }
}

/* ---- forward Euler update: src/phi_cahn_hilliard.c:1018-1049 ----------------------------- */

void orc_phi_update(const orc_geom_t * g, const double * flux, double * phi) {
  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const double wz = (g->nlocal[Z] == 1) ? 0.0 : 1.0;
  const double * fw = flux + 0*ns;
  const double * fe = flux + 1*ns;
  const double * fy = flux + 2*ns;
  const double * fz = flux + 3*ns;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double ph = phi[index];
	ph -= (+ fe[index] - fw[index] + fy[index] - fy[index - ys]
	       + wz*fz[index] - wz*fz[index - 1]);
	phi[index] = ph;
      }
    }
  }
}

/* ---- cahn_hilliard_options_conserve 2 (all-fluid lattices) ------------------------------------------------------
 * phi_ch_subtract_sum_phi_after_forward_step, src/phi_cahn_hilliard.c:1102-1169: kernel1 (:1225-1282) adds phi over the fluid
 * sites with plain += (here: one thread, i.e. (ic, jc, kc) order), kernel2 (:1290-1319) subtracts (sum - phi0)/nfluid.
 * phi0 = phi->field_init_sum from cahn_hilliard_stats_time0 (src/cahn_hilliard_stats.c:58-76): with conserve != 0 the
 * statistics use the doubly compensated (Klein) sum of csum (zero at time 0) and phi, src/cahn_hilliard_stats.c:270-330,
 * src/util_sum.c:108-180 -- restated for one thread. */

typedef struct { double sum, cs, ccs; } orc_klein_t;

static void orc_klein_add_double(orc_klein_t * k, double val) {        /* klein_add_double, src/util_sum.c:144-166 */
  double t, c, cc;
  t = k->sum + val;
  if (fabs(k->sum) >= fabs(val)) c = (k->sum - t) + val;
  else                           c = (val - t) + k->sum;
  k->sum = t;
  t = k->cs + c;
  if (fabs(k->cs) >= fabs(c)) cc = (k->cs - t) + c;
  else                        cc = (c - t) + k->cs;
  k->cs = t;
  k->ccs = k->ccs + cc;
}

double orc_phi_sum_time0(const orc_geom_t * g, const double * phi) {
  orc_klein_t thread = {0.0, 0.0, 0.0}, block = {0.0, 0.0, 0.0}, total = {0.0, 0.0, 0.0};
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	orc_klein_add_double(&thread, 0.0);                                  /* the compensation field, zero at time 0 */
	orc_klein_add_double(&thread, phi[orc_index(g, ic, jc, kc)]);
      }
  /* the per-thread, per-block and global accumulations (klein_add, src/util_sum.c:168-180) for one thread, one block */
  orc_klein_add_double(&block, thread.sum); orc_klein_add_double(&block, thread.cs); orc_klein_add_double(&block, thread.ccs);
  orc_klein_add_double(&total, block.sum); orc_klein_add_double(&total, block.cs); orc_klein_add_double(&total, block.ccs);
  return total.sum + total.cs + total.ccs;                                   /* klein_sum */
}

void orc_phi_subtract_sum(const orc_geom_t * g, double phi_init_sum, double * phi) {
  double sum = 0.0;
  int nfluid = 0;
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) { sum += phi[orc_index(g, ic, jc, kc)]; nfluid += 1; }
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	const int index = orc_index(g, ic, jc, kc);
	phi[index] -= (sum - phi_init_sum)/nfluid;
      }
}

/* ---- phi_ch_update_conserve -> phi_ch_csum_kernel: src/phi_cahn_hilliard.c:1059-1094, 1181-1215, with
 * kahan_add_double (src/util_sum.c:30-40: y = val + cs; t = sum + y; cs = y - (t - sum); sum = t).
 * csum is the per-site compensation carried from step to step (pch->csum, zero at creation). ---- */

void orc_phi_update_conserve(const orc_geom_t * g, const double * flux, double * csum, double * phi) {
  int nall[3];
  const size_t ns = (size_t) orc_nsites(g);
  orc_nall(g, nall);
  const int ys = nall[Z];
  const double wz = (g->nlocal[Z] == 1) ? 0.0 : 1.0;
  const double * fw = flux + 0*ns;
  const double * fe = flux + 1*ns;
  const double * fy = flux + 2*ns;
  const double * fz = flux + 3*ns;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	volatile double sum = phi[index];
	volatile double cs  = csum[index];
	const double val[6] = {-fe[index], fw[index], -fy[index], fy[index - ys], -wz*fz[index], wz*fz[index - 1]};
	for (int n = 0; n < 6; n++) {
	  volatile double y = val[n] + cs;
	  volatile double t = sum + y;
	  cs  = y - (t - sum);
	  sum = t;
	}
	csum[index] = cs;
	phi[index] = sum;
      }
    }
  }
}

/* ---- hydro_f_zero / hydro_u_zero: src/hydro.c:217-263, 319-345 ---------------------------- */

void orc_field_set(const orc_geom_t * g, int nf, double * data, const double * values) {
  const size_t ns = (size_t) orc_nsites(g);
  for (int n = 0; n < nf; n++)
    for (size_t i = 0; i < ns; i++) data[(size_t) n*ns + i] = values[n];
}

/* ---- one time step in the reference driver's order: src/ludwig.c:528-860 ------------------- */

void orc_step(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
	      const orc_symm_param_t * sp, int binary, int halo_reduced, int nsteps,
	      double * f, double * phi, double * u, double * rho, double * force,
	      double * grad, double * delsq) {

  const size_t ns = (size_t) orc_nsites(g);
  const double zero[3] = {0.0, 0.0, 0.0};
  double * fprime = (double *) calloc(ns*m->nvel, sizeof(double));
  double * str = NULL;
  double * flux = NULL;
  double * csum = NULL;     /* pch->csum: zero when the phi_ch_t is created, i.e. at the start of this run */

  assert(fprime);
  /* fprime's never-written x-halo planes hold stale data in the reference too; start equal */
  memcpy(fprime, f, ns*m->nvel*sizeof(double));

  if (binary) {
    str  = (double *) calloc(ns*9, sizeof(double));
    flux = (double *) calloc(ns*4, sizeof(double));
    assert(str && flux);
    if (sp->conserve == 1) { csum = (double *) calloc(ns, sizeof(double)); assert(csum); }
  }

  for (int n = 0; n < nsteps; n++) {
    orc_field_set(g, 3, force, zero);
    if (binary) {
      orc_field_halo(g, 1, phi);
      if (sp->grad_7pt) orc_grad_7pt(g, 1, phi, grad, delsq);          /* grad_3d_7pt_fluid_d2 */
      else              orc_grad_27pt(g, phi, grad, delsq);
      if (sp->force_method == 1) orc_phi_force_gradmu(g, sp, phi, delsq, force);
      else {
	orc_stress_symm(g, sp, phi, grad, delsq, str);
	orc_force_divergence(g, str, force);
      }
      orc_field_halo(g, 3, u);
      orc_advection(g, sp->adv_order, u, phi, flux);
      orc_flux_mu(g, sp, phi, delsq, flux);
      orc_flux_mu_ext(g, sp, flux);
      orc_no_flux(g, NULL, flux);
      if (csum) orc_phi_update_conserve(g, flux, csum, phi);
      else      orc_phi_update(g, flux, phi);
      if (sp->conserve == 2) orc_phi_subtract_sum(g, sp->phi_init_sum, phi);
    }
    orc_field_set(g, 3, u, zero);
    orc_collide(g, m, cp, NULL, 0, f, force, rho, u);
    orc_lb_halo(g, m, 1, halo_reduced, f);
    orc_propagation(g, m, 1, f, fprime);
    memcpy(f, fprime, ns*m->nvel*sizeof(double));
  }

  free(fprime);
  free(str);
  free(flux);
  free(csum);
}

/* ---- symmetric_lb: two-distribution binary fluid ------------------------------------------------
 * phi_lb_to_field / phi_lb_from_field: src/phi_lb_coupler.c:39-137
 * lb_collision_binary -> lb_collision_mrt2_site: src/collision.c:604-1013 (interior sites only, no
 * status test, hydro->rho NOT written), relaxation rates from lb->param->rtau[]
 * (src/collision.c:1163-1246), rtau2 = 2/(1 + 2M) (:1949-1950), thermodynamic stress fe_symm_str_v and
 * chemical potential fe_symm_mu (src/symmetric.c:307-319, 371-416), order-parameter reprojection
 * d3q19_mode2f_phi (src/collision.c:2856-3135) / generic loop (:974-1008).                          */

void orc_phi_lb_to_field(const orc_geom_t * g, const orc_model_t * m, const double * f, double * phi) {
  const size_t ns = (size_t) orc_nsites_lb(g);          /* the distributions carry no Lees-Edwards buffer planes */
  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double phi0 = 0.0;
	for (int p = 0; p < m->nvel; p++) phi0 += f[(size_t) (m->nvel + p)*ns + index];
	phi[index] = phi0;
      }
}

void orc_phi_lb_from_field(const orc_geom_t * g, const orc_model_t * m, const double * phi, double * f) {
  const size_t ns = (size_t) orc_nsites_lb(g);
  for (int ic = 1; ic <= g->nlocal[X]; ic++)
    for (int jc = 1; jc <= g->nlocal[Y]; jc++)
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	f[(size_t) (m->nvel + 0)*ns + index] = phi[index];
	for (int p = 1; p < m->nvel; p++) f[(size_t) (m->nvel + p)*ns + index] = 0.0;
      }
}

static void collide_binary_site(const orc_model_t * m, const orc_collide_param_t * cp,
				const orc_symm_param_t * sp, double * fs, double * gs,
				const double hforce[3], double phi, const double grad[3], double delsq,
				double u_out[3]) {

  const int nvel = m->nvel;
  const int nhydro = 10;
  const double cs2 = (1.0/3.0);
  const double r3 = 1.0/3.0;
  const signed char d[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double mode[27];
  double rho, rrho;
  double u[3], s[3][3], seq[3][3], sth[3][3], sphi[3][3], force[3], jphi[3];
  double rtau[27];
  double tr_s, tr_seq, mu;

  /* lb_collision_relaxation_times_set */
  {
    double rtau_shear = 1.0/(0.5 + cp->eta_shear/(cp->rho0*cs2));
    double rtau_bulk  = 1.0/(0.5 + cp->eta_bulk/(cp->rho0*cs2));
    for (int p = 0; p < 27; p++) rtau[p] = 0.0;
    if (cp->nrelax == ORC_RELAX_M10) {
      rtau[5] = rtau_shear; rtau[4] = rtau_bulk;
      for (int p = nhydro; p < nvel; p++) rtau[p] = 1.0;
    }
    if (cp->nrelax == ORC_RELAX_BGK) {
      for (int p = 0; p < nvel; p++) rtau[p] = rtau_shear;
    }
    if (cp->nrelax == ORC_RELAX_TRT) {
      double tau = cp->eta_shear/(cp->rho0*cs2);
      double rg = 0.5 + 2.0*tau/(tau + 3.0/8.0);
      if (rg > 2.0) rg = 2.0;
      rtau[5] = rtau_shear; rtau[4] = rtau_bulk;
      if (nvel == 15) {
	rtau[10] = rtau_shear; rtau[11] = rg; rtau[12] = rg; rtau[13] = rg; rtau[14] = rtau_shear;
      }
      if (nvel == 19) {
	rtau[10] = rtau_shear; rtau[14] = rtau_shear; rtau[18] = rtau_shear;
	rtau[11] = rg; rtau[12] = rg; rtau[13] = rg; rtau[15] = rg; rtau[16] = rg; rtau[17] = rg;
      }
    }
  }

  if (nvel == 19) {
    for (int mm = 0; mm < 19; mm++) {
      mode[mm] = 0.0;
      for (int p = 0; p < 19; p++) {
	if (d3q19_fwd[mm][p] != 0.0) mode[mm] += fs[p]*d3q19_fwd[mm][p];
      }
    }
  }
  else {
    for (int mm = 0; mm < nvel; mm++) {
      mode[mm] = 0.0;
      for (int p = 0; p < nvel; p++) mode[mm] += m->ma[mm][p]*fs[p];
    }
  }

  rho = mode[0];
  for (int ia = 0; ia < 3; ia++) u[ia] = mode[1 + ia];
  {
    int k = 0;
    for (int ia = 0; ia < 3; ia++)
      for (int ib = ia; ib < 3; ib++) { s[ia][ib] = mode[4 + k]; k++; }
    for (int ia = 1; ia < 3; ia++)
      for (int ib = 0; ib < ia; ib++) s[ia][ib] = s[ib][ia];
  }

  rrho = 1.0/rho;
  for (int ia = 0; ia < 3; ia++) {
    force[ia] = cp->force_global[ia] + hforce[ia];
    u[ia] = rrho*(u[ia] + 0.5*force[ia]);
  }
  for (int ia = 0; ia < 3; ia++) u_out[ia] = u[ia];

  /* fe_symm_str_v */
  {
    double p0 = 0.5*sp->a*phi*phi + 0.75*sp->b*phi*phi*phi*phi - sp->kappa*phi*delsq
      - 0.5*sp->kappa*(grad[X]*grad[X] + grad[Y]*grad[Y] + grad[Z]*grad[Z]);
    for (int ia = 0; ia < 3; ia++)
      for (int ib = 0; ib < 3; ib++) sth[ia][ib] = p0*d[ia][ib] + sp->kappa*grad[ia]*grad[ib];
  }

  tr_s = 0.0;
  tr_seq = 0.0;
  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) seq[ia][ib] = rho*u[ia]*u[ib] + sth[ia][ib];
    tr_s   += s[ia][ia];
    tr_seq += seq[ia][ia];
  }
  for (int ia = 0; ia < 3; ia++) {
    s[ia][ia]   -= r3*tr_s;
    seq[ia][ia] -= r3*tr_seq;
  }
  tr_s = tr_s - rtau[4]*(tr_s - tr_seq);

  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      s[ia][ib] -= rtau[5]*(s[ia][ib] - seq[ia][ib]);
      s[ia][ib] += d[ia][ib]*r3*tr_s;
      s[ia][ib] += (2.0 - rtau[5])*(u[ia]*force[ib] + force[ia]*u[ib]);
    }
  }

  for (int ia = 0; ia < 3; ia++) mode[1 + ia] += force[ia];
  {
    int k = 0;
    for (int ia = 0; ia < 3; ia++)
      for (int ib = ia; ib < 3; ib++) { mode[4 + k] = s[ia][ib] + 0.0; k++; }
  }
  for (int mm = nhydro; mm < nvel; mm++) {
    mode[mm] = mode[mm] - rtau[mm]*(mode[mm] - 0.0) + 0.0;
  }

  if (nvel == 19) {
    for (int p = 0; p < 19; p++) {
      double ftmp = 0.0;
      for (int mm = 0; mm < 19; mm++) {
	if (d3q19_bwd[p][mm] != 0.0) ftmp += d3q19_bwd[p][mm]*mode[mm];
      }
      fs[p] = ftmp;
    }
  }
  else {
    for (int p = 0; p < nvel; p++) {
      double ftmp = 0.0;
      for (int mm = 0; mm < nvel; mm++) ftmp += m->mi[p][mm]*mode[mm];
      fs[p] = ftmp;
    }
  }

  /* the order parameter distribution */
  mu = sp->a*phi + sp->b*phi*phi*phi - sp->kappa*delsq;
  jphi[X] = 0.0; jphi[Y] = 0.0; jphi[Z] = 0.0;
  for (int p = 1; p < nvel; p++)
    for (int ia = 0; ia < 3; ia++) jphi[ia] += m->cv[p][ia]*gs[p];

  {
    const double rtau2 = 2.0/(1.0 + 2.0*sp->mobility);
    for (int ia = 0; ia < 3; ia++) {
      for (int ib = 0; ib < 3; ib++) sphi[ia][ib] = phi*u[ia]*u[ib] + mu*d[ia][ib];
      jphi[ia] = jphi[ia] - rtau2*(jphi[ia] - phi*u[ia]);
    }
  }

  if (nvel == 19) {
    /* unrolled: only the non-zero terms, literal constants 2/3, -1/3, +-1 (src/collision.c:2856-3135) */
    const double q23 = 6.6666666666666663e-01, q13 = -3.3333333333333331e-01;
    for (int p = 0; p < 19; p++) {
      double jdotc = 0.0, sphidotq = 0.0;
      for (int ia = 0; ia < 3; ia++) {
	if (m->cv[p][ia] > 0) jdotc += jphi[ia];
	if (m->cv[p][ia] < 0) jdotc -= jphi[ia];
      }
      for (int ia = 0; ia < 3; ia++) {
	for (int ib = 0; ib < 3; ib++) {
	  int cc = m->cv[p][ia]*m->cv[p][ib];
	  if (ia == ib) sphidotq += sphi[ia][ib]*(cc ? q23 : q13);
	  else if (cc != 0) sphidotq += sphi[ia][ib]*(cc > 0 ? 1.0000000000000000e+00 : -1.0000000000000000e+00);
	}
      }
      gs[p] = m->wv[p]*(jdotc*3.0 + sphidotq*(9.0/2.0));
      if (p == 0) gs[p] = m->wv[p]*(jdotc*3.0 + sphidotq*(9.0/2.0)) + phi;
    }
  }
  else {
    for (int p = 0; p < nvel; p++) {
      int dp0 = (p == 0);
      double jdotc = 0.0, sphidotq = 0.0;
      for (int ia = 0; ia < 3; ia++) {
	jdotc += jphi[ia]*m->cv[p][ia];
	for (int ib = 0; ib < 3; ib++) {
	  sphidotq += sphi[ia][ib]*(m->cv[p][ia]*m->cv[p][ib] - cs2*d[ia][ib]);
	}
      }
      gs[p] = m->wv[p]*(jdotc*3.0 + sphidotq*4.5) + phi*dp0;
    }
  }
}

void orc_collide_binary(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
			const orc_symm_param_t * sp, double * f, const double * force,
			const double * phi, const double * grad, const double * delsq, double * u) {

  const size_t ns = (size_t) orc_nsites(g);
  const size_t nsf = (size_t) orc_nsites_lb(g);         /* distributions (no Lees-Edwards buffer planes) */
  const int nvel = m->nvel;

  #pragma omp parallel for collapse(2) schedule(static)
  for (int ic = 1; ic <= g->nlocal[X]; ic++) {
    for (int jc = 1; jc <= g->nlocal[Y]; jc++) {
      for (int kc = 1; kc <= g->nlocal[Z]; kc++) {
	int index = orc_index(g, ic, jc, kc);
	double fs[27], gs[27], hf[3], gr[3], uu[3];
	for (int p = 0; p < nvel; p++) {
	  fs[p] = f[(size_t) p*nsf + index];
	  gs[p] = f[(size_t) (nvel + p)*nsf + index];
	}
	for (int ia = 0; ia < 3; ia++) {
	  hf[ia] = force[(size_t) ia*ns + index];
	  gr[ia] = grad[(size_t) ia*ns + index];
	}
	collide_binary_site(m, cp, sp, fs, gs, hf, phi[index], gr, delsq[index], uu);
	for (int p = 0; p < nvel; p++) {
	  f[(size_t) p*nsf + index] = fs[p];
	  f[(size_t) (nvel + p)*nsf + index] = gs[p];
	}
	for (int ia = 0; ia < 3; ia++) u[(size_t) ia*ns + index] = uu[ia];
      }
    }
  }
}

/* nsteps of the symmetric_lb time step, reference order src/ludwig.c:528-860 with ndist == 2:
 * hydro_f_zero; phi_lb_to_field; field_halo(phi); field_grad_compute; hydro_u_zero; lb_collide
 * (binary); lb_halo; lb_propagation.  f holds both distributions. */
void orc_step_lb2(const orc_geom_t * g, const orc_model_t * m, const orc_collide_param_t * cp,
		  const orc_symm_param_t * sp, int halo_reduced, int nsteps,
		  double * f, double * phi, double * u, double * force, double * grad, double * delsq) {

  const size_t ns = (size_t) orc_nsites(g);
  const size_t nf = ns*2*m->nvel;
  const double zero[3] = {0.0, 0.0, 0.0};
  double * fprime = (double *) calloc(nf, sizeof(double));
  assert(fprime);
  memcpy(fprime, f, nf*sizeof(double));

  for (int n = 0; n < nsteps; n++) {
    orc_field_set(g, 3, force, zero);
    orc_phi_lb_to_field(g, m, f, phi);
    orc_field_halo(g, 1, phi);
    orc_grad_27pt(g, phi, grad, delsq);
    orc_field_set(g, 3, u, zero);
    orc_collide_binary(g, m, cp, sp, f, force, phi, grad, delsq, u);
    orc_lb_halo(g, m, 2, halo_reduced, f);
    orc_propagation(g, m, 2, f, fprime);
    memcpy(f, fprime, nf*sizeof(double));
  }
  free(fprime);
}
