/*
 * ludwig_host.c -- Ludwig's host-side names for the hot path on top of the lb200 C-ABI
 * (declarations and reference citations: include/ludwig_host.h).
 *
 * One device context (lb200_t) per coordinate system, created at the first device operation from
 * the objects registered with that cs_t by their constructors (lb_t, hydro_t, the order-parameter
 * field_t, its field_grad_t, map_t) -- the reference likewise builds every object before the first
 * *_memcpy(HostToDevice) (src/ludwig.c:202-429, 501-506).
 */

#include <assert.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ludwig_host.h"

/* ---- pe ------------------------------------------------------------------------------------ */

struct pe_s {int quiet; int nref;};

int pe_create(MPI_Comm parent, pe_enum_t flag, pe_t ** ppe) {
  pe_t * pe = (pe_t *) calloc(1, sizeof(pe_t));
  (void) parent;
  assert(ppe);
  if (pe == NULL) return -1;
  pe->quiet = (flag == PE_QUIET);
  pe->nref = 1;
  *ppe = pe;
  return 0;
}

int pe_free(pe_t * pe) { free(pe); return 0; }
int pe_mpi_rank(pe_t * pe) { (void) pe; return 0; }
int pe_mpi_size(pe_t * pe) { (void) pe; return 1; }

int pe_info(pe_t * pe, const char * fmt, ...) {
  va_list args;
  if (pe && pe->quiet) return 0;
  va_start(args, fmt);
  vprintf(fmt, args);
  va_end(args);
  return 0;
}

/* reference: print, then MPI_Abort (src/pe.c:226-240) */
int pe_fatal(pe_t * pe, const char * fmt, ...) {
  va_list args;
  (void) pe;
  printf("[0] ");
  va_start(args, fmt);
  vprintf(fmt, args);
  va_end(args);
  printf("[0] aborting\n");
  fflush(stdout);
  exit(1);
  return 0;
}

/* ---- cs -------------------------------------------------------------------------------------- */

struct cs_s {
  pe_t * pe;
  int ntotal[3];
  int nlocal[3];
  int noffset[3];
  int nhalo;
  int periodic[3];
  int nall[3];
  int nsites;
  int initialised;
  /* objects living on this coordinate system, and the device lattice behind them */
  lb_t * lb;
  hydro_t * hydro;
  field_t * phi;
  field_grad_t * phi_grad;
  map_t * map;
  lees_edw_t * le;
  field_t * q;              /* liquid crystal: the tensor order parameter */
  field_grad_t * q_grad;
  lb200_t * ctx;
};

int cs_create(pe_t * pe, cs_t ** pcs) {
  cs_t * cs = (cs_t *) calloc(1, sizeof(cs_t));
  assert(pe);
  assert(pcs);
  if (cs == NULL) pe_fatal(pe, "calloc(cs_t) failed\n");
  cs->pe = pe;
  cs->ntotal[X] = cs->ntotal[Y] = cs->ntotal[Z] = 64;     /* reference defaults, src/coords.c:60-75 */
  cs->periodic[X] = cs->periodic[Y] = cs->periodic[Z] = 1;
  cs->nhalo = 1;
  *pcs = cs;
  return 0;
}

int cs_free(cs_t * cs) {
  if (cs == NULL) return 0;
  if (cs->ctx) lb200_free(cs->ctx);
  free(cs);
  return 0;
}

int cs_ntotal_set(cs_t * cs, const int ntotal[3]) { for (int a = 0; a < 3; a++) cs->ntotal[a] = ntotal[a]; return 0; }
int cs_nhalo_set(cs_t * cs, int nhalo) { cs->nhalo = nhalo; return 0; }
int cs_periodicity_set(cs_t * cs, const int iper[3]) { for (int a = 0; a < 3; a++) cs->periodic[a] = iper[a]; return 0; }

int cs_init(cs_t * cs) {
  assert(cs);
  cs->nsites = 1;
  for (int a = 0; a < 3; a++) {
    cs->nlocal[a] = cs->ntotal[a];                         /* single process: src/coords.c:211-215 */
    cs->noffset[a] = 0;
    cs->nall[a] = cs->nlocal[a] + 2*cs->nhalo;
    cs->nsites *= cs->nall[a];
  }
  cs->initialised = 1;
  return 0;
}

int cs_ntotal(cs_t * cs, int n[3]) { for (int a = 0; a < 3; a++) n[a] = cs->ntotal[a]; return 0; }
int cs_nlocal(cs_t * cs, int n[3]) { for (int a = 0; a < 3; a++) n[a] = cs->nlocal[a]; return 0; }
int cs_nlocal_offset(cs_t * cs, int n[3]) { for (int a = 0; a < 3; a++) n[a] = cs->noffset[a]; return 0; }
int cs_nall(cs_t * cs, int n[3]) { for (int a = 0; a < 3; a++) n[a] = cs->nall[a]; return 0; }
int cs_periodic(cs_t * cs, int p[3]) { for (int a = 0; a < 3; a++) p[a] = cs->periodic[a]; return 0; }
int cs_nhalo(cs_t * cs, int * nhalo) { *nhalo = cs->nhalo; return 0; }
int cs_nsites(cs_t * cs, int * nsites) { *nsites = cs->nsites; return 0; }
int cs_cartsz(cs_t * cs, int sz[3]) { (void) cs; sz[X] = sz[Y] = sz[Z] = 1; return 0; }
int cs_cart_coords(cs_t * cs, int c[3]) { (void) cs; c[X] = c[Y] = c[Z] = 0; return 0; }

/* src/coords.c:617-631 */
int cs_index(cs_t * cs, int ic, int jc, int kc) {
  return (ic + cs->nhalo - 1)*cs->nall[Y]*cs->nall[Z] + (jc + cs->nhalo - 1)*cs->nall[Z] + (kc + cs->nhalo - 1);
}

static int cs_index_is_interior(cs_t * cs, int index) {
  const int kc = index % cs->nall[Z] - cs->nhalo + 1;
  const int jc = (index/cs->nall[Z]) % cs->nall[Y] - cs->nhalo + 1;
  const int ic = index/(cs->nall[Y]*cs->nall[Z]) - cs->nhalo + 1;
  return ic >= 1 && ic <= cs->nlocal[X] && jc >= 1 && jc <= cs->nlocal[Y] && kc >= 1 && kc <= cs->nlocal[Z];
}

int cs_strides(cs_t * cs, int * xs, int * ys, int * zs) {
  *xs = cs->nall[Y]*cs->nall[Z]; *ys = cs->nall[Z]; *zs = 1;
  return 0;
}

struct beris_edw_s {pe_t * pe; cs_t * cs; lees_edw_t * le; beris_edw_param_t param;};
enum {FE_LC_ID = 1000};   /* tag of the liquid-crystal free energy in fe_t.id (the reference's enum value is not part of this interface) */
static void b200_le_options(cs_t * cs, lb200_options_t * o);
static void b200_time_sync(cs_t * cs);

lb200_t * cs_b200_context(cs_t * cs) {
  assert(cs);
  assert(cs->initialised);
  if (cs->ctx == NULL) {
    lb200_options_t o;
    const char * math = getenv("LB200_MATH");
    memset(&o, 0, sizeof(o));
    for (int a = 0; a < 3; a++) { o.nlocal[a] = cs->nlocal[a]; o.periodic[a] = cs->periodic[a]; }
    o.nhalo = cs->nhalo;
    o.nvel = cs->lb ? cs->lb->nvel : 19;
    o.ndist = cs->lb ? cs->lb->ndist : 1;
    o.have_phi = (cs->phi != NULL);
    o.have_q = (cs->q != NULL);
    if (o.have_q && o.have_phi) pe_fatal(cs->pe, "a scalar order parameter and the Q tensor on one lattice are outside this build\n");
    o.halo_scheme = cs->lb ? (int) cs->lb->haloscheme : LB200_HALO_FULL;
    o.math = (math && strcmp(math, "strict") == 0) ? LB200_MATH_STRICT : LB200_MATH_FAST;
    o.device = -1;
    o.cart_size = 1;
    o.cart_rank = 0;
    b200_le_options(cs, &o);
    if (lb200_create(&o, &cs->ctx) != 0) pe_fatal(cs->pe, "lb200_create: %s\n", lb200_last_error());
  }
  return cs->ctx;
}

static void b200_check(pe_t * pe, int rc, const char * what) {
  if (rc != 0) pe_fatal(pe, "%s: %s\n", what, lb200_last_error());
}

static int b200_kind(tdpMemcpyKind flag) {
  return (flag == tdpMemcpyHostToDevice) ? LB200_HOST_TO_DEVICE : LB200_DEVICE_TO_HOST;
}

/* ---- physics (singleton) ---------------------------------------------------------------------- */

struct physics_s {
  double rho0, eta_shear, eta_bulk, fbody[3], mobility, grad_mu[3];
  int t_start, nsteps, t_current;
};

static physics_t * physics_static = NULL;

int physics_create(pe_t * pe, physics_t ** phys) {
  physics_t * p = (physics_t *) calloc(1, sizeof(physics_t));
  if (p == NULL) pe_fatal(pe, "calloc(physics_t) failed\n");
  p->rho0 = 1.0;                         /* reference defaults, src/physics.c:60-80 */
  p->eta_shear = 1.0/6.0;
  p->eta_bulk = 1.0/6.0;
  physics_static = p;
  *phys = p;
  return 0;
}

int physics_free(physics_t * phys) { if (phys == physics_static) physics_static = NULL; free(phys); return 0; }
int physics_ref(physics_t ** phys) { assert(physics_static); *phys = physics_static; return 0; }
int physics_rho0_set(physics_t * p, double rho0) { p->rho0 = rho0; return 0; }
int physics_eta_shear_set(physics_t * p, double eta) { p->eta_shear = eta; return 0; }
int physics_eta_bulk_set(physics_t * p, double zeta) { p->eta_bulk = zeta; return 0; }
int physics_fbody_set(physics_t * p, double f[3]) { for (int a = 0; a < 3; a++) p->fbody[a] = f[a]; return 0; }
int physics_mobility_set(physics_t * p, double m) { p->mobility = m; return 0; }
int physics_grad_mu_set(physics_t * p, double gm[3]) { for (int a = 0; a < 3; a++) p->grad_mu[a] = gm[a]; return 0; }
int physics_rho0(physics_t * p, double * rho0) { *rho0 = p->rho0; return 0; }
int physics_eta_shear(physics_t * p, double * eta) { *eta = p->eta_shear; return 0; }
int physics_eta_bulk(physics_t * p, double * eta) { *eta = p->eta_bulk; return 0; }
int physics_fbody(physics_t * p, double f[3]) { for (int a = 0; a < 3; a++) f[a] = p->fbody[a]; return 0; }
int physics_mobility(physics_t * p, double * m) { *m = p->mobility; return 0; }
int physics_grad_mu(physics_t * p, double gm[3]) { for (int a = 0; a < 3; a++) gm[a] = p->grad_mu[a]; return 0; }

/* src/physics.c:600-670 */
int physics_control_init_time(physics_t * p, int nstart, int nstep) { p->t_start = nstart; p->nsteps = nstep; p->t_current = nstart; return 0; }
int physics_control_next_step(physics_t * p) { p->t_current += 1; return (p->t_start + p->nsteps - p->t_current + 1); }
int physics_control_timestep(physics_t * p) { return p->t_current; }
int physics_control_time(physics_t * p, double * t) { *t = 1.0*(p->t_start + p->t_current - 1.0); return 0; }

/* ---- Lees-Edwards planes (steady shear): src/leesedwards.c -------------------------------------- */

struct lees_edw_s {pe_t * pe; cs_t * cs; lees_edw_options_t opts; lb200_options_t o;};

/* the plane / buffer arithmetic is the library's (lb200_le_plane_location, lb200_le_ic_to_buff) */
static void le_fill_options(cs_t * cs, const lees_edw_options_t * opts, lb200_options_t * o) {
  memset(o, 0, sizeof(*o));
  for (int a = 0; a < 3; a++) { o->nlocal[a] = cs->nlocal[a]; o->periodic[a] = cs->periodic[a]; }
  o->nhalo = cs->nhalo;
  o->cart_size = 1;
  o->le_nplanes = opts->nplanes; o->le_uy = opts->uy; o->le_nt0 = opts->nt0;
}

int lees_edw_create(pe_t * pe, cs_t * cs, const lees_edw_options_t * opts, lees_edw_t ** ple) {
  lees_edw_t * le = (lees_edw_t *) calloc(1, sizeof(lees_edw_t));
  if (le == NULL) pe_fatal(pe, "calloc(lees_edw_t) failed\n");
  le->pe = pe; le->cs = cs;
  if (opts) le->opts = *opts;
  if (le->opts.nplanes > 0) {
    if (cs->ctx != NULL) pe_fatal(pe, "lees_edw_create: create the planes before the first device use of the coordinate system\n");
    if (le->opts.type == LE_SHEAR_TYPE_OSCILLATORY) pe_fatal(pe, "Oscillatory shear is outside this build\n");
    if (cs->ntotal[X] % le->opts.nplanes) pe_fatal(pe, "Number of planes must divide system size\n");     /* :249-253 */
    le_fill_options(cs, &le->opts, &le->o);
    for (int p = 0; p < le->opts.nplanes; p++) {
      int ic = lb200_le_plane_location(&le->o, p);
      if (ic <= cs->nhalo || ic > cs->nlocal[X] - cs->nhalo) pe_fatal(pe, "Wall at domain boundary\n");    /* :449-460 */
    }
    cs->le = le;
  }
  else {
    le_fill_options(cs, &le->opts, &le->o);
  }
  *ple = le;
  return 0;
}
int lees_edw_free(lees_edw_t * le) { if (le && le->cs && le->cs->le == le) le->cs->le = NULL; free(le); return 0; }
int lees_edw_nplane_total(lees_edw_t * le) { return le ? le->opts.nplanes : 0; }
int lees_edw_nplane_local(lees_edw_t * le) { return le ? le->opts.nplanes : 0; }
int lees_edw_plane_uy(lees_edw_t * le, double * uy) { *uy = le->opts.uy; return 0; }
int lees_edw_nxbuffer(lees_edw_t * le, int * nxb) { *nxb = 2*le->cs->nhalo*le->opts.nplanes; return 0; }
int lees_edw_nsites(lees_edw_t * le, int * nsites) {
  cs_t * cs = le->cs;
  *nsites = (cs->nall[X] + 2*cs->nhalo*le->opts.nplanes)*cs->nall[Y]*cs->nall[Z];        /* :485-495 */
  return 0;
}
int lees_edw_index(lees_edw_t * le, int ic, int jc, int kc) { return cs_index(le->cs, ic, jc, kc); }    /* :783-798 */
int lees_edw_plane_location(lees_edw_t * le, int np) { return lb200_le_plane_location(&le->o, np); }
int lees_edw_ic_to_buff(lees_edw_t * le, int ic, int di) { return lb200_le_ic_to_buff(&le->o, ic, di); }
int lees_edw_ibuff_to_real(lees_edw_t * le, int ib) {                                                   /* :1008-1022 */
  const int nh = le->cs->nhalo;
  return lees_edw_plane_location(le, ib/(2*nh)) - (nh - 1) + ib % (2*nh);
}
int lees_edw_shear_rate(lees_edw_t * le, double * gammadot) {                                           /* :759-769 */
  *gammadot = le->opts.uy*le->opts.nplanes/(1.0*le->cs->ntotal[X]);
  return 0;
}
int lees_edw_steady_uy(lees_edw_t * le, int ic, double * uy) {                                          /* :508-533 */
  const double dx_sep = 1.0*le->cs->ntotal[X]/le->opts.nplanes, dx_min = 0.5*dx_sep;
  double gammadot, xglobal = le->cs->noffset[X] + (double) ic - 0.5;
  int nplane = (int) ((dx_min + xglobal)/dx_sep);
  lees_edw_shear_rate(le, &gammadot);
  *uy = xglobal*gammadot - le->opts.uy*nplane;
  return 0;
}

static void b200_le_options(cs_t * cs, lb200_options_t * o) {
  if (cs->le == NULL) return;
  o->le_nplanes = cs->le->opts.nplanes; o->le_uy = cs->le->opts.uy; o->le_nt0 = cs->le->opts.nt0;
}

/* the device context keeps its own copy of the step counter (the reference's kernels read the singleton) */
static void b200_time_sync(cs_t * cs) {
  if (cs->le == NULL || physics_static == NULL) return;
  b200_check(cs->pe, lb200_physics_control_time_set(cs_b200_context(cs), physics_static->t_start, physics_static->t_current),
	     "physics_control_time");
}

static int le_nsites_or(cs_t * cs, lees_edw_t * le) {
  int ns = cs->nsites;
  if (le && le->opts.nplanes > 0) lees_edw_nsites(le, &ns);
  return ns;
}

/* ---- lb_t -------------------------------------------------------------------------------------------- */

static const signed char cv19_[19][3] = {
  { 0,  0,  0},
  { 1,  1,  0}, { 1,  0,  1}, { 1,  0,  0}, { 1,  0, -1}, { 1, -1,  0}, { 0,  1,  1},
  { 0,  1,  0}, { 0,  1, -1}, { 0,  0,  1}, { 0,  0, -1}, { 0, -1,  1}, { 0, -1,  0},
  { 0, -1, -1}, {-1,  1,  0}, {-1,  0,  1}, {-1,  0,  0}, {-1,  0, -1}, {-1, -1,  0}};
static const signed char cv15_[15][3] = {
  { 0,  0,  0},
  { 1,  1,  1}, { 1,  1, -1}, { 1,  0,  0}, { 1, -1,  1}, { 1, -1, -1}, { 0,  1,  0},
  { 0,  0,  1}, { 0,  0, -1}, { 0, -1,  0}, {-1,  1,  1}, {-1,  1, -1}, {-1,  0,  0},
  {-1, -1,  1}, {-1, -1, -1}};

static int lb_model_init(lb_model_t * m, int nvel) {
  m->ndim = 3;
  m->nvel = nvel;
  m->cs2 = (1.0/3.0);
  m->cv = (signed char (*)[3]) calloc(nvel, sizeof(signed char[3]));
  m->wv = (double *) calloc(nvel, sizeof(double));
  if (m->cv == NULL || m->wv == NULL) return -1;
  if (nvel == 19 || nvel == 15) {
    for (int p = 0; p < nvel; p++) {
      int c1 = 0;
      for (int a = 0; a < 3; a++) { m->cv[p][a] = (nvel == 19) ? cv19_[p][a] : cv15_[p][a]; c1 += abs(m->cv[p][a]); }
      if (nvel == 19) m->wv[p] = (c1 == 0) ? 12.0/36.0 : (c1 == 1) ? 2.0/36.0 : 1.0/36.0;
      else            m->wv[p] = (c1 == 0) ? 16.0/72.0 : (c1 == 1) ? 8.0/72.0 : 1.0/72.0;
    }
  }
  else if (nvel == 27) {
    int p = 1;
    m->wv[0] = 64.0/216.0;
    for (int i = -1; i <= 1; i++)
      for (int j = -1; j <= 1; j++)
	for (int k = -1; k <= 1; k++) {
	  int c1 = abs(i) + abs(j) + abs(k);
	  if (c1 == 0) continue;
	  m->cv[p][X] = i; m->cv[p][Y] = j; m->cv[p][Z] = k;
	  m->wv[p] = (c1 == 1) ? 16.0/216.0 : (c1 == 2) ? 4.0/216.0 : 1.0/216.0;
	  p++;
	}
  }
  else return -1;
  return 0;
}

/* src/io_options.c:36-120, src/io_info_args.c:24-36 */
io_options_t io_options_with_format(io_mode_enum_t mode, io_record_format_enum_t iorf) {
  io_options_t o = {.mode = mode, .iorformat = iorf, .metadata_version = IO_METADATA_V2, .report = 0, .asynchronous = 0,
		    .compression_levl = 0, .iogrid = {1, 1, 1}};
  return o;
}
io_options_t io_options_default(void) { return io_options_with_format(IO_MODE_MPIIO, IO_RECORD_BINARY); }
io_info_args_t io_info_args_default(void) {
  io_info_args_t a = {.input = io_options_default(), .output = io_options_default(), .grid = {1, 1, 1}, .iofreq = 100000};
  return a;
}

lb_data_options_t lb_data_options_default(void) {
  lb_data_options_t o = {.ndim = 3, .nvel = 19, .ndist = 1, .nrelax = LB_RELAXATION_M10, .halo = LB_HALO_FULL,
			 .reportimbalance = 0, .usefirsttouch = 0, .iodata = io_info_args_default()};
  return o;
}

lb_data_options_t lb_data_options_ndim_nvel_ndist(int ndim, int nvel, int ndist) {
  lb_data_options_t o = lb_data_options_default();
  o.ndim = ndim; o.nvel = nvel; o.ndist = ndist;
  return o;
}

int lb_data_create(pe_t * pe, cs_t * cs, const lb_data_options_t * opts, lb_t ** plb) {
  lb_t * lb = (lb_t *) calloc(1, sizeof(lb_t));
  assert(pe); assert(cs); assert(opts); assert(plb);
  if (lb == NULL) pe_fatal(pe, "calloc(1, lb_t) failed\n");
  if (opts->ndist != 1 && opts->ndist != 2) pe_fatal(pe, "ndist = %d\n", opts->ndist);
  lb->pe = pe; lb->cs = cs;
  lb->ndim = opts->ndim; lb->nvel = opts->nvel; lb->ndist = opts->ndist;
  lb->nrelax = opts->nrelax; lb->haloscheme = opts->halo; lb->opts = *opts;
  lb->nsite = cs->nsites;
  if (lb_model_init(&lb->model, opts->nvel) != 0) pe_fatal(pe, "unsupported nvel %d\n", opts->nvel);
  /* reference guard: src/lb_data.c:116-120 */
  if ((long long) lb->nsite*lb->ndist*lb->nvel > 2147483647LL) pe_fatal(pe, "local lattice too large for int indexing\n");
  lb->f = (double *) calloc((size_t) lb->nsite*lb->ndist*lb->nvel, sizeof(double));
  if (lb->f == NULL) pe_fatal(pe, "calloc(lb->f) failed\n");
  lb->target = lb;
  cs->lb = lb;
  *plb = lb;
  return 0;
}

int lb_free(lb_t * lb) {
  if (lb == NULL) return 0;
  if (lb->cs && lb->cs->lb == lb) lb->cs->lb = NULL;
  free(lb->model.cv); free(lb->model.wv); free(lb->f);
  free(lb);
  return 0;
}

int lb_memcpy(lb_t * lb, tdpMemcpyKind flag) {
  b200_check(lb->pe, lb200_memcpy(cs_b200_context(lb->cs), LB200_F, lb->f, b200_kind(flag)), "lb_memcpy");
  return 0;
}

int lb_halo(lb_t * lb) {
  b200_check(lb->pe, lb200_lb_halo(cs_b200_context(lb->cs)), "lb_halo");
  return 0;
}

int lb_propagation(lb_t * lb) {
  b200_check(lb->pe, lb200_lb_propagation(cs_b200_context(lb->cs)), "lb_propagation");
  return 0;
}

/* src/model_le.c:78-180 */
int lb_data_apply_le_boundary_conditions(lb_t * lb, lees_edw_t * le) {
  assert(lb); assert(le);
  if (lees_edw_nplane_total(le) == 0) return 0;
  b200_time_sync(lb->cs);
  b200_check(lb->pe, lb200_lb_le_apply_boundary_conditions(cs_b200_context(lb->cs)), "lb_data_apply_le_boundary_conditions");
  return 0;
}

/* src/model_le.c:652-714: host distributions consistent with a steady linear shear profile */
int lb_le_init_shear_profile(lb_t * lb, lees_edw_t * le) {
  physics_t * phys = NULL;
  int nlocal[3];
  double rho0, eta, u[3] = {0.0, 0.0, 0.0}, gradu[3][3] = {{0.0}};
  const double cs2 = lb->model.cs2, rcs2 = 1.0/cs2;
  assert(lb); assert(le);
  physics_ref(&phys);
  physics_rho0(phys, &rho0);
  physics_eta_shear(phys, &eta);
  cs_nlocal(lb->cs, nlocal);
  lees_edw_shear_rate(le, &gradu[X][Y]);
  for (int ic = 1; ic <= nlocal[X]; ic++) {
    lees_edw_steady_uy(le, ic, &u[Y]);
    for (int jc = 1; jc <= nlocal[Y]; jc++) {
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(lb->cs, ic, jc, kc);
	for (int p = 0; p < lb->nvel; p++) {
	  double f, cdotu = 0.0, sdotq = 0.0;
	  for (int i = 0; i < 3; i++) {
	    cdotu += lb->model.cv[p][i]*u[i];
	    for (int j = 0; j < 3; j++) {
	      double dij = (i == j);
	      double qij = lb->model.cv[p][i]*lb->model.cv[p][j] - cs2*dij;
	      sdotq += (rho0*u[i]*u[j] - eta*gradu[i][j])*qij;
	    }
	  }
	  f = lb->model.wv[p]*(rho0 + rcs2*rho0*cdotu + 0.5*rcs2*rcs2*sdotq);
	  lb_f_set(lb, index, p, 0, f);
	}
      }
    }
  }
  return 0;
}

int lb_collision_relaxation_set(lb_t * lb, lb_relaxation_enum_t nrelax) { lb->nrelax = nrelax; return 0; }

int lb_f(lb_t * lb, int index, int p, int n, double * f) {
  *f = lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, n, p)];
  return 0;
}

int lb_f_set(lb_t * lb, int index, int p, int n, double f) {
  lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, n, p)] = f;
  return 0;
}

int lb_0th_moment(lb_t * lb, int index, lb_dist_enum_t nd, double * rho) {
  *rho = 0.0;
  for (int p = 0; p < lb->nvel; p++) *rho += lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, nd, p)];
  return 0;
}

int lb_1st_moment(lb_t * lb, int index, lb_dist_enum_t nd, double g[3]) {
  for (int a = 0; a < 3; a++) g[a] = 0.0;
  for (int p = 0; p < lb->nvel; p++)
    for (int a = 0; a < 3; a++)
      g[a] += lb->model.cv[p][a]*lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, nd, p)];
  return 0;
}

/* src/lb_data.c:809-834 */
int lb_1st_moment_equilib_set(lb_t * lb, int index, double rho, double u[3]) {
  for (int p = 0; p < lb->model.nvel; p++) {
    double cs2 = lb->model.cs2;
    double rcs2 = 1.0/cs2;
    double udotc = 0.0;
    double sdotq = 0.0;
    for (int ia = 0; ia < 3; ia++) {
      udotc += u[ia]*lb->model.cv[p][ia];
      for (int ib = 0; ib < 3; ib++) {
	double dab = (ia == ib);
	sdotq += (lb->model.cv[p][ia]*lb->model.cv[p][ib] - cs2*dab)*u[ia]*u[ib];
      }
    }
    lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, LB_RHO, p)]
      = rho*lb->model.wv[p]*(1.0 + rcs2*udotc + 0.5*rcs2*rcs2*sdotq);
  }
  return 0;
}

/* src/lb_data.c:659-680 */
int lb_init_rest_f(lb_t * lb, double rho0) {
  int nlocal[3];
  cs_nlocal(lb->cs, nlocal);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	double u0[3] = {0.0, 0.0, 0.0};
	lb_1st_moment_equilib_set(lb, cs_index(lb->cs, ic, jc, kc), rho0, u0);
      }
  return 0;
}

/* ---- field_t ------------------------------------------------------------------------------------------- */

field_options_t field_options_default(void) { field_options_t o = {.ndata = 1, .nhcomm = 0, .iodata = io_info_args_default()}; return o; }
field_options_t field_options_ndata_nhalo(int ndata, int nhalo) {
  field_options_t o = {.ndata = ndata, .nhcomm = nhalo, .iodata = io_info_args_default()};
  return o;
}

static int field_create_tagged(pe_t * pe, cs_t * cs, lees_edw_t * le, const char * name,
			       const field_options_t * opts, int tag, field_t ** pobj) {
  field_t * obj = (field_t *) calloc(1, sizeof(field_t));
  if (obj == NULL) pe_fatal(pe, "calloc(field_t) failed\n");
  obj->nf = opts->ndata; obj->nhcomm = opts->nhcomm; obj->opts = *opts;
  obj->pe = pe; obj->cs = cs; obj->le = le;
  obj->nsites = le_nsites_or(cs, le);                 /* src/field.c:196-197 */
  obj->name = strdup(name);
  obj->data = (double *) calloc((size_t) obj->nf*obj->nsites, sizeof(double));
  if (obj->data == NULL) pe_fatal(pe, "calloc(field->data) failed\n");
  obj->b200_array = tag;
  obj->target = obj;
  *pobj = obj;
  return 0;
}

int field_create(pe_t * pe, cs_t * cs, lees_edw_t * le, const char * name, const field_options_t * opts, field_t ** pobj) {
  int tag = -1;
  assert(pe); assert(cs); assert(opts); assert(pobj);
  /* the scalar order parameter of the symmetric free energy is the one device-backed user field */
  if (opts->ndata == 1 && cs->phi == NULL) tag = LB200_PHI;
  if (opts->ndata == NQAB && cs->q == NULL) tag = LB200_Q;          /* the liquid-crystal tensor order parameter */
  field_create_tagged(pe, cs, le, name, opts, tag, pobj);
  if (tag == LB200_PHI) cs->phi = *pobj;
  if (tag == LB200_Q) cs->q = *pobj;
  return 0;
}

int field_free(field_t * obj) {
  if (obj == NULL) return 0;
  if (obj->cs && obj->cs->phi == obj) obj->cs->phi = NULL;
  if (obj->cs && obj->cs->q == obj) obj->cs->q = NULL;
  free(obj->name); free(obj->data); free(obj);
  return 0;
}

int field_memcpy(field_t * obj, tdpMemcpyKind flag) {
  if (obj->b200_array < 0) return 0;
  b200_check(obj->pe, lb200_memcpy(cs_b200_context(obj->cs), obj->b200_array, obj->data, b200_kind(flag)), "field_memcpy");
  return 0;
}

int field_halo(field_t * obj) {
  lb200_t * ctx = cs_b200_context(obj->cs);
  if (obj->b200_array == LB200_PHI) b200_check(obj->pe, lb200_phi_halo(ctx), "field_halo");
  else if (obj->b200_array == LB200_U) b200_check(obj->pe, lb200_hydro_u_halo(ctx), "field_halo");
  else if (obj->b200_array == LB200_Q) b200_check(obj->pe, lb200_q_halo(ctx), "field_halo");
  else pe_fatal(obj->pe, "field_halo: field \"%s\" has no device halo in this build\n", obj->name);
  return 0;
}

/* src/field.c:418-510 */
int field_leesedwards(field_t * obj) {
  if (obj->le == NULL || lees_edw_nplane_total(obj->le) == 0) return 0;
  if (obj->b200_array != LB200_PHI) pe_fatal(obj->pe, "field_leesedwards: field \"%s\" is not device backed\n", obj->name);
  b200_time_sync(obj->cs);
  b200_check(obj->pe, lb200_field_leesedwards(cs_b200_context(obj->cs)), "field_leesedwards");
  return 0;
}

/* src/field.c:1541-1560: the swap with an explicit scheme; one device scheme here */
int field_halo_swap(field_t * obj, field_halo_enum_t flag) {
  (void) flag;
  return field_halo(obj);
}

int field_nf(field_t * obj, int * nop) { *nop = obj->nf; return 0; }
int field_scalar(field_t * obj, int index, double * phi) { *phi = obj->data[addr_rank1(obj->nsites, 1, index, 0)]; return 0; }
int field_scalar_set(field_t * obj, int index, double phi) { obj->data[addr_rank1(obj->nsites, 1, index, 0)] = phi; return 0; }
int field_vector(field_t * obj, int index, double p[3]) {
  for (int a = 0; a < 3; a++) p[a] = obj->data[addr_rank1(obj->nsites, 3, index, a)];
  return 0;
}
int field_vector_set(field_t * obj, int index, const double p[3]) {
  for (int a = 0; a < 3; a++) obj->data[addr_rank1(obj->nsites, 3, index, a)] = p[a];
  return 0;
}

/* ---- field_grad_t ------------------------------------------------------------------------------------------ */

int field_grad_create(pe_t * pe, field_t * f, int level, field_grad_t ** pobj) {
  field_grad_t * obj = (field_grad_t *) calloc(1, sizeof(field_grad_t));
  if (obj == NULL) pe_fatal(pe, "calloc(field_grad_t) failed\n");
  obj->pe = pe; obj->field = f; obj->nf = f->nf; obj->level = level; obj->nsite = f->nsites;
  obj->grad = (double *) calloc((size_t) 3*obj->nf*obj->nsite, sizeof(double));
  obj->delsq = (double *) calloc((size_t) obj->nf*obj->nsite, sizeof(double));
  if (obj->grad == NULL || obj->delsq == NULL) pe_fatal(pe, "calloc(field_grad) failed\n");
  if (level >= 4) {
    /* src/field_grad.c:112-135 */
    obj->grad_delsq = (double *) calloc((size_t) 3*obj->nf*obj->nsite, sizeof(double));
    obj->delsq_delsq = (double *) calloc((size_t) obj->nf*obj->nsite, sizeof(double));
    if (obj->grad_delsq == NULL || obj->delsq_delsq == NULL) pe_fatal(pe, "calloc(field_grad d4) failed\n");
  }
  obj->target = obj;
  if (f->cs->phi == f) f->cs->phi_grad = obj;
  if (f->cs->q == f) f->cs->q_grad = obj;
  *pobj = obj;
  return 0;
}

void field_grad_free(field_grad_t * obj) {
  if (obj == NULL) return;
  if (obj->field && obj->field->cs && obj->field->cs->phi_grad == obj) obj->field->cs->phi_grad = NULL;
  free(obj->grad); free(obj->delsq); free(obj->grad_delsq); free(obj->delsq_delsq); free(obj);
}

int field_grad_set(field_grad_t * obj, grad_ft d2, grad_ft d4) { obj->d2 = d2; obj->d4 = d4; return 0; }

/* src/field_grad.c:319-340 */
int field_grad_compute(field_grad_t * obj) {
  assert(obj);
  assert(obj->d2);
  obj->d2(obj);
  if (obj->level >= 4) {
    assert(obj->d4);
    obj->d4(obj);
  }
  return 0;
}

int grad_3d_27pt_fluid_d2(field_grad_t * fg) {
  b200_check(fg->pe, lb200_set_knob(cs_b200_context(fg->field->cs), LB200_KNOB_GRAD_7PT, 0), "grad_3d_27pt_fluid_d2");
  b200_time_sync(fg->field->cs);       /* with planes: field_leesedwards + d2 + grad_3d_27pt_fluid_le on the device */
  b200_check(fg->pe, lb200_phi_grad_compute(cs_b200_context(fg->field->cs)), "grad_3d_27pt_fluid_d2");
  return 0;
}

/* src/gradient_3d_27pt_fluid.c:112-134 */
int grad_3d_27pt_fluid_d4(field_grad_t * fg) {
  b200_check(fg->pe, lb200_phi_grad_compute_d4(cs_b200_context(fg->field->cs)), "grad_3d_27pt_fluid_d4");
  fg->d4_on_device = 1;
  return 0;
}

/* src/gradient_3d_7pt_fluid.c:76-99 */
int grad_3d_7pt_fluid_d2(field_grad_t * fg) {
  if (fg->field->b200_array == LB200_PHI) {
    /* the scalar order parameter with fd_gradient_calculation 3d_7pt_fluid: field_grad_set(obj, grad_3d_7pt_fluid_d2, ...) */
    lb200_t * ctx = cs_b200_context(fg->field->cs);
    b200_time_sync(fg->field->cs);
    b200_check(fg->pe, lb200_set_knob(ctx, LB200_KNOB_GRAD_7PT, 1), "grad_3d_7pt_fluid_d2");
    b200_check(fg->pe, lb200_phi_grad_compute(ctx), "grad_3d_7pt_fluid_d2");
    return 0;
  }
  if (fg->field->b200_array != LB200_Q) pe_fatal(fg->pe, "grad_3d_7pt_fluid_d2: only phi and the Q tensor field are device backed\n");
  b200_check(fg->pe, lb200_q_grad_compute(cs_b200_context(fg->field->cs)), "grad_3d_7pt_fluid_d2");
  return 0;
}

int field_grad_memcpy(field_grad_t * obj, tdpMemcpyKind flag) {
  if (obj->field->b200_array == LB200_Q) {
    lb200_t * cq = cs_b200_context(obj->field->cs);
    b200_check(obj->pe, lb200_memcpy(cq, LB200_QGRAD, obj->grad, b200_kind(flag)), "field_grad_memcpy");
    b200_check(obj->pe, lb200_memcpy(cq, LB200_QDELSQ, obj->delsq, b200_kind(flag)), "field_grad_memcpy");
    return 0;
  }
  lb200_t * ctx = cs_b200_context(obj->field->cs);
  b200_check(obj->pe, lb200_memcpy(ctx, LB200_GRAD, obj->grad, b200_kind(flag)), "field_grad_memcpy");
  b200_check(obj->pe, lb200_memcpy(ctx, LB200_DELSQ, obj->delsq, b200_kind(flag)), "field_grad_memcpy");
  if (obj->level >= 4 && obj->d4_on_device && flag == tdpMemcpyDeviceToHost) {
    b200_check(obj->pe, lb200_memcpy(ctx, LB200_GRAD_DELSQ, obj->grad_delsq, b200_kind(flag)), "field_grad_memcpy");
    b200_check(obj->pe, lb200_memcpy(ctx, LB200_DELSQ_DELSQ, obj->delsq_delsq, b200_kind(flag)), "field_grad_memcpy");
  }
  return 0;
}

int field_grad_scalar_grad(field_grad_t * obj, int index, double grad[3]) {
  for (int a = 0; a < 3; a++) grad[a] = obj->grad[addr_rank2(obj->nsite, 1, 3, index, 0, a)];
  return 0;
}
int field_grad_scalar_delsq(field_grad_t * obj, int index, double * delsq) {
  *delsq = obj->delsq[addr_rank1(obj->nsite, 1, index, 0)];
  return 0;
}

/* ---- hydro_t -------------------------------------------------------------------------------------------------- */

hydro_options_t hydro_options_nhalo(int nhalo) {
  hydro_options_t o = {.nhcomm = nhalo, .rho = field_options_ndata_nhalo(1, nhalo), .u = field_options_ndata_nhalo(3, nhalo),
		       .force = field_options_ndata_nhalo(3, nhalo), .eta = field_options_ndata_nhalo(1, nhalo)};
  return o;
}
hydro_options_t hydro_options_default(void) { return hydro_options_nhalo(1); }

int hydro_create(pe_t * pe, cs_t * cs, lees_edw_t * le, const hydro_options_t * opts, hydro_t ** pobj) {
  hydro_t * obj = (hydro_t *) calloc(1, sizeof(hydro_t));
  if (obj == NULL) pe_fatal(pe, "calloc(hydro) failed\n");
  obj->pe = pe; obj->cs = cs; obj->le = le; obj->nhcomm = opts->nhcomm; obj->nsite = le_nsites_or(cs, le);
  field_create_tagged(pe, cs, le, "rho", &opts->rho, LB200_RHO, &obj->rho);
  field_create_tagged(pe, cs, le, "u", &opts->u, LB200_U, &obj->u);
  field_create_tagged(pe, cs, le, "force", &opts->force, LB200_FORCE, &obj->force);
  field_create_tagged(pe, cs, le, "eta", &opts->eta, -1, &obj->eta);
  obj->target = obj;
  cs->hydro = obj;
  *pobj = obj;
  return 0;
}

int hydro_free(hydro_t * obj) {
  if (obj == NULL) return 0;
  if (obj->cs && obj->cs->hydro == obj) obj->cs->hydro = NULL;
  field_free(obj->rho); field_free(obj->u); field_free(obj->force); field_free(obj->eta);
  free(obj);
  return 0;
}

/* src/hydro.c:123-160: rho, u, force */
int hydro_memcpy(hydro_t * obj, tdpMemcpyKind flag) {
  field_memcpy(obj->rho, flag);
  field_memcpy(obj->u, flag);
  field_memcpy(obj->force, flag);
  return 0;
}

int hydro_u_halo(hydro_t * obj) { return field_halo(obj->u); }

/* src/hydro.c:350-440 */
int hydro_lees_edwards(hydro_t * obj) {
  if (obj->le == NULL || lees_edw_nplane_total(obj->le) == 0) return 0;
  b200_time_sync(obj->cs);
  b200_check(obj->pe, lb200_hydro_lees_edwards(cs_b200_context(obj->cs)), "hydro_lees_edwards");
  return 0;
}

int hydro_f_zero(hydro_t * obj, const double fzero[3]) {
  if (fzero[X] != 0.0 || fzero[Y] != 0.0 || fzero[Z] != 0.0) pe_fatal(obj->pe, "hydro_f_zero: non-zero value not supported\n");
  b200_check(obj->pe, lb200_hydro_f_zero(cs_b200_context(obj->cs)), "hydro_f_zero");
  return 0;
}

int hydro_u_zero(hydro_t * obj, const double uzero[3]) {
  if (uzero[X] != 0.0 || uzero[Y] != 0.0 || uzero[Z] != 0.0) pe_fatal(obj->pe, "hydro_u_zero: non-zero value not supported\n");
  b200_check(obj->pe, lb200_hydro_u_zero(cs_b200_context(obj->cs)), "hydro_u_zero");
  return 0;
}

int hydro_u(hydro_t * obj, int index, double u[3]) { return field_vector(obj->u, index, u); }
int hydro_u_set(hydro_t * obj, int index, const double u[3]) { return field_vector_set(obj->u, index, u); }
int hydro_f_local(hydro_t * obj, int index, double f[3]) { return field_vector(obj->force, index, f); }
int hydro_f_local_set(hydro_t * obj, int index, const double f[3]) { return field_vector_set(obj->force, index, f); }
int hydro_rho(hydro_t * obj, int index, double * rho) { return field_scalar(obj->rho, index, rho); }

/* ---- map_t ------------------------------------------------------------------------------------------------------ */

map_options_t map_options_default(void) { map_options_t o = {.ndata = 0, .is_porous_media = 0}; return o; }

int map_create(pe_t * pe, cs_t * cs, const map_options_t * options, map_t ** pmap) {
  map_t * map = (map_t *) calloc(1, sizeof(map_t));
  (void) options;
  if (map == NULL) pe_fatal(pe, "calloc(map_t) failed\n");
  map->pe = pe; map->cs = cs; map->nsite = cs->nsites;
  map->status = (char *) calloc(map->nsite, sizeof(char));     /* MAP_FLUID everywhere */
  if (map->status == NULL) pe_fatal(pe, "calloc(map->status) failed\n");
  map->target = map;
  cs->map = map;
  *pmap = map;
  return 0;
}

int map_free(map_t ** pmap) {
  map_t * map = *pmap;
  if (map == NULL) return 0;
  if (map->cs && map->cs->map == map) map->cs->map = NULL;
  free(map->status); free(map);
  *pmap = NULL;
  return 0;
}

int map_memcpy(map_t * map, tdpMemcpyKind flag) {
  double * tmp = (double *) malloc((size_t) map->nsite*sizeof(double));
  if (tmp == NULL) pe_fatal(map->pe, "malloc failed\n");
  if (flag == tdpMemcpyHostToDevice) for (int i = 0; i < map->nsite; i++) tmp[i] = (double) map->status[i];
  b200_check(map->pe, lb200_memcpy(cs_b200_context(map->cs), LB200_MAP, tmp, b200_kind(flag)), "map_memcpy");
  if (flag == tdpMemcpyDeviceToHost) for (int i = 0; i < map->nsite; i++) map->status[i] = (char) tmp[i];
  free(tmp);
  return 0;
}

int map_status(map_t * map, int index, int * status) { *status = (int) map->status[index]; return 0; }
int map_status_set(map_t * map, int index, int status) { map->status[index] = (char) status; return 0; }

/* ---- symmetric free energy ------------------------------------------------------------------------------------------ */

int fe_symm_create(pe_t * pe, cs_t * cs, field_t * f, field_grad_t * grd, fe_symm_t ** p) {
  fe_symm_t * fe = (fe_symm_t *) calloc(1, sizeof(fe_symm_t));
  if (fe == NULL) pe_fatal(pe, "calloc(fe_symm_t) failed\n");
  fe->param = (fe_symm_param_t *) calloc(1, sizeof(fe_symm_param_t));
  fe->pe = pe; fe->cs = cs; fe->phi = f; fe->dphi = grd;
  fe->super.id = 1;      /* FE_SYMMETRIC */
  fe->target = fe;
  *p = fe;
  return 0;
}
/* src/symmetric.c:160-170: the device-side object; parameters travel with every call here, so the host object is it */
int fe_symm_target(fe_symm_t * fe, fe_t ** target) { *target = (fe_t *) fe->target; return 0; }

/* src/lb_data.c:585-600: constants are passed to the kernels with every launch; nothing to commit */
int lb_collide_param_commit(lb_t * lb) { assert(lb); return 0; }

int fe_symm_free(fe_symm_t * fe) { if (fe) { free(fe->param); free(fe); } return 0; }
int fe_symm_param_set(fe_symm_t * fe, fe_symm_param_t values) { *fe->param = values; return 0; }
int fe_symm_param(fe_symm_t * fe, fe_symm_param_t * values) { *values = *fe->param; return 0; }

/* src/symmetric.c:284-299 (host arrays) */
int fe_symm_fed(fe_symm_t * fe, int index, double * fed) {
  double phi, dphi[3];
  field_scalar(fe->phi, index, &phi);
  field_grad_scalar_grad(fe->dphi, index, dphi);
  *fed = (0.5*fe->param->a + 0.25*fe->param->b*phi*phi)*phi*phi
    + 0.5*fe->param->kappa*(dphi[X]*dphi[X] + dphi[Y]*dphi[Y] + dphi[Z]*dphi[Z]);
  return 0;
}

/* src/symmetric.c:333-362 (host arrays): P_ab = p0 delta_ab + kappa d_a phi d_b phi */
int fe_symm_str(fe_symm_t * fe, int index, double s[3][3]) {
  double phi, delsq, dphi[3], p0;
  const double kappa = fe->param->kappa;
  field_scalar(fe->phi, index, &phi);
  field_grad_scalar_grad(fe->dphi, index, dphi);
  delsq = fe->dphi->delsq[addr_rank0(fe->phi->nsites, index)];
  p0 = 0.5*fe->param->a*phi*phi + 0.75*fe->param->b*phi*phi*phi*phi
    - kappa*phi*delsq - 0.5*kappa*(dphi[X]*dphi[X] + dphi[Y]*dphi[Y] + dphi[Z]*dphi[Z]);
  for (int ia = 0; ia < 3; ia++) {
    for (int ib = 0; ib < 3; ib++) {
      const double d_ab = (ia == ib);
      s[ia][ib] = p0*d_ab + kappa*dphi[ia]*dphi[ib];
    }
  }
  return 0;
}

/* src/symmetric.c:307-319 */
int fe_symm_mu(fe_symm_t * fe, int index, double * mu) {
  double phi = fe->phi->data[addr_rank0(fe->phi->nsites, index)];
  double delsq = fe->dphi->delsq[addr_rank0(fe->phi->nsites, index)];
  *mu = fe->param->a*phi + fe->param->b*phi*phi*phi - fe->param->kappa*delsq;
  return 0;
}

static void symm_param_from(fe_t * fe, lb200_symm_param_t * sp) {
  fe_symm_t * fs = (fe_symm_t *) fe;
  physics_t * phys = NULL;
  memset(sp, 0, sizeof(*sp));
  physics_ref(&phys);
  sp->a = fs->param->a; sp->b = fs->param->b; sp->kappa = fs->param->kappa;
  physics_mobility(phys, &sp->mobility);
  physics_grad_mu(phys, sp->gradmu);
  advection_order(&sp->adv_order);
}

/* ---- phi_force ---------------------------------------------------------------------------------------------------------- */

int pth_create(pe_t * pe, cs_t * cs, int method, pth_t ** ppth) {
  pth_t * pth = (pth_t *) calloc(1, sizeof(pth_t));
  if (pth == NULL) pe_fatal(pe, "calloc(pth_t) failed\n");
  pth->pe = pe; pth->cs = cs; pth->method = method; pth->nsites = cs->nsites;
  pth->target = pth;
  *ppth = pth;
  return 0;
}
int pth_free(pth_t * pth) { free(pth); return 0; }

/* src/phi_force_stress.c:171-217 */
static void lc_param_from(fe_t * fe, const struct beris_edw_s * be, lb200_lc_param_t * lc);

int pth_stress_compute(pth_t * pth, fe_t * fe) {
  lb200_symm_param_t sp;
  assert(pth); assert(fe);
  if (fe->id == FE_LC_ID) {
    /* fe->func->stress_v == fe_lc_stress_v, src/phi_force_stress.c:256-284 */
    lb200_lc_param_t lc;
    lc_param_from(fe, NULL, &lc);
    b200_check(pth->pe, lb200_lc_stress_compute(cs_b200_context(pth->cs), &lc), "pth_stress_compute");
    return 0;
  }
  symm_param_from(fe, &sp);
  b200_check(pth->pe, lb200_pth_stress_compute(cs_b200_context(pth->cs), &sp), "pth_stress_compute");
  return 0;
}

/* src/phi_force_colloid.c:274-301 */
int pth_force_fluid_driver(pth_t * pth, hydro_t * hydro) {
  assert(pth); assert(hydro);
  b200_check(pth->pe, lb200_pth_force_fluid_driver(cs_b200_context(pth->cs)), "pth_force_fluid_driver");
  return 0;
}

/* src/phi_force.c:74-137 */
int phi_force_calculation(pe_t * pe, cs_t * cs, lees_edw_t * le, wall_t * wall, pth_t * pth, fe_t * fe,
			  map_t * map, field_t * phi, hydro_t * hydro) {
  lb200_symm_param_t sp;
  (void) map; (void) phi;
  assert(pth);
  if (hydro == NULL) return 0;
  if (pth->method == FE_FORCE_METHOD_NO_FORCE) return 0;
  if (wall != NULL) pe_fatal(pe, "phi_force_calculation: walls are outside this build\n");
  (void) le;                       /* with planes the library uses the flux form, src/phi_force.c:91-97 */
  if (pth->method != FE_FORCE_METHOD_STRESS_DIVERGENCE) pe_fatal(pe, "Bad force method\n");
  if (fe->id == FE_LC_ID) {
    lb200_lc_param_t lc;
    lc_param_from(fe, NULL, &lc);
    b200_check(pe, lb200_lc_force_calculation(cs_b200_context(cs), &lc), "phi_force_calculation");
    return 0;
  }
  symm_param_from(fe, &sp);
  b200_check(pe, lb200_phi_force_calculation(cs_b200_context(cs), &sp), "phi_force_calculation");
  return 0;
}

/* ---- Cahn-Hilliard ---------------------------------------------------------------------------------------------------------- */

static int advection_order_ = 1;     /* src/advection.c:74 */
int advection_order_set(const int order) { advection_order_ = order; return 0; }
int advection_order(int * order) { *order = advection_order_; return 0; }

int phi_ch_create(pe_t * pe, cs_t * cs, lees_edw_t * le, phi_ch_info_t * info, phi_ch_t ** ppch) {
  phi_ch_t * pch = (phi_ch_t *) calloc(1, sizeof(phi_ch_t));
  if (pch == NULL) pe_fatal(pe, "calloc(phi_ch_t) failed\n");
  /* PHI_CONSERVE_COMPENSATED_SUM (1): the compensation field (pch->csum) lives in the device context;
   * PHI_CONSERVE_GLOBAL_SUBTRACT (2) needs the initial global sum and an all-reduce: outside this build */
  if (info->conserve != 0 && info->conserve != 1) pe_fatal(pe, "cahn_hilliard_options_conserve %d is outside this build\n", info->conserve);
  pch->pe = pe; pch->cs = cs; pch->le = le; pch->info = *info;
  *ppch = pch;
  return 0;
}
int phi_ch_free(phi_ch_t * pch) { free(pch); return 0; }

/* src/phi_cahn_hilliard.c:213-288 */
int phi_cahn_hilliard(phi_ch_t * pch, fe_t * fe, field_t * phi, hydro_t * hydro, map_t * map, noise_t * noise) {
  lb200_symm_param_t sp;
  (void) phi; (void) map;
  assert(pch); assert(fe);
  if (noise != NULL || pch->info.noise) pe_fatal(pch->pe, "phi_cahn_hilliard: noise is outside this build\n");
  if (hydro == NULL) pe_fatal(pch->pe, "phi_cahn_hilliard: hydro == NULL is outside this build\n");
  symm_param_from(fe, &sp);
  sp.conserve = pch->info.conserve;
  b200_time_sync(pch->cs);
  b200_check(pch->pe, lb200_phi_cahn_hilliard(cs_b200_context(pch->cs), &sp), "phi_cahn_hilliard");
  return 0;
}


/* ---- liquid crystal -------------------------------------------------------------------------------------------- */

int field_tensor(field_t * obj, int index, double q[3][3]) {           /* src/field.c: compressed XX XY XZ YY YZ */
  const double * d = obj->data;
  const int ns = obj->nsites;
  q[X][X] = d[addr_rank1(ns, NQAB, index, XX)]; q[X][Y] = d[addr_rank1(ns, NQAB, index, XY)];
  q[X][Z] = d[addr_rank1(ns, NQAB, index, XZ)]; q[Y][X] = q[X][Y];
  q[Y][Y] = d[addr_rank1(ns, NQAB, index, YY)]; q[Y][Z] = d[addr_rank1(ns, NQAB, index, YZ)];
  q[Z][X] = q[X][Z]; q[Z][Y] = q[Y][Z]; q[Z][Z] = 0.0 - q[X][X] - q[Y][Y];
  return 0;
}

int field_tensor_set(field_t * obj, int index, double q[3][3]) {
  double * d = obj->data;
  const int ns = obj->nsites;
  d[addr_rank1(ns, NQAB, index, XX)] = q[X][X]; d[addr_rank1(ns, NQAB, index, XY)] = q[X][Y];
  d[addr_rank1(ns, NQAB, index, XZ)] = q[X][Z]; d[addr_rank1(ns, NQAB, index, YY)] = q[Y][Y];
  d[addr_rank1(ns, NQAB, index, YZ)] = q[Y][Z];
  return 0;
}

int fe_lc_create(pe_t * pe, cs_t * cs, lees_edw_t * le, field_t * q, field_grad_t * dq, fe_lc_t ** pfe) {
  fe_lc_t * fe = (fe_lc_t *) calloc(1, sizeof(fe_lc_t));
  (void) le;
  if (fe == NULL) pe_fatal(pe, "calloc(fe_lc_t) failed\n");
  fe->param = (fe_lc_param_t *) calloc(1, sizeof(fe_lc_param_t));
  if (fe->param == NULL) pe_fatal(pe, "calloc(fe_lc_param_t) failed\n");
  fe->pe = pe; fe->cs = cs; fe->q = q; fe->dq = dq;
  fe->super.id = FE_LC_ID;
  fe->param->redshift = 1.0; fe->param->rredshift = 1.0; fe->param->coswt = 1.0;
  fe->target = fe;
  *pfe = fe;
  return 0;
}
int fe_lc_free(fe_lc_t * fe) { if (fe) { free(fe->param); free(fe); } return 0; }

/* src/blue_phase.c:241-258 */
int fe_lc_param_set(fe_lc_t * fe, const fe_lc_param_t * values) {
  const double pi = 3.1415926535897932385;           /* PI_DOUBLE, src/util.h */
  *fe->param = *values;
  fe->param->epsilon *= (1.0/(12.0*pi));
  if (fe->param->is_redshift_updated) pe_fatal(fe->pe, "liquid crystal: lc_redshift_update is outside this build\n");
  if (fe->param->is_active && fe->param->zeta2 != 0.0) pe_fatal(fe->pe, "liquid crystal: lc_active_zeta2 != 0 is outside this build\n");
  if (fe->param->redshift == 0.0) fe->param->redshift = 1.0;
  fe->param->rredshift = 1.0/fe->param->redshift;            /* fe_lc_redshift_set, src/blue_phase.c:1357-1366 */
  if (fe->param->coswt == 0.0) fe->param->coswt = 1.0;
  return 0;
}
int fe_lc_param(fe_lc_t * fe, fe_lc_param_t * vals) { *vals = *fe->param; return 0; }

/* src/blue_phase.c:1406-1418 */
int fe_lc_q_uniaxial(fe_lc_param_t * param, const double n[3], double q[3][3]) {
  for (int ia = 0; ia < 3; ia++)
    for (int ib = 0; ib < 3; ib++) q[ia][ib] = 0.5*param->amplitude0*(3.0*n[ia]*n[ib] - (ia == ib));
  return 0;
}

/* src/blue_phase_init.c:763-823 */
int blue_phase_twist_init(cs_t * cs, fe_lc_param_t * param, field_t * fq, int helical_axis) {
  double n[3] = {0.0, 0.0, 0.0}, q[3][3];
  const double q0 = param->q0;
  for (int ic = 1; ic <= cs->nlocal[X]; ic++) {
    if (helical_axis == X) { double x = cs->noffset[X] + ic; n[Y] = cos(q0*x); n[Z] = sin(q0*x); }
    for (int jc = 1; jc <= cs->nlocal[Y]; jc++) {
      if (helical_axis == Y) { double y = cs->noffset[Y] + jc; n[X] = cos(q0*y); n[Z] = -sin(q0*y); }
      for (int kc = 1; kc <= cs->nlocal[Z]; kc++) {
	if (helical_axis == Z) { double z = cs->noffset[Z] + kc; n[X] = cos(q0*z); n[Y] = sin(q0*z); }
	fe_lc_q_uniaxial(param, n, q);
	field_tensor_set(fq, cs_index(cs, ic, jc, kc), q);
      }
    }
  }
  return 0;
}

int beris_edw_create(pe_t * pe, cs_t * cs, lees_edw_t * le, beris_edw_t ** pobj) {
  beris_edw_t * be = (beris_edw_t *) calloc(1, sizeof(beris_edw_t));
  if (be == NULL) pe_fatal(pe, "calloc(beris_edw_t) failed\n");
  be->pe = pe; be->cs = cs; be->le = le;
  *pobj = be;
  return 0;
}
int beris_edw_free(beris_edw_t * be) { free(be); return 0; }
int beris_edw_param_set(beris_edw_t * be, beris_edw_param_t * values) { be->param = *values; return 0; }

static void lc_param_from(fe_t * fe, const beris_edw_t * be, lb200_lc_param_t * lc) {
  const fe_lc_param_t * p = ((fe_lc_t *) fe)->param;
  memset(lc, 0, sizeof(*lc));
  lc->a0 = p->a0; lc->q0 = p->q0; lc->gamma = p->gamma; lc->kappa0 = p->kappa0; lc->kappa1 = p->kappa1; lc->xi = p->xi;
  lc->epsilon = p->epsilon;
  for (int a = 0; a < 3; a++) lc->e0[a] = p->e0[a]*p->coswt;
  lc->Gamma = be ? be->param.gamma : 0.0;
  lc->adv_order = advection_order_;
  lc->is_active = p->is_active; lc->zeta0 = p->zeta0; lc->zeta1 = p->zeta1; lc->zeta2 = p->zeta2;
  lc->redshift = p->redshift;
}

/* src/blue_phase_beris_edwards.c:266-296 */
int beris_edw_update(beris_edw_t * be, fe_t * fe, field_t * fq, field_grad_t * fq_grad, hydro_t * hydro,
		     colloids_info_t * cinfo, map_t * map, noise_t * noise) {
  lb200_lc_param_t lc;
  (void) fq; (void) fq_grad; (void) map;
  if (hydro == NULL) pe_fatal(be->pe, "beris_edw_update: hydro == NULL is outside this build\n");
  if (cinfo != NULL) pe_fatal(be->pe, "beris_edw_update: colloids are outside this build\n");
  if (noise != NULL || be->param.noise) pe_fatal(be->pe, "beris_edw_update: noise is outside this build\n");
  lc_param_from(fe, be, &lc);
  b200_check(be->pe, lb200_beris_edw_update(cs_b200_context(be->cs), &lc), "beris_edw_update");
  return 0;
}

/* ---- on-disk formats --------------------------------------------------------------------------------------- */

/* src/io_subfile.c:186-204 (one file: index 0 of 1) */
static void io_file_name(const char * stub, int it, char * filename, size_t bufsz) {
  snprintf(filename, bufsz, "%s-%9.9d.%3.3d-%3.3d", stub, it, 1, 1);
}

/* the reference's metadata file for the default options (mode mpiio, binary records, one file): same text */
static int io_metadata_write_default(cs_t * cs, const char * stub, int count, int ascii) {
  char filename[BUFSIZ];
  FILE * fp = NULL;
  int nplanes = cs->le ? lees_edw_nplane_total(cs->le) : 0;
  snprintf(filename, BUFSIZ, "%s-metadata.%3.3d-%3.3d", stub, 1, 1);
  fp = fopen(filename, "w");
  if (fp == NULL) return -1;
  fprintf(fp, "{\n\t\"coords\":\t{\n\t\t\"options\":\t{\n");
  fprintf(fp, "\t\t\t\"System size (total)\":\t[%d, %d, %d],\n", cs->ntotal[X], cs->ntotal[Y], cs->ntotal[Z]);
  fprintf(fp, "\t\t\t\"Periodic boundaries\":\t[%d, %d, %d],\n", cs->periodic[X], cs->periodic[Y], cs->periodic[Z]);
  fprintf(fp, "\t\t\t\"Left-end limit Lmin\":\t[0.5, 0.5, 0.5]\n\t\t},\n");
  fprintf(fp, "\t\t\"lees_edwards\":\t{\n\t\t\t\"Number of planes\":\t%d", nplanes);
  if (nplanes > 0) {
    /* lees_edw_opts_to_json, src/lees_edwards_options.c; numbers as cJSON prints them (%1.15g, else %1.17g) */
    char num[64];
    double back = 0.0;
    snprintf(num, sizeof(num), "%1.15g", cs->le->opts.uy);
    if (sscanf(num, "%lg", &back) != 1 || back != cs->le->opts.uy) snprintf(num, sizeof(num), "%1.17g", cs->le->opts.uy);
    fprintf(fp, ",\n\t\t\t\"Shear type\":\t\"STEADY\",\n\t\t\t\"Reference time\":\t%d,\n\t\t\t\"Plane speed\":\t%s",
	    cs->le->opts.nt0, num);
  }
  fprintf(fp, "\n\t\t}\n\t},\n");
  fprintf(fp, "\t\"io_options\":\t{\n\t\t\"Mode\":\t\"mpiio\",\n\t\t\"Record format\":\t\"%s\",\n\t\t\"Metadata version\":\t3,\n"
	  "\t\t\"Report\":\tfalse,\n\t\t\"Asynchronous\":\tfalse,\n\t\t\"Compression level\":\t0,\n\t\t\"I/O grid\":\t[1, 1, 1]\n\t},\n", ascii ? "ascii" : "binary");
  fprintf(fp, "\t\"io_element\":\t{\n\t\t\"MPI_Datatype\":\t\"%s\",\n\t\t\"Size (bytes)\":\t%d,\n\t\t\"Count\":\t%d,\n"
	  "\t\t\"Endianness\":\t\"LITTLE_ENDIAN\"\n\t},\n", ascii ? "MPI_CHAR" : "MPI_DOUBLE", ascii ? 1 : 8, count);
  fprintf(fp, "\t\"io_subfile\":\t{\n\t\t\"Number of files\":\t1,\n\t\t\"File index\":\t0,\n\t\t\"Topology\":\t[1, 1, 1],\n"
	  "\t\t\"Coordinate\":\t[0, 0, 0],\n\t\t\"Data ndims\":\t3,\n\t\t\"File size (sites)\":\t[%d, %d, %d],\n"
	  "\t\t\"File offset (sites)\":\t[0, 0, 0]\n\t}\n}", cs->nlocal[X], cs->nlocal[Y], cs->nlocal[Z]);
  fclose(fp);
  return 0;
}

/* src/lb_data.c:1533-1575 */
int lb_write_buf(const lb_t * lb, int index, char * buf) {
  for (int n = 0; n < lb->ndist; n++) {
    size_t sz = lb->nvel*sizeof(double);
    double data[27];
    for (int p = 0; p < lb->nvel; p++) data[p] = lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, n, p)];
    memcpy(buf + n*sz, data, sz);
  }
  return 0;
}

int lb_read_buf(lb_t * lb, int index, const char * buf) {
  for (int n = 0; n < lb->ndist; n++) {
    size_t sz = lb->nvel*sizeof(double);
    double data[27];
    memcpy(data, buf + n*sz, sz);
    for (int p = 0; p < lb->nvel; p++) lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, n, p)] = data[p];
  }
  return 0;
}

/* src/field.c:896-930 */
int field_write_buf(field_t * field, int index, char * buf) {
  double array[9];
  for (int n = 0; n < field->nf; n++) array[n] = field->data[addr_rank1(field->nsites, field->nf, index, n)];
  memcpy(buf, array, field->nf*sizeof(double));
  return 0;
}

int field_read_buf(field_t * field, int index, const char * buf) {
  double array[9];
  memcpy(array, buf, field->nf*sizeof(double));
  for (int n = 0; n < field->nf; n++) field->data[addr_rank1(field->nsites, field->nf, index, n)] = array[n];
  return 0;
}

/* src/lb_data.c:1579-1640: ndist values per line, one line per velocity */
int lb_write_buf_ascii(const lb_t * lb, int index, char * buf) {
  const int nbyte = LB_RECORD_LENGTH_ASCII;
  int ifail = 0;
  for (int p = 0; p < lb->nvel; p++) {
    char tmp[64];
    const int poffset = p*(lb->ndist*nbyte + 1);
    for (int n = 0; n < lb->ndist; n++) {
      int np = snprintf(tmp, sizeof(tmp), " %22.15e", lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, n, p)]);
      if (np != nbyte) ifail = 1;
      memcpy(buf + poffset + n*nbyte, tmp, nbyte);
    }
    buf[poffset + lb->ndist*nbyte] = '\n';
  }
  return ifail;
}

int lb_read_buf_ascii(lb_t * lb, int index, const char * buf) {
  const int nbyte = LB_RECORD_LENGTH_ASCII;
  int ifail = 0;
  for (int p = 0; p < lb->nvel; p++) {
    const int poffset = p*(lb->ndist*nbyte + 1);
    for (int n = 0; n < lb->ndist; n++) {
      char tmp[64] = {0};
      memcpy(tmp, buf + poffset + n*nbyte, nbyte);
      if (sscanf(tmp, "%le", lb->f + LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, n, p)) != 1) ifail = 1;
    }
  }
  return ifail;
}

/* src/field.c:931-1000: nf values and a newline per site */
int field_write_buf_ascii(field_t * field, int index, char * buf) {
  const int nbyte = 23;
  int ifail = 0;
  for (int n = 0; n < field->nf; n++) {
    char tmp[64];
    int np = snprintf(tmp, sizeof(tmp), " %22.15e", field->data[addr_rank1(field->nsites, field->nf, index, n)]);
    if (np != nbyte) ifail = 1;
    memcpy(buf + n*nbyte, tmp, nbyte);
  }
  buf[field->nf*nbyte] = '\n';
  return ifail;
}

int field_read_buf_ascii(field_t * field, int index, const char * buf) {
  const int nbyte = 23;
  int ifail = 0;
  for (int n = 0; n < field->nf; n++) {
    char tmp[64] = {0};
    memcpy(tmp, buf + n*nbyte, nbyte);
    if (sscanf(tmp, "%le", field->data + addr_rank1(field->nsites, field->nf, index, n)) != 1) ifail = 1;
  }
  return ifail;
}

/* the aggregator: interior sites in (ic, jc, kc) order (cs_limits, src/lb_data.c:1646-1673, src/field.c:1590-1625) */
static int io_file_transfer(cs_t * cs, const char * filename, size_t szelement, int write,
			    void * obj, int (* wbuf)(void *, int, char *), int (* rbuf)(void *, int, const char *)) {
  const size_t nsite = (size_t) cs->nlocal[X]*cs->nlocal[Y]*cs->nlocal[Z];
  char * buf = (char *) malloc(nsite*szelement);
  FILE * fp = NULL;
  size_t ib = 0, nio;
  if (buf == NULL) return -1;
  if (!write) {
    fp = fopen(filename, "rb");
    if (fp == NULL) { free(buf); return -1; }
    nio = fread(buf, szelement, nsite, fp);
    fclose(fp);
    if (nio != nsite) { free(buf); return -1; }
  }
  for (int ic = 1; ic <= cs->nlocal[X]; ic++)
    for (int jc = 1; jc <= cs->nlocal[Y]; jc++)
      for (int kc = 1; kc <= cs->nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	if (write) wbuf(obj, index, buf + ib*szelement); else rbuf(obj, index, buf + ib*szelement);
	ib++;
      }
  if (write) {
    fp = fopen(filename, "wb");
    if (fp == NULL) { free(buf); return -1; }
    nio = fwrite(buf, szelement, nsite, fp);
    fclose(fp);
    if (nio != nsite) { free(buf); return -1; }
  }
  free(buf);
  return 0;
}

static int lb_wbuf_(void * o, int index, char * buf) { return lb_write_buf((const lb_t *) o, index, buf); }
static int lb_rbuf_(void * o, int index, const char * buf) { return lb_read_buf((lb_t *) o, index, buf); }
static int field_wbuf_(void * o, int index, char * buf) { return field_write_buf((field_t *) o, index, buf); }
static int field_rbuf_(void * o, int index, const char * buf) { return field_read_buf((field_t *) o, index, buf); }
static int lb_wbufa_(void * o, int index, char * buf) { return lb_write_buf_ascii((const lb_t *) o, index, buf); }
static int lb_rbufa_(void * o, int index, const char * buf) { return lb_read_buf_ascii((lb_t *) o, index, buf); }
static int field_wbufa_(void * o, int index, char * buf) { return field_write_buf_ascii((field_t *) o, index, buf); }
static int field_rbufa_(void * o, int index, const char * buf) { return field_read_buf_ascii((field_t *) o, index, buf); }

/* src/lb_data.c:1716-1830.  The device copy is brought to the host first (and pushed back after a read) when the
 * lattice has one; a host-only lb_t (no device use yet) is written / read as it stands. */
int lb_io_write(lb_t * lb, int timestep, io_event_t * event) {
  char filename[BUFSIZ];
  (void) event;
  if (lb->cs->ctx) lb_memcpy(lb, tdpMemcpyDeviceToHost);
  const int ascii = (lb->opts.iodata.output.iorformat == IO_RECORD_ASCII);
  const size_t szel = ascii ? (size_t) lb->nvel*(1 + LB_RECORD_LENGTH_ASCII*lb->ndist) : sizeof(double)*lb->ndist*lb->nvel;
  if (io_metadata_write_default(lb->cs, "dist", ascii ? (int) szel : lb->ndist*lb->nvel, ascii) != 0) pe_fatal(lb->pe, "Could not write dist metadata\n");
  io_file_name("dist", timestep, filename, BUFSIZ);
  if (io_file_transfer(lb->cs, filename, szel, 1, lb, ascii ? lb_wbufa_ : lb_wbuf_, ascii ? lb_rbufa_ : lb_rbuf_) != 0) {
    pe_fatal(lb->pe, "Error: could not write distribution file: %s\n", filename);
  }
  return 0;
}

int lb_io_read(lb_t * lb, int timestep, io_event_t * event) {
  char filename[BUFSIZ];
  (void) event;
  const int ascii = (lb->opts.iodata.input.iorformat == IO_RECORD_ASCII);
  const size_t szel = ascii ? (size_t) lb->nvel*(1 + LB_RECORD_LENGTH_ASCII*lb->ndist) : sizeof(double)*lb->ndist*lb->nvel;
  io_file_name("dist", timestep, filename, BUFSIZ);
  if (io_file_transfer(lb->cs, filename, szel, 0, lb, ascii ? lb_wbufa_ : lb_wbuf_, ascii ? lb_rbufa_ : lb_rbuf_) != 0) {
    pe_fatal(lb->pe, "Error: could not read distribuiion file: %s\n", filename);
  }
  if (lb->cs->ctx) lb_memcpy(lb, tdpMemcpyHostToDevice);
  return 0;
}

/* src/field.c:1633-1740 */
int field_io_write(field_t * field, int timestep, io_event_t * event) {
  char filename[BUFSIZ];
  (void) event;
  if (field->cs->ctx && field->b200_array >= 0) field_memcpy(field, tdpMemcpyDeviceToHost);
  const int ascii = (field->opts.iodata.output.iorformat == IO_RECORD_ASCII);
  const size_t szel = ascii ? (size_t) (1 + 23*field->nf) : sizeof(double)*field->nf;
  if (io_metadata_write_default(field->cs, field->name, ascii ? (int) szel : field->nf, ascii) != 0) pe_fatal(field->pe, "Could not write %s metadata\n", field->name);
  io_file_name(field->name, timestep, filename, BUFSIZ);
  if (io_file_transfer(field->cs, filename, szel, 1, field, ascii ? field_wbufa_ : field_wbuf_, ascii ? field_rbufa_ : field_rbuf_) != 0) {
    pe_fatal(field->pe, "Error: could not write file: %s\n", filename);
  }
  return 0;
}

int field_io_read(field_t * field, int timestep, io_event_t * event) {
  char filename[BUFSIZ];
  (void) event;
  const int ascii = (field->opts.iodata.input.iorformat == IO_RECORD_ASCII);
  const size_t szel = ascii ? (size_t) (1 + 23*field->nf) : sizeof(double)*field->nf;
  io_file_name(field->name, timestep, filename, BUFSIZ);
  if (io_file_transfer(field->cs, filename, szel, 0, field, ascii ? field_wbufa_ : field_wbuf_, ascii ? field_rbufa_ : field_rbuf_) != 0) {
    pe_fatal(field->pe, "Error: could not read file: %s\n", filename);
  }
  if (field->cs->ctx && field->b200_array >= 0) field_memcpy(field, tdpMemcpyHostToDevice);
  return 0;
}

/* ---- collision ------------------------------------------------------------------------------------------------------------------ */

/* src/collision.c:143-162 with the per-call parameter refresh of :1163-1246 and :1906-1958 */
int lb_collide(lb_t * lb, hydro_t * hydro, map_t * map, noise_t * noise, fe_t * fe, visc_t * visc) {
  physics_t * phys = NULL;
  lb200_collide_param_t cp;
  if (hydro == NULL) return 0;
  assert(lb);
  assert(map);
  if (noise != NULL) pe_fatal(lb->pe, "lb_collide: fluctuations are outside this build\n");
  if (visc != NULL) pe_fatal(lb->pe, "lb_collide: viscosity models are outside this build\n");
  memset(&cp, 0, sizeof(cp));
  physics_ref(&phys);
  cp.nrelax = (int) lb->nrelax;
  physics_rho0(phys, &cp.rho0);
  physics_eta_shear(phys, &cp.eta_shear);
  physics_eta_bulk(phys, &cp.eta_bulk);
  physics_fbody(phys, cp.force_global);
  if (lb->ndist == 2) {
    /* lb_collision_binary(lb, hydro, noise, (fe_symm_t *) fe, visc), src/collision.c:157-159 */
    lb200_symm_param_t sp;
    if (fe == NULL) pe_fatal(lb->pe, "lb_collide: ndist = 2 needs the symmetric free energy\n");
    symm_param_from(fe, &sp);
    b200_check(lb->pe, lb200_lb_collision_binary(cs_b200_context(lb->cs), &cp, &sp), "lb_collision_binary");
    return 0;
  }
  b200_check(lb->pe, lb200_lb_collide(cs_b200_context(lb->cs), &cp), "lb_collide");
  return 0;
}

/* src/phi_lb_coupler.c:39-137 */
int phi_lb_to_field(field_t * phi, lb_t * lb) {
  assert(phi); assert(lb);
  b200_check(lb->pe, lb200_phi_lb_to_field(cs_b200_context(lb->cs)), "phi_lb_to_field");
  return 0;
}

int phi_lb_from_field(field_t * phi, lb_t * lb) {
  /* host operation in the reference too: move phi into the non-propagating population */
  assert(phi); assert(lb);
  for (int index = 0; index < lb->nsite; index++) {
    if (!cs_index_is_interior(lb->cs, index)) continue;
    lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, LB_PHI, 0)] = phi->data[addr_rank0(phi->nsites, index)];
    for (int p = 1; p < lb->nvel; p++) lb->f[LB_ADDR(lb->nsite, lb->ndist, lb->nvel, index, LB_PHI, p)] = 0.0;
  }
  return 0;
}
