/*
 * ludwig_host.h -- C host layer with Ludwig's own object and function names for the hot path,
 * implemented on top of the C-ABI in ludwig_b200.h (ludwig_b200/host/ludwig_host.c, linked into
 * libludwig_b200.so).
 *
 * It mirrors the slice of the reference's host interface (ludwig-cf/ludwig v0.23.0) that the
 * per-timestep path and its unit tests use: same function names, same argument order and meaning,
 * same return convention (0 on success), same "host array is authoritative only after
 * *_memcpy(..., tdpMemcpyDeviceToHost)" contract, same fatal-on-device-error behaviour
 * (pe_fatal prints and aborts, reference src/pe.c:226-240).  Each declaration cites the reference
 * declaration it mirrors (paths relative to the reference root).  Host arrays use the reference's
 * -DADDR_SOA addressing (src/memory.h:182-195): lb->f[LB_ADDR(nsite, ndist, nvel, index, n, p)].
 *
 * Mirrored beyond the bare time step: Lees-Edwards planes (lees_edw_t with planes), the symmetric_lb and liquid-crystal
 * sectors, the reference's restart-file writers / readers (lb_io_*, field_io_*).  Not mirrored (outside SURVEY.md
 * section 8): run-time input parsing, statistics, colloids, walls, noise, viscosity models.
 *
 * This header is a stand-alone, self-contained statement of that interface (its structs are cut-down look-alikes).
 * The proof that the library is a drop-in for the reference's OWN structs and callers is integration/: the reference's
 * src/ludwig.c and tests/unit/test_*.c, unchanged, linked against libludwig_b200.so through integration/ludwig_b200_shim.c,
 * which is compiled against the reference's own headers (tests/test_gpu_reference_callers.py).
 */
#ifndef LUDWIG_HOST_H
#define LUDWIG_HOST_H

#include "ludwig_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {X = 0, Y = 1, Z = 2};

/* mpi_s/mpi.h: single process */
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0

/* target/target.h */
typedef enum tdpMemcpyKind_enum {
  tdpMemcpyHostToHost = 0, tdpMemcpyHostToDevice = 1, tdpMemcpyDeviceToHost = 2
} tdpMemcpyKind;

/* ---- pe: src/pe.h:20-43 ---------------------------------------------------------------- */
typedef struct pe_s pe_t;
typedef enum {PE_QUIET = 0, PE_VERBOSE, PE_OPTION_MAX} pe_enum_t;
int pe_create(MPI_Comm parent, pe_enum_t flag, pe_t ** ppe);
int pe_free(pe_t * pe);
int pe_info(pe_t * pe, const char * fmt, ...);
int pe_fatal(pe_t * pe, const char * fmt, ...);
int pe_mpi_rank(pe_t * pe);
int pe_mpi_size(pe_t * pe);

/* ---- cs: src/coords.h:73-109 -------------------------------------------------------------- */
typedef struct cs_s cs_t;
int cs_create(pe_t * pe, cs_t ** pcs);
int cs_free(cs_t * cs);
int cs_init(cs_t * cs);
int cs_ntotal_set(cs_t * cs, const int ntotal[3]);
int cs_nhalo_set(cs_t * cs, int nhalo);
int cs_periodicity_set(cs_t * cs, const int iper[3]);
int cs_ntotal(cs_t * cs, int ntotal[3]);
int cs_nlocal(cs_t * cs, int n[3]);
int cs_nlocal_offset(cs_t * cs, int n[3]);
int cs_nhalo(cs_t * cs, int * nhalo);
int cs_nsites(cs_t * cs, int * nsites);
int cs_nall(cs_t * cs, int nall[3]);
int cs_periodic(cs_t * cs, int period[3]);
int cs_index(cs_t * cs, int ic, int jc, int kc);
int cs_strides(cs_t * cs, int * xs, int * ys, int * zs);
int cs_cartsz(cs_t * cs, int cartsz[3]);
int cs_cart_coords(cs_t * cs, int coords[3]);

/* ---- physics: src/physics.h:20-52 (singleton, as in the reference) ------------------------- */
typedef struct physics_s physics_t;
int physics_create(pe_t * pe, physics_t ** phys);
int physics_free(physics_t * phys);
int physics_ref(physics_t ** phys);
int physics_rho0_set(physics_t * phys, double rho0);
int physics_eta_shear_set(physics_t * phys, double eta);
int physics_eta_bulk_set(physics_t * phys, double zeta);
int physics_fbody_set(physics_t * phys, double f[3]);
int physics_mobility_set(physics_t * phys, double mobility);
int physics_grad_mu_set(physics_t * phys, double gm[3]);
int physics_rho0(physics_t * phys, double * rho0);
int physics_eta_shear(physics_t * phys, double * eta);
int physics_eta_bulk(physics_t * phys, double * eta);
int physics_fbody(physics_t * phys, double f[3]);
int physics_mobility(physics_t * phys, double * mobility);
int physics_grad_mu(physics_t * phys, double gm[3]);
/* time control, src/physics.c:600-670: the Lees-Edwards plane displacement is a function of these */
int physics_control_init_time(physics_t * phys, int nstart, int nstep);
int physics_control_next_step(physics_t * phys);
int physics_control_timestep(physics_t * phys);
int physics_control_time(physics_t * phys, double * t);

/* ---- Lees-Edwards planes, steady shear: src/leesedwards.h:26-83, src/lees_edwards_options.h:23-38 ----------
 * Create the lees_edw_t straight after cs_init (as src/ludwig.c does): the device lattice of the coordinate
 * system is laid out with the buffer planes of the planes it finds. */
typedef struct lees_edw_s lees_edw_t;
typedef enum lees_edw_enum {LE_SHEAR_TYPE_INVALID, LE_SHEAR_TYPE_STEADY, LE_SHEAR_TYPE_OSCILLATORY} lees_edw_enum_t;
typedef struct lees_edw_options_s {int nplanes; int type; int period; int nt0; double uy;} lees_edw_options_t;
int lees_edw_create(pe_t * pe, cs_t * cs, const lees_edw_options_t * opts, lees_edw_t ** le);
int lees_edw_free(lees_edw_t * le);
int lees_edw_nplane_total(lees_edw_t * le);
int lees_edw_nplane_local(lees_edw_t * le);
int lees_edw_plane_uy(lees_edw_t * le, double * uy);
int lees_edw_nxbuffer(lees_edw_t * le, int * nxb);
int lees_edw_nsites(lees_edw_t * le, int * nsites);
int lees_edw_index(lees_edw_t * le, int ic, int jc, int kc);
int lees_edw_plane_location(lees_edw_t * le, int plane);
int lees_edw_ic_to_buff(lees_edw_t * le, int ic, int di);
int lees_edw_ibuff_to_real(lees_edw_t * le, int ib);
int lees_edw_shear_rate(lees_edw_t * le, double * gammadot);
int lees_edw_steady_uy(lees_edw_t * le, int ic, double * uy);

/* ---- memory.h addressing, SOA: src/memory.h:182-195 ---------------------------------------- */
#define addr_rank0(nsites, index) (index)
#define addr_rank1(nsites, na, index, ia) ((nsites)*(ia) + (index))
#define addr_rank2(nsites, na, nb, index, ia, ib) ((nb)*(nsites)*(ia) + (nsites)*(ib) + (index))
#define LB_ADDR(nsites, ndist, nvel, index, n, p) addr_rank2(nsites, ndist, nvel, index, n, p)

/* ---- lb_t: src/lb_data.h:107-284, src/lb_data_options.h:21-47 ------------------------------- */
typedef enum lb_relaxation_enum {LB_RELAXATION_M10, LB_RELAXATION_BGK, LB_RELAXATION_TRT} lb_relaxation_enum_t;
typedef enum lb_halo_enum {LB_HALO_FULL = 1, LB_HALO_REDUCED = 2} lb_halo_enum_t;
typedef enum lb_dist_enum_type {LB_RHO = 0, LB_PHI = 1} lb_dist_enum_t;

/* src/io_options.h:22-52, src/io_info_args.h:33-38 (record format of the files lb_io_* / field_io_* write and read) */
typedef enum io_mode_enum {IO_MODE_INVALID, IO_MODE_MPIIO} io_mode_enum_t;
typedef enum io_record_format_enum {IO_RECORD_INVALID, IO_RECORD_ASCII, IO_RECORD_BINARY} io_record_format_enum_t;
typedef enum io_metadata_version_enum {IO_METADATA_INVALID, IO_METADATA_SINGLE_V1, IO_METADATA_MULTI_V1, IO_METADATA_V2} io_metadata_version_enum_t;
typedef struct io_options_s {
  io_mode_enum_t mode;
  io_record_format_enum_t iorformat;      /* IO_RECORD_ASCII: " %22.15e" per datum; IO_RECORD_BINARY (default): doubles */
  io_metadata_version_enum_t metadata_version;
  int report;
  int asynchronous;
  int compression_levl;
  int iogrid[3];                          /* {1, 1, 1}: one file (the only decomposition of a one-process host layer) */
} io_options_t;
typedef struct io_info_args_s {io_options_t input; io_options_t output; int grid[3]; int iofreq;} io_info_args_t;
io_options_t io_options_default(void);
io_options_t io_options_with_format(io_mode_enum_t mode, io_record_format_enum_t iorf);
io_info_args_t io_info_args_default(void);

typedef struct lb_data_options_s {
  int ndim;
  int nvel;
  int ndist;
  lb_relaxation_enum_t nrelax;
  lb_halo_enum_t halo;
  int reportimbalance;
  int usefirsttouch;
  io_info_args_t iodata;                  /* src/lb_data_options.h:41 */
} lb_data_options_t;

typedef struct lb_model_s {
  int ndim;
  int nvel;
  signed char (* cv)[3];
  double * wv;
  double cs2;
} lb_model_t;

typedef struct lb_data_s lb_t;
struct lb_data_s {
  int ndim;
  int nvel;
  int ndist;
  int nsite;
  pe_t * pe;
  cs_t * cs;
  lb_model_t model;
  double * f;                    /* host distributions, LB_ADDR addressing */
  double * fprime;               /* not used on the host here */
  lb_relaxation_enum_t nrelax;
  lb_halo_enum_t haloscheme;
  lb_data_options_t opts;
  lb_t * target;
};

lb_data_options_t lb_data_options_default(void);
lb_data_options_t lb_data_options_ndim_nvel_ndist(int ndim, int nvel, int ndist);
int lb_data_create(pe_t * pe, cs_t * cs, const lb_data_options_t * opts, lb_t ** lb);
int lb_free(lb_t * lb);
int lb_memcpy(lb_t * lb, tdpMemcpyKind flag);
int lb_halo(lb_t * lb);
int lb_init_rest_f(lb_t * lb, double rho0);
int lb_f(lb_t * lb, int index, int p, int n, double * f);
int lb_f_set(lb_t * lb, int index, int p, int n, double f);
int lb_0th_moment(lb_t * lb, int index, lb_dist_enum_t nd, double * rho);
int lb_1st_moment(lb_t * lb, int index, lb_dist_enum_t nd, double g[3]);
int lb_1st_moment_equilib_set(lb_t * lb, int index, double rho, double u[3]);
/* src/model_le.h: plane-crossing populations after the collision; steady shear profile initial condition */
int lb_data_apply_le_boundary_conditions(lb_t * lb, lees_edw_t * le);
int lb_le_init_shear_profile(lb_t * lb, lees_edw_t * le);
int lb_collision_relaxation_set(lb_t * lb, lb_relaxation_enum_t nrelax);     /* src/collision.h:30 */
int lb_collide_param_commit(lb_t * lb);                                      /* src/lb_data.h:166 */

/* ---- field_t: src/field.h:68-130, src/field_options.h ---------------------------------------- */
typedef struct field_options_s {int ndata; int nhcomm; int haloscheme; int haloverbose; int usefirsttouch;
				io_info_args_t iodata;   /* src/field_options.h:36 */
} field_options_t;
field_options_t field_options_default(void);
field_options_t field_options_ndata_nhalo(int ndata, int nhalo);

typedef struct field_s field_t;
struct field_s {
  int nf;
  int nhcomm;
  int nsites;
  double * data;                 /* host data, addr_rank1(nsites, nf, index, n) */
  char * name;
  pe_t * pe;
  cs_t * cs;
  lees_edw_t * le;
  field_options_t opts;
  int b200_array;                /* which device array backs this field (LB200_PHI, LB200_U, ...) */
  field_t * target;
};

int field_create(pe_t * pe, cs_t * cs, lees_edw_t * le, const char * name, const field_options_t * opts, field_t ** pobj);
int field_free(field_t * obj);
int field_memcpy(field_t * obj, tdpMemcpyKind flag);
int field_halo(field_t * obj);
typedef enum field_halo_enum {FIELD_HALO_HOST, FIELD_HALO_TARGET, FIELD_HALO_OPENMP} field_halo_enum_t;   /* src/field_options.h:22-24 */
int field_halo_swap(field_t * obj, field_halo_enum_t flag);
int field_leesedwards(field_t * obj);                                        /* src/field.h:106 */
int field_nf(field_t * obj, int * nop);
int field_scalar(field_t * obj, int index, double * phi);
int field_scalar_set(field_t * obj, int index, double phi);
int field_vector(field_t * obj, int index, double p[3]);
int field_vector_set(field_t * obj, int index, const double p[3]);

/* ---- field_grad_t: src/field_grad.h:24-56, src/gradient_3d_27pt_fluid.h:20 ------------------- */
typedef struct field_grad_s field_grad_t;
typedef int (* grad_ft)(field_grad_t * fgrad);
struct field_grad_s {
  pe_t * pe;
  field_t * field;
  int nf;
  int level;
  int nsite;
  double * grad;                 /* addr_rank2(nsite, nf, 3, index, n, ia) */
  double * delsq;                /* addr_rank1(nsite, nf, index, n) */
  double * grad_delsq;           /* level >= 4: addr_rank2(nsite, nf, 3, index, n, ia) */
  double * delsq_delsq;          /* level >= 4 */
  int d4_on_device;
  grad_ft d2;
  grad_ft d4;
  field_grad_t * target;
};

int field_grad_create(pe_t * pe, field_t * f, int level, field_grad_t ** pobj);
void field_grad_free(field_grad_t * obj);
int field_grad_set(field_grad_t * obj, grad_ft d2, grad_ft d4);
int field_grad_compute(field_grad_t * obj);
int field_grad_memcpy(field_grad_t * obj, tdpMemcpyKind flag);
int field_grad_scalar_grad(field_grad_t * obj, int index, double grad[3]);
int field_grad_scalar_delsq(field_grad_t * obj, int index, double * delsq);
int grad_3d_27pt_fluid_d2(field_grad_t * fg);
int grad_3d_27pt_fluid_d4(field_grad_t * fg);

/* ---- hydro_t: src/hydro.h:32-62, src/hydro_options.h ------------------------------------------ */
typedef struct hydro_options_s {int nhcomm; field_options_t rho; field_options_t u; field_options_t force; field_options_t eta;} hydro_options_t;
hydro_options_t hydro_options_default(void);
hydro_options_t hydro_options_nhalo(int nhalo);

typedef struct hydro_s hydro_t;
struct hydro_s {
  int nsite;
  int nhcomm;
  pe_t * pe;
  cs_t * cs;
  lees_edw_t * le;
  field_t * rho;
  field_t * u;
  field_t * force;
  field_t * eta;
  hydro_t * target;
};

int hydro_create(pe_t * pe, cs_t * cs, lees_edw_t * le, const hydro_options_t * opts, hydro_t ** pobj);
int hydro_free(hydro_t * obj);
int hydro_memcpy(hydro_t * obj, tdpMemcpyKind flag);
int hydro_u_halo(hydro_t * obj);
int hydro_lees_edwards(hydro_t * obj);                                       /* src/hydro.h:57 */
int hydro_f_zero(hydro_t * obj, const double fzero[3]);
int hydro_u_zero(hydro_t * obj, const double uzero[3]);
int hydro_u(hydro_t * obj, int index, double u[3]);           /* src/hydro_impl.h */
int hydro_u_set(hydro_t * obj, int index, const double u[3]);
int hydro_f_local(hydro_t * obj, int index, double force[3]);
int hydro_f_local_set(hydro_t * obj, int index, const double force[3]);
int hydro_rho(hydro_t * obj, int index, double * rho);

/* ---- map_t: src/map.h:26-63 ------------------------------------------------------------------ */
enum map_status {MAP_FLUID = 0, MAP_BOUNDARY, MAP_COLLOID, MAP_STATUS_MAX};
typedef struct map_options_s {int ndata; int is_porous_media;} map_options_t;
map_options_t map_options_default(void);
typedef struct map_s map_t;
struct map_s {
  int nsite;
  pe_t * pe;
  cs_t * cs;
  char * status;
  map_t * target;
};
int map_create(pe_t * pe, cs_t * cs, const map_options_t * options, map_t ** map);
int map_free(map_t ** map);
int map_memcpy(map_t * map, tdpMemcpyKind flag);
int map_status(map_t * map, int index, int * status);
int map_status_set(map_t * map, int index, int status);

/* ---- free energy (symmetric): src/free_energy.h:36-80, src/symmetric.h:28-66 ------------------- */
typedef struct fe_s fe_t;
struct fe_s {int id; int use_stress_relaxation;};
typedef struct fe_symm_param_s {double a; double b; double kappa; double c; double h;} fe_symm_param_t;
typedef struct fe_symm_s fe_symm_t;
struct fe_symm_s {
  fe_t super;
  pe_t * pe;
  cs_t * cs;
  fe_symm_param_t * param;
  field_t * phi;
  field_grad_t * dphi;
  fe_symm_t * target;
};
int fe_symm_create(pe_t * pe, cs_t * cs, field_t * f, field_grad_t * grd, fe_symm_t ** p);
int fe_symm_free(fe_symm_t * fe);
int fe_symm_target(fe_symm_t * fe, fe_t ** target);
int fe_symm_param_set(fe_symm_t * fe, fe_symm_param_t values);
int fe_symm_param(fe_symm_t * fe, fe_symm_param_t * values);
int fe_symm_fed(fe_symm_t * fe, int index, double * fed);     /* host arrays */
int fe_symm_mu(fe_symm_t * fe, int index, double * mu);
int fe_symm_str(fe_symm_t * fe, int index, double s[3][3]);   /* src/symmetric.c:333-362, host arrays */

/* ---- force from the phi sector: src/phi_force_stress.h:26-39, src/phi_force.h:24, fe_force_method.h */
typedef enum {
  FE_FORCE_METHOD_INVALID, FE_FORCE_METHOD_NO_FORCE, FE_FORCE_METHOD_STRESS_DIVERGENCE, FE_FORCE_METHOD_PHI_GRADMU,
  FE_FORCE_METHOD_PHI_GRADMU_CORRECTION, FE_FORCE_METHOD_RELAXATION_SYMM, FE_FORCE_METHOD_RELAXATION_ANTI,
  FE_FORCE_METHOD_MAX
} fe_force_method_enum_t;
typedef struct pth_s pth_t;
struct pth_s {pe_t * pe; cs_t * cs; int method; int nsites; pth_t * target;};
typedef struct wall_s wall_t;          /* never dereferenced here: pass NULL (no walls in scope) */
int pth_create(pe_t * pe, cs_t * cs, int method, pth_t ** pth);
int pth_free(pth_t * pth);
int pth_stress_compute(pth_t * pth, fe_t * fe);
int pth_force_fluid_driver(pth_t * pth, hydro_t * hydro);
int phi_force_calculation(pe_t * pe, cs_t * cs, lees_edw_t * le, wall_t * wall, pth_t * pth, fe_t * fe,
			  map_t * map, field_t * phi, hydro_t * hydro);

/* ---- Cahn-Hilliard: src/phi_cahn_hilliard.h:38-56, src/advection.h:40-41 ---------------------- */
typedef struct phi_ch_info_s {int conserve; int noise;} phi_ch_info_t;
typedef struct phi_ch_s phi_ch_t;
struct phi_ch_s {pe_t * pe; cs_t * cs; lees_edw_t * le; phi_ch_info_t info;};
typedef struct noise_s noise_t;        /* pass NULL */
typedef struct visc_s visc_t;          /* pass NULL */
int phi_ch_create(pe_t * pe, cs_t * cs, lees_edw_t * le, phi_ch_info_t * info, phi_ch_t ** pch);
int phi_ch_free(phi_ch_t * pch);
int phi_cahn_hilliard(phi_ch_t * pch, fe_t * fe, field_t * phi, hydro_t * hydro, map_t * map, noise_t * noise);
int advection_order_set(const int order);

/* ---- symmetric_lb coupling: src/phi_lb_coupler.h (lb_collide dispatches on lb->ndist, src/collision.c:157-159) */
int phi_lb_to_field(field_t * phi, lb_t * lb);
int phi_lb_from_field(field_t * phi, lb_t * lb);
int advection_order(int * order);

/* ---- liquid crystal: src/blue_phase.h:32-75, src/blue_phase_beris_edwards.h:27-50, src/blue_phase_init.h,
 * src/gradient_3d_7pt_fluid.h.  The tensor order parameter is the field_t "q" created with ndata = NQAB. -------- */
#define NQAB 5
enum {XX = 0, XY = 1, XZ = 2, YY = 3, YZ = 4};
typedef struct fe_lc_param_s {
  double a0, q0, gamma, kappa0, kappa1;
  double xi, zeta0, zeta1, zeta2;
  double redshift, rredshift;
  double epsilon;
  double amplitude0;
  double e0[3];
  double coswt;
  int is_redshift_updated;
  int is_active;
} fe_lc_param_t;
typedef struct fe_lc_s fe_lc_t;
struct fe_lc_s {
  fe_t super;
  pe_t * pe;
  cs_t * cs;
  fe_lc_param_t * param;
  field_t * q;
  field_grad_t * dq;
  fe_lc_t * target;
};
int fe_lc_create(pe_t * pe, cs_t * cs, lees_edw_t * le, field_t * q, field_grad_t * dq, fe_lc_t ** fe);
int fe_lc_free(fe_lc_t * fe);
int fe_lc_param_set(fe_lc_t * fe, const fe_lc_param_t * values);      /* epsilon *= 1/12pi, as the reference */
int fe_lc_param(fe_lc_t * fe, fe_lc_param_t * vals);
int fe_lc_q_uniaxial(fe_lc_param_t * param, const double n[3], double q[3][3]);
int field_tensor(field_t * obj, int index, double q[3][3]);
int field_tensor_set(field_t * obj, int index, double q[3][3]);
int blue_phase_twist_init(cs_t * cs, fe_lc_param_t * param, field_t * fq, int helical_axis);
int grad_3d_7pt_fluid_d2(field_grad_t * fg);

typedef struct beris_edw_param_s {double xi; double gamma; int noise; double var;} beris_edw_param_t;
typedef struct beris_edw_s beris_edw_t;
typedef struct colloids_info_s colloids_info_t;   /* never dereferenced here: pass NULL (no colloids in scope) */
int beris_edw_create(pe_t * pe, cs_t * cs, lees_edw_t * le, beris_edw_t ** pobj);
int beris_edw_free(beris_edw_t * be);
int beris_edw_param_set(beris_edw_t * be, beris_edw_param_t * values);
int beris_edw_update(beris_edw_t * be, fe_t * fe, field_t * fq, field_grad_t * fq_grad, hydro_t * hydro,
		     colloids_info_t * cinfo, map_t * map, noise_t * noise);

/* ---- on-disk formats: src/lb_data.c:1533-1575, 1716-1830, src/field.c:896-930, 1633-1740,
 * src/io_subfile.c:186-204, src/io_metadata.c.  One file per quantity and time step, "<stub>-%9.9d.001-001": the
 * interior sites in (ic, jc, kc) order with kc fastest, one binary record per site (distributions: ndist*nvel
 * doubles in (n, p) order; fields: nf doubles), plus "<stub>-metadata.001-001" (JSON, written once).  Files are
 * byte-identical with the reference's default (mpiio, binary, single file) output, so either code restarts from
 * the other's.  options.iodata.{input,output}.iorformat = IO_RECORD_ASCII selects the reference's text records
 * (src/lb_data.c:1579-1640, src/field.c:931-1000: " %22.15e" per datum; distributions: one line of ndist values per
 * velocity; fields: one line of nf values per site), again byte-identical. */
typedef struct io_event_s {int unused;} io_event_t;
#define LB_RECORD_LENGTH_ASCII 23        /* src/lb_data.h:45 */
int lb_write_buf(const lb_t * lb, int index, char * buf);
int lb_read_buf(lb_t * lb, int index, const char * buf);
int lb_write_buf_ascii(const lb_t * lb, int index, char * buf);
int lb_read_buf_ascii(lb_t * lb, int index, const char * buf);
int field_write_buf_ascii(field_t * field, int index, char * buf);
int field_read_buf_ascii(field_t * field, int index, const char * buf);
int lb_io_write(lb_t * lb, int timestep, io_event_t * event);
int lb_io_read(lb_t * lb, int timestep, io_event_t * event);
int field_write_buf(field_t * field, int index, char * buf);
int field_read_buf(field_t * field, int index, const char * buf);
int field_io_write(field_t * field, int timestep, io_event_t * event);
int field_io_read(field_t * field, int timestep, io_event_t * event);

/* ---- collision and propagation: src/collision.h:27-28, src/propagation.h:21 ------------------- */
int lb_collide(lb_t * lb, hydro_t * hydro, map_t * map, noise_t * noise, fe_t * fe, visc_t * visc);
int lb_propagation(lb_t * lb);

/* the device context behind a coordinate system (created on first device use) */
lb200_t * cs_b200_context(cs_t * cs);

#ifdef __cplusplus
}
#endif
#endif
