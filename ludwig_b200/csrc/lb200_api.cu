// lb200_api.cu -- context object and C-ABI of libludwig_b200.so (see include/ludwig_b200.h).
//
// Host-side logic only: owns the device arrays, tracks lazily-applied operations (pending
// propagation, logically-zero force / velocity), derives the per-call collision constants the way
// the reference's host code does, and launches the kernels in lb200_kernels.cu.

#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <new>
#include <vector>

#include <cuda.h>             // types of the driver's green-context API (entry points through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#ifndef LB200_NO_NCCL
#include <nccl.h>
#endif

#include "ludwig_b200.h"
#include "lb200_kernels.h"

// ------------------------------------------------------------------------------------------------

static thread_local char g_err[512] = "";

static int fail(int code, const char * fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
  return fail(LB200_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); } while (0)

enum ZeroState {ARRAY_CLEAN = 0, ZERO_PENDING = 1, INTERIOR_ONLY = 2};
enum {SRC_NONE = 0, SRC_EVENT = 1, SRC_FLAG = 2};
enum {FLAG_PS_LO = 0, FLAG_PS_HI = 1, FLAG_COL_LO = 2, FLAG_COL_HI = 3, FLAG_COUNT = 16};

struct PeerLink {            // one neighbour's arrays, in ITS allocation order
  double * f[2];
  double * phi[2];
  double * u[2];
  unsigned int * flags;
};

enum { LB200_PIPE_MAXS = 16 };

struct lb200_s {
  lb200_options_t opt;
  Lb200Geom g;
  int device;
  cudaStream_t stream;
  const Lb200Kernels * k;

  int nvel, ndist;
  Lb200ModelDev model_h;
  Lb200ModelDev * model_d;
  int unrolled19;            // use the unrolled D3Q19 projections (reference -D_D3Q19_ build)

  double * f;                // current distributions
  double * fprime;           // propagation target
  double * phi;
  double * phinew;
  double * u;
  double * rho;
  double * force;
  double * grad;
  double * delsq;
  double * grad_delsq;       // field_grad level 4 (grad_3d_27pt_fluid_d4), allocated on first use
  double * delsq_delsq;
  double * csum;             // phi_ch_t.csum: per-site compensation of cahn_hilliard_options_conserve 1, allocated on first use
  double * str;              // pth->str (9 x nsites), allocated on first use of lb200_pth_stress_compute
  double * q;                // liquid crystal: Q (5 x nsites) and its update target
  double * qnew;
  double * qgrad;            // 15 x nsites, allocated on first use of lb200_q_grad_compute
  double * qdelsq;           // 5 x nsites
  char * status;             // device copy of map->status, nullptr if all fluid
  int map_all_fluid;

  // x-plane staging for slab decomposition
  double * xlo;              // receive staging (planes of the low / high neighbour)
  double * xhi;
  double * slo;              // send staging (my low / high boundary planes, all components contiguous)
  double * shi;
  size_t stage_doubles;

  // second stream: halo exchanges of lb200_step run here, overlapped with the compute kernels
  cudaStream_t comm;
  cudaEvent_t ev_main;       // last producer on the main stream
  cudaEvent_t ev_phi, ev_u, ev_f;   // halo of phi / u / f complete (comm stream)
  int phi_halo_valid;        // lb200_step already exchanged the halo of the current phi / u
  int u_halo_valid;

  int prop_pending;          // lb_propagation requested, not yet applied (fused into next collide)
  double * psum;             // conserve 2: scratch of the phi sums (partials | result[2] | all ranks | total[2])
  double phi_init_sum;       // phi->field_init_sum (cahn_hilliard_stats_time0): the sum the global correction restores
  int phi_init_sum_set;
  double * halo_snap;        // snapshot of an array for halo swaps on lattices thinner than the halo
  size_t halo_snap_size;
  int fused_ready;           // the last thing that touched the lattice was a one-kernel step: it left the y / z (and peer x) halos of
                             // f (the populations the next pull reads), phi (depth nhalo) and u (depth 1) filled for the next one
  int f_halo_stale;          // halo-free lb200_step: lb_halo(f) was folded into the kernels' reads and has
                             // not been applied to the halo sites of f (done on demand with the propagation)
  int wrap_x_valid;          // halo-free lb200_step on slabs: x-planes of phi, u_x, f already exchanged
  int force_state;           // ZeroState
  int u_state;

  long long launches;
  void * nccl;               // ncclComm_t
  int knob_wrap;             // lb200_set_knob
  int knob_phi_sector;
  int knob_peer;
  int knob_grad7;            // fd_gradient_calculation 3d_7pt_fluid for the scalar order parameter (0: 3d_27pt_fluid)
  int knob_qgrad2d;          // fd_gradient_calculation 2d_5pt_fluid for the Q tensor (0: 3d_7pt_fluid)
  int knob_lazy_diag;        // rho / grad / delsq stored by the last step of an lb200_step call only (LB200_LAZY_DIAG, default 1)
  int knob_f32;              // FP32 storage of the distributions inside lb200_step (0: off)
  int knob_fused;            // one kernel per binary-fluid step where it applies (LB200_FUSED, default 1)
  int knob_fused_le;         // ... also with Lees-Edwards planes (LB200_FUSED_LE: 1 = default, the patch chain after the sweep;
                             // 2: next to it on its own stream (measured: no faster, the chain's grids take whole SMs from
                             // the sweep); 0: two kernels + patches)
  float * f32[2];            // float(f_p - w_p), allocated on first use
  int knob_pipe;             // slab pipeline of lb200_step: number of x-slabs (0: off)
  int knob_pipe_sms;         // SMs of the phi-sector partition (the collision gets the rest)

  // slab pipeline (step_pipe): the two kernels of a binary-fluid step on disjoint SM partitions (green contexts)
  int pipe_state;            // 0: not set up, 1: green contexts, 2: plain priority streams, -1: unavailable
  cudaStream_t pipe_a, pipe_b;            // phi sector / collision
  cudaEvent_t pipe_ev_ps[LB200_PIPE_MAXS], pipe_ev_c[LB200_PIPE_MAXS], pipe_ev_join[2];
  void * pipe_gctx[2];       // CUgreenCtx
  int pipe_sm[2];            // SMs provisioned for the phi sector / the collision

  // peer-store exchange of lb200_step on x-slabs: the neighbours' arrays mapped into this process (cudaIpc)
  int peer_state;            // 0: not set up yet, 1: active, -1: unavailable (NCCL exchange instead)
  double * f_alloc[2], * phi_alloc[2], * u_alloc[2];   // my buffers in allocation order (the neighbours swap in lockstep)
  double * u2;               // second velocity buffer: the collision writes the one the neighbours' phi sector is not reading
  unsigned int * flags;      // [0] phi sector done on the low / [1] high neighbour, [2] collision low / [3] high
  int * spin_err;
  int spin_used;             // the spinning wait kernel has been launched in this context
  int joint_signal;          // the last signal set the phi flags and the f / u flags together (signal2)
  PeerLink lo, hi;
  void * mapped[14];         // everything opened with cudaIpcOpenMemHandle
  int nmapped;
  unsigned int n_ps, n_col;  // producer kernels run with peer stores (identical on every rank)
  int phi_src, u_src, f_src; // how the boundary data the next consumer needs arrives: SRC_*
  void * wait_value32;       // cuStreamWaitValue32, if the driver has it

  // Lees-Edwards planes (options.le_nplanes > 0)
  Lb200LeDev le;             // planes of this slab
  int nxbuf;                 // buffer x-planes appended to every hydro / field array
  int nsites_lb;             // nall[0]*nall[1]*nall[2]: host stride of LB200_F / LB200_MAP
  int t_start, t_current;    // physics_control_* (src/physics.c:600-647)
  int * le_trip;             // device: (x-1, x, x+1) plane triples of the gradient patch
  int le_ntrip;
  cudaStream_t le_stream;    // the patch chain of the one-kernel step runs here, next to the sweep (high priority)
  cudaEvent_t ev_le_a, ev_le_b;
  int * le_trip_fused;       // the same + the four real planes beyond (one-kernel step: the sweep keeps grad / delsq in registers)
  int le_ntrip_fused;
  int * le_xlist;            // device: the x-planes within nhalo of a plane (force / Cahn-Hilliard patch)
  int le_nxlist;
  double * le_term;          // 3*nplane*Ny*Nz: summands of the per-plane force flux correction
  double * le_fcor;          // 3*nplane
  double * le_chx;           // 2*nplane*Ny*Nz: raw Cahn-Hilliard x-face fluxes either side of the planes
  double * le_sbuf;          // 2*nplane*ndist*nprop*Ny*Nz: re-projected plane-crossing populations

  // optional per-kernel-class timing with CUDA events on the launching stream
  int profile;
  std::vector<cudaEvent_t> * ev[LB200_KCLASS_MAX];   // pairs (start, stop)
};

// bracket one launch (or a short sequence) with events when profiling is on
struct ProfScope {
  lb200_t * c; int cls; cudaEvent_t stop; cudaStream_t st;
  ProfScope(lb200_t * c_, int cls_, cudaStream_t st_ = nullptr) : c(c_), cls(cls_), stop(nullptr), st(st_ ? st_ : c_->stream) {
    if (!c->profile) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    c->ev[cls]->push_back(a); c->ev[cls]->push_back(b);
    stop = b;
  }
  ~ProfScope() { if (stop) cudaEventRecord(stop, st); }
};

const char * lb200_last_error(void) { return g_err; }
int lb200_version(void) { return 100; }

// ---- model tables: same double-precision expressions as the reference host code --------------
// (src/lb_d3q19.c:108-150, src/lb_d3q15.c:150-178, src/lb_d3q27.c:155-195, normalisers
//  src/lb_d3q19.c:72-78, inverse src/lb_data.c:640-646)

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

static int model_init(int nvel, Lb200ModelDev * md) {
  const double cs2 = (1.0/3.0);
  double wv[27], na[27];
  memset(md, 0, sizeof(*md));
  md->nvel = nvel;

  if (nvel == 19) {
    for (int p = 0; p < 19; p++) {
      int c1 = 0;
      for (int a = 0; a < 3; a++) { md->cv[p][a] = cv19[p][a]; c1 += abs(cv19[p][a]); }
      wv[p] = (c1 == 0) ? 12.0/36.0 : (c1 == 1) ? 2.0/36.0 : 1.0/36.0;
    }
  }
  else if (nvel == 15) {
    for (int p = 0; p < 15; p++) {
      int c1 = 0;
      for (int a = 0; a < 3; a++) { md->cv[p][a] = cv15[p][a]; c1 += abs(cv15[p][a]); }
      wv[p] = (c1 == 0) ? 16.0/72.0 : (c1 == 1) ? 8.0/72.0 : 1.0/72.0;
    }
  }
  else if (nvel == 27) {
    int p = 1;
    wv[0] = 64.0/216.0;
    for (int i = -1; i <= 1; i++)
      for (int j = -1; j <= 1; j++)
	for (int k = -1; k <= 1; k++) {
	  int c1 = abs(i) + abs(j) + abs(k);
	  if (c1 == 0) continue;
	  md->cv[p][0] = i; md->cv[p][1] = j; md->cv[p][2] = k;
	  wv[p] = (c1 == 1) ? 16.0/216.0 : (c1 == 2) ? 4.0/216.0 : 1.0/216.0;
	  p++;
	}
  }
  else {
    return -1;
  }

  for (int p = 0; p < nvel; p++) {
    double rho = 1.0;
    double cx = rho*md->cv[p][0];
    double cy = rho*md->cv[p][1];
    double cz = rho*md->cv[p][2];
    double (*ma)[27] = md->ma;
    ma[0][p] = rho;  ma[1][p] = cx;  ma[2][p] = cy;  ma[3][p] = cz;
    ma[4][p] = cx*cx - cs2;  ma[5][p] = cx*cy;  ma[6][p] = cx*cz;
    ma[7][p] = cy*cy - cs2;  ma[8][p] = cy*cz;  ma[9][p] = cz*cz - cs2;
    if (nvel == 19) {
      double c2   = cx*cx + cy*cy + cz*cz;
      double chi1 = (2.0*c2 - 3.0)*(3.0*cz*cz - c2);
      double chi2 = (2.0*c2 - 3.0)*(cy*cy - cx*cx);
      double chi3 = 3.0*c2*c2 - 6.0*c2 + 1;
      ma[10][p] = chi1; ma[11][p] = chi1*cx; ma[12][p] = chi1*cy; ma[13][p] = chi1*cz;
      ma[14][p] = chi2; ma[15][p] = chi2*cx; ma[16][p] = chi2*cy; ma[17][p] = chi2*cz;
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
      ma[10][p] = 3.0*(cx*cx - cs2)*cy;  ma[11][p] = 3.0*(cx*cx - cs2)*cz;
      ma[12][p] = 3.0*(cy*cy - cs2)*cz;  ma[13][p] = 3.0*(cy*cy - cs2)*cx;
      ma[14][p] = 3.0*(cz*cz - cs2)*cx;  ma[15][p] = 3.0*(cz*cz - cs2)*cy;
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
  for (int p = 0; p < nvel; p++) md->wv[p] = wv[p];
  for (int m = 0; m < nvel; m++) {
    double wip = 0.0;
    for (int p = 0; p < nvel; p++) wip += wv[p]*md->ma[m][p]*md->ma[m][p];
    na[m] = 1.0/wip;
  }
  for (int p = 0; p < nvel; p++)
    for (int m = 0; m < nvel; m++) {
      double maba = md->ma[m][p];
      md->mi[p][m] = wv[p]*na[m]*maba;
    }
  return 0;
}

// ---- collision constants per call: src/collision.c:1269-1520 (rates), 1906-1958 ---------------

static int collide_dev(const lb200_t * c, const lb200_collide_param_t * cp, Lb200CollideDev * d) {
  const double cs2 = (1.0/3.0);
  memset(d, 0, sizeof(*d));
  for (int a = 0; a < 3; a++) d->fg[a] = cp->force_global[a];
  d->rtau = 1.0/(0.5 + cp->eta_shear/(cp->rho0*cs2));
  if (cp->nrelax == LB200_RELAXATION_BGK) d->rtau_bulk = 1.0/(0.5 + cp->eta_shear/(cp->rho0*cs2));
  else d->rtau_bulk = 1.0/(0.5 + cp->eta_bulk/(cp->rho0*cs2));
  d->tmr = 2.0 - d->rtau;
  for (int m = 0; m < 27; m++) d->rtau_ghost[m] = 1.0;
  d->ghost = 0;
  if (cp->nrelax == LB200_RELAXATION_BGK) {
    for (int m = 10; m < c->nvel; m++) d->rtau_ghost[m] = d->rtau;
    d->ghost = 1;
  }
  else if (cp->nrelax == LB200_RELAXATION_TRT) {
    double tau = cp->eta_shear/(cp->rho0*cs2);
    double rg = 0.5 + 2.0*tau/(tau + 3.0/8.0);
    if (rg > 2.0) rg = 2.0;
    if (c->nvel == 15) {
      d->rtau_ghost[10] = d->rtau; d->rtau_ghost[14] = d->rtau;
      d->rtau_ghost[11] = rg; d->rtau_ghost[12] = rg; d->rtau_ghost[13] = rg;
    }
    else if (c->nvel == 19) {
      d->rtau_ghost[10] = d->rtau; d->rtau_ghost[14] = d->rtau; d->rtau_ghost[18] = d->rtau;
      d->rtau_ghost[11] = rg; d->rtau_ghost[12] = rg; d->rtau_ghost[13] = rg;
      d->rtau_ghost[15] = rg; d->rtau_ghost[16] = rg; d->rtau_ghost[17] = rg;
    }
    else {
      return fail(LB200_EINVAL, "TRT is defined for D3Q15 and D3Q19 only (reference src/collision.c:1219-1242)");
    }
    d->ghost = 1;
  }
  else if (cp->nrelax != LB200_RELAXATION_M10) {
    return fail(LB200_EINVAL, "unknown relaxation scheme %d", cp->nrelax);
  }
  return 0;
}

static void symm_dev(const lb200_t * c, const lb200_symm_param_t * sp, Lb200SymmDev * d) {
  d->a = sp->a; d->b = sp->b; d->kappa = sp->kappa; d->mobility = sp->mobility;
  for (int a = 0; a < 3; a++) d->gm[a] = sp->gradmu[a];
  d->order = sp->adv_order;
  d->wz = (c->g.nl[2] == 1) ? 0.0 : 1.0;
  d->rtau2 = 2.0/(1.0 + 2.0*sp->mobility);          // src/collision.c:1949-1950
  d->csum = (sp->conserve == 1) ? c->csum : nullptr;   // allocated by conserve_prepare
  d->force_method = sp->force_method;
}

// ------------------------------------------------------------------------------------------------

static int alloc_d(double ** p, size_t n) {
  CUDA_TRY(cudaMalloc((void **) p, n*sizeof(double)));
  CUDA_TRY(cudaMemset(*p, 0, n*sizeof(double)));
  // cudaMemset on device memory returns before the fill has run, on the legacy stream, which is not ordered with
  // the contexts' non-blocking streams: without this wait the fill can land after the first upload / kernel
  CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
  return 0;
}

// cahn_hilliard_options_conserve: validate, and create the compensation field on first use (pch->csum is created,
// zero, with the phi_ch_t: src/phi_cahn_hilliard.c:146-153)
static int conserve_prepare(lb200_t * c, const lb200_symm_param_t * sp) {
  if (sp->conserve == 0) return 0;
  if (sp->conserve == 2) {
    // PHI_CONSERVE_GLOBAL_SUBTRACT: needs the initial sum (the reference takes it from its statistics code at time 0)
    if (!c->phi_init_sum_set) return fail(LB200_ESTATE, "cahn_hilliard_options_conserve 2: no initial sum (lb200_phi_conserve_sum / lb200_phi_init_sum_set first)");
    return 0;
  }
  if (sp->conserve != 1) return fail(LB200_EINVAL, "cahn_hilliard_options_conserve %d: 0, 1 (compensated sum) and 2 (global subtraction) are built", sp->conserve);
  if (c->csum == nullptr && alloc_d(&c->csum, (size_t) c->g.nsites) != 0) return LB200_ECUDA;
  return 0;
}

// ---- Lees-Edwards planes: host-side geometry (reference src/leesedwards.c) ---------------------------------

static int le_nplane_local(const lb200_options_t * o) {
  return (o->le_nplanes > 0 && o->cart_size > 0) ? o->le_nplanes/o->cart_size : 0;
}

// lees_edw_plane_location, src/leesedwards.c:615-634: dx_sep = ntotal[X]/nplanetotal, dx_min = dx_sep/2 (:256-257)
int lb200_le_plane_location(const lb200_options_t * o, int np) {
  if (o == nullptr || o->le_nplanes <= 0) return fail(LB200_EINVAL, "no Lees-Edwards planes");
  const int npl = le_nplane_local(o);
  if (np < 0 || np >= npl) return fail(LB200_EINVAL, "plane %d of %d", np, npl);
  const double dx_sep = 1.0*(o->nlocal[0]*o->cart_size)/o->le_nplanes;
  const double dx_min = 0.5*dx_sep;
  const int offset = o->cart_rank*o->nlocal[0];
  const int nplane_offset = o->cart_rank*npl;
  int ix = dx_min + (np + nplane_offset)*dx_sep - offset;
  return ix;
}

// lees_edw_ic_to_buff, src/leesedwards.c:1030-1065
int lb200_le_ic_to_buff(const lb200_options_t * o, int ic, int di) {
  if (o == nullptr) return fail(LB200_EINVAL, "null argument");
  const int npl = le_nplane_local(o);
  if (npl > 0) {
    const int nh = o->nhalo;
    int p = ic/(o->nlocal[0]/npl);
    p = p < 0 ? 0 : (p > npl - 1 ? npl - 1 : p);
    int ip = lb200_le_plane_location(o, p) - (nh - 1);
    if (di > 0 && (ic >= ip && ic < ip + nh) && (ic + di >= ip + nh)) return o->nlocal[0] + (1 + 2*p)*nh + (ic - ip + 1) + di;
    ip = lb200_le_plane_location(o, p) + 1;
    if (di < 0 && (ic >= ip && ic < ip + nh) && (ic + di < ip)) return o->nlocal[0] + (2 + 2*p)*nh + (ic - ip + 1) + di;
  }
  return ic + di;
}

// lees_edw_init / lees_edw_init_tables / lees_edw_checks: src/leesedwards.c:233-279, 384-470
static int le_setup(lb200_t * c) {
  const lb200_options_t * o = &c->opt;
  memset(&c->le, 0, sizeof(c->le));
  c->nxbuf = 0;
  if (o->le_nplanes < 0) return fail(LB200_EINVAL, "le_nplanes = %d", o->le_nplanes);
  if (o->le_nplanes == 0) return 0;
  const int ntotal_x = o->nlocal[0]*o->cart_size;
  if (ntotal_x % o->le_nplanes) return fail(LB200_EINVAL, "Number of planes must divide system size (src/leesedwards.c:249-253)");
  if (o->le_nplanes % o->cart_size) return fail(LB200_EINVAL, "Must have a uniform number of planes per process (src/leesedwards.c:463-468)");
  const int npl = o->le_nplanes/o->cart_size;
  if (npl > LB200_LE_MAXPLANES) return fail(LB200_EINVAL, "more than %d planes per slab", LB200_LE_MAXPLANES);
  if (!o->periodic[1]) return fail(LB200_EINVAL, "Lees-Edwards planes need a periodic y direction");
  c->le.nplane = npl;
  c->le.xblock = o->nlocal[0]/npl;
  c->le.uy = o->le_uy;
  for (int p = 0; p < npl; p++) {
    const int ic = lb200_le_plane_location(o, p);
    if (ic <= o->nhalo || ic > o->nlocal[0] - o->nhalo) return fail(LB200_EINVAL, "Wall at domain boundary (lees_edw_checks, src/leesedwards.c:449-460)");
    c->le.loc[p] = ic;
  }
  c->nxbuf = 2*o->nhalo*npl;
  return 0;
}

static int le_alloc(lb200_t * c) {
  const lb200_options_t * o = &c->opt;
  const int npl = c->le.nplane, nh = o->nhalo;
  const size_t nyz = (size_t) o->nlocal[1]*o->nlocal[2];
  int nprop = 0;
  for (int p = 1; p < c->nvel; p++) if (c->model_h.cv[p][0] == 1) nprop++;
  c->le.nprop = nprop;
  c->le.nvel = c->nvel;

  // gradient patch: the real planes either side of each plane and the nextra = nhalo - 1 buffer planes beyond
  // (grad_3d_27pt_fluid_le, src/gradient_3d_27pt_fluid.c:421-425, 535-541)
  std::vector<int> trip, xl;
  const int ne = nh - 1;
  for (int p = 0; p < npl; p++) {
    const int ic = c->le.loc[p];
    trip.insert(trip.end(), {ic - 1, ic, lb200_le_ic_to_buff(o, ic, +1)});
    trip.insert(trip.end(), {lb200_le_ic_to_buff(o, ic + 1, -1), ic + 1, ic + 2});
    for (int n = 1; n <= ne; n++) {
      trip.insert(trip.end(), {lb200_le_ic_to_buff(o, ic, n - 1), lb200_le_ic_to_buff(o, ic, n), lb200_le_ic_to_buff(o, ic, n + 1)});
      trip.insert(trip.end(), {lb200_le_ic_to_buff(o, ic + 1, -n - 1), lb200_le_ic_to_buff(o, ic + 1, -n), lb200_le_ic_to_buff(o, ic + 1, -n + 1)});
    }
    // force / Cahn-Hilliard patch of the fused fast path: every plane whose stencil (+-2 in x, or +-1 of a
    // site whose gradient was patched) crosses the plane
    for (int x = ic - nh + 1; x <= ic + nh; x++) xl.push_back(x);
  }
  c->le_ntrip = (int) trip.size()/3;
  c->le_nxlist = (int) xl.size();
  // one-kernel step: the force / Cahn-Hilliard patch of planes ic-1 .. ic+2 also reads grad and delsq of ic-2, ic-1, ic+2, ic+3,
  // which the sweep does not store on intermediate steps
  std::vector<int> tripf(trip);
  for (int p = 0; p < npl; p++) {
    const int ic = c->le.loc[p];
    for (int x : {ic - 2, ic - 1, ic + 2, ic + 3}) tripf.insert(tripf.end(), {x - 1, x, x + 1});
  }
  c->le_ntrip_fused = (int) tripf.size()/3;
  if (cudaMalloc((void **) &c->le_trip_fused, tripf.size()*sizeof(int)) != cudaSuccess) return -1;
  if (cudaMemcpyAsync(c->le_trip_fused, tripf.data(), tripf.size()*sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -1;
  {
    // the patch chain's own stream: above the sweep's priority, so that its small grids take the SMs the sweep's CTAs free
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess
	|| cudaStreamCreateWithPriority(&c->le_stream, cudaStreamNonBlocking, hi) != cudaSuccess) { cudaGetLastError(); c->le_stream = nullptr; }
    if (cudaEventCreateWithFlags(&c->ev_le_a, cudaEventDisableTiming) != cudaSuccess
	|| cudaEventCreateWithFlags(&c->ev_le_b, cudaEventDisableTiming) != cudaSuccess) return -1;
  }
  if (cudaMalloc((void **) &c->le_trip, trip.size()*sizeof(int)) != cudaSuccess) return -1;
  if (cudaMalloc((void **) &c->le_xlist, xl.size()*sizeof(int)) != cudaSuccess) return -1;
  // on the context's own (non-blocking) stream, which orders them before every kernel of this context: a legacy-stream
  // cudaMemcpy from pageable memory may return before the DMA has landed and is not ordered with c->stream
  if (cudaMemcpyAsync(c->le_trip, trip.data(), trip.size()*sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -1;
  if (cudaMemcpyAsync(c->le_xlist, xl.data(), xl.size()*sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  int rc = 0;
  rc |= alloc_d(&c->le_term, (size_t) 3*npl*nyz);
  rc |= alloc_d(&c->le_fcor, (size_t) 3*npl);
  rc |= alloc_d(&c->le_chx, (size_t) 2*npl*nyz);
  rc |= alloc_d(&c->le_sbuf, (size_t) 2*npl*c->ndist*nprop*nyz);
  return rc;
}

// physics_control_time, src/physics.c:622-630
static double le_time(const lb200_t * c) { return 1.0*(c->t_start + c->t_current - 1.0); }

// lees_edw_buffer_displacement (steady shear), src/leesedwards.c:649-673, for a buffer with jump sign duy
static double le_buffer_displacement(const lb200_t * c, int duy, double t) {
  if (t < 0.0) t = 0.0;
  const double tle = t - 1.0*c->opt.le_nt0;
  return tle*c->opt.le_uy*duy;
}

// field_leesedwards (src/field.c:466-470, 492-497): 4-point Lagrange weights in the reference's evaluation order
static void le_interp_cubic(const lb200_t * c, Lb200LeInterp * ip) {
  const double r6 = (1.0/6.0);
  const double ltot_y = 1.0*c->g.nl[1];
  for (int s = 0; s < 2; s++) {
    double dy = le_buffer_displacement(c, s ? +1 : -1, le_time(c) + 0.0);
    dy = fmod(dy, ltot_y);
    const int jdy = (int) floor(dy);
    const double fr = 1.0 - (dy - jdy);
    ip->jdy[s] = jdy;
    ip->w[s][0] = r6*fr*(fr-1.0)*(fr-2.0);
    ip->w[s][1] = 0.5*(fr*fr-1.0)*(fr-2.0);
    ip->w[s][2] = 0.5*fr*(fr+1.0)*(fr-2.0);
    ip->w[s][3] = r6*fr*(fr*fr-1.0);
  }
}

// hydro_lees_edwards (src/hydro.c:398-404): the displacement is taken at time + 1
static void le_interp_linear(const lb200_t * c, Lb200LeInterp * ip) {
  const double ltot_y = 1.0*c->g.nl[1];
  memset(ip, 0, sizeof(*ip));
  for (int s = 0; s < 2; s++) {
    double dy = le_buffer_displacement(c, s ? +1 : -1, le_time(c) + 1.0);
    dy = fmod(dy, ltot_y);
    const int jdy = (int) floor(dy);
    const double fr = dy - jdy;
    ip->jdy[s] = jdy;
    ip->w[s][0] = fr;
    ip->w[s][1] = 1.0 - fr;
  }
}

// phi_ch_le_fix_fluxes (src/phi_cahn_hilliard.c:667-671, 692-696): lees_edw_plane_dy = time*uy, looked at from
// below (+dy) and from above (-dy); and the force flux normaliser (src/phi_force.c:647)
static void le_fix_param(const lb200_t * c, Lb200LeFix * fx) {
  const double ltot_y = 1.0*c->g.nl[1], ltot_z = 1.0*c->g.nl[2];
  const double dy0 = le_time(c)*c->opt.le_uy;
  for (int s = 0; s < 2; s++) {
    double dy = fmod(s ? -dy0 : +dy0, ltot_y);
    const int jdy = (int) floor(dy);
    fx->jdy[s] = jdy;
    fx->fr[s] = dy - jdy;
  }
  fx->ra = 0.5/(ltot_y*ltot_z);
}

// lb_data_apply_le_boundary_conditions (src/model_le.c:106-110, 375-380, 612-616): displacement of buffer
// ib = nhalo (jump +uy) at t = the time STEP, seen by the populations going up (cx = +1) and down (cx = -1)
static void le_lb_param(const lb200_t * c, Lb200LeFix * fx) {
  const double ltot_y = 1.0*c->g.nl[1];
  const double dy_le = le_buffer_displacement(c, +1, 1.0*c->t_current);
  for (int s = 0; s < 2; s++) {
    const int cx = 1 - 2*s;
    double dy = fmod(dy_le*cx, ltot_y);
    const int jdy = (int) floor(dy);
    fx->jdy[s] = jdy;
    fx->fr[s] = dy - jdy;
  }
  fx->ra = 0.0;
}

int lb200_create(const lb200_options_t * o, lb200_t ** pctx) {
  if (o == nullptr || pctx == nullptr) return fail(LB200_EINVAL, "null argument");
  *pctx = nullptr;
  for (int a = 0; a < 3; a++) if (o->nlocal[a] < 1) return fail(LB200_EINVAL, "nlocal[%d] = %d", a, o->nlocal[a]);
  if (o->nhalo < 1 || o->nhalo > 3) return fail(LB200_EINVAL, "nhalo = %d", o->nhalo);
  if (o->ndist != 1 && o->ndist != 2) return fail(LB200_EINVAL, "ndist = %d", o->ndist);
  if (o->ndist == 2 && !o->have_phi) return fail(LB200_EINVAL, "ndist = 2 (symmetric_lb) needs have_phi (reference src/ludwig.c:1340-1383)");
  if (o->have_phi && o->ndist == 1 && o->nhalo < 2) return fail(LB200_EINVAL, "the symmetric FD route needs nhalo >= 2 (reference src/ludwig.c:1198)");
  if (o->nvel != 15 && o->nvel != 19 && o->nvel != 27) return fail(LB200_EINVAL, "nvel = %d", o->nvel);
  if (o->cart_size < 1 || o->cart_rank < 0 || o->cart_rank >= o->cart_size) return fail(LB200_EINVAL, "bad cart_size/cart_rank");
  if (o->have_q && (o->have_phi || o->ndist != 1)) return fail(LB200_EINVAL, "have_q (lc_blue_phase) excludes have_phi / ndist = 2");
  if (o->have_q && o->nhalo < 2) return fail(LB200_EINVAL, "the liquid crystal needs nhalo >= 2 (reference src/ludwig.c:1605)");
  if (o->have_q && o->le_nplanes != 0) return fail(LB200_EINVAL, "have_q: no Lees-Edwards planes in this build");
  if (o->halo_scheme != LB200_HALO_FULL && o->halo_scheme != LB200_HALO_REDUCED) return fail(LB200_EINVAL, "halo_scheme = %d", o->halo_scheme);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(LB200_ENODEVICE, "no CUDA device: libludwig_b200 has no CPU fallback");
  }

  lb200_t * c = new (std::nothrow) lb200_t();
  if (c == nullptr) return fail(LB200_ENOMEM, "host allocation failed");
  memset(c, 0, sizeof(*c));
  c->opt = *o;
  if (o->device >= 0) { CUDA_TRY(cudaSetDevice(o->device)); }
  CUDA_TRY(cudaGetDevice(&c->device));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->comm, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_phi, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_u, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_f, cudaEventDisableTiming));
  c->k = (o->math == LB200_MATH_STRICT) ? &lb200_kernels_strict : &lb200_kernels_fast;
  c->nvel = o->nvel;
  c->ndist = o->ndist;

  Lb200Geom & g = c->g;
  for (int a = 0; a < 3; a++) { g.nl[a] = o->nlocal[a]; g.nall[a] = o->nlocal[a] + 2*o->nhalo; g.per[a] = o->periodic[a] != 0; }
  g.nh = o->nhalo;
  g.ys = g.nall[2];
  g.xs = g.nall[1]*g.nall[2];
  if (le_setup(c) != 0) { delete c; return LB200_EINVAL; }
  c->nsites_lb = g.nall[0]*g.nall[1]*g.nall[2];
  const long long ns = (long long) (g.nall[0] + c->nxbuf)*g.nall[1]*g.nall[2];
  if (ns*c->nvel*c->ndist > 2147483647LL) {
    delete c;
    return fail(LB200_EINVAL, "nsites*nvel exceeds INT_MAX (reference guard src/lb_data.c:116-120)");
  }
  g.nsites = (int) ns;
  g.remote_x = (o->cart_size > 1);
  g.has_lo = g.per[0] || o->cart_rank > 0;
  g.has_hi = g.per[0] || o->cart_rank < o->cart_size - 1;

  if (model_init(c->nvel, &c->model_h) != 0) { delete c; return fail(LB200_EINVAL, "model"); }
  CUDA_TRY(cudaMalloc((void **) &c->model_d, sizeof(Lb200ModelDev)));
  CUDA_TRY(cudaMemcpyAsync(c->model_d, &c->model_h, sizeof(Lb200ModelDev), cudaMemcpyHostToDevice, c->stream));   // ordered before every kernel
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->unrolled19 = (c->nvel == 19);
  {
    const char * e = getenv("LB200_GENERIC_D3Q19");     // model matrices instead of coded constants
    if (e && atoi(e) != 0) c->unrolled19 = 0;
  }

  const size_t nsz = (size_t) g.nsites;
  int rc = 0;
  rc |= alloc_d(&c->f, nsz*c->nvel*c->ndist);
  rc |= alloc_d(&c->fprime, nsz*c->nvel*c->ndist);
  rc |= alloc_d(&c->u, nsz*3);
  rc |= alloc_d(&c->rho, nsz);
  rc |= alloc_d(&c->force, nsz*3);
  if (o->have_phi) {
    rc |= alloc_d(&c->phi, nsz);
    rc |= alloc_d(&c->phinew, nsz);
    rc |= alloc_d(&c->grad, nsz*3);
    rc |= alloc_d(&c->delsq, nsz);
  }
  if (o->have_q) {
    rc |= alloc_d(&c->q, nsz*5);
    rc |= alloc_d(&c->qnew, nsz*5);
    rc |= alloc_d(&c->str, nsz*9);
  }
  if (g.remote_x) {
    // staging: nvel planes of depth 1 (f) | 3 components of depth nhalo (u) | nhalo planes (phi)
    c->stage_doubles = (size_t) c->nvel*c->ndist*g.xs + (size_t) 3*g.nh*g.xs + (size_t) g.nh*g.xs
      + (o->have_q ? (size_t) 5*g.nh*g.xs : 0);
    rc |= alloc_d(&c->xlo, c->stage_doubles);
    rc |= alloc_d(&c->xhi, c->stage_doubles);
    rc |= alloc_d(&c->slo, c->stage_doubles);
    rc |= alloc_d(&c->shi, c->stage_doubles);
  }
  if (rc == 0 && c->le.nplane > 0) rc = le_alloc(c);
  if (rc != 0) { lb200_free(c); return LB200_ECUDA; }
  CUDA_TRY(cudaMalloc((void **) &c->status, nsz));
  CUDA_TRY(cudaMemsetAsync(c->status, 0, nsz, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->map_all_fluid = 1;

  c->force_state = ARRAY_CLEAN;
  c->u_state = ARRAY_CLEAN;
  c->knob_wrap = getenv("LB200_WRAP") ? atoi(getenv("LB200_WRAP")) : 1;
  c->knob_phi_sector = getenv("LB200_PHI_SECTOR") ? atoi(getenv("LB200_PHI_SECTOR")) : 1;
  c->knob_peer = getenv("LB200_PEER") ? atoi(getenv("LB200_PEER")) : 1;
  c->knob_pipe = getenv("LB200_PIPE") ? atoi(getenv("LB200_PIPE")) : 0;
  c->knob_f32 = getenv("LB200_F32") ? atoi(getenv("LB200_F32")) : 0;
  c->knob_fused = getenv("LB200_FUSED") ? atoi(getenv("LB200_FUSED")) : 1;
  c->knob_fused_le = getenv("LB200_FUSED_LE") ? atoi(getenv("LB200_FUSED_LE")) : 1;
  c->knob_grad7 = getenv("LB200_GRAD_7PT") ? atoi(getenv("LB200_GRAD_7PT")) : 0;
  c->knob_lazy_diag = getenv("LB200_LAZY_DIAG") ? atoi(getenv("LB200_LAZY_DIAG")) : 1;
  c->knob_pipe_sms = getenv("LB200_PIPE_SMS") ? atoi(getenv("LB200_PIPE_SMS")) : 56;
  c->f_alloc[0] = c->f; c->f_alloc[1] = c->fprime;
  c->phi_alloc[0] = c->phi; c->phi_alloc[1] = c->phinew;
  if (o->have_q) { c->phi_alloc[0] = c->q; c->phi_alloc[1] = c->qnew; }     // the order-parameter slots of the peer links carry Q
  c->u_alloc[0] = c->u; c->u_alloc[1] = nullptr;
  for (int i = 0; i < LB200_KCLASS_MAX; i++) c->ev[i] = new std::vector<cudaEvent_t>();
  *pctx = c;
  return 0;
}

int lb200_set_knob(lb200_t * c, int knob, int value) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  if (knob == LB200_KNOB_WRAP) c->knob_wrap = (value != 0);
  else if (knob == LB200_KNOB_PHI_SECTOR) c->knob_phi_sector = (value != 0);
  else if (knob == LB200_KNOB_PEER) { c->knob_peer = (value != 0); c->wrap_x_valid = 0; }
  else if (knob == LB200_KNOB_F32) c->knob_f32 = (value != 0);
  else if (knob == LB200_KNOB_FUSED) c->knob_fused = (value != 0);
  else if (knob == LB200_KNOB_GRAD_7PT) {
    if (value != 0 && c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
    c->knob_grad7 = (value != 0);
  }
  else if (knob == LB200_KNOB_QGRAD_2D5) {
    if (value != 0 && c->q == nullptr) return fail(LB200_ESTATE, "no q in this context");
    if (value != 0 && c->g.nl[2] != 1) return fail(LB200_EINVAL, "2d_5pt_fluid needs a lattice with one plane in z");
    c->knob_qgrad2d = (value != 0);
  }
  else if (knob == LB200_KNOB_PIPE) c->knob_pipe = (value < 0) ? 0 : (value > LB200_PIPE_MAXS ? LB200_PIPE_MAXS : value);
  else if (knob == LB200_KNOB_PIPE_SMS) {
    if (c->pipe_state != 0) return fail(LB200_ESTATE, "the SM partitions of the slab pipeline are already provisioned");
    c->knob_pipe_sms = value;
  }
  else return fail(LB200_EINVAL, "unknown knob %d", knob);
  return 0;
}

int lb200_profile(lb200_t * c, int on) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < LB200_KCLASS_MAX; i++) {
    for (cudaEvent_t e : *c->ev[i]) cudaEventDestroy(e);
    c->ev[i]->clear();
  }
  c->profile = on;
  return 0;
}

int lb200_profile_get(lb200_t * c, int kclass, double * total_ms, int * count) {
  if (c == nullptr || kclass < 0 || kclass >= LB200_KCLASS_MAX) return fail(LB200_EINVAL, "bad argument");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  double t = 0.0;
  const std::vector<cudaEvent_t> & v = *c->ev[kclass];
  for (size_t i = 0; i + 1 < v.size(); i += 2) {
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
    t += ms;
  }
  if (total_ms) *total_ms = t;
  if (count) *count = (int) (v.size()/2);
  return 0;
}

static void pipe_teardown(lb200_t * c);

int lb200_free(lb200_t * c) {
  if (c == nullptr) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->f); cudaFree(c->fprime); cudaFree(c->u_alloc[0] ? c->u_alloc[0] : c->u); cudaFree(c->u_alloc[1]); cudaFree(c->rho); cudaFree(c->force);
  cudaFree(c->phi); cudaFree(c->phinew); cudaFree(c->grad); cudaFree(c->delsq);
  cudaFree(c->grad_delsq); cudaFree(c->delsq_delsq); cudaFree(c->str);
  cudaFree(c->q); cudaFree(c->qnew); cudaFree(c->qgrad); cudaFree(c->qdelsq);
  if (c->le_stream) cudaStreamDestroy(c->le_stream);
  if (c->ev_le_a) cudaEventDestroy(c->ev_le_a);
  if (c->ev_le_b) cudaEventDestroy(c->ev_le_b);
  cudaFree(c->le_trip); cudaFree(c->le_trip_fused); cudaFree(c->le_xlist); cudaFree(c->le_term); cudaFree(c->le_fcor); cudaFree(c->le_chx); cudaFree(c->le_sbuf);
  for (int i = 0; i < c->nmapped; i++) cudaIpcCloseMemHandle(c->mapped[i]);
  cudaFree(c->f32[0]); cudaFree(c->f32[1]); cudaFree(c->csum);
  cudaFree(c->flags); cudaFree(c->spin_err); cudaFree(c->halo_snap); cudaFree(c->psum);
  cudaFree(c->status); cudaFree(c->xlo); cudaFree(c->xhi); cudaFree(c->slo); cudaFree(c->shi); cudaFree(c->model_d);
  for (int i = 0; i < LB200_KCLASS_MAX; i++) {
    if (c->ev[i]) { for (cudaEvent_t e : *c->ev[i]) cudaEventDestroy(e); delete c->ev[i]; }
  }
  pipe_teardown(c);
  cudaStreamSynchronize(c->comm);
  cudaEventDestroy(c->ev_main); cudaEventDestroy(c->ev_phi); cudaEventDestroy(c->ev_u); cudaEventDestroy(c->ev_f);
  cudaStreamDestroy(c->comm);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int lb200_nsites(const lb200_t * c) { return c ? c->g.nsites : LB200_EINVAL; }
int lb200_nsites_lb(const lb200_t * c) { return c ? c->nsites_lb : LB200_EINVAL; }
long long lb200_launch_count(const lb200_t * c) { return c ? c->launches : 0; }
void * lb200_stream(lb200_t * c) { return c ? (void *) c->stream : nullptr; }

// The fallback flag wait (a spinning kernel, when the driver has no stream memory operations) gives up after 20 s:
// whatever ran after it then read halo planes that never arrived.  Checked wherever the host waits for the stream.
static int spin_check(lb200_t * c) {
  if (!c->spin_used) return 0;
  int e = 0;
  CUDA_TRY(cudaMemcpy(&e, c->spin_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e != 0) return fail(LB200_ECOMM, "a neighbour GPU's boundary planes did not arrive within 20 s (peer-store flag wait timed out)");
  return 0;
}

int lb200_sync(lb200_t * c) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return spin_check(c);
}

static const char * status_ptr(const lb200_t * c) { return c->map_all_fluid ? nullptr : c->status; }
static const Lb200ModelDev * model_ptr(const lb200_t * c) { return c->unrolled19 ? nullptr : c->model_d; }

// ---- lazily applied operations --------------------------------------------------------------------

static int halo_field(lb200_t * c, double * data, int ncomp, int depth, int reduced, cudaStream_t st);

// a halo-free lb200_step leaves "lb_halo(f); lb_propagation(f)" pending: apply the halo swap first
static int ensure_f_halo(lb200_t * c) {
  if (!c->f_halo_stale) return 0;
  c->f_halo_stale = 0;
  return halo_field(c, c->f, c->nvel*c->ndist, 1, c->opt.halo_scheme == LB200_HALO_REDUCED, nullptr);
}

// apply a pending lb_propagation as a stand-alone sweep (reference semantics, src/propagation.c)
static int materialise_propagation(lb200_t * c) {
  if (!c->prop_pending) { c->f_halo_stale = 0; return 0; }
  {
    int rc = ensure_f_halo(c);
    if (rc != 0) return rc;
  }
  {
    ProfScope ps(c, LB200_K_PROPAGATE);
    c->launches += c->k->propagate(c->stream, c->g, c->model_d, c->nvel, c->ndist, c->f, c->fprime);
  }
  double * t = c->f; c->f = c->fprime; c->fprime = t;       // lb_model_swapf, src/propagation.c:211-240
  c->prop_pending = 0;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int materialise_zero(lb200_t * c, double * a, int * state) {
  if (*state == ZERO_PENDING) {
    CUDA_TRY(cudaMemsetAsync(a, 0, (size_t) 3*c->g.nsites*sizeof(double), c->stream));
  }
  else if (*state == INTERIOR_ONLY) {
    c->launches += c->k->zero_outside(c->stream, c->g, 3, a);
  }
  *state = ARRAY_CLEAN;
  return 0;
}

// ---- memcpy --------------------------------------------------------------------------------------

static int array_info(lb200_t * c, int array, double ** dev, size_t * ncomp) {
  switch (array) {
  case LB200_F:     *dev = c->f;     *ncomp = (size_t) c->nvel*c->ndist; break;
  case LB200_PHI:   *dev = c->phi;   *ncomp = 1; break;
  case LB200_U:     *dev = c->u;     *ncomp = 3; break;
  case LB200_RHO:   *dev = c->rho;   *ncomp = 1; break;
  case LB200_FORCE: *dev = c->force; *ncomp = 3; break;
  case LB200_GRAD:  *dev = c->grad;  *ncomp = 3; break;
  case LB200_DELSQ: *dev = c->delsq; *ncomp = 1; break;
  case LB200_GRAD_DELSQ:  *dev = c->grad_delsq;  *ncomp = 3; break;
  case LB200_DELSQ_DELSQ: *dev = c->delsq_delsq; *ncomp = 1; break;
  case LB200_STR:   *dev = c->str;   *ncomp = 9; break;
  case LB200_Q:      *dev = c->q;      *ncomp = 5; break;
  case LB200_QGRAD:  *dev = c->qgrad;  *ncomp = 15; break;
  case LB200_QDELSQ: *dev = c->qdelsq; *ncomp = 5; break;
  default: return fail(LB200_EINVAL, "unknown array id %d", array);
  }
  if (*dev == nullptr) return fail(LB200_ESTATE, "array %d is not allocated in this context (have_phi = 0?)", array);
  return 0;
}

static int do_memcpy(lb200_t * c, int array, double * host, int kind, int async) {
  if (c == nullptr || host == nullptr) return fail(LB200_EINVAL, "null argument");
  if (kind != LB200_HOST_TO_DEVICE && kind != LB200_DEVICE_TO_HOST) return fail(LB200_EINVAL, "memcpy kind %d", kind);
  CUDA_TRY(cudaSetDevice(c->device));

  if (array == LB200_MAP) {
    // status is exchanged as doubles at this interface and held as bytes on the device
    // (map->status has cs_nsites entries, a prefix of the device array when there are Lees-Edwards buffer planes)
    const size_t ns = (size_t) c->nsites_lb;
    std::vector<char> tmp(ns);
    if (kind == LB200_HOST_TO_DEVICE) {
      int all_fluid = 1;
      for (size_t i = 0; i < ns; i++) { tmp[i] = (char) host[i]; if (tmp[i] != 0) all_fluid = 0; }
      CUDA_TRY(cudaMemcpyAsync(c->status, tmp.data(), ns, cudaMemcpyHostToDevice, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      c->map_all_fluid = all_fluid;
    }
    else {
      CUDA_TRY(cudaMemcpyAsync(tmp.data(), c->status, ns, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      for (size_t i = 0; i < ns; i++) host[i] = (double) tmp[i];
    }
    return 0;
  }

  double * dev = nullptr;
  size_t ncomp = 0;
  int rc = array_info(c, array, &dev, &ncomp);
  if (rc != 0) return rc;
  c->fused_ready = 0;          // uploads replace the state; downloads materialise zeros / a pending propagation in the halos

  if (array == LB200_F) {
    rc = materialise_propagation(c);
    if (rc != 0) return rc;
    dev = c->f;
  }
  if (kind == LB200_DEVICE_TO_HOST) {
    if (array == LB200_FORCE) materialise_zero(c, c->force, &c->force_state);
    if (array == LB200_U) materialise_zero(c, c->u, &c->u_state);
  }
  else {
    if (array == LB200_FORCE) c->force_state = ARRAY_CLEAN;
    if (array == LB200_U) { c->u_state = ARRAY_CLEAN; c->u_halo_valid = 0; }
    if (array == LB200_PHI) c->phi_halo_valid = 0;
    if (array == LB200_F) c->f_halo_stale = 0;
    c->wrap_x_valid = 0;
  }

  const size_t bytes = ncomp*(size_t) c->g.nsites*sizeof(double);
  if (array == LB200_F && c->nxbuf > 0) {
    // lb->f has cs_nsites per population on the host (src/lb_data.c:102-115), the device arrays share one stride
    const size_t hpitch = (size_t) c->nsites_lb*sizeof(double), dpitch = (size_t) c->g.nsites*sizeof(double);
    if (kind == LB200_HOST_TO_DEVICE) CUDA_TRY(cudaMemcpy2DAsync(dev, dpitch, host, hpitch, hpitch, ncomp, cudaMemcpyHostToDevice, c->stream));
    else                              CUDA_TRY(cudaMemcpy2DAsync(host, hpitch, dev, dpitch, hpitch, ncomp, cudaMemcpyDeviceToHost, c->stream));
  }
  else if (kind == LB200_HOST_TO_DEVICE) {
    CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
  }
  else {
    CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  if (!async) CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

int lb200_memcpy(lb200_t * c, int array, double * host, int kind) {
  const int rc = do_memcpy(c, array, host, kind, 0);
  return (rc == 0 && kind == LB200_DEVICE_TO_HOST) ? spin_check(c) : rc;
}
int lb200_memcpy_async(lb200_t * c, int array, double * host, int kind) { return do_memcpy(c, array, host, kind, 1); }

int lb200_device_ptr(lb200_t * c, int array, void ** ptr) {
  if (c == nullptr || ptr == nullptr) return fail(LB200_EINVAL, "null argument");
  if (array == LB200_MAP) { *ptr = c->status; return 0; }
  double * dev = nullptr;
  size_t ncomp = 0;
  int rc = array_info(c, array, &dev, &ncomp);
  if (rc != 0) return rc;
  if (array == LB200_F) { rc = materialise_propagation(c); dev = c->f; }
  *ptr = dev;
  return rc;
}

// staging areas: f | u | phi are disjoint so that exchanges in flight at the same time never alias
static double * stage_ptr(lb200_t * c, double * base, const double * data) {
  if (base == nullptr) return nullptr;
  const size_t xs = (size_t) c->g.xs;
  if (data == c->u) return base + (size_t) c->nvel*c->ndist*xs;
  if (data == c->phi || data == c->phinew) return base + (size_t) c->nvel*c->ndist*xs + (size_t) 3*c->g.nh*xs;
  if (c->q != nullptr && (data == c->q || data == c->qnew)) return base + (size_t) c->nvel*c->ndist*xs + (size_t) 4*c->g.nh*xs;
  return base;
}
static double * stage_lo(lb200_t * c, const double * data) { return stage_ptr(c, c->xlo, data); }
static double * stage_hi(lb200_t * c, const double * data) { return stage_ptr(c, c->xhi, data); }

int lb200_slab_plan(const lb200_options_t * o, int ncomp, int depth, lb200_slab_plan_t * p) {
  if (o == nullptr || p == nullptr) return fail(LB200_EINVAL, "null argument");
  if (depth < 1 || depth > o->nhalo || ncomp < 1) return fail(LB200_EINVAL, "bad depth/ncomp");
  if (o->cart_size < 1 || o->cart_rank < 0 || o->cart_rank >= o->cart_size) return fail(LB200_EINVAL, "bad cart_size/cart_rank");
  const long long nay = o->nlocal[1] + 2*o->nhalo, naz = o->nlocal[2] + 2*o->nhalo, nax = o->nlocal[0] + 2*o->nhalo;
  const long long xs = nay*naz;
  p->left = (o->cart_rank - 1 + o->cart_size) % o->cart_size;
  p->right = (o->cart_rank + 1) % o->cart_size;
  p->has_lo = (o->periodic[0] != 0) || o->cart_rank > 0;
  p->has_hi = (o->periodic[0] != 0) || o->cart_rank < o->cart_size - 1;
  p->nsites = (nax + 2*o->nhalo*le_nplane_local(o))*xs;
  p->chunk = depth*xs;
  p->off_lo = (long long) (1 + o->nhalo - 1)*xs;
  p->off_hi = (long long) (o->nlocal[0] - depth + 1 + o->nhalo - 1)*xs;
  p->halo_lo = (long long) (1 - depth + o->nhalo - 1)*xs;
  p->halo_hi = (long long) (o->nlocal[0] + 1 + o->nhalo - 1)*xs;
  p->count = ncomp*p->chunk;
  return 0;
}

int lb200_step_plan(const lb200_options_t * o, int what, lb200_step_plan_t * p) {
  if (o == nullptr || p == nullptr) return fail(LB200_EINVAL, "null argument");
  if (o->cart_size < 1 || o->cart_rank < 0 || o->cart_rank >= o->cart_size) return fail(LB200_EINVAL, "bad cart_size/cart_rank");
  if (what != LB200_STEP_PHI && what != LB200_STEP_UX && what != LB200_STEP_F) return fail(LB200_EINVAL, "unknown step array %d", what);
  memset(p, 0, sizeof(*p));
  const long long nay = o->nlocal[1] + 2*o->nhalo, naz = o->nlocal[2] + 2*o->nhalo, nax = o->nlocal[0] + 2*o->nhalo;
  const long long xs = nay*naz;
  p->left = (o->cart_rank - 1 + o->cart_size) % o->cart_size;
  p->right = (o->cart_rank + 1) % o->cart_size;
  p->depth = (what == LB200_STEP_PHI) ? o->nhalo : 1;
  p->nsites = (nax + 2*o->nhalo*le_nplane_local(o))*xs;
  p->chunk = p->depth*xs;
  p->src_up = (long long) (o->nlocal[0] - p->depth + o->nhalo)*xs;
  p->dst_up = (long long) (o->nhalo - p->depth)*xs;
  p->src_down = (long long) o->nhalo*xs;
  p->dst_down = (long long) (o->nlocal[0] + o->nhalo)*xs;
  p->peer_shift = (long long) o->nlocal[0]*xs;
  if (what == LB200_STEP_F) {
    Lb200ModelDev md;
    if (model_init(o->nvel, &md) != 0) return fail(LB200_EINVAL, "nvel = %d", o->nvel);
    for (int q = 0; q < o->nvel; q++) {
      if (md.cv[q][0] > 0) p->comp_up[p->ncomp_up++] = q;
      if (md.cv[q][0] < 0) p->comp_down[p->ncomp_down++] = q;
    }
  }
  else {
    p->ncomp_up = p->ncomp_down = 1;
  }
  return 0;
}

// ---- x-plane exchange between slabs (NCCL send/recv over NVLink) --------------------------------
// Replaces MPI_Isend/Irecv/Waitall of lb_halo_post / field_halo_post.  For each component the
// `depth` boundary planes at either end of the slab are contiguous in the SOA layout, so no pack
// kernel is needed on the send side.

static int exchange_x(lb200_t * c, cudaStream_t st, const double * data, int ncomp, int depth) {
  if (!c->g.remote_x) return 0;
#ifdef LB200_NO_NCCL
  return fail(LB200_ECOMM, "library built without NCCL");
#else
  if (c->nccl == nullptr) return fail(LB200_ECOMM, "cart_size > 1 but no NCCL communicator attached (lb200_attach_nccl)");
  ncclComm_t comm = (ncclComm_t) c->nccl;
  const Lb200Geom & g = c->g;
  lb200_slab_plan_t plan;
  int prc = lb200_slab_plan(&c->opt, ncomp, depth, &plan);
  if (prc != 0) return prc;
  const int left = plan.left, right = plan.right;
  const size_t chunk = (size_t) plan.chunk;
  const size_t ns = (size_t) plan.nsites;
  // my top planes i in [N-d+1, N] go right (-> neighbour's xlo); my bottom planes i in [1, d] go left
  const size_t off_hi = (size_t) plan.off_hi;
  const size_t off_lo = (size_t) plan.off_lo;
  double * xlo = stage_lo(c, data);
  double * xhi = stage_hi(c, data);

  // One message per direction: the boundary planes of all components are first gathered into a
  // contiguous send buffer with one strided device copy (each component's planes are already
  // contiguous), laid out exactly as the receiver's staging area.  (One ncclSend per component cost
  // 0.70 ms for the 19 populations at 256^2 planes; see profiles/r01_halo_exchange.md.)
  const double * send_hi = data + off_hi;
  const double * send_lo = data + off_lo;
  if (ncomp > 1) {
    double * shi = stage_ptr(c, c->shi, data);
    double * slo = stage_ptr(c, c->slo, data);
    if (g.has_hi) CUDA_TRY(cudaMemcpy2DAsync(shi, chunk*sizeof(double), data + off_hi, ns*sizeof(double),
					     chunk*sizeof(double), ncomp, cudaMemcpyDeviceToDevice, st));
    if (g.has_lo) CUDA_TRY(cudaMemcpy2DAsync(slo, chunk*sizeof(double), data + off_lo, ns*sizeof(double),
					     chunk*sizeof(double), ncomp, cudaMemcpyDeviceToDevice, st));
    send_hi = shi;
    send_lo = slo;
  }
  const size_t count = (size_t) ncomp*chunk;

  ncclResult_t r = ncclGroupStart();
  if (g.has_hi && r == ncclSuccess) r = ncclSend(send_hi, count, ncclDouble, right, comm, st);
  if (g.has_lo && r == ncclSuccess) r = ncclSend(send_lo, count, ncclDouble, left, comm, st);
  if (g.has_lo && r == ncclSuccess) r = ncclRecv(xlo, count, ncclDouble, left, comm, st);
  if (g.has_hi && r == ncclSuccess) r = ncclRecv(xhi, count, ncclDouble, right, comm, st);
  ncclResult_t r2 = ncclGroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) {
    return fail(LB200_ECOMM, "NCCL halo exchange: %s", ncclGetErrorString(r != ncclSuccess ? r : r2));
  }
  return 0;
#endif
}

int lb200_attach_nccl(lb200_t * c, void * comm) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  c->nccl = comm;
  return 0;
}

int lb200_nccl_unique_id(void * id128) {
#ifdef LB200_NO_NCCL
  return fail(LB200_ECOMM, "library built without NCCL");
#else
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclResult_t r = ncclGetUniqueId((ncclUniqueId *) id128);
  if (r != ncclSuccess) return fail(LB200_ECOMM, "ncclGetUniqueId: %s", ncclGetErrorString(r));
  return 0;
#endif
}

int lb200_nccl_comm_create(const void * id128, int nranks, int rank, void ** comm) {
#ifdef LB200_NO_NCCL
  return fail(LB200_ECOMM, "library built without NCCL");
#else
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t cm;
  ncclResult_t r = ncclCommInitRank(&cm, nranks, id, rank);
  if (r != ncclSuccess) return fail(LB200_ECOMM, "ncclCommInitRank: %s", ncclGetErrorString(r));
  *comm = (void *) cm;
  return 0;
#endif
}

int lb200_nccl_comm_destroy(void * comm) {
#ifndef LB200_NO_NCCL
  if (comm) ncclCommDestroy((ncclComm_t) comm);
#endif
  return 0;
}

// ---- operators ------------------------------------------------------------------------------------

#define CTX_ENTER(c) do { if ((c) == nullptr) return fail(LB200_EINVAL, "null context"); \
  CUDA_TRY(cudaSetDevice((c)->device)); (c)->phi_halo_valid = 0; (c)->u_halo_valid = 0; (c)->wrap_x_valid = 0; (c)->fused_ready = 0; } while (0)
#define CTX_LEAVE_SYNC(c) do { CUDA_TRY(cudaGetLastError()); CUDA_TRY(cudaStreamSynchronize((c)->stream)); return 0; } while (0)

int lb200_hydro_f_zero(lb200_t * c) {
  CTX_ENTER(c);
  c->force_state = ZERO_PENDING;         // folded into the next producer / consumer of the force
  return 0;
}

int lb200_hydro_u_zero(lb200_t * c) {
  CTX_ENTER(c);
  c->u_state = ZERO_PENDING;
  return 0;
}

// a periodic dimension thinner than the halo (the reference's 2-d runs): halo swaps then deliver pre-swap halo content,
// so what the halos hold BEFORE a swap matters and must be what the reference's arrays hold
static bool thin_lattice(const lb200_t * c) {
  for (int a = 0; a < 3; a++) if (c->g.per[a] && c->g.nl[a] < c->g.nh) return true;
  return false;
}

static int halo_field(lb200_t * c, double * data, int ncomp, int depth, int reduced, cudaStream_t st) {
  if (st == nullptr) st = c->stream;
  // A periodic lattice thinner than the swap depth (e.g. 64 x 64 x 1 with nhalo 2, the reference's 2-d runs): the
  // reference's send regions then reach into the halo, and because it packs every send buffer before it unpacks any
  // (src/field.c:1329-1355, 1412-1531) what arrives in the outer halo layers is the halo's content from BEFORE the swap.
  // Reproduced by reading the local sources from a snapshot of the array.
  bool thin = false;
  for (int a = 0; a < 3; a++) thin = thin || (c->g.per[a] && c->g.nl[a] < depth);
  const double * snapshot = nullptr;
  if (thin) {
    if (c->g.remote_x && c->g.nl[0] < depth) return fail(LB200_EINVAL, "x-slabs thinner than the halo");
    const size_t need = (size_t) ncomp*c->g.nsites;
    if (c->halo_snap_size < need) {
      cudaFree(c->halo_snap);
      c->halo_snap = nullptr; c->halo_snap_size = 0;
      if (alloc_d(&c->halo_snap, need) != 0) return LB200_ECUDA;
      c->halo_snap_size = need;
    }
    CUDA_TRY(cudaMemcpyAsync(c->halo_snap, data, need*sizeof(double), cudaMemcpyDeviceToDevice, st));
    snapshot = c->halo_snap;
  }
  ProfScope ps(c, LB200_K_HALO, st);
  int rc = exchange_x(c, st, data, ncomp, depth);
  if (rc != 0) return rc;
  c->launches += c->k->halo(st, c->g, c->model_d, ncomp, depth, reduced, data, stage_lo(c, data), stage_hi(c, data), snapshot);
  return 0;
}

static int u_halo_async(lb200_t * c) {
  if (c->u_state == ZERO_PENDING || (c->u_state == INTERIOR_ONLY && thin_lattice(c))) materialise_zero(c, c->u, &c->u_state);
  int rc = halo_field(c, c->u, 3, c->g.nh, 0, nullptr);
  c->u_state = ARRAY_CLEAN;              // every halo site within nhalo has just been written
  return rc;
}

int lb200_hydro_u_halo(lb200_t * c) {
  CTX_ENTER(c);
  int rc = u_halo_async(c);
  if (rc != 0) return rc;
  CTX_LEAVE_SYNC(c);
}

int lb200_phi_halo(lb200_t * c) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  int rc = halo_field(c, c->phi, 1, c->g.nh, 0, nullptr);
  if (rc != 0) return rc;
  CTX_LEAVE_SYNC(c);
}

// field_leesedwards, src/field.c:418-510
static int le_field_async(lb200_t * c, double * phi, const Lb200Geom * gw = nullptr, cudaStream_t st = nullptr) {
  if (c->le.nplane == 0) return 0;
  Lb200LeInterp ip;
  le_interp_cubic(c, &ip);
  ProfScope ps(c, LB200_K_LE, st);
  c->launches += c->k->le_interp(st ? st : c->stream, gw ? *gw : c->g, c->le, ip, 1, 1, c->g.nh, phi);
  return 0;
}

// hydro_lees_edwards, src/hydro.c:350-440
static int le_hydro_async(lb200_t * c, const Lb200Geom * gw = nullptr, cudaStream_t st = nullptr) {
  if (c->le.nplane == 0) return 0;
  Lb200LeInterp ip;
  le_interp_linear(c, &ip);
  ProfScope ps(c, LB200_K_LE, st);
  c->launches += c->k->le_interp(st ? st : c->stream, gw ? *gw : c->g, c->le, ip, 0, 3, c->g.nh, c->u);
  return 0;
}

// grad_3d_27pt_fluid_d2 with planes: the two real planes next to each plane again, through the buffer planes,
// and the buffer planes themselves (src/gradient_3d_27pt_fluid.c:94-95, 250-253, 375-651)
static int le_grad_async(lb200_t * c, const Lb200Geom * gw = nullptr) {
  if (c->le.nplane == 0) return 0;
  ProfScope ps(c, LB200_K_LE);
  // halo-free steps: every reader takes its y / z neighbours from the interior, so only interior (j, k) are needed
  const int ne = (gw && gw->wrap[1] && gw->wrap[2]) ? 0 : c->g.nh - 1;
  // (fd_gradient_calculation 3d_7pt_fluid: the launcher takes -ne - 1)
  c->launches += c->k->le_grad_planes(c->stream, gw ? *gw : c->g, c->knob_grad7 ? -ne - 1 : ne, c->le_ntrip, c->le_trip, c->phi, c->grad, c->delsq);
  return 0;
}

int lb200_phi_grad_compute(lb200_t * c) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  le_field_async(c, c->phi);              // field_grad_compute -> field_leesedwards, src/field_grad.c:324
  {
    ProfScope ps(c, LB200_K_GRAD);
    if (c->knob_grad7) c->launches += c->k->grad7(c->stream, c->g, c->g.nh - 1, 1, c->phi, c->grad, c->delsq);   // grad_3d_7pt_fluid_d2
    else               c->launches += c->k->grad27(c->stream, c->g, c->g.nh - 1, c->phi, c->grad, c->delsq);
  }
  le_grad_async(c);
  CTX_LEAVE_SYNC(c);
}

int lb200_field_leesedwards(lb200_t * c) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  le_field_async(c, c->phi);
  CTX_LEAVE_SYNC(c);
}

int lb200_hydro_lees_edwards(lb200_t * c) {
  CTX_ENTER(c);
  if (c->u_state == ZERO_PENDING) materialise_zero(c, c->u, &c->u_state);
  le_hydro_async(c);
  CTX_LEAVE_SYNC(c);
}

int lb200_physics_control_time_set(lb200_t * c, int t_start, int t_current) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  c->t_start = t_start;
  c->t_current = t_current;
  return 0;
}

int lb200_physics_control_timestep(const lb200_t * c) { return c ? c->t_current : LB200_EINVAL; }

// phi_force_flux / phi_cahn_hilliard with planes: the generic kernels of lb200_le.cuh on nx planes
// (xlist == nullptr: the whole lattice)
static int le_force_ch_async(lb200_t * c, const Lb200SymmDev & sd, int nx, const int * xlist, int do_force, int do_ch,
			     int accumulate, double * phinew, const Lb200Geom * gw = nullptr, cudaStream_t st = nullptr) {
  Lb200LeFix fx;
  le_fix_param(c, &fx);
  const Lb200Geom & g = gw ? *gw : c->g;
  if (st == nullptr) st = c->stream;
  ProfScope ps(c, LB200_K_LE, st);
  if (do_force && do_ch) {
    c->launches += c->k->le_prep_both(st, g, c->le, sd, c->phi, c->grad, c->delsq, c->u, status_ptr(c), c->le_term, c->le_fcor, c->le_chx);
  }
  else {
    if (do_force) c->launches += c->k->le_force_prep(st, g, c->le, sd, c->phi, c->grad, c->delsq, c->le_term, c->le_fcor);
    if (do_ch)    c->launches += c->k->le_ch_prep(st, g, c->le, sd, c->phi, c->delsq, c->u, status_ptr(c), c->le_chx);
  }
  c->launches += c->k->le_force_ch(st, g, c->le, sd, fx, nx, xlist, do_force, do_ch, accumulate, c->phi, c->grad,
				   c->delsq, c->u, status_ptr(c), c->le_fcor, c->le_chx, c->force, phinew);
  return 0;
}

// lb_data_apply_le_boundary_conditions, src/model_le.c:78-180 (in place on the post-collision distributions)
static int le_lb_bc_async(lb200_t * c, double * f = nullptr, cudaStream_t st = nullptr) {
  if (c->le.nplane == 0) return 0;
  Lb200LeFix fx;
  le_lb_param(c, &fx);
  ProfScope ps(c, LB200_K_LE, st);
  c->launches += c->k->le_lb_bc(st ? st : c->stream, c->g, c->le, fx, c->model_d, c->ndist, f ? f : c->f, c->le_sbuf);
  return 0;
}

int lb200_lb_le_apply_boundary_conditions(lb200_t * c) {
  CTX_ENTER(c);
  int rc = materialise_propagation(c);
  if (rc != 0) return rc;
  le_lb_bc_async(c);
  c->f_halo_stale = 0;
  CTX_LEAVE_SYNC(c);
}

static int phi_force_async(lb200_t * c, const Lb200SymmDev & sd) {
  const int accumulate = (c->force_state != ZERO_PENDING);
  if (c->le.nplane > 0) {
    // "Must use the flux method for LE planes", src/phi_force.c:91-97
    le_force_ch_async(c, sd, c->g.nl[0], nullptr, 1, 0, accumulate, nullptr);
    if (c->force_state == ZERO_PENDING) c->force_state = INTERIOR_ONLY;
    return 0;
  }
  ProfScope ps(c, LB200_K_FORCE_CH);
  c->launches += c->k->phi_force(c->stream, c->g, sd, accumulate, c->phi, c->grad, c->delsq, c->force);
  if (c->force_state == ZERO_PENDING) c->force_state = INTERIOR_ONLY;
  return 0;
}

// grad_3d_27pt_fluid_d4, src/gradient_3d_27pt_fluid.c:112-134: the same operator applied to delsq
int lb200_phi_grad_compute_d4(lb200_t * c) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  if (c->g.nh < 2) return fail(LB200_ESTATE, "grad_3d_27pt_fluid_d4 needs nhalo >= 2 (reference asserts nhalo - 2 >= 0)");
  if (c->le.nplane > 0) return fail(LB200_ESTATE, "grad_3d_27pt_fluid_d4 with Lees-Edwards planes is not implemented");
  if (c->grad_delsq == nullptr) {
    if (alloc_d(&c->grad_delsq, (size_t) 3*c->g.nsites) != 0 || alloc_d(&c->delsq_delsq, (size_t) c->g.nsites) != 0) return LB200_ECUDA;
  }
  {
    ProfScope ps(c, LB200_K_GRAD);
    c->launches += c->k->grad27(c->stream, c->g, c->g.nh - 2, c->delsq, c->grad_delsq, c->delsq_delsq);
  }
  CTX_LEAVE_SYNC(c);
}

// pth_stress_compute, src/phi_force_stress.c:171-217 (fe_symm_str_v per site)
int lb200_pth_stress_compute(lb200_t * c, const lb200_symm_param_t * sp) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  if (sp == nullptr) return fail(LB200_EINVAL, "null parameters");
  if (c->str == nullptr && alloc_d(&c->str, (size_t) 9*c->g.nsites) != 0) return LB200_ECUDA;
  Lb200SymmDev sd;
  symm_dev(c, sp, &sd);
  {
    ProfScope ps(c, LB200_K_FORCE_CH);
    c->launches += c->k->stress(c->stream, c->g, sd, c->phi, c->grad, c->delsq, c->str);
  }
  CTX_LEAVE_SYNC(c);
}

// pth_force_fluid_driver, src/phi_force_colloid.c:274-301, 315-465
int lb200_pth_force_fluid_driver(lb200_t * c) {
  CTX_ENTER(c);
  if (c->str == nullptr) return fail(LB200_ESTATE, "no stress: call lb200_pth_stress_compute first");
  {
    const int accumulate = (c->force_state != ZERO_PENDING);
    ProfScope ps(c, LB200_K_FORCE_CH);
    c->launches += c->k->force_from_stress(c->stream, c->g, accumulate, c->str, c->force);
    if (c->force_state == ZERO_PENDING) c->force_state = INTERIOR_ONLY;
  }
  CTX_LEAVE_SYNC(c);
}

int lb200_phi_force_calculation(lb200_t * c, const lb200_symm_param_t * sp) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  if (sp == nullptr) return fail(LB200_EINVAL, "null parameters");
  if (sp->force_method < 0 || sp->force_method > 1) return fail(LB200_EINVAL, "fe_force_method %d: 0 (stress_divergence) and 1 (phi_gradmu) are built", sp->force_method);
  Lb200SymmDev sd;
  symm_dev(c, sp, &sd);
  phi_force_async(c, sd);
  CTX_LEAVE_SYNC(c);
}

// ---- cahn_hilliard_options_conserve 2 (src/phi_cahn_hilliard.c:1102-1169) --------------------------------------
// psum layout (doubles): [0, 2B] partials + fluid count | [2B+2, 2B+4) this GPU's (sum, nfluid) | [2B+4, 2B+4+2P) every
// rank's | [2B+4+2P, +2) the global (sum, nfluid)
static int phi_sum_async(lb200_t * c, const double * phi, cudaStream_t st, double ** total) {
  const int B = c->k->psum_blocks, P = c->opt.cart_size;
  if (c->psum == nullptr) {
    if (alloc_d(&c->psum, (size_t) 2*B + 8 + 2*P) != 0) return LB200_ECUDA;
    CUDA_TRY(cudaMemsetAsync(c->psum, 0, ((size_t) 2*B + 8 + 2*P)*sizeof(double), st));
  }
  double * mine = c->psum + 2*B + 2, * all = c->psum + 2*B + 4, * tot = all + 2*P;
  c->launches += c->k->phi_sum(st, c->g, phi, status_ptr(c), c->psum, mine);
  if (P > 1) {
#ifdef LB200_NO_NCCL
    return fail(LB200_ECOMM, "library built without NCCL");
#else
    if (c->nccl == nullptr) return fail(LB200_ECOMM, "cart_size > 1 but no NCCL communicator attached (lb200_attach_nccl)");
    // the one exchange of the path that is not nearest-neighbour: all-gather, then every GPU adds in rank order
    if (ncclAllGather(mine, all, 2, ncclDouble, (ncclComm_t) c->nccl, st) != ncclSuccess) return fail(LB200_ECOMM, "ncclAllGather of the phi sums failed");
    c->launches += c->k->phi_sum_ranks(st, all, P, tot);
#endif
  }
  else {
    tot = mine;
  }
  *total = tot;
  return 0;
}

// phi_ch_subtract_sum_phi_after_forward_step on the array that has just been advanced
static int phi_conserve_subtract_async(lb200_t * c, double * phi, cudaStream_t st) {
  double * tot = nullptr;
  int rc = phi_sum_async(c, phi, st, &tot);
  if (rc != 0) return rc;
  c->launches += c->k->phi_subtract(st, c->g, tot, c->phi_init_sum, status_ptr(c), phi);
  return 0;
}

// cahn_hilliard_stats_time0 (src/cahn_hilliard_stats.c:58-76): the sum of phi over the fluid sites of the whole lattice,
// compensated, kept as the value conserve 2 restores after every step
int lb200_phi_conserve_sum(lb200_t * c, double * sum) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  double * tot = nullptr;
  int rc = phi_sum_async(c, c->phi, c->stream, &tot);
  if (rc != 0) return rc;
  double h[2] = {0.0, 0.0};
  CUDA_TRY(cudaMemcpyAsync(h, tot, 2*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->phi_init_sum = h[0];
  c->phi_init_sum_set = 1;
  if (sum != nullptr) *sum = h[0];
  return 0;
}

int lb200_phi_init_sum_set(lb200_t * c, double phi0) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  c->phi_init_sum = phi0;
  c->phi_init_sum_set = 1;
  return 0;
}

int lb200_phi_cahn_hilliard(lb200_t * c, const lb200_symm_param_t * sp) {
  CTX_ENTER(c);
  if (c->phi == nullptr) return fail(LB200_ESTATE, "no phi in this context");
  if (sp == nullptr) return fail(LB200_EINVAL, "null parameters");
  if (sp->adv_order < 1 || sp->adv_order > 4) return fail(LB200_EINVAL, "advection order %d: device kernels exist for 1-4 (reference src/advection.c:456-480)", sp->adv_order);
  int rc = conserve_prepare(c, sp);
  if (rc != 0) return rc;
  Lb200SymmDev sd;
  symm_dev(c, sp, &sd);
  rc = u_halo_async(c);                   // hydro_u_halo inside phi_cahn_hilliard, src/phi_cahn_hilliard.c:229
  if (rc != 0) return rc;
  // phinew holds phi everywhere (halo included) so that the swap keeps the reference's view
  CUDA_TRY(cudaMemcpyAsync(c->phinew, c->phi, (size_t) c->g.nsites*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (c->le.nplane > 0) {
    le_hydro_async(c);                    // hydro_lees_edwards, src/phi_cahn_hilliard.c:230
    le_force_ch_async(c, sd, c->g.nl[0], nullptr, 0, 1, 0, c->phinew);
  }
  else {
    ProfScope ps(c, LB200_K_FORCE_CH);
    c->launches += c->k->cahn_hilliard(c->stream, c->g, sd, c->phi, c->delsq, c->u, status_ptr(c), c->phinew);
  }
  if (sp->conserve == 2 && (rc = phi_conserve_subtract_async(c, c->phinew, c->stream)) != 0) return rc;
  double * t = c->phi; c->phi = c->phinew; c->phinew = t;
  CTX_LEAVE_SYNC(c);
}

static int collide_async(lb200_t * c, const Lb200CollideDev & cd, const Lb200Geom * gwrap = nullptr) {
  if (c->prop_pending && gwrap == nullptr) {
    int rc = ensure_f_halo(c);
    if (rc != 0) return rc;
  }
  ProfScope ps(c, LB200_K_COLLIDE);
  const double * force = (c->force_state == ZERO_PENDING) ? nullptr : c->force;
  if (c->prop_pending) {
    // lb_propagation(t) fused with lb_collide(t+1): one read and one write of every population
    // (gwrap: lb_halo(t) folded in as well, the periodic images are read from the interior)
    c->launches += c->k->collide(c->stream, gwrap ? *gwrap : c->g, cd, model_ptr(c), c->nvel, 1, c->f, c->fprime,
				 force, status_ptr(c), c->rho, c->u);
    double * t = c->f; c->f = c->fprime; c->fprime = t;
    c->prop_pending = 0;
  }
  else {
    c->launches += c->k->collide(c->stream, c->g, cd, model_ptr(c), c->nvel, 0, c->f, c->f,
				 force, status_ptr(c), c->rho, c->u);
  }
  if (c->u_state == ZERO_PENDING) c->u_state = INTERIOR_ONLY;
  return 0;
}

int lb200_lb_collide(lb200_t * c, const lb200_collide_param_t * cp) {
  CTX_ENTER(c);
  if (cp == nullptr) return fail(LB200_EINVAL, "null parameters");
  if (c->ndist == 2) return fail(LB200_ESTATE, "ndist = 2: the collision needs the free energy, call lb200_lb_collision_binary (reference lb_collide -> lb_collision_binary, src/collision.c:157-159)");
  Lb200CollideDev cd;
  int rc = collide_dev(c, cp, &cd);
  if (rc != 0) return rc;
  collide_async(c, cd);
  CTX_LEAVE_SYNC(c);
}

static int lb_halo_async(lb200_t * c) {
  int rc = materialise_propagation(c);
  if (rc != 0) return rc;
  c->f_halo_stale = 0;
  return halo_field(c, c->f, c->nvel*c->ndist, 1, c->opt.halo_scheme == LB200_HALO_REDUCED, nullptr);
}

int lb200_lb_halo(lb200_t * c) {
  CTX_ENTER(c);
  int rc = lb_halo_async(c);
  if (rc != 0) return rc;
  CTX_LEAVE_SYNC(c);
}

int lb200_lb_propagation(lb200_t * c) {
  CTX_ENTER(c);
  int rc = materialise_propagation(c);   // two propagations in a row: apply the first
  if (rc != 0) return rc;
  c->prop_pending = 1;
  return 0;
}

// ---- symmetric_lb (ndist = 2) -------------------------------------------------------------------------

// phi_lb_to_field, src/phi_lb_coupler.c:39-96.  With a propagation pending the sum runs over the pulled
// populations (the state the reference's phi_lb_to_field sees after lb_propagation).
static int phi_lb_to_field_async(lb200_t * c) {
  if (c->prop_pending) {
    int rc = ensure_f_halo(c);
    if (rc != 0) return rc;
  }
  ProfScope ps(c, LB200_K_GRAD);
  c->launches += c->k->phi_from_g(c->stream, c->g, c->model_d, c->prop_pending, c->f, c->phi);
  return 0;
}

int lb200_phi_lb_to_field(lb200_t * c) {
  CTX_ENTER(c);
  if (c->ndist != 2) return fail(LB200_ESTATE, "phi_lb_to_field needs ndist = 2");
  int rc = phi_lb_to_field_async(c);
  if (rc != 0) return rc;
  CTX_LEAVE_SYNC(c);
}

// phi_lb_from_field, src/phi_lb_coupler.c:98-137
int lb200_phi_lb_from_field(lb200_t * c) {
  CTX_ENTER(c);
  if (c->ndist != 2) return fail(LB200_ESTATE, "phi_lb_from_field needs ndist = 2");
  int rc = materialise_propagation(c);
  if (rc != 0) return rc;
  c->launches += c->k->phi_to_g(c->stream, c->g, c->nvel, c->phi, c->f);
  CTX_LEAVE_SYNC(c);
}

static int collide_binary_async(lb200_t * c, const Lb200CollideDev & cd, const Lb200SymmDev & sd) {
  if (c->prop_pending) {
    int rc = ensure_f_halo(c);
    if (rc != 0) return rc;
  }
  ProfScope ps(c, LB200_K_COLLIDE);
  const double * force = (c->force_state == ZERO_PENDING) ? nullptr : c->force;
  if (c->prop_pending) {
    c->launches += c->k->collide_binary(c->stream, c->g, cd, sd, c->model_d, c->unrolled19, 1, c->f, c->fprime,
					force, c->phi, c->grad, c->delsq, c->u);
    double * t = c->f; c->f = c->fprime; c->fprime = t;
    c->prop_pending = 0;
  }
  else {
    c->launches += c->k->collide_binary(c->stream, c->g, cd, sd, c->model_d, c->unrolled19, 0, c->f, c->f,
					force, c->phi, c->grad, c->delsq, c->u);
  }
  if (c->u_state == ZERO_PENDING) c->u_state = INTERIOR_ONLY;
  return 0;
}

// lb_collide with ndist == 2 -> lb_collision_binary, src/collision.c:157-159, 604-1013
int lb200_lb_collision_binary(lb200_t * c, const lb200_collide_param_t * cp, const lb200_symm_param_t * sp) {
  CTX_ENTER(c);
  if (c->ndist != 2) return fail(LB200_ESTATE, "lb_collision_binary needs ndist = 2");
  if (cp == nullptr || sp == nullptr) return fail(LB200_EINVAL, "null parameters");
  Lb200CollideDev cd;
  Lb200SymmDev sd;
  int rc = collide_dev(c, cp, &cd);
  if (rc != 0) return rc;
  symm_dev(c, sp, &sd);
  rc = collide_binary_async(c, cd, sd);
  if (rc != 0) return rc;
  CTX_LEAVE_SYNC(c);
}

// whole symmetric_lb time steps, reference order (src/ludwig.c:528-860 with ndist == 2): hydro_f_zero;
// phi_lb_to_field; field_halo(phi); field_grad_compute; hydro_u_zero; lb_collide (binary); lb_halo;
// lb_propagation.  The propagation is fused into the next phi_lb_to_field and collision (pull).
static int step_lb2(lb200_t * c, const Lb200CollideDev & cd, const Lb200SymmDev & sd, int nsteps) {
  int rc = 0;
  for (int n = 0; n < nsteps; n++) {
    c->t_current += 1;                                                   // physics_control_next_step
    c->force_state = ZERO_PENDING;
    rc = phi_lb_to_field_async(c);
    if (rc != 0) return rc;
    rc = halo_field(c, c->phi, 1, c->g.nh, 0, nullptr);
    if (rc != 0) return rc;
    le_field_async(c, c->phi);                                           // field_grad_compute -> field_leesedwards (planes only)
    {
      ProfScope ps(c, LB200_K_GRAD);
      c->launches += c->k->grad27(c->stream, c->g, c->g.nh - 1, c->phi, c->grad, c->delsq);
    }
    le_grad_async(c);                                                    // the planes next to a Lees-Edwards plane, and the buffers
    c->u_state = ZERO_PENDING;
    rc = collide_binary_async(c, cd, sd);
    if (rc != 0) return rc;
    le_lb_bc_async(c);                                                   // lb_data_apply_le_boundary_conditions, both distributions
    rc = halo_field(c, c->f, c->nvel*c->ndist, 1, c->opt.halo_scheme == LB200_HALO_REDUCED, nullptr);
    if (rc != 0) return rc;
    c->prop_pending = 1;
  }
  c->phi_halo_valid = 0;
  c->u_halo_valid = 0;
  c->wrap_x_valid = 0;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- peer-store exchange: set-up (collective over the NCCL communicator, first lb200_step only) --------------
// Every rank exports its two distribution buffers, two phi buffers, two velocity buffers and its flag words
// with cudaIpcGetMemHandle; the handles are all-gathered and each rank maps its two neighbours' arrays.
// From then on the kernels of lb200_step store their boundary planes directly in the neighbour's halo planes
// (Lb200Geom::peer_*) and the only other traffic is one 4-byte flag per kernel and neighbour.

static int peer_setup(lb200_t * c) {
  if (c->peer_state != 0) return 0;
  c->peer_state = -1;
#ifdef LB200_NO_NCCL
  return 0;
#else
  if (!c->g.remote_x || c->nccl == nullptr || c->opt.cart_size < 2) return 0;
  ncclComm_t comm = (ncclComm_t) c->nccl;
  const int P = c->opt.cart_size, rank = c->opt.cart_rank;
  const int left = (rank - 1 + P) % P, right = (rank + 1) % P;
  const size_t ns = (size_t) c->g.nsites;
  enum {NH = 7};
  int ok = 1;

  if (alloc_d(&c->u2, 3*ns) != 0) return LB200_ECUDA;
  c->u_alloc[0] = c->u; c->u_alloc[1] = c->u2;
  // (a whole 2 MiB granule: an exported allocation exposes the granule it lives in)
  CUDA_TRY(cudaMalloc((void **) &c->flags, (size_t) 2 << 20));
  CUDA_TRY(cudaMemset(c->flags, 0, FLAG_COUNT*sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc((void **) &c->spin_err, sizeof(int)));
  CUDA_TRY(cudaMemset(c->spin_err, 0, sizeof(int)));
  CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));

  void * mine[NH] = {c->f_alloc[0], c->f_alloc[1], c->phi_alloc[0], c->phi_alloc[1], c->u_alloc[0], c->u_alloc[1], c->flags};
  std::vector<cudaIpcMemHandle_t> hs(NH), all((size_t) NH*P);
  for (int i = 0; i < NH; i++) {
    if (mine[i] == nullptr) { memset(&hs[i], 0, sizeof(hs[i])); continue; }      // single fluid: no phi
    if (cudaIpcGetMemHandle(&hs[i], mine[i]) != cudaSuccess) { cudaGetLastError(); ok = 0; memset(&hs[i], 0, sizeof(hs[i])); }
  }
  char * dsend = nullptr, * drecv = nullptr;
  const size_t nb = NH*sizeof(cudaIpcMemHandle_t);
  CUDA_TRY(cudaMalloc((void **) &dsend, nb + sizeof(int)));
  CUDA_TRY(cudaMalloc((void **) &drecv, (nb + sizeof(int))*P));
  CUDA_TRY(cudaMemcpyAsync(dsend, hs.data(), nb, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(dsend + nb, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (ncclAllGather(dsend, drecv, nb + sizeof(int), ncclChar, comm, c->stream) != ncclSuccess) {
    return fail(LB200_ECOMM, "ncclAllGather of the IPC handles failed");
  }
  std::vector<char> raw((nb + sizeof(int))*P);
  CUDA_TRY(cudaMemcpyAsync(raw.data(), drecv, raw.size(), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < P; r++) {
    int okr = 0;
    memcpy(&all[(size_t) r*NH], raw.data() + (size_t) r*(nb + sizeof(int)), nb);
    memcpy(&okr, raw.data() + (size_t) r*(nb + sizeof(int)) + nb, sizeof(int));
    ok = ok && okr;
  }

  // map the neighbours (the same rank on both sides when P == 2: map once)
  auto map_rank = [&](int r, PeerLink * L) -> int {
    void * ptr[NH];
    for (int i = 0; i < NH; i++) {
      ptr[i] = nullptr;
      if (mine[i] == nullptr) continue;
      if (cudaIpcOpenMemHandle(&ptr[i], all[(size_t) r*NH + i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
	cudaGetLastError();
	return 0;
      }
      c->mapped[c->nmapped++] = ptr[i];
    }
    L->f[0] = (double *) ptr[0]; L->f[1] = (double *) ptr[1];
    L->phi[0] = (double *) ptr[2]; L->phi[1] = (double *) ptr[3];
    L->u[0] = (double *) ptr[4]; L->u[1] = (double *) ptr[5];
    L->flags = (unsigned int *) ptr[6];
    return 1;
  };
  int mapped_ok = ok;
  if (mapped_ok) mapped_ok = map_rank(left, &c->lo);
  if (mapped_ok) { if (right == left) c->hi = c->lo; else mapped_ok = map_rank(right, &c->hi); }

  // every rank must reach the same verdict
  {
    int * dflag = (int *) dsend;
    CUDA_TRY(cudaMemcpyAsync(dflag, &mapped_ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (ncclAllReduce(dflag, dflag, 1, ncclInt, ncclMin, comm, c->stream) != ncclSuccess) {
      return fail(LB200_ECOMM, "ncclAllReduce of the peer set-up verdict failed");
    }
    CUDA_TRY(cudaMemcpyAsync(&mapped_ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  cudaFree(dsend); cudaFree(drecv);

  {
    // stream memory operations for the consumer side of the flags (the polling kernel is the fallback)
    void * fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess
	&& q == cudaDriverEntryPointSuccess && getenv("LB200_SPIN_WAIT") == nullptr) c->wait_value32 = fn;
    else cudaGetLastError();
  }
  c->peer_state = mapped_ok ? 1 : -1;
  return 0;
#endif
}

int lb200_exchange_mode(const lb200_t * c) {
  if (c == nullptr || !c->g.remote_x) return 0;
  return (c->peer_state == 1 && c->knob_peer) ? 2 : 1;
}

typedef int (* wait_value32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);

// wait until the neighbours on both sides have signalled `value` on the flag pair (lo, hi)
static int flags_wait(lb200_t * c, cudaStream_t st, int ilo, unsigned int value) {
  for (int i = ilo; i <= ilo + 1; i++) {
    if (c->wait_value32 != nullptr) {
      // CU_STREAM_WAIT_VALUE_GEQ = 0: (int32_t)(*addr - value) >= 0
      int r = ((wait_value32_fn) c->wait_value32)(st, (unsigned long long) (uintptr_t) (c->flags + i), value, 0u);
      if (r != 0) return fail(LB200_ECUDA, "cuStreamWaitValue32 failed (%d)", r);
    }
    else {
      c->launches += c->k->spin_wait(st, c->flags + i, value, 20000, c->spin_err);
      c->spin_used = 1;
    }
  }
  return 0;
}

// ---- halo-free whole time steps ---------------------------------------------------------------------
// On a periodic lattice every halo site is the image of an interior site of the same GPU (y, z; x too
// on one GPU), so the kernels of lb200_step read those images straight from the interior (Lb200Geom::wrap)
// and the three halo sweeps of the reference step (field_halo(phi), hydro_u_halo, lb_halo: 26 pack
// kernels + messages + unpack kernels each) disappear.  With x-slabs on several GPUs only what the next
// kernels read crosses NVLink: 2 planes of phi, 1 plane of u_x and the populations with c_x = +-1
// (5 of 19 for D3Q19) per direction, received directly in the halo planes.  Results on the interior are
// those of the reference order; the halo sites of f are brought up to date on demand (f_halo_stale).

struct XMsg { const double * send_hi; const double * send_lo; double * recv_lo; double * recv_hi; size_t count; };

static int nccl_exchange(lb200_t * c, cudaStream_t st, const XMsg * m, int nm) {
#ifdef LB200_NO_NCCL
  return fail(LB200_ECOMM, "library built without NCCL");
#else
  if (c->nccl == nullptr) return fail(LB200_ECOMM, "cart_size > 1 but no NCCL communicator attached (lb200_attach_nccl)");
  ncclComm_t comm = (ncclComm_t) c->nccl;
  const int left = (c->opt.cart_rank - 1 + c->opt.cart_size) % c->opt.cart_size;
  const int right = (c->opt.cart_rank + 1) % c->opt.cart_size;
  ncclResult_t r = ncclGroupStart();
  for (int i = 0; i < nm && r == ncclSuccess; i++) {
    r = ncclSend(m[i].send_hi, m[i].count, ncclDouble, right, comm, st);
    if (r == ncclSuccess) r = ncclSend(m[i].send_lo, m[i].count, ncclDouble, left, comm, st);
    if (r == ncclSuccess) r = ncclRecv(m[i].recv_lo, m[i].count, ncclDouble, left, comm, st);
    if (r == ncclSuccess) r = ncclRecv(m[i].recv_hi, m[i].count, ncclDouble, right, comm, st);
  }
  ncclResult_t r2 = ncclGroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) {
    return fail(LB200_ECOMM, "NCCL plane exchange: %s", ncclGetErrorString(r != ncclSuccess ? r : r2));
  }
  return 0;
#endif
}

// planes of phi (depth nhalo) straight from the array into the neighbour's halo planes (contiguous: no staging)
static int wrap_exchange_phi(lb200_t * c, cudaStream_t st) {
  const Lb200Geom & g = c->g;
  const size_t xs = (size_t) g.xs;
  const int d = g.nh;
  XMsg m;
  m.send_hi = c->phi + (size_t) (g.nl[0] - d + g.nh)*xs;      // planes N-d+1 .. N
  m.send_lo = c->phi + (size_t) g.nh*xs;                      // planes 1 .. d
  m.recv_lo = c->phi + (size_t) (g.nh - d)*xs;                // planes 1-d .. 0
  m.recv_hi = c->phi + (size_t) (g.nl[0] + g.nh)*xs;          // planes N+1 .. N+d
  m.count = (size_t) d*xs;
  ProfScope ps(c, LB200_K_HALO, st);
  return nccl_exchange(c, st, &m, 1);
}

// after a collision: u_x of planes N / 1 (the only velocity the next phi sector reads across the slab
// boundary), straight into the neighbour's halo planes -- first, because the phi sector waits for it
static int wrap_exchange_ux(lb200_t * c, cudaStream_t st) {
  const Lb200Geom & g = c->g;
  const size_t xs = (size_t) g.xs;
  XMsg m;
  m.send_hi = c->u + (size_t) (g.nl[0] + g.nh - 1)*xs;   // u_x, plane N
  m.send_lo = c->u + (size_t) g.nh*xs;                   // u_x, plane 1
  m.recv_lo = c->u + (size_t) (g.nh - 1)*xs;             // plane 0
  m.recv_hi = c->u + (size_t) (g.nl[0] + g.nh)*xs;       // plane N+1
  m.count = xs;
  ProfScope ps(c, LB200_K_HALO, st);
  return nccl_exchange(c, st, &m, 1);
}

// ... and the populations moving in +x / -x of planes N / 1 (overlaps the phi sector); what travels is
// lb200_step_plan(LB200_STEP_F): runs of consecutive populations are gathered / scattered with one strided copy
static int wrap_exchange_f(lb200_t * c, cudaStream_t st) {
  const size_t ns = (size_t) c->g.nsites;
  lb200_step_plan_t pl;
  int rc = lb200_step_plan(&c->opt, LB200_STEP_F, &pl);
  if (rc != 0) return rc;
  if (pl.ncomp_up != pl.ncomp_down) return fail(LB200_ESTATE, "velocity set not symmetric in x");
  const size_t chunk = (size_t) pl.chunk;
  ProfScope ps(c, LB200_K_HALO, st);
  double * shi = c->shi, * slo = c->slo, * rlo = c->xlo, * rhi = c->xhi;

  // dir 0: up (my plane N -> staging -> the high neighbour's plane 0), dir 1: down
  auto strided = [&](int dir, bool pack) -> int {
    const int * comp = dir == 0 ? pl.comp_up : pl.comp_down;
    const int ncomp = dir == 0 ? pl.ncomp_up : pl.ncomp_down;
    for (int a = 0; a < ncomp; ) {
      int b = a;
      while (b + 1 < ncomp && comp[b + 1] == comp[b] + 1) b++;
      const int nrun = b - a + 1;
      if (pack) {
	double * dst = (dir == 0 ? shi : slo) + (size_t) a*chunk;
	const double * src = c->f + (size_t) comp[a]*ns + (size_t) (dir == 0 ? pl.src_up : pl.src_down);
	CUDA_TRY(cudaMemcpy2DAsync(dst, chunk*sizeof(double), src, ns*sizeof(double), chunk*sizeof(double), nrun,
				   cudaMemcpyDeviceToDevice, st));
      }
      else {
	// what arrives from the LOW neighbour are its `up` components -> my low halo plane; from the high one its `down`
	const double * src = (dir == 0 ? rlo : rhi) + (size_t) a*chunk;
	double * dst = c->f + (size_t) comp[a]*ns + (size_t) (dir == 0 ? pl.dst_up : pl.dst_down);
	CUDA_TRY(cudaMemcpy2DAsync(dst, ns*sizeof(double), src, chunk*sizeof(double), chunk*sizeof(double), nrun,
				   cudaMemcpyDeviceToDevice, st));
      }
      a = b + 1;
    }
    return 0;
  };
  if ((rc = strided(0, true)) != 0 || (rc = strided(1, true)) != 0) return rc;
  XMsg m;
  m.send_hi = shi; m.send_lo = slo; m.recv_lo = rlo; m.recv_hi = rhi; m.count = (size_t) pl.ncomp_up*chunk;
  rc = nccl_exchange(c, st, &m, 1);
  if (rc != 0) return rc;
  if ((rc = strided(0, false)) != 0 || (rc = strided(1, false)) != 0) return rc;
  return 0;
}

static int idx2(const double * ptr, double * const alloc[2]) { return (ptr == alloc[0]) ? 0 : 1; }

// make the boundary data of one kind available to the next kernel on S
static int src_wait(lb200_t * c, cudaStream_t S, int src, cudaEvent_t ev, int flag_lo, unsigned int value) {
  if (src == SRC_EVENT) CUDA_TRY(cudaStreamWaitEvent(S, ev, 0));
  else if (src == SRC_FLAG) return flags_wait(c, S, flag_lo, value);
  return 0;
}

static int step_wrap(lb200_t * c, const Lb200CollideDev & cd, const Lb200SymmDev * sd, int nsteps) {
  const int binary = (sd != nullptr);
  const int remote = c->g.remote_x;
  Lb200Geom gw = c->g;
  gw.wrap[0] = !remote; gw.wrap[1] = 1; gw.wrap[2] = 1;
  cudaStream_t S = c->stream, C = c->profile ? c->stream : c->comm;
  int rc = 0;

  if (remote) {
    rc = peer_setup(c);                  // collective, does something the first time only
    if (rc != 0) return rc;
  }
  // peer stores from inside the kernels + flags instead of NCCL messages
  const bool peer = remote && c->peer_state == 1 && c->knob_peer;

  if (remote) {
    CUDA_TRY(cudaEventRecord(c->ev_main, S));
    CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
    if (!c->wrap_x_valid) {
      // The state did not come out of a previous lb200_step: one round of messages brings the planes the
      // first kernels read (and orders this step after whatever the neighbours were doing before).
      if (binary) {
	if (c->u_state == ZERO_PENDING) materialise_zero(c, c->u, &c->u_state);
	CUDA_TRY(cudaEventRecord(c->ev_main, S));
	CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
	rc = wrap_exchange_phi(c, C);
	if (rc != 0) return rc;
      }
      rc = wrap_exchange_ux(c, C);
      if (rc != 0) return rc;
      CUDA_TRY(cudaEventRecord(c->ev_phi, C));
      CUDA_TRY(cudaEventRecord(c->ev_u, C));
      if (c->prop_pending) {
	rc = wrap_exchange_f(c, C);
	if (rc != 0) return rc;
      }
      CUDA_TRY(cudaEventRecord(c->ev_f, C));
      c->phi_src = c->u_src = c->f_src = SRC_EVENT;
    }
  }
  if (c->u_state == ZERO_PENDING && binary) materialise_zero(c, c->u, &c->u_state);

  const bool le = (c->le.nplane > 0);
  // FP32 storage of the distributions for the steps of this call (LB200_KNOB_F32; single GPU, D3Q19, no planes)
  const bool f32 = c->knob_f32 && !remote && !le && c->nvel == 19 && c->unrolled19 && c->ndist == 1 && c->map_all_fluid;
  int f32_cur = -1;                                                      // >= 0: the state lives in c->f32[f32_cur]
  if (f32 && c->f32[0] == nullptr) {
    CUDA_TRY(cudaMalloc((void **) &c->f32[0], (size_t) 19*c->g.nsites*sizeof(float)));
    CUDA_TRY(cudaMalloc((void **) &c->f32[1], (size_t) 19*c->g.nsites*sizeof(float)));
  }
  // one kernel per step (LB200_KNOB_FUSED): binary fluid, D3Q19 with the coded matrices, all-fluid, no planes
  // (fast arithmetic mode; the TMA boxes of the populations need 16-byte aligned rows: even extents in z)
  // (with Lees-Edwards planes: the sweep runs over the whole lattice as if there were none, then the 2*nhalo x-planes per
  // plane whose stencils cross it are produced again by the patch kernels through the buffer planes)
  bool le_uniform = (c->le.xblock >= 2*c->g.nh + 4);                     // one plane every xblock x-planes (the sweep's store mask)
  for (int p = 1; p < c->le.nplane; p++) le_uniform = le_uniform && (c->le.loc[p] == c->le.loc[0] + p*c->le.xblock);
  const bool fuse_ok = binary && c->knob_fused && (!le || (c->knob_fused_le && le_uniform)) && !f32 && c->nvel == 19 && c->unrolled19 && c->ndist == 1
    && c->map_all_fluid && c->knob_pipe < 2 && sd->order <= 3 && sd->csum == nullptr
    && c->g.nh == 2 && (c->g.nall[2] & 1) == 0 && (c->g.nsites & 1) == 0
    && c->g.nl[1] >= 2 && c->g.nl[2] >= 2;
  if (fuse_ok && c->u_alloc[1] == nullptr) {
    if (alloc_d(&c->u2, (size_t) 3*c->g.nsites) != 0) return LB200_ECUDA;
    c->u_alloc[0] = c->u; c->u_alloc[1] = c->u2;
  }
  for (int n = 0; n < nsteps; n++) {
    c->t_current += 1;                                                   // physics_control_next_step
    c->force_state = ZERO_PENDING;                                       // hydro_f_zero
    // hydro->rho, grad and delsq are read by nobody inside the step (the Lees-Edwards patches excepted, which read grad
    // and delsq): only the last step of the call stores them
    gw.skip_diag = (c->knob_lazy_diag && !le && n < nsteps - 1) ? 1 : 0;
    if (fuse_ok && c->knob_fused && c->prop_pending) {
      // The whole step in one sweep (LB200_KNOB_FUSED): phi sector + pull-stream + collision of the same plane, the
      // force stays in registers.  (The first step after an upload of the distributions collides in place -- no
      // propagation is pending -- and takes the two-kernel route below.)
      if (le) gw.skip_diag = (c->knob_lazy_diag && n < nsteps - 1) ? 1 : 0;  // (the patches form the gradients they read themselves)
      if (remote) {
	// (after a one-kernel step the neighbours set the phi flags and then the f / u flags from one thread, with a fence in
	// between: the second pair implies the first)
	const bool joint = c->joint_signal && c->phi_src == SRC_FLAG && c->u_src == SRC_FLAG && c->f_src == SRC_FLAG;
	if (!joint) rc = src_wait(c, S, c->phi_src, c->ev_phi, FLAG_PS_LO, c->n_ps);
	if (rc == 0) rc = src_wait(c, S, c->u_src, c->ev_u, FLAG_COL_LO, c->n_col);
	if (rc == 0 && !(c->f_src == SRC_FLAG && c->u_src == SRC_FLAG)) rc = src_wait(c, S, c->f_src, c->ev_f, FLAG_COL_LO, c->n_col);
	if (rc != 0) return rc;
      }
      // the TMA boxes of the populations read the y / z halos (and the rims of the x halo planes): valid after a
      // one-kernel step, else one lb_halo brings them up to date
      if (!c->fused_ready) {
	rc = ensure_f_halo(c);
	if (rc == 0) rc = halo_field(c, c->phi, 1, c->g.nh, 0, S);
	if (rc != 0) return rc;
	if (c->u_state != ARRAY_CLEAN) materialise_zero(c, c->u, &c->u_state);
	rc = halo_field(c, c->u, 3, c->g.nh, 0, S);
	if (rc != 0) return rc;
	c->f_halo_stale = 1;                 // (still true for the reference's view: only what the pull reads is kept up to date)
      }
      // the phi sector of other CTAs (and of the neighbour GPUs) reads u(t-1) while this kernel writes u(t)
      double * u_out = (c->u == c->u_alloc[0]) ? c->u_alloc[1] : c->u_alloc[0];
      // Lees-Edwards planes.  The sweep runs over the whole lattice as if there were none but stores nothing for the
      // 2*nhalo x-planes per plane whose stencils cross it; the PATCH CHAIN produces those through the buffer planes:
      // field_leesedwards + hydro_lees_edwards (buffer planes), gradients of planes loc-2 .. loc+3 and of the buffers,
      // flux-form force with the plane correction + Cahn-Hilliard with the averaged plane fluxes (-> force, phinew),
      // pull-stream + collision of the planes with that force, lb_data_apply_le_boundary_conditions, y / z images.
      // It reads only what the previous step left (phi, u, f) and writes only what the sweep leaves out, so it may run
      // NEXT TO the sweep on its own high-priority stream (LB200_FUSED_LE=2, intermediate steps; on the last step of a
      // call the sweep also stores grad / delsq everywhere, which the chain must overwrite next to the planes).  Measured
      // at 256^3, one plane: 1.321 ms next to the sweep, 1.332 after it; 512 x 256 x 256, 8 planes: 3.01 / 2.97 ms -- the
      // sweep's CTAs fill the SMs (228 kB of shared memory each), so the chain's grids run on SMs taken from the sweep,
      // not beside it.  Default: after the sweep.
      const bool le_conc = le && c->le_stream != nullptr && !c->profile && gw.skip_diag && c->knob_fused_le >= 2;
      auto le_chain = [&](cudaStream_t T) {
	Lb200Geom gl = gw;
	gl.peer_phi_lo = gl.peer_phi_hi = gl.peer_f_lo = gl.peer_f_hi = gl.peer_u_lo = gl.peer_u_hi = nullptr;
	{
	  Lb200LeInterp ipc, ipl;
	  le_interp_cubic(c, &ipc);
	  le_interp_linear(c, &ipl);
	  ProfScope ps(c, LB200_K_LE, T);
	  c->launches += c->k->le_interp_both(T, gl, c->le, ipc, ipl, c->g.nh, c->phi, c->u);
	}
	{
	  ProfScope ps(c, LB200_K_LE, T);
	  c->launches += c->k->le_grad_planes(T, gl, 0, c->le_ntrip_fused, c->le_trip_fused, c->phi, c->grad, c->delsq);
	}
	le_force_ch_async(c, *sd, c->le_nxlist, c->le_xlist, 1, 1, 0, c->phinew, &gl, T);
	{
	  ProfScope ps(c, LB200_K_LE, T);
	  for (int p = 0; p < c->le.nplane; p++) {
	    Lb200Geom gc = gl;
	    gc.xoff = c->le.loc[p] - c->g.nh; gc.xcnt = 2*c->g.nh;
	    c->launches += c->k->collide(T, gc, cd, model_ptr(c), c->nvel, 1, c->f, c->fprime, c->force, status_ptr(c), c->rho, u_out);
	  }
	}
	le_lb_bc_async(c, c->fprime, T);
	ProfScope ps(c, LB200_K_LE, T);
	c->launches += c->k->le_yz_images(T, c->g, c->le_nxlist, c->le_xlist, c->fprime, c->nvel, 1, c->phinew, 1, c->g.nh, u_out, 3, 1);
      };
      if (le_conc) {
	CUDA_TRY(cudaEventRecord(c->ev_le_a, S));
	CUDA_TRY(cudaStreamWaitEvent(c->le_stream, c->ev_le_a, 0));
	le_chain(c->le_stream);
	CUDA_TRY(cudaEventRecord(c->ev_le_b, c->le_stream));
      }
      gw.peer_phi_lo = peer ? c->lo.phi[idx2(c->phinew, c->phi_alloc)] : nullptr;
      gw.peer_phi_hi = peer ? c->hi.phi[idx2(c->phinew, c->phi_alloc)] : nullptr;
      gw.peer_f_lo = peer ? c->lo.f[idx2(c->fprime, c->f_alloc)] : nullptr;
      gw.peer_f_hi = peer ? c->hi.f[idx2(c->fprime, c->f_alloc)] : nullptr;
      gw.peer_u_lo = peer ? c->lo.u[idx2(u_out, c->u_alloc)] : nullptr;
      gw.peer_u_hi = peer ? c->hi.u[idx2(u_out, c->u_alloc)] : nullptr;
      int launched = 0;
      {
	ProfScope ps(c, LB200_K_STEP_FUSED);
	launched = c->k->step_fused(S, gw, *sd, cd, c->phi, c->u, c->f, c->fprime, c->grad, c->delsq, c->force,
				    c->phinew, c->rho, u_out, le ? c->le.loc[0] - c->g.nh + 1 : 0, le ? c->le.xblock : 0);
      }
      if (le_conc) CUDA_TRY(cudaStreamWaitEvent(S, c->ev_le_b, 0));
      if (le_conc && launched <= 0) return fail(LB200_ECUDA, "one-kernel step: launch refused after the Lees-Edwards patch chain was issued");
      if (launched > 0) {
	c->launches += launched;
	c->force_state = gw.skip_diag ? ZERO_PENDING : INTERIOR_ONLY;      // the force array is written by the last step only
	if (le && !le_conc) le_chain(S);
	{ double * t = c->phi; c->phi = c->phinew; c->phinew = t; }
	{ double * t = c->f; c->f = c->fprime; c->fprime = t; }
	c->u = u_out;
	c->u_state = INTERIOR_ONLY;
	c->prop_pending = 1;
	c->f_halo_stale = 1; c->fused_ready = 1;
	if (peer) {
	  c->n_ps++; c->n_col++;
	  c->launches += c->k->signal2(S, c->hi.flags + FLAG_PS_LO, c->lo.flags + FLAG_PS_HI, c->n_ps,
				       c->hi.flags + FLAG_COL_LO, c->lo.flags + FLAG_COL_HI, c->n_col);
	  c->phi_src = c->f_src = c->u_src = SRC_FLAG;
	  c->joint_signal = 1;
	}
	else if (remote) {
	  CUDA_TRY(cudaEventRecord(c->ev_main, S));
	  CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
	  rc = wrap_exchange_phi(c, C);
	  if (rc != 0) return rc;
	  CUDA_TRY(cudaEventRecord(c->ev_phi, C));
	  rc = wrap_exchange_ux(c, C);
	  if (rc != 0) return rc;
	  CUDA_TRY(cudaEventRecord(c->ev_u, C));
	  rc = wrap_exchange_f(c, C);
	  if (rc != 0) return rc;
	  CUDA_TRY(cudaEventRecord(c->ev_f, C));
	  c->phi_src = c->f_src = c->u_src = SRC_EVENT;
	}
	continue;
      }
      // no such kernel after all (no tensor-map entry point in this driver): the two-kernel step from now on
      gw.peer_f_lo = gw.peer_f_hi = gw.peer_u_lo = gw.peer_u_hi = nullptr;
      c->knob_fused = 0;
      if (le) gw.skip_diag = 0;
    }
    if (binary) {
      if (remote) {
	rc = src_wait(c, S, c->phi_src, c->ev_phi, FLAG_PS_LO, c->n_ps);
	if (rc == 0) rc = src_wait(c, S, c->u_src, c->ev_u, FLAG_COL_LO, c->n_col);
	if (rc != 0) return rc;
      }
      gw.peer_phi_lo = peer ? c->lo.phi[idx2(c->phinew, c->phi_alloc)] : nullptr;
      gw.peer_phi_hi = peer ? c->hi.phi[idx2(c->phinew, c->phi_alloc)] : nullptr;
      if (le) {
	// field_leesedwards, hydro_lees_edwards: the buffer planes, from interior columns (no halo needed)
	le_field_async(c, c->phi, &gw);
	le_hydro_async(c, &gw);
      }
      {
	// field_halo(phi) + field_grad_compute + phi_force_calculation + phi_cahn_hilliard (hydro_u_halo inside)
	ProfScope ps(c, LB200_K_PHI_SECTOR);
	c->launches += c->k->phi_sector(S, gw, *sd, c->phi, c->u, c->grad, c->delsq, c->force, c->phinew);
      }
      if (le) {
	// the 2*nhalo planes per Lees-Edwards plane whose stencils cross it, redone through the buffer planes
	Lb200Geom gl = gw;                 // no peer stores from the patch kernels (they never touch boundary planes)
	gl.peer_phi_lo = gl.peer_phi_hi = nullptr;
	le_grad_async(c, &gl);
	le_force_ch_async(c, *sd, c->le_nxlist, c->le_xlist, 1, 1, 0, c->phinew, &gl);
      }
      c->force_state = INTERIOR_ONLY;
      double * t = c->phi; c->phi = c->phinew; c->phinew = t;
      if (peer) {
	// the new phi planes are already in the neighbours' halo planes: tell them
	c->n_ps++;
	c->launches += c->k->signal(S, c->hi.flags + FLAG_PS_LO, c->lo.flags + FLAG_PS_HI, c->n_ps);
	c->joint_signal = 0;
	c->phi_src = SRC_FLAG;
      }
      else if (remote) {
	CUDA_TRY(cudaEventRecord(c->ev_main, S));
	CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
	rc = wrap_exchange_phi(c, C);                                    // overlaps the collision
	if (rc != 0) return rc;
	CUDA_TRY(cudaEventRecord(c->ev_phi, C));
	c->phi_src = SRC_EVENT;
      }
    }
    c->u_state = ZERO_PENDING;                                           // hydro_u_zero
    if (remote) {
      rc = src_wait(c, S, c->f_src, c->ev_f, FLAG_COL_LO, c->n_col);
      if (rc != 0) return rc;
    }
    {
      // lb_halo + lb_propagation + lb_collide (in place the first time after an lb_memcpy / explicit propagation)
      double * u_out = c->u;
      if (peer) {
	// the neighbours' phi sector may still be reading the u_x plane it got last step: write the other buffer
	u_out = (c->u == c->u_alloc[0]) ? c->u_alloc[1] : c->u_alloc[0];
	const double * fdst = c->prop_pending ? c->fprime : c->f;
	gw.peer_f_lo = c->lo.f[idx2(fdst, c->f_alloc)];
	gw.peer_f_hi = c->hi.f[idx2(fdst, c->f_alloc)];
	gw.peer_u_lo = binary ? c->lo.u[idx2(u_out, c->u_alloc)] : nullptr;
	gw.peer_u_hi = binary ? c->hi.u[idx2(u_out, c->u_alloc)] : nullptr;
      }
      if (f32 && c->prop_pending && f32_cur < 0) {
	c->launches += c->k->f_convert(S, c->g, 1, c->f, c->f32[0]);
	f32_cur = 0;
      }
      ProfScope ps(c, LB200_K_COLLIDE);
      const double * force = (c->force_state == ZERO_PENDING) ? nullptr : c->force;
      if (f32 && c->prop_pending) {
	c->launches += c->k->collide_f32(S, gw, cd, c->f32[f32_cur], c->f32[1 - f32_cur], force, c->rho, u_out);
	f32_cur = 1 - f32_cur;
      }
      else if (c->prop_pending) {
	c->launches += c->k->collide(S, gw, cd, model_ptr(c), c->nvel, 1, c->f, c->fprime, force, status_ptr(c), c->rho, u_out);
	double * t = c->f; c->f = c->fprime; c->fprime = t;
	c->prop_pending = 0;
      }
      else {
	Lb200Geom gi = c->g;                                             // in place: no pull, but the peer stores
	gi.peer_f_lo = gw.peer_f_lo; gi.peer_f_hi = gw.peer_f_hi; gi.peer_u_lo = gw.peer_u_lo; gi.peer_u_hi = gw.peer_u_hi;
	c->launches += c->k->collide(S, gi, cd, model_ptr(c), c->nvel, 0, c->f, c->f, force, status_ptr(c), c->rho, u_out);
      }
      c->u = u_out;
      c->u_state = INTERIOR_ONLY;
    }
    if (le) le_lb_bc_async(c);                                           // lb_data_apply_le_boundary_conditions
    c->prop_pending = 1;                                                 // lb_halo; lb_propagation (lazy)
    c->f_halo_stale = 1; c->fused_ready = 0;
    if (peer) {
      c->n_col++;
      c->launches += c->k->signal(S, c->hi.flags + FLAG_COL_LO, c->lo.flags + FLAG_COL_HI, c->n_col);
      c->joint_signal = 0;
      c->f_src = c->u_src = SRC_FLAG;
    }
    else if (remote) {
      CUDA_TRY(cudaEventRecord(c->ev_main, S));
      CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
      if (binary) {
	rc = wrap_exchange_ux(c, C);                                     // the next phi sector waits for this one only
	if (rc != 0) return rc;
	CUDA_TRY(cudaEventRecord(c->ev_u, C));
      }
      rc = wrap_exchange_f(c, C);                                        // overlaps the next phi sector
      if (rc != 0) return rc;
      CUDA_TRY(cudaEventRecord(c->ev_f, C));
      c->f_src = c->u_src = SRC_EVENT;
    }
  }
  if (f32_cur >= 0) c->launches += c->k->f_convert(S, c->g, 0, c->f, c->f32[f32_cur]);   // back to the FP64 array
  c->phi_halo_valid = 0;
  c->u_halo_valid = 0;
  c->wrap_x_valid = 1;
  if (remote) {
    // rejoin: whatever is still in flight on the comm stream is ordered before what follows on the main stream
    if (c->f_src == SRC_EVENT) CUDA_TRY(cudaStreamWaitEvent(S, c->ev_f, 0));
    if (c->u_src == SRC_EVENT) CUDA_TRY(cudaStreamWaitEvent(S, c->ev_u, 0));
    if (binary && c->phi_src == SRC_EVENT) CUDA_TRY(cudaStreamWaitEvent(S, c->ev_phi, 0));
    // peer stores: the neighbours' last phi sector / collision write into THIS rank's halo planes; what follows on the
    // main stream (a copy to the host, an upload, lb200_free) must come after them (the flags stay set: the next
    // step's own wait on the same values passes at once)
    if (binary && c->phi_src == SRC_FLAG && (rc = flags_wait(c, S, FLAG_PS_LO, c->n_ps)) != 0) return rc;
    if (c->f_src == SRC_FLAG && (rc = flags_wait(c, S, FLAG_COL_LO, c->n_col)) != 0) return rc;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}


// ---- slab pipeline of the single-GPU binary-fluid step (LB200_KNOB_PIPE) ----------------------------------------
// The two kernels of a step bound on different resources: the collision on HBM bandwidth with most issue slots
// idle, the phi sector on shared-memory instruction issue with half the bandwidth idle (profiles/).  They depend
// on each other only through neighbouring x-planes:
//   phi_sector(n, slab s) reads u(n-1) of slabs s-1, s, s+1 and overwrites force on slab s   -> after collide(n-1, s-1 .. s+1)
//   collide(n, slab s)    reads force(n) of slab s                                           -> after phi_sector(n, s)
// (phi, f: each is produced and consumed on one stream, in order; u is double-buffered.)  So with the lattice cut
// into S x-slabs, stream A runs the phi sector slab by slab and stream B the collision slab by slab, linked by
// one event per slab, and the phi sector of step n+1 runs next to the collision of step n.  Each stream belongs
// to a green context with its own SM partition, so that neither kernel's CTAs queue behind the other's.

typedef CUresult (*fn_cuDeviceGet_t)(CUdevice *, int);
typedef CUresult (*fn_cuDeviceGetDevResource_t)(CUdevice, CUdevResource *, CUdevResourceType);
typedef CUresult (*fn_cuDevSmResourceSplitByCount_t)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *,
						     unsigned int, unsigned int);
typedef CUresult (*fn_cuDevResourceGenerateDesc_t)(CUdevResourceDesc *, CUdevResource *, unsigned int);
typedef CUresult (*fn_cuGreenCtxCreate_t)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*fn_cuGreenCtxStreamCreate_t)(CUstream *, CUgreenCtx, unsigned int, int);
typedef CUresult (*fn_cuGreenCtxDestroy_t)(CUgreenCtx);

template <class F> static bool driver_entry(const char * name, F * fn) {
  void * p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return false;
  }
  *fn = (F) p;
  return true;
}

// two green contexts: >= knob_pipe_sms SMs for the phi sector, the remaining SMs for the collision
static bool pipe_green(lb200_t * c) {
  fn_cuDeviceGet_t dev_get; fn_cuDeviceGetDevResource_t get_res; fn_cuDevSmResourceSplitByCount_t split;
  fn_cuDevResourceGenerateDesc_t gen_desc; fn_cuGreenCtxCreate_t gctx_create; fn_cuGreenCtxStreamCreate_t gstream;
  fn_cuGreenCtxDestroy_t gctx_destroy;
  if (!driver_entry("cuDeviceGet", &dev_get) || !driver_entry("cuDeviceGetDevResource", &get_res)
      || !driver_entry("cuDevSmResourceSplitByCount", &split) || !driver_entry("cuDevResourceGenerateDesc", &gen_desc)
      || !driver_entry("cuGreenCtxCreate", &gctx_create) || !driver_entry("cuGreenCtxStreamCreate", &gstream)
      || !driver_entry("cuGreenCtxDestroy", &gctx_destroy)) return false;
  CUdevice dev;
  CUdevResource all, part[2];
  unsigned int ngroups = 1;
  memset(&all, 0, sizeof(all)); memset(part, 0, sizeof(part));
  if (dev_get(&dev, c->device) != CUDA_SUCCESS) return false;
  if (get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  const int want = c->knob_pipe_sms;
  if (want < 8 || want + 8 > (int) all.sm.smCount) return false;
  if (split(&part[0], &ngroups, &all, &part[1], 0, (unsigned int) want) != CUDA_SUCCESS || ngroups != 1) return false;
  if (part[0].sm.smCount == 0 || part[1].sm.smCount == 0) return false;
  CUgreenCtx g[2] = {nullptr, nullptr};
  CUstream st[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; i++) {
    CUdevResourceDesc desc = nullptr;
    if (gen_desc(&desc, &part[i], 1) != CUDA_SUCCESS || gctx_create(&g[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS
	|| gstream(&st[i], g[i], CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
      for (int j = 0; j <= i; j++) { if (st[j]) cudaStreamDestroy((cudaStream_t) st[j]); if (g[j]) gctx_destroy(g[j]); }
      cudaGetLastError();
      return false;
    }
  }
  c->pipe_a = (cudaStream_t) st[0]; c->pipe_b = (cudaStream_t) st[1];
  c->pipe_gctx[0] = g[0]; c->pipe_gctx[1] = g[1];
  c->pipe_sm[0] = (int) part[0].sm.smCount; c->pipe_sm[1] = (int) part[1].sm.smCount;
  return true;
}

static void pipe_teardown(lb200_t * c) {
  if (c->pipe_state <= 0) return;
  cudaStreamSynchronize(c->pipe_a); cudaStreamSynchronize(c->pipe_b);
  for (int i = 0; i < LB200_PIPE_MAXS; i++) { cudaEventDestroy(c->pipe_ev_ps[i]); cudaEventDestroy(c->pipe_ev_c[i]); }
  cudaEventDestroy(c->pipe_ev_join[0]); cudaEventDestroy(c->pipe_ev_join[1]);
  cudaStreamDestroy(c->pipe_a); cudaStreamDestroy(c->pipe_b);
  fn_cuGreenCtxDestroy_t gctx_destroy;
  if (c->pipe_state == 1 && driver_entry("cuGreenCtxDestroy", &gctx_destroy)) {
    gctx_destroy((CUgreenCtx) c->pipe_gctx[0]); gctx_destroy((CUgreenCtx) c->pipe_gctx[1]);
  }
  c->pipe_state = 0;
}

static int pipe_setup(lb200_t * c) {
  if (c->pipe_state != 0) return 0;
  const char * e = getenv("LB200_PIPE_GREEN");
  const bool try_green = (e == nullptr) || atoi(e) != 0;
  if (try_green && pipe_green(c)) {
    // a kernel of this library's module must launch on a green-context stream and events must link the two
    // partitions: check once, and fall back to plain streams if the driver refuses
    c->pipe_state = 1;
    unsigned int * scratch = nullptr;
    cudaEvent_t ev = nullptr;
    bool ok = cudaMalloc((void **) &scratch, 2*sizeof(unsigned int)) == cudaSuccess
      && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
      c->k->signal(c->pipe_a, scratch, scratch + 1, 1u);
      ok = cudaGetLastError() == cudaSuccess && cudaEventRecord(ev, c->pipe_a) == cudaSuccess
	&& cudaStreamWaitEvent(c->pipe_b, ev, 0) == cudaSuccess;
      if (ok) { c->k->signal(c->pipe_b, scratch, scratch + 1, 2u); ok = cudaGetLastError() == cudaSuccess; }
      ok = ok && cudaStreamSynchronize(c->pipe_b) == cudaSuccess && cudaStreamSynchronize(c->pipe_a) == cudaSuccess;
    }
    if (ev) cudaEventDestroy(ev);
    cudaFree(scratch);
    if (!ok) {
      cudaGetLastError();
      cudaStreamDestroy(c->pipe_a); cudaStreamDestroy(c->pipe_b);
      fn_cuGreenCtxDestroy_t gctx_destroy;
      if (driver_entry("cuGreenCtxDestroy", &gctx_destroy)) { gctx_destroy((CUgreenCtx) c->pipe_gctx[0]); gctx_destroy((CUgreenCtx) c->pipe_gctx[1]); }
      c->pipe_state = 0;
    }
  }
  if (c->pipe_state == 0) {
    // no partition API: two streams, the phi sector (one CTA needs a whole SM's registers) at the higher priority
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&c->pipe_a, cudaStreamNonBlocking, hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&c->pipe_b, cudaStreamNonBlocking, lo));
    c->pipe_sm[0] = c->pipe_sm[1] = 0;
    c->pipe_state = 2;
  }
  for (int i = 0; i < LB200_PIPE_MAXS; i++) {
    CUDA_TRY(cudaEventCreateWithFlags(&c->pipe_ev_ps[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->pipe_ev_c[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&c->pipe_ev_join[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->pipe_ev_join[1], cudaEventDisableTiming));
  if (c->u_alloc[1] == nullptr) {
    if (alloc_d(&c->u2, (size_t) 3*c->g.nsites) != 0) return LB200_ECUDA;
    c->u_alloc[0] = c->u; c->u_alloc[1] = c->u2;
  }
  return 0;
}

int lb200_pipe_state(const lb200_t * c, int sms[2]) {
  if (c == nullptr) return LB200_EINVAL;
  if (sms != nullptr) { sms[0] = c->pipe_sm[0]; sms[1] = c->pipe_sm[1]; }
  return c->pipe_state;
}

static int step_wrap(lb200_t * c, const Lb200CollideDev & cd, const Lb200SymmDev * sd, int nsteps);

static bool pipe_eligible(const lb200_t * c, int nsteps) {
  const Lb200Geom & g = c->g;
  return c->knob_pipe >= 2 && !g.remote_x && c->le.nplane == 0 && c->nvel == 19 && c->unrolled19 && c->ndist == 1
    && c->map_all_fluid && nsteps >= 2 && g.nl[0] >= 8*c->knob_pipe && c->pipe_state >= 0
    && !c->profile           // per-kernel timing wants each kernel alone on the device
    && !c->knob_f32;         // FP32 storage lives in the serial step
}

static int step_pipe(lb200_t * c, const Lb200CollideDev & cd, const Lb200SymmDev & sd, int nsteps) {
  int rc = pipe_setup(c);
  if (rc != 0) return rc;
  // the first step after an upload (in-place collision, u = 0 to materialise) takes the serial path
  if (!c->prop_pending || c->u_state == ZERO_PENDING || c->u_state == ARRAY_CLEAN) {
    rc = step_wrap(c, cd, &sd, 1);
    if (rc != 0) return rc;
    nsteps -= 1;
  }
  if (nsteps <= 0) return 0;

  const int S = c->knob_pipe;
  Lb200Geom gw = c->g;
  gw.wrap[0] = gw.wrap[1] = gw.wrap[2] = 1;
  cudaStream_t A = c->pipe_a, B = c->pipe_b;
  int x0[LB200_PIPE_MAXS + 1];
  for (int s = 0; s <= S; s++) x0[s] = (int) (((long long) s*gw.nl[0])/S);
  static const int chunks_per_slab = getenv("LB200_PIPE_CHUNKS") ? atoi(getenv("LB200_PIPE_CHUNKS")) : 1;

  CUDA_TRY(cudaEventRecord(c->ev_main, c->stream));
  CUDA_TRY(cudaStreamWaitEvent(A, c->ev_main, 0));
  CUDA_TRY(cudaStreamWaitEvent(B, c->ev_main, 0));

  bool linked = false;          // pipe_ev_c[] hold the collisions of the previous pipelined step
  for (int n = 0; n < nsteps; n++) {
    c->t_current += 1;
    double * u_in = c->u;
    double * u_out = (c->u == c->u_alloc[0]) ? c->u_alloc[1] : c->u_alloc[0];
    for (int s = 0; s < S; s++) {
      if (linked) {
	const int sm = (s + S - 1) % S, sp = (s + 1) % S;
	CUDA_TRY(cudaStreamWaitEvent(A, c->pipe_ev_c[s], 0));
	CUDA_TRY(cudaStreamWaitEvent(A, c->pipe_ev_c[sm], 0));
	if (sp != sm) CUDA_TRY(cudaStreamWaitEvent(A, c->pipe_ev_c[sp], 0));
      }
      Lb200Geom gs = gw;
      gs.skip_diag = (c->knob_lazy_diag && n < nsteps - 1) ? 1 : 0;
      gs.xoff = x0[s]; gs.xcnt = x0[s + 1] - x0[s];
      gs.xchunk = (chunks_per_slab >= 1) ? (gs.xcnt + chunks_per_slab - 1)/chunks_per_slab : 0;
      {
	ProfScope ps(c, LB200_K_PHI_SECTOR, A);
	c->launches += c->k->phi_sector(A, gs, sd, c->phi, u_in, c->grad, c->delsq, c->force, c->phinew);
      }
      CUDA_TRY(cudaEventRecord(c->pipe_ev_ps[s], A));
    }
    { double * t = c->phi; c->phi = c->phinew; c->phinew = t; }
    for (int s = 0; s < S; s++) {
      CUDA_TRY(cudaStreamWaitEvent(B, c->pipe_ev_ps[s], 0));
      Lb200Geom gs = gw;
      gs.skip_diag = (c->knob_lazy_diag && n < nsteps - 1) ? 1 : 0;
      gs.xoff = x0[s]; gs.xcnt = x0[s + 1] - x0[s];
      {
	ProfScope ps(c, LB200_K_COLLIDE, B);
	c->launches += c->k->collide(B, gs, cd, nullptr, 19, 1, c->f, c->fprime, c->force, nullptr, c->rho, u_out);
      }
      CUDA_TRY(cudaEventRecord(c->pipe_ev_c[s], B));
    }
    { double * t = c->f; c->f = c->fprime; c->fprime = t; }
    c->u = u_out;
    linked = true;
  }
  c->force_state = INTERIOR_ONLY;
  c->u_state = INTERIOR_ONLY;
  c->prop_pending = 1;
  c->f_halo_stale = 1; c->fused_ready = 0;
  c->phi_halo_valid = 0;
  c->u_halo_valid = 0;
  c->wrap_x_valid = 1;
  // rejoin the context's stream
  CUDA_TRY(cudaEventRecord(c->pipe_ev_join[0], A));
  CUDA_TRY(cudaEventRecord(c->pipe_ev_join[1], B));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->pipe_ev_join[0], 0));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->pipe_ev_join[1], 0));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- liquid crystal (options.have_q): lb200_lc.cuh ---------------------------------------------------------------

static int lc_dev(lb200_t * c, const lb200_lc_param_t * lc, Lb200LcDev * d) {
  if (c->q == nullptr) return fail(LB200_ESTATE, "no q in this context (options.have_q)");
  if (lc == nullptr) return fail(LB200_EINVAL, "null parameters");
  if (lc->adv_order < 1 || lc->adv_order > 4) return fail(LB200_EINVAL, "advection order %d: device kernels exist for 1-4", lc->adv_order);
  if (!c->map_all_fluid) return fail(LB200_ESTATE, "liquid crystal: all-fluid lattices only in this build");
  d->a0 = lc->a0; d->q0 = lc->q0; d->gamma = lc->gamma; d->kappa0 = lc->kappa0; d->kappa1 = lc->kappa1; d->xi = lc->xi;
  d->Gamma = lc->Gamma; d->epsilon = lc->epsilon;
  for (int a = 0; a < 3; a++) d->e0[a] = lc->e0[a];
  d->order = lc->adv_order;
  if (lc->is_active && lc->zeta2 != 0.0) return fail(LB200_EINVAL, "lc_active_zeta2 != 0 (polarisation-gradient active stress, fe_lc_active_stress) is outside this build");
  d->is_active = (lc->is_active != 0); d->zeta0 = lc->zeta0; d->zeta1 = lc->zeta1;
  d->redshift = (lc->redshift != 0.0) ? lc->redshift : 1.0;
  if (fabs(d->redshift) < 1.0e-5) return fail(LB200_EINVAL, "redshift %g below FE_REDSHIFT_MIN (src/blue_phase.c:29-32)", d->redshift);
  d->rredshift = 1.0/d->redshift;
  d->g2d = c->knob_qgrad2d;
  return 0;
}

int lb200_q_halo(lb200_t * c) {
  CTX_ENTER(c);
  if (c->q == nullptr) return fail(LB200_ESTATE, "no q in this context");
  int rc = halo_field(c, c->q, 5, c->g.nh, 0, nullptr);
  if (rc != 0) return rc;
  CTX_LEAVE_SYNC(c);
}

// field_grad_compute(q_grad), d2 = grad_3d_7pt_fluid_d2: [1 - (nhalo - 1), N + nhalo - 1]^3
int lb200_q_grad_compute(lb200_t * c) {
  CTX_ENTER(c);
  if (c->q == nullptr) return fail(LB200_ESTATE, "no q in this context");
  if (c->qgrad == nullptr) {
    if (alloc_d(&c->qgrad, (size_t) 15*c->g.nsites) != 0 || alloc_d(&c->qdelsq, (size_t) 5*c->g.nsites) != 0) return LB200_ECUDA;
  }
  {
    ProfScope ps(c, LB200_K_GRAD);
    c->launches += c->k->grad7(c->stream, c->g, c->g.nh - 1, c->knob_qgrad2d ? -5 : 5, c->q, c->qgrad, c->qdelsq);   // (-5: 2d_5pt_fluid)
  }
  CTX_LEAVE_SYNC(c);
}

int lb200_lc_stress_compute(lb200_t * c, const lb200_lc_param_t * lc) {
  CTX_ENTER(c);
  Lb200LcDev d;
  int rc = lc_dev(c, lc, &d);
  if (rc != 0) return rc;
  {
    ProfScope ps(c, LB200_K_LC_STRESS);
    c->launches += c->k->lc_stress(c->stream, c->g, d, 1, 1, c->q, c->str);
  }
  CTX_LEAVE_SYNC(c);
}

int lb200_lc_force_calculation(lb200_t * c, const lb200_lc_param_t * lc) {
  CTX_ENTER(c);
  Lb200LcDev d;
  int rc = lc_dev(c, lc, &d);
  if (rc != 0) return rc;
  {
    ProfScope ps(c, LB200_K_LC_STRESS);
    c->launches += c->k->lc_stress(c->stream, c->g, d, 1, 1, c->q, c->str);
  }
  {
    const int accumulate = (c->force_state != ZERO_PENDING);
    ProfScope ps(c, LB200_K_LC_BE);
    c->launches += c->k->lc_force_be(c->stream, c->g, d, 1, 0, accumulate, c->q, c->str, c->u, c->force, nullptr);
    if (c->force_state == ZERO_PENDING) c->force_state = INTERIOR_ONLY;
  }
  CTX_LEAVE_SYNC(c);
}

int lb200_beris_edw_update(lb200_t * c, const lb200_lc_param_t * lc) {
  CTX_ENTER(c);
  Lb200LcDev d;
  int rc = lc_dev(c, lc, &d);
  if (rc != 0) return rc;
  if (c->u_state != ARRAY_CLEAN) materialise_zero(c, c->u, &c->u_state);
  // qnew holds q everywhere (halo included) so that the swap keeps the reference's view of the array
  CUDA_TRY(cudaMemcpyAsync(c->qnew, c->q, (size_t) 5*c->g.nsites*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  {
    ProfScope ps(c, LB200_K_LC_BE);
    c->launches += c->k->lc_force_be(c->stream, c->g, d, 0, 1, 0, c->q, nullptr, c->u, nullptr, c->qnew);
  }
  { double * t = c->q; c->q = c->qnew; c->qnew = t; }
  CTX_LEAVE_SYNC(c);
}

// `depth` boundary x-planes of every component of a field straight into the neighbours' halo planes (each component's
// planes are contiguous: no staging), one NCCL group
static int wrap_exchange_field(lb200_t * c, cudaStream_t st, double * data, int ncomp, int depth) {
  const Lb200Geom & g = c->g;
  const size_t xs = (size_t) g.xs, ns = (size_t) g.nsites;
  XMsg m[8];
  if (ncomp > 8) return fail(LB200_EINVAL, "wrap_exchange_field: ncomp = %d", ncomp);
  for (int n = 0; n < ncomp; n++) {
    double * a = data + n*ns;
    m[n].send_hi = a + (size_t) (g.nl[0] - depth + g.nh)*xs;
    m[n].send_lo = a + (size_t) g.nh*xs;
    m[n].recv_lo = a + (size_t) (g.nh - depth)*xs;
    m[n].recv_hi = a + (size_t) (g.nl[0] + g.nh)*xs;
    m[n].count = (size_t) depth*xs;
  }
  ProfScope ps(c, LB200_K_HALO, st);
  return nccl_exchange(c, st, m, ncomp);
}

// Peer copies of boundary x-planes (x-slabs, liquid-crystal step): `ncomp` consecutive components starting at `comp0`,
// `depth` planes, from my array `src` straight into the halo planes of the neighbours' arrays (cudaIpc-mapped, NVLink),
// as strided device-to-device copies on the comm stream -- no SM is taken from the compute kernels (the NCCL send/recv
// kernels they replace compete with 0.1 ms sweeps for SMs).  up != 0: my planes [N-d+1, N] -> the high neighbour's
// planes [1-d, 0]; down != 0: my planes [1, d] -> the low neighbour's planes [N+1, N+d].
static int peer_copy_planes(lb200_t * c, cudaStream_t st, const double * src, double * dst_lo, double * dst_hi,
			    int comp0, int ncomp, int depth, int up, int down) {
  const Lb200Geom & g = c->g;
  const size_t xs = (size_t) g.xs, ns = (size_t) g.nsites;
  const size_t width = (size_t) depth*xs*sizeof(double), pitch = ns*sizeof(double);
  if (up) {
    CUDA_TRY(cudaMemcpy2DAsync(dst_hi + comp0*ns + (size_t) (g.nh - depth)*xs, pitch,
			       src + comp0*ns + (size_t) (g.nl[0] - depth + g.nh)*xs, pitch, width, ncomp, cudaMemcpyDeviceToDevice, st));
  }
  if (down) {
    CUDA_TRY(cudaMemcpy2DAsync(dst_lo + comp0*ns + (size_t) (g.nl[0] + g.nh)*xs, pitch,
			       src + comp0*ns + (size_t) g.nh*xs, pitch, width, ncomp, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// the populations moving in +x go up, those moving in -x go down (runs of consecutive populations: one copy each)
static int peer_copy_f(lb200_t * c, cudaStream_t st, const double * f, double * dst_lo, double * dst_hi) {
  lb200_step_plan_t pl;
  int rc = lb200_step_plan(&c->opt, LB200_STEP_F, &pl);
  if (rc != 0) return rc;
  for (int dir = 0; dir < 2; dir++) {
    const int * comp = dir == 0 ? pl.comp_up : pl.comp_down;
    const int ncomp = dir == 0 ? pl.ncomp_up : pl.ncomp_down;
    for (int a = 0; a < ncomp; ) {
      int b = a;
      while (b + 1 < ncomp && comp[b + 1] == comp[b] + 1) b++;
      rc = peer_copy_planes(c, st, f, dst_lo, dst_hi, comp[a], b - a + 1, 1, dir == 0, dir == 1);
      if (rc != 0) return rc;
      a = b + 1;
    }
  }
  return 0;
}

int lb200_step_lc(lb200_t * c, const lb200_collide_param_t * cp, const lb200_lc_param_t * lc, int nsteps) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  if (cp == nullptr) return fail(LB200_EINVAL, "null collision parameters");
  Lb200CollideDev cd;
  Lb200LcDev d;
  int rc = collide_dev(c, cp, &cd);
  if (rc != 0) return rc;
  rc = lc_dev(c, lc, &d);
  if (rc != 0) return rc;
  if (nsteps <= 0) return 0;
  const Lb200Geom & g = c->g;
  // halo-free steps on periodic lattices: every y / z (one GPU: also x) periodic image is read in-kernel; x-slabs
  // exchange only the x-planes the kernels read (q: nhalo, u: 1, f: the populations crossing) on the comm stream,
  // overlapped with the kernels that do not need them -- as peer copies into the neighbours' mapped arrays with one
  // flag per exchange (default), or NCCL send/recv.  Otherwise the reference's structure with halo kernels.
  bool wrap = c->knob_wrap && g.per[0] && g.per[1] && g.per[2];
  for (int a = 0; a < 3; a++) wrap = wrap && (g.nl[a] >= 2*g.nh);
  const bool remote = wrap && g.remote_x;
  if (remote) {
    rc = peer_setup(c);                  // collective, does something the first time only
    if (rc != 0) return rc;
  }
  const bool peer = remote && c->peer_state == 1 && c->knob_peer;
  cudaStream_t S = c->stream, C = (remote && !c->profile) ? c->comm : c->stream;
  Lb200Geom gw = c->g;
  if (wrap) { gw.wrap[0] = !g.remote_x; gw.wrap[1] = gw.wrap[2] = 1; }
  if (!wrap) {
    rc = ensure_f_halo(c);
    if (rc != 0) return rc;
  }
  if (c->u_state == ZERO_PENDING) materialise_zero(c, c->u, &c->u_state);
  int q_src = SRC_NONE, u_src = SRC_NONE, f_src = SRC_NONE;
  if (remote) {
    // bring the planes the first kernels read (the state did not come out of a previous halo-free step); this round
    // of messages also orders this call after whatever the neighbours were doing before
    CUDA_TRY(cudaEventRecord(c->ev_main, S));
    CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
    if ((rc = wrap_exchange_field(c, C, c->q, 5, g.nh)) != 0) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_phi, C));
    if ((rc = wrap_exchange_field(c, C, c->u, 3, 1)) != 0) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_u, C));
    if (c->prop_pending && (rc = wrap_exchange_f(c, C)) != 0) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_f, C));
    q_src = u_src = f_src = SRC_EVENT;
  }

  for (int n = 0; n < nsteps; n++) {
    c->force_state = ZERO_PENDING;                                       // hydro_f_zero
    if (!wrap) {
      rc = halo_field(c, c->q, 5, g.nh, 0, S);                           // field_halo(q)
      if (rc != 0) return rc;
    }
    if (remote && (rc = src_wait(c, S, q_src, c->ev_phi, FLAG_PS_LO, c->n_ps)) != 0) return rc;
    {
      // field_grad_compute + pth_stress_compute: gradients in registers, only the stress is stored
      ProfScope ps(c, LB200_K_LC_STRESS);
      c->launches += c->k->lc_stress(S, gw, d, (!wrap || remote) ? 1 : 0, wrap ? 0 : 1, c->q, c->str);
    }
    if (!wrap) {
      if (c->u_state == ZERO_PENDING) materialise_zero(c, c->u, &c->u_state);
      rc = halo_field(c, c->u, 3, g.nh, 0, S);                           // hydro_u_halo
      if (rc != 0) return rc;
      c->u_state = ARRAY_CLEAN;
    }
    if (remote && (rc = src_wait(c, S, u_src, c->ev_u, FLAG_COL_LO, c->n_col)) != 0) return rc;
    {
      // pth_force_fluid_driver + beris_edw_update: two sweeps (default; the 64-register force gather runs at high
      // occupancy on its own: 1.19 vs 1.52 ms at 256^3) or one (LB200_LC_SPLIT=0)
      static const int split = getenv("LB200_LC_SPLIT") ? atoi(getenv("LB200_LC_SPLIT")) : 1;
      ProfScope ps(c, LB200_K_LC_BE);
      if (split) {
	c->launches += c->k->lc_force_be(S, gw, d, 1, 0, 0, c->q, c->str, c->u, c->force, c->qnew);
	c->launches += c->k->lc_force_be(S, gw, d, 0, 1, 0, c->q, c->str, c->u, c->force, c->qnew);
      }
      else {
	c->launches += c->k->lc_force_be(S, gw, d, 1, 1, 0, c->q, c->str, c->u, c->force, c->qnew);
      }
    }
    c->force_state = INTERIOR_ONLY;
    { double * t = c->q; c->q = c->qnew; c->qnew = t; }
    if (remote) {
      // the new q planes travel while the collision runs
      CUDA_TRY(cudaEventRecord(c->ev_main, S));
      CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
      if (peer) {
	const int iq = idx2(c->q, c->phi_alloc);
	ProfScope ps(c, LB200_K_HALO, C);
	if ((rc = peer_copy_planes(c, C, c->q, c->lo.phi[iq], c->hi.phi[iq], 0, 5, g.nh, 1, 1)) != 0) return rc;
	c->n_ps++;
	c->launches += c->k->signal(C, c->hi.flags + FLAG_PS_LO, c->lo.flags + FLAG_PS_HI, c->n_ps);
	c->joint_signal = 0;
	q_src = SRC_FLAG;
      }
      else {
	if ((rc = wrap_exchange_field(c, C, c->q, 5, g.nh)) != 0) return rc;
	CUDA_TRY(cudaEventRecord(c->ev_phi, C));
	q_src = SRC_EVENT;
      }
    }
    c->u_state = ZERO_PENDING;                                           // hydro_u_zero
    if (wrap) {
      if (remote && (rc = src_wait(c, S, f_src, c->ev_f, FLAG_COL_LO, c->n_col)) != 0) return rc;
      // with peer copies the neighbours may still be reading the u planes they got last step: write the other buffer
      double * u_out = peer ? ((c->u == c->u_alloc[0]) ? c->u_alloc[1] : c->u_alloc[0]) : c->u;
      {
	ProfScope ps(c, LB200_K_COLLIDE);
	if (c->prop_pending) {
	  gw.skip_diag = (c->knob_lazy_diag && n < nsteps - 1) ? 1 : 0;      // hydro->rho: the last step of the call stores it
	  c->launches += c->k->collide(S, gw, cd, model_ptr(c), c->nvel, 1, c->f, c->fprime, c->force, status_ptr(c), c->rho, u_out);
	  double * t = c->f; c->f = c->fprime; c->fprime = t;
	  c->prop_pending = 0;
	}
	else {
	  c->launches += c->k->collide(S, c->g, cd, model_ptr(c), c->nvel, 0, c->f, c->f, c->force, status_ptr(c), c->rho, u_out);
	}
      }
      c->u = u_out;
      c->u_state = INTERIOR_ONLY;
      c->prop_pending = 1;                                               // lb_halo; lb_propagation (lazy)
      c->f_halo_stale = 1; c->fused_ready = 0;
      if (remote) {
	CUDA_TRY(cudaEventRecord(c->ev_main, S));
	CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
	if (peer) {
	  const int iu = idx2(c->u, c->u_alloc), jf = idx2(c->f, c->f_alloc);
	  ProfScope ps(c, LB200_K_HALO, C);
	  if ((rc = peer_copy_planes(c, C, c->u, c->lo.u[iu], c->hi.u[iu], 0, 3, 1, 1, 1)) != 0) return rc;
	  if ((rc = peer_copy_f(c, C, c->f, c->lo.f[jf], c->hi.f[jf])) != 0) return rc;
	  c->n_col++;
	  c->launches += c->k->signal(C, c->hi.flags + FLAG_COL_LO, c->lo.flags + FLAG_COL_HI, c->n_col);
	  c->joint_signal = 0;
	  u_src = f_src = SRC_FLAG;
	}
	else {
	  if ((rc = wrap_exchange_field(c, C, c->u, 3, 1)) != 0) return rc;       // the next Beris-Edwards sweep waits for this
	  CUDA_TRY(cudaEventRecord(c->ev_u, C));
	  if ((rc = wrap_exchange_f(c, C)) != 0) return rc;                       // overlaps the next two LC sweeps
	  CUDA_TRY(cudaEventRecord(c->ev_f, C));
	  u_src = f_src = SRC_EVENT;
	}
      }
    }
    else {
      rc = collide_async(c, cd);
      if (rc != 0) return rc;
      rc = halo_field(c, c->f, c->nvel*c->ndist, 1, c->opt.halo_scheme == LB200_HALO_REDUCED, S);   // lb_halo
      if (rc != 0) return rc;
      c->prop_pending = 1;
      c->f_halo_stale = 0;
    }
  }
  if (remote) {
    // rejoin: what is still in flight on the comm stream (it reads this rank's boundary planes) is ordered before
    // whatever follows on the main stream, and the planes the neighbours sent have arrived
    CUDA_TRY(cudaEventRecord(c->ev_f, C));
    CUDA_TRY(cudaStreamWaitEvent(S, c->ev_f, 0));
    if (q_src == SRC_FLAG && (rc = flags_wait(c, S, FLAG_PS_LO, c->n_ps)) != 0) return rc;
    if (f_src == SRC_FLAG && (rc = flags_wait(c, S, FLAG_COL_LO, c->n_col)) != 0) return rc;
  }
  c->phi_halo_valid = 0;
  c->u_halo_valid = 0;
  c->wrap_x_valid = 0;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- whole time steps with Lees-Edwards planes -----------------------------------------------------------------
// Reference order (src/ludwig.c:528-860 with ludwig->le): hydro_f_zero; field_halo(phi); field_grad_compute
// (field_leesedwards + d2 + buffer-region gradients); phi_force_calculation (flux form); phi_cahn_hilliard
// (hydro_u_halo, hydro_lees_edwards, fluxes, phi_ch_le_fix_fluxes, update); hydro_u_zero; lb_collide;
// lb_data_apply_le_boundary_conditions; lb_halo; lb_propagation.
//   strict: the reference's operations everywhere (flux-form force on the whole lattice): bit-identical.
//   fast  : the one-sweep phi-sector kernel for the bulk, then the planes whose stencils cross a Lees-Edwards
//           plane (2*nhalo per plane) are redone by the generic plane kernels through the buffer planes.
static int step_le(lb200_t * c, const Lb200CollideDev & cd, const Lb200SymmDev & sd, int nsteps, bool conserve2) {
  cudaStream_t S = c->stream;
  const Lb200Geom & g = c->g;
  int rc = ensure_f_halo(c);
  if (rc != 0) return rc;
  const bool fused = (c->opt.math == LB200_MATH_FAST) && c->knob_phi_sector && c->map_all_fluid && sd.order <= 3 && sd.csum == nullptr
    && !c->knob_grad7;

  for (int n = 0; n < nsteps; n++) {
    c->t_current += 1;                                                   // physics_control_next_step
    c->force_state = ZERO_PENDING;                                       // hydro_f_zero
    rc = halo_field(c, c->phi, 1, g.nh, 0, S);                           // field_halo(phi)
    if (rc != 0) return rc;
    le_field_async(c, c->phi);                                           // field_leesedwards
    if (c->u_state == ZERO_PENDING || (c->u_state == INTERIOR_ONLY && thin_lattice(c))) materialise_zero(c, c->u, &c->u_state);
    rc = halo_field(c, c->u, 3, g.nh, 0, S);                             // hydro_u_halo
    if (rc != 0) return rc;
    c->u_state = ARRAY_CLEAN;
    le_hydro_async(c);                                                   // hydro_lees_edwards
    if (thin_lattice(c)) {
      // the reference updates phi in place: its halo sites keep their content until the next swap
      CUDA_TRY(cudaMemcpyAsync(c->phinew, c->phi, (size_t) c->g.nsites*sizeof(double), cudaMemcpyDeviceToDevice, S));
    }
    if (fused) {
      {
	ProfScope ps(c, LB200_K_PHI_SECTOR);
	c->launches += c->k->phi_sector(S, g, sd, c->phi, c->u, c->grad, c->delsq, c->force, c->phinew);
      }
      le_grad_async(c);
      le_force_ch_async(c, sd, c->le_nxlist, c->le_xlist, 1, 1, 0, c->phinew);
    }
    else {
      {
	ProfScope ps(c, LB200_K_GRAD);
	if (c->knob_grad7) c->launches += c->k->grad7(S, g, g.nh - 1, 1, c->phi, c->grad, c->delsq);   // grad_3d_7pt_fluid_d2
	else               c->launches += c->k->grad27(S, g, g.nh - 1, c->phi, c->grad, c->delsq);
      }
      le_grad_async(c);
      le_force_ch_async(c, sd, g.nl[0], nullptr, 1, 1, 0, c->phinew);
    }
    // cahn_hilliard_options_conserve 2: the same correction with or without planes (src/phi_cahn_hilliard.c:281-285)
    if (conserve2 && (rc = phi_conserve_subtract_async(c, c->phinew, S)) != 0) return rc;
    c->force_state = INTERIOR_ONLY;
    { double * t = c->phi; c->phi = c->phinew; c->phinew = t; }
    c->u_state = ZERO_PENDING;                                           // hydro_u_zero
    rc = collide_async(c, cd);                                           // (lb_propagation +) lb_collide
    if (rc != 0) return rc;
    le_lb_bc_async(c);                                                   // lb_data_apply_le_boundary_conditions
    rc = halo_field(c, c->f, c->nvel*c->ndist, 1, c->opt.halo_scheme == LB200_HALO_REDUCED, S);   // lb_halo
    if (rc != 0) return rc;
    c->prop_pending = 1;                                                 // lb_propagation (lazy)
    c->f_halo_stale = 0;
  }
  c->phi_halo_valid = 0;
  c->u_halo_valid = 0;
  c->wrap_x_valid = 0;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- whole time steps ---------------------------------------------------------------------------------
// Order of operations of the reference driver (src/ludwig.c:528-860):
//   hydro_f_zero; field_halo(phi); field_grad_compute; phi_force_calculation; phi_cahn_hilliard
//   (hydro_u_halo inside); hydro_u_zero; lb_collide; lb_halo; lb_propagation.
// Fusions (results unchanged): propagation(t) + collide(t+1) in one pull kernel; stress + force
// divergence + Cahn-Hilliard fluxes + update in one kernel; the two zeroing sweeps folded away.

int lb200_step(lb200_t * c, const lb200_collide_param_t * cp, const lb200_symm_param_t * sp, int nsteps) {
  if (c == nullptr) return fail(LB200_EINVAL, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  if (cp == nullptr) return fail(LB200_EINVAL, "null collision parameters");
  const int binary = (sp != nullptr && c->phi != nullptr);
  if (binary && (sp->adv_order < 1 || sp->adv_order > 4)) return fail(LB200_EINVAL, "advection order %d", sp->adv_order);
  if (binary && (sp->force_method < 0 || sp->force_method > 1)) return fail(LB200_EINVAL, "fe_force_method %d: 0 (stress_divergence) and 1 (phi_gradmu) are built", sp->force_method);
  if (binary && sp->force_method == 1 && c->ndist != 1) return fail(LB200_EINVAL, "fe_force_method phi_gradmu: the finite-difference binary fluid (ndist = 1)");
  Lb200CollideDev cd;
  Lb200SymmDev sd;
  int rc = collide_dev(c, cp, &cd);
  if (rc != 0) return rc;
  if (binary) {
    if (c->ndist == 1 && (rc = conserve_prepare(c, sp)) != 0) return rc;
    symm_dev(c, sp, &sd);
  }
  // conserve 2: a global sum sits between the Cahn-Hilliard update and everything that reads the new phi -- the
  // reference-structured step (halo kernels), not the halo-free one whose kernels hand planes to the neighbours as they go
  const bool conserve2 = binary && c->ndist == 1 && sp->conserve == 2;
  if (nsteps <= 0) return 0;
  if (binary && c->knob_grad7 && c->ndist != 1)
    return fail(LB200_EINVAL, "fd_gradient_calculation 3d_7pt_fluid: built for the finite-difference binary fluid (ndist = 1)");
  if (c->le.nplane > 0) {
    if (!binary) return fail(LB200_EINVAL, "Lees-Edwards planes: lb200_step is implemented for the binary fluid (free_energy symmetric / symmetric_lb)");
    if (c->ndist == 2) return step_lb2(c, cd, sd, nsteps);              // symmetric_lb: two distributions, no finite-difference sector
    // fast mode on a fully periodic all-fluid lattice: the halo-free step with plane patches, provided every plane
    // patch (and its +-2 stencil) stays clear of the slab boundary planes that the neighbours exchange
    const Lb200Geom & g = c->g;
    bool ok = (c->opt.math == LB200_MATH_FAST) && c->knob_wrap && c->knob_phi_sector && c->map_all_fluid
      && g.per[0] && g.per[1] && g.per[2] && sd.order <= 3 && sd.csum == nullptr && !conserve2 && !c->knob_grad7;
    for (int a = 0; a < 3; a++) ok = ok && (g.nl[a] >= 2*g.nh);
    for (int p = 0; p < c->le.nplane; p++) ok = ok && (c->le.loc[p] - g.nh - 1 >= 1) && (c->le.loc[p] + g.nh + 2 <= g.nl[0]);
    if (ok) return step_wrap(c, cd, &sd, nsteps);
    return step_le(c, cd, sd, nsteps, conserve2);
  }
  if (c->ndist == 2) {
    if (!binary) return fail(LB200_EINVAL, "ndist = 2 needs the free-energy parameters");
    return step_lb2(c, cd, sd, nsteps);
  }

  {
    // Halo-free time steps (fully periodic lattices; binary fluid: all-fluid map, the one-sweep phi sector)
    const int wrap_enabled = c->knob_wrap, ps_on = c->knob_phi_sector;
    const Lb200Geom & g = c->g;
    bool ok = wrap_enabled && g.per[0] && g.per[1] && g.per[2] && c->ndist == 1;
    for (int a = 0; a < 3; a++) ok = ok && (g.nl[a] >= 2*g.nh);
    if (binary) ok = ok && ps_on && c->map_all_fluid && sd.order <= 3 && sd.csum == nullptr && !c->knob_grad7 && !conserve2
		  && sd.force_method == 0;   // the one-sweep phi sector: orders 1-3, plain update
    if (ok && binary && pipe_eligible(c, nsteps)) return step_pipe(c, cd, sd, nsteps);
    if (ok) return step_wrap(c, cd, binary ? &sd : nullptr, nsteps);
  }
  {
    int rc2 = ensure_f_halo(c);
    if (rc2 != 0) return rc2;
  }

  // The three halo exchanges run on the comm stream, each overlapped with a compute kernel that does
  // not touch the array in flight:   phi(t+1) halo || collide(t);   f(t) and u(t) halos || grad(t+1),
  // force+CH(t+1).  Same operations on the same data as the serial order, so results are unchanged.
  // (while per-kernel profiling is on, everything runs on one stream so that launch durations are clean)
  cudaStream_t S = c->stream, C = c->profile ? c->stream : c->comm;

  // anything still pending on the main stream must be visible to the comm stream
  CUDA_TRY(cudaEventRecord(c->ev_main, S));
  CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));

  for (int n = 0; n < nsteps; n++) {
    c->force_state = ZERO_PENDING;                                       // hydro_f_zero
    if (binary) {
      if (!c->phi_halo_valid) {                                          // field_halo(phi)
	rc = halo_field(c, c->phi, 1, c->g.nh, 0, C);
	if (rc != 0) return rc;
	CUDA_TRY(cudaEventRecord(c->ev_phi, C));
      }
      // all-fluid lattices: gradient + force + Cahn-Hilliard in one sweep (LB200_PHI_SECTOR=0 disables)
      const bool use_ps = c->knob_phi_sector && c->map_all_fluid && sd.order <= 3 && sd.csum == nullptr && !c->knob_grad7
	&& sd.force_method == 0;
      CUDA_TRY(cudaStreamWaitEvent(S, c->ev_phi, 0));
      if (!use_ps) {
	ProfScope ps(c, LB200_K_GRAD);
	if (c->knob_grad7) c->launches += c->k->grad7(S, c->g, c->g.nh - 1, 1, c->phi, c->grad, c->delsq);   // grad_3d_7pt_fluid_d2
	else               c->launches += c->k->grad27(S, c->g, c->g.nh - 1, c->phi, c->grad, c->delsq);    // field_grad_compute
      }
      if (!c->u_halo_valid) {                                            // hydro_u_halo
	if (c->u_state == ZERO_PENDING || (c->u_state == INTERIOR_ONLY && thin_lattice(c))) materialise_zero(c, c->u, &c->u_state);
	CUDA_TRY(cudaEventRecord(c->ev_main, S));
	CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
	rc = halo_field(c, c->u, 3, c->g.nh, 0, C);
	if (rc != 0) return rc;
	CUDA_TRY(cudaEventRecord(c->ev_u, C));
	c->u_state = ARRAY_CLEAN;
      }
      CUDA_TRY(cudaStreamWaitEvent(S, c->ev_u, 0));
      if (thin_lattice(c)) {
	// the reference updates phi in place: its halo sites keep their content until the next swap
	CUDA_TRY(cudaMemcpyAsync(c->phinew, c->phi, (size_t) c->g.nsites*sizeof(double), cudaMemcpyDeviceToDevice, S));
      }
      if (use_ps) {
	// field_grad_compute + phi_force_calculation + phi_cahn_hilliard
	ProfScope ps(c, LB200_K_PHI_SECTOR);
	c->launches += c->k->phi_sector(S, c->g, sd, c->phi, c->u, c->grad, c->delsq, c->force, c->phinew);
      }
      else {
	// phi_force_calculation + phi_cahn_hilliard: one sweep, or two lighter ones (LB200_SPLIT_FCH=1)
	static const int split = getenv("LB200_SPLIT_FCH") ? atoi(getenv("LB200_SPLIT_FCH")) : 0;
	ProfScope ps(c, LB200_K_FORCE_CH);
	if (split || sd.force_method != 0) {                             // (phi_gradmu: its own force kernel)
	  c->launches += c->k->phi_force(S, c->g, sd, 0, c->phi, c->grad, c->delsq, c->force);
	  c->launches += c->k->cahn_hilliard(S, c->g, sd, c->phi, c->delsq, c->u, status_ptr(c), c->phinew);
	}
	else {
	  c->launches += c->k->force_ch(S, c->g, sd, 0, c->phi, c->grad, c->delsq, c->u,
					status_ptr(c), c->force, c->phinew);
	}
      }
      c->force_state = INTERIOR_ONLY;
      if (conserve2 && (rc = phi_conserve_subtract_async(c, c->phinew, S)) != 0) return rc;
      double * t = c->phi; c->phi = c->phinew; c->phinew = t;
      // halo of the NEW phi (needed by the next step's gradient) while the collision runs
      CUDA_TRY(cudaEventRecord(c->ev_main, S));
      CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
      rc = halo_field(c, c->phi, 1, c->g.nh, 0, C);
      if (rc != 0) return rc;
      CUDA_TRY(cudaEventRecord(c->ev_phi, C));
      c->phi_halo_valid = 1;
    }
    c->u_state = ZERO_PENDING;                                           // hydro_u_zero
    if (c->prop_pending) CUDA_TRY(cudaStreamWaitEvent(S, c->ev_f, 0));   // halo of f from the previous step
    collide_async(c, cd);                                                // (lb_propagation +) lb_collide
    // lb_halo and the next step's hydro_u_halo while the next gradient / force kernels run
    rc = materialise_propagation(c);
    if (rc != 0) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_main, S));
    CUDA_TRY(cudaStreamWaitEvent(C, c->ev_main, 0));
    rc = halo_field(c, c->f, c->nvel*c->ndist, 1, c->opt.halo_scheme == LB200_HALO_REDUCED, C);   // lb_halo
    if (rc != 0) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_f, C));
    if (binary) {
      rc = halo_field(c, c->u, 3, c->g.nh, 0, C);
      if (rc != 0) return rc;
      CUDA_TRY(cudaEventRecord(c->ev_u, C));
      c->u_state = ARRAY_CLEAN;
      c->u_halo_valid = 1;
    }
    c->prop_pending = 1;                                                 // lb_propagation (lazy)
  }
  // rejoin: everything issued on the comm stream is ordered before whatever follows on the main stream
  CUDA_TRY(cudaStreamWaitEvent(S, c->ev_f, 0));
  if (binary) {
    CUDA_TRY(cudaStreamWaitEvent(S, c->ev_u, 0));
    CUDA_TRY(cudaStreamWaitEvent(S, c->ev_phi, 0));
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}
