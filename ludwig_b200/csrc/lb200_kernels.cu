// lb200_kernels.cu -- hand-written sm_100a kernels for Ludwig's LB hot path.
//
// Compiled twice (see __graft_entry__.build / Makefile):
//   default        -> lb200_kernels_fast    (FMA contraction on)
//   -DLB200_STRICT -> lb200_kernels_strict  (-fmad=false: every operation rounded as on the CPU,
//                     written in the reference's operation order => bit-identical results)
//
// Device layout: structure of arrays on the reference's allocated lattice, z fastest
// (site index = reference cs_index, src/coords.c:617-631), one thread per lattice site with
// threadIdx.x along z so every global access of a warp is one contiguous 256-byte run.
// Everything here is HBM-bandwidth bound FP64 work: no tensor cores by design.
//
// Reference behaviour each kernel reproduces is cited at the kernel (paths relative to the
// reference root).

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <cuda.h>
#include "lb200_kernels.h"
#include "d3q19_proj.cuh"

#ifdef LB200_STRICT
#define LB200_TABLE lb200_kernels_strict
namespace lb200_strict {
#else
#define LB200_TABLE lb200_kernels_fast
namespace lb200_fast {
#endif

namespace {

constexpr int TPB = 128;
constexpr int TPB_MAX = 256;
constexpr int LB200_MAX_DEVICES = 64;     // per-device launch state (function attributes, SM counts)

// threads per block of the site-per-thread kernels; LB200_TPB (64/128/256) overrides for tuning runs
__host__ inline int tuned_tpb() {
  static int tpb = 0;
  if (tpb == 0) {
    const char * e = getenv("LB200_TPB");
    tpb = e ? atoi(e) : TPB;
    if (tpb != 64 && tpb != 128 && tpb != 256) tpb = TPB;
  }
  return tpb;
}

__host__ inline int tuned_flag(const char * name, int dflt) {
  const char * e = getenv(name);
  return e ? atoi(e) : dflt;
}

__host__ inline void block_shape_n(int nz, int tpb, dim3 & blk) {
  int bx = ((nz + 31)/32)*32;
  if (bx > tpb) bx = tpb;
  blk = dim3(bx, tpb/bx, 1);
}

__host__ inline void block_shape(int nz, dim3 & blk) { block_shape_n(nz, tuned_tpb(), blk); }

// streaming (evict-first) store: every population is written once per step and not re-read before
// 2.8 GB of other traffic has passed through the 126 MB L2
template <bool STREAM>
__device__ __forceinline__ void store_f(double * p, double v) {
  if (STREAM) __stcs(p, v); else *p = v;
}

// D3Q19 velocity set, reference src/lb_d3q19.h:26-39
__device__ constexpr int CV19[19][3] = {
  { 0,  0,  0},
  { 1,  1,  0}, { 1,  0,  1}, { 1,  0,  0}, { 1,  0, -1}, { 1, -1,  0}, { 0,  1,  1},
  { 0,  1,  0}, { 0,  1, -1}, { 0,  0,  1}, { 0,  0, -1}, { 0, -1,  1}, { 0, -1,  0},
  { 0, -1, -1}, {-1,  1,  0}, {-1,  0,  1}, {-1,  0,  0}, {-1,  0, -1}, {-1, -1,  0}};

// ---------------------------------------------------------------------------------------------
// Collision (single distribution), reference src/collision.c:253-593.  Shared by the unrolled
// D3Q19 path and the generic path: takes the 10 hydrodynamic modes + force, returns the relaxed
// hydrodynamic modes and writes rho, u.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void relax_hydro(double * __restrict__ mode, const double force[3],
					    const Lb200CollideDev & cp, double & rho_out,
					    double u[3]) {
  const double rdim = (1.0/3);
  const double rho = mode[0];
  const double rrho = 1.0/rho;

  for (int ia = 0; ia < 3; ia++) u[ia] = rrho*(mode[1 + ia] + 0.5*force[ia]);

  // stress, upper triangle xx xy xz yy yz zz <-> modes 4..9
  double sxx = mode[4], sxy = mode[5], sxz = mode[6], syy = mode[7], syz = mode[8], szz = mode[9];
  double qxx = rho*u[0]*u[0], qxy = rho*u[0]*u[1], qxz = rho*u[0]*u[2];
  double qyy = rho*u[1]*u[1], qyz = rho*u[1]*u[2], qzz = rho*u[2]*u[2];

  double tr_s = sxx + syy + szz;
  double tr_seq = qxx + qyy + qzz;

  sxx -= rdim*tr_s;  syy -= rdim*tr_s;  szz -= rdim*tr_s;
  qxx -= rdim*tr_seq; qyy -= rdim*tr_seq; qzz -= rdim*tr_seq;

  tr_s = tr_s - cp.rtau_bulk*(tr_s - tr_seq);

  sxx -= cp.rtau*(sxx - qxx); sxx += rdim*tr_s; sxx += cp.tmr*(u[0]*force[0] + force[0]*u[0]);
  sxy -= cp.rtau*(sxy - qxy);                   sxy += cp.tmr*(u[0]*force[1] + force[0]*u[1]);
  sxz -= cp.rtau*(sxz - qxz);                   sxz += cp.tmr*(u[0]*force[2] + force[0]*u[2]);
  syy -= cp.rtau*(syy - qyy); syy += rdim*tr_s; syy += cp.tmr*(u[1]*force[1] + force[1]*u[1]);
  syz -= cp.rtau*(syz - qyz);                   syz += cp.tmr*(u[1]*force[2] + force[1]*u[2]);
  szz -= cp.rtau*(szz - qzz); szz += rdim*tr_s; szz += cp.tmr*(u[2]*force[2] + force[2]*u[2]);

  for (int ia = 0; ia < 3; ia++) mode[1 + ia] += force[ia];
  mode[4] = sxx; mode[5] = sxy; mode[6] = sxz; mode[7] = syy; mode[8] = syz; mode[9] = szz;
  rho_out = rho;
}

// One thread per interior site.  PULL: read the 19 populations from the upwind neighbours of
// fsrc (lb_propagation, src/propagation.c:153-200) and write the post-collision state to fdst,
// so each population is read once and written once per time step.  !PULL: in-place collision.
// WRAP (whole time steps on periodic lattices, lb200_step): the periodic images are read straight from
// the interior sites they mirror instead of from a halo shell filled by a separate kernel -- the halo
// swap of lb_halo for the local dimensions costs nothing.  Dimensions with g.wrap[d] == 0 (x with
// slab neighbours on other GPUs) still read the halo planes, which the exchange has filled.
template <bool PULL, bool GHOST, bool HAS_FORCE, bool HAS_MAP, bool STREAM, bool WRAP>
__global__ void __launch_bounds__(TPB_MAX, GHOST ? 3 : 4)
collide_d3q19_kernel(const Lb200Geom g, const Lb200CollideDev cp,
		     const double * __restrict__ fsrc, double * __restrict__ fdst,
		     const double * __restrict__ hforce, const char * __restrict__ status,
		     double * __restrict__ rho_out, double * __restrict__ u_out) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + g.xoff + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;

  double f[19];
  double mode[19];
  double force[3];
  double u[3];
  double rho;

  if (PULL && WRAP) {
    // offset of the site one step DOWN (m) / UP (p) each axis, through the periodic boundary if need be
    const int oxm = (g.wrap[0] && ic == 1)       ?  (g.nl[0] - 1)*g.xs : -g.xs;
    const int oxp = (g.wrap[0] && ic == g.nl[0]) ? -(g.nl[0] - 1)*g.xs :  g.xs;
    const int oym = (g.wrap[1] && jc == 1)       ?  (g.nl[1] - 1)*g.ys : -g.ys;
    const int oyp = (g.wrap[1] && jc == g.nl[1]) ? -(g.nl[1] - 1)*g.ys :  g.ys;
    const int ozm = (g.wrap[2] && kc == 1)       ?  (g.nl[2] - 1) : -1;
    const int ozp = (g.wrap[2] && kc == g.nl[2]) ? -(g.nl[2] - 1) :  1;
#pragma unroll
    for (int p = 0; p < 19; p++) {
      // population p arrives from the site at -c_p
      const int off = (CV19[p][0] > 0 ? oxm : CV19[p][0] < 0 ? oxp : 0)
	+ (CV19[p][1] > 0 ? oym : CV19[p][1] < 0 ? oyp : 0)
	+ (CV19[p][2] > 0 ? ozm : CV19[p][2] < 0 ? ozp : 0);
      f[p] = fsrc[p*ns + (index + off)];
    }
  }
  else {
#pragma unroll
    for (int p = 0; p < 19; p++) {
      const int off = PULL ? (CV19[p][0]*g.xs + CV19[p][1]*g.ys + CV19[p][2]) : 0;
      f[p] = fsrc[p*ns + (index - off)];
    }
  }

  if (HAS_MAP) {
    if (status[index] != 0) {
      // non-fluid site: propagation still moves the populations; no collision, no rho/u
      if (PULL) {
#pragma unroll
	for (int p = 0; p < 19; p++) fdst[p*ns + index] = f[p];
      }
      if (ic == 1 && g.peer_f_lo != nullptr) {
#pragma unroll
	for (int p = 0; p < 19; p++) if (CV19[p][0] < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = f[p];
      }
      if (ic == g.nl[0] && g.peer_f_hi != nullptr) {
#pragma unroll
	for (int p = 0; p < 19; p++) if (CV19[p][0] > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = f[p];
      }
      return;
    }
  }

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    force[ia] = HAS_FORCE ? (cp.fg[ia] + hforce[ia*ns + index]) : (cp.fg[ia] + 0.0);
  }

  d3q19_f2mode<GHOST>(f, mode);
  relax_hydro(mode, force, cp, rho, u);

  if (GHOST) {
#pragma unroll
    for (int m = 10; m < 19; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);
  }

  d3q19_mode2f<GHOST>(mode, f);

#pragma unroll
  for (int p = 0; p < 19; p++) store_f<STREAM>(fdst + p*ns + index, f[p]);

  if (!g.skip_diag) rho_out[index] = rho;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];

  // boundary planes straight into the neighbour GPUs' halo planes (NVLink peer stores): what their next
  // pull-collision and phi sector read from this slab (block-uniform branches, 2 of N planes)
  if (ic == 1 && g.peer_f_lo != nullptr) {
    const size_t dst = (size_t) index + (size_t) g.nl[0]*g.xs;            // my plane 1 -> its plane N+1
#pragma unroll
    for (int p = 0; p < 19; p++) if (CV19[p][0] < 0) g.peer_f_lo[p*ns + dst] = f[p];
    if (g.peer_u_lo != nullptr) g.peer_u_lo[dst] = u[0];
  }
  if (ic == g.nl[0] && g.peer_f_hi != nullptr) {
    const size_t dst = (size_t) index - (size_t) g.nl[0]*g.xs;            // my plane N -> its plane 0
#pragma unroll
    for (int p = 0; p < 19; p++) if (CV19[p][0] > 0) g.peer_f_hi[p*ns + dst] = f[p];
    if (g.peer_u_hi != nullptr) g.peer_u_hi[dst] = u[0];
  }
}


// ---------------------------------------------------------------------------------------------
// FP32 STORAGE of the distributions (not the reference's arithmetic: an
// opt-in mode with a stated error bound).  Inside lb200_step the two distribution arrays hold
// d_p = float(f_p - w_p), the deviation from the rest-state weights, so the 24-bit significand is spent on
// the O(Ma) part of f_p; every load widens to FP64 and adds w_p back, the collision itself is the FP64
// collision above, every store rounds f_p - w_p to float once.  Rounding per population per step:
// <= 2^-24 |f_p - w_p|.  Bytes per site: 19 x 4 x 2 + 56 = 208 instead of 360.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ constexpr double w19(int p) {
  return (CV19[p][0]*CV19[p][0] + CV19[p][1]*CV19[p][1] + CV19[p][2]*CV19[p][2] == 0) ? (12.0/36.0)
    : (CV19[p][0]*CV19[p][0] + CV19[p][1]*CV19[p][1] + CV19[p][2]*CV19[p][2] == 1) ? (2.0/36.0) : (1.0/36.0);
}

__global__ void __launch_bounds__(TPB_MAX)
f_convert_kernel(size_t ns, int to_f32, double * __restrict__ f64, float * __restrict__ f32) {
  const size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= ns) return;
#pragma unroll
  for (int p = 0; p < 19; p++) {
    if (to_f32) f32[p*ns + i] = (float) (f64[p*ns + i] - w19(p));
    else        f64[p*ns + i] = w19(p) + (double) f32[p*ns + i];
  }
}

int launch_f_convert(cudaStream_t st, const Lb200Geom & g, int to_f32, double * f64, float * f32) {
  const size_t ns = (size_t) g.nsites;
  f_convert_kernel<<<(unsigned int) ((ns + TPB_MAX - 1)/TPB_MAX), TPB_MAX, 0, st>>>(ns, to_f32, f64, f32);
  return 1;
}

template <bool GHOST, bool HAS_FORCE>
__global__ void __launch_bounds__(TPB_MAX, GHOST ? 3 : 4)
collide_d3q19_f32_kernel(const Lb200Geom g, const Lb200CollideDev cp,
			 const float * __restrict__ fsrc, float * __restrict__ fdst,
			 const double * __restrict__ hforce,
			 double * __restrict__ rho_out, double * __restrict__ u_out) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;

  double f[19];
  double mode[19];
  double force[3];
  double u[3];
  double rho;

  // pull, periodic images read from the interior (as collide_d3q19_kernel<PULL, ..., WRAP>)
  const int oxm = (ic == 1)       ?  (g.nl[0] - 1)*g.xs : -g.xs;
  const int oxp = (ic == g.nl[0]) ? -(g.nl[0] - 1)*g.xs :  g.xs;
  const int oym = (jc == 1)       ?  (g.nl[1] - 1)*g.ys : -g.ys;
  const int oyp = (jc == g.nl[1]) ? -(g.nl[1] - 1)*g.ys :  g.ys;
  const int ozm = (kc == 1)       ?  (g.nl[2] - 1) : -1;
  const int ozp = (kc == g.nl[2]) ? -(g.nl[2] - 1) :  1;
#pragma unroll
  for (int p = 0; p < 19; p++) {
    const int off = (CV19[p][0] > 0 ? oxm : CV19[p][0] < 0 ? oxp : 0)
      + (CV19[p][1] > 0 ? oym : CV19[p][1] < 0 ? oyp : 0)
      + (CV19[p][2] > 0 ? ozm : CV19[p][2] < 0 ? ozp : 0);
    f[p] = w19(p) + (double) fsrc[p*ns + (index + off)];
  }

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    force[ia] = HAS_FORCE ? (cp.fg[ia] + hforce[ia*ns + index]) : (cp.fg[ia] + 0.0);
  }

  d3q19_f2mode<GHOST>(f, mode);
  relax_hydro(mode, force, cp, rho, u);
  if (GHOST) {
#pragma unroll
    for (int m = 10; m < 19; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);
  }
  d3q19_mode2f<GHOST>(mode, f);

#pragma unroll
  for (int p = 0; p < 19; p++) __stcs(fdst + p*ns + index, (float) (f[p] - w19(p)));

  if (!g.skip_diag) rho_out[index] = rho;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];
}

int launch_collide_f32(cudaStream_t st, const Lb200Geom & g, const Lb200CollideDev & cp, const float * fsrc,
		       float * fdst, const double * force, double * rho, double * u) {
  dim3 blk;
  block_shape_n(g.nl[2], 256, blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
  if (cp.ghost) {
    if (force) collide_d3q19_f32_kernel<true, true><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, rho, u);
    else       collide_d3q19_f32_kernel<true, false><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, rho, u);
  }
  else {
    if (force) collide_d3q19_f32_kernel<false, true><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, rho, u);
    else       collide_d3q19_f32_kernel<false, false><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, rho, u);
  }
  return 1;
}

// Generic velocity set (D3Q15, D3Q27; also D3Q19 with the model matrices instead of the coded
// constants): reference src/collision.c:335-342, 541-551.
template <bool PULL, bool WRAP>
__global__ void __launch_bounds__(TPB_MAX)
collide_generic_kernel(const Lb200Geom g, const Lb200CollideDev cp,
		       const Lb200ModelDev * __restrict__ md,
		       const double * __restrict__ fsrc, double * __restrict__ fdst,
		       const double * __restrict__ hforce, const char * __restrict__ status,
		       double * __restrict__ rho_out, double * __restrict__ u_out) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  const int nvel = md->nvel;

  double f[27];
  double mode[27];
  double force[3];
  double u[3];
  double rho;

  if (PULL && WRAP) {
    int o[3][3];          // o[axis][1 - c]: offset of the site at -c along the axis
    o[0][0] = (g.wrap[0] && ic == 1)       ?  (g.nl[0] - 1)*g.xs : -g.xs;
    o[0][2] = (g.wrap[0] && ic == g.nl[0]) ? -(g.nl[0] - 1)*g.xs :  g.xs;
    o[1][0] = (g.wrap[1] && jc == 1)       ?  (g.nl[1] - 1)*g.ys : -g.ys;
    o[1][2] = (g.wrap[1] && jc == g.nl[1]) ? -(g.nl[1] - 1)*g.ys :  g.ys;
    o[2][0] = (g.wrap[2] && kc == 1)       ?  (g.nl[2] - 1) : -1;
    o[2][2] = (g.wrap[2] && kc == g.nl[2]) ? -(g.nl[2] - 1) :  1;
    o[0][1] = o[1][1] = o[2][1] = 0;
    for (int p = 0; p < nvel; p++) {
      const int off = o[0][1 - md->cv[p][0]] + o[1][1 - md->cv[p][1]] + o[2][1 - md->cv[p][2]];
      f[p] = fsrc[p*ns + (index + off)];
    }
  }
  else {
    for (int p = 0; p < nvel; p++) {
      const int off = PULL ? (md->cv[p][0]*g.xs + md->cv[p][1]*g.ys + md->cv[p][2]) : 0;
      f[p] = fsrc[p*ns + (index - off)];
    }
  }

  if (status != nullptr && status[index] != 0) {
    if (PULL) {
      for (int p = 0; p < nvel; p++) fdst[p*ns + index] = f[p];
    }
    for (int p = 0; p < nvel; p++) {
      if (ic == 1 && g.peer_f_lo != nullptr && md->cv[p][0] < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = f[p];
      if (ic == g.nl[0] && g.peer_f_hi != nullptr && md->cv[p][0] > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = f[p];
    }
    return;
  }

  for (int ia = 0; ia < 3; ia++) {
    force[ia] = cp.fg[ia] + (hforce ? hforce[ia*ns + index] : 0.0);
  }

  for (int m = 0; m < nvel; m++) {
    double s = 0.0;
    for (int p = 0; p < nvel; p++) s += f[p]*md->ma[m][p];
    mode[m] = s;
  }

  relax_hydro(mode, force, cp, rho, u);

  for (int m = 10; m < nvel; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);

  const bool to_lo = (ic == 1 && g.peer_f_lo != nullptr), to_hi = (ic == g.nl[0] && g.peer_f_hi != nullptr);
  for (int p = 0; p < nvel; p++) {
    double s = 0.0;
    for (int m = 0; m < nvel; m++) s += md->mi[p][m]*mode[m];
    fdst[p*ns + index] = s;
    if (to_lo && md->cv[p][0] < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = s;
    if (to_hi && md->cv[p][0] > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = s;
  }

  rho_out[index] = rho;
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];
  if (to_lo && g.peer_u_lo != nullptr) g.peer_u_lo[(size_t) index + (size_t) g.nl[0]*g.xs] = u[0];
  if (to_hi && g.peer_u_hi != nullptr) g.peer_u_hi[(size_t) index - (size_t) g.nl[0]*g.xs] = u[0];
}

// The same operations with the velocity-set size fixed at compile time (D3Q15, D3Q27): the populations and modes stay
// in registers (the run-time-sized kernel above indexes them dynamically, i.e. in local memory) and the model matrices
// travel in the kernel parameter space (constant bank: every thread reads the same entry, with immediate offsets).
// Loop order unchanged (m outer, p inner; p outer, m inner), so the results are those of the generic kernel bit for bit.
// 3.7 -> 12+ GLUPS for D3Q15, 1.3 -> 5+ GLUPS for D3Q27 at 256^3 (tools/bench_models.py).
template <int NVEL, bool PULL, bool WRAP>
__global__ void __launch_bounds__(TPB_MAX)
collide_model_kernel(const Lb200Geom g, const Lb200CollideDev cp, const __grid_constant__ Lb200ModelDev md,
		     const double * __restrict__ fsrc, double * __restrict__ fdst,
		     const double * __restrict__ hforce, const char * __restrict__ status,
		     double * __restrict__ rho_out, double * __restrict__ u_out) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;

  double f[NVEL];
  double mode[NVEL];
  double force[3];
  double u[3];
  double rho;

  if (PULL && WRAP) {
    const int oxm = (g.wrap[0] && ic == 1)       ?  (g.nl[0] - 1)*g.xs : -g.xs;
    const int oxp = (g.wrap[0] && ic == g.nl[0]) ? -(g.nl[0] - 1)*g.xs :  g.xs;
    const int oym = (g.wrap[1] && jc == 1)       ?  (g.nl[1] - 1)*g.ys : -g.ys;
    const int oyp = (g.wrap[1] && jc == g.nl[1]) ? -(g.nl[1] - 1)*g.ys :  g.ys;
    const int ozm = (g.wrap[2] && kc == 1)       ?  (g.nl[2] - 1) : -1;
    const int ozp = (g.wrap[2] && kc == g.nl[2]) ? -(g.nl[2] - 1) :  1;
#pragma unroll
    for (int p = 0; p < NVEL; p++) {
      // c in {-1, 0, 1}: (c + 1) >> 1 selects the upwind offset of c = +1, (1 - c) >> 1 that of c = -1 -- integer
      // arithmetic, so that the 15 / 27 loads are issued back to back (ptxas compiles the ternary form as one
      // uniform branch per population when it feels like it: D3Q15 lost 5 %)
      const int cx = md.cv[p][0], cy = md.cv[p][1], cz = md.cv[p][2];
      const int off = ((cx + 1) >> 1)*oxm + ((1 - cx) >> 1)*oxp + ((cy + 1) >> 1)*oym + ((1 - cy) >> 1)*oyp
	+ ((cz + 1) >> 1)*ozm + ((1 - cz) >> 1)*ozp;
      f[p] = fsrc[p*ns + (index + off)];
    }
  }
  else {
#pragma unroll
    for (int p = 0; p < NVEL; p++) {
      const int off = PULL ? (md.cv[p][0]*g.xs + md.cv[p][1]*g.ys + md.cv[p][2]) : 0;
      f[p] = fsrc[p*ns + (index - off)];
    }
  }

  if (status != nullptr && status[index] != 0) {
    if (PULL) {
#pragma unroll
      for (int p = 0; p < NVEL; p++) fdst[p*ns + index] = f[p];
    }
#pragma unroll
    for (int p = 0; p < NVEL; p++) {
      if (ic == 1 && g.peer_f_lo != nullptr && md.cv[p][0] < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = f[p];
      if (ic == g.nl[0] && g.peer_f_hi != nullptr && md.cv[p][0] > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = f[p];
    }
    return;
  }

#pragma unroll
  for (int ia = 0; ia < 3; ia++) force[ia] = cp.fg[ia] + (hforce ? hforce[ia*ns + index] : 0.0);

#pragma unroll
  for (int m = 0; m < NVEL; m++) {
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < NVEL; p++) s += f[p]*md.ma[m][p];
    mode[m] = s;
  }

  relax_hydro(mode, force, cp, rho, u);

#pragma unroll
  for (int m = 10; m < NVEL; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);

  const bool to_lo = (ic == 1 && g.peer_f_lo != nullptr), to_hi = (ic == g.nl[0] && g.peer_f_hi != nullptr);
#pragma unroll
  for (int p = 0; p < NVEL; p++) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < NVEL; m++) s += md.mi[p][m]*mode[m];
    store_f<PULL>(fdst + p*ns + index, s);
    if (to_lo && md.cv[p][0] < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = s;
    if (to_hi && md.cv[p][0] > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = s;
  }

  rho_out[index] = rho;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];
  if (to_lo && g.peer_u_lo != nullptr) g.peer_u_lo[(size_t) index + (size_t) g.nl[0]*g.xs] = u[0];
  if (to_hi && g.peer_u_hi != nullptr) g.peer_u_hi[(size_t) index - (size_t) g.nl[0]*g.xs] = u[0];
}

#ifndef LB200_STRICT
// ---------------------------------------------------------------------------------------------
// D3Q27, fast mode: the reference's D3Q27 basis (src/lb_d3q27.c:100-190) is the complete tensor-product Hermite basis
// m_abc = k_abc sum_p h_a(c_px) h_b(c_py) h_c(c_pz) f_p with h_0 = 1, h_1 = c, h_2 = c^2 - 1/3 and k = 1, 3, 9, 27 for
// polynomial order per axis (see D27_MODE), so both projections factorise into three passes of 3-point transforms:
// 2 x 243 multiply-adds per site instead of the 2 x 729 of the dense matrix products (which make the generic kernel
// FP64-bound: 4.9 GLUPS).  Same relaxation as every other collision kernel; within 1e-12 of the dense form (tested).
// The back projection uses mi[p][m] = w_p N_m ma[m][p], N_m = 1/(k_m^2 n_a n_b n_c), n = (1, 1/3, 2/9), w_p = w(c_x) w(c_y) w(c_z).
// ---------------------------------------------------------------------------------------------

// population index of velocity (i, j, k): rest first, then i, j, k ascending with the rest skipped (src/lb_d3q27.h:26-32)
__host__ __device__ constexpr int d27_p(int i, int j, int k) {
  const int lin = (i + 1)*9 + (j + 1)*3 + (k + 1);          // 0..26, the rest velocity at 13
  return (lin == 13) ? 0 : (lin < 13 ? lin + 1 : lin);
}
// mode m of the reference <-> polynomial orders (a, b, c) along x, y, z and scale k (src/lb_d3q27.c:160-186)
__device__ constexpr int D27_MODE[27][4] = {
  {0, 0, 0, 1}, {1, 0, 0, 1}, {0, 1, 0, 1}, {0, 0, 1, 1},
  {2, 0, 0, 1}, {1, 1, 0, 1}, {1, 0, 1, 1}, {0, 2, 0, 1}, {0, 1, 1, 1}, {0, 0, 2, 1},
  {2, 1, 0, 3}, {2, 0, 1, 3}, {0, 2, 1, 3}, {1, 2, 0, 3}, {1, 0, 2, 3}, {0, 1, 2, 3},
  {1, 1, 1, 1},
  {2, 2, 0, 9}, {0, 2, 2, 9}, {2, 0, 2, 9},
  {2, 1, 1, 9}, {1, 2, 1, 9}, {1, 1, 2, 9},
  {2, 2, 1, 9}, {1, 2, 2, 9}, {2, 1, 2, 9},
  {2, 2, 2, 27}};

// forward 3-point transform along one axis: values at c = -1, 0, +1 -> orders 0, 1, 2
__device__ __forceinline__ void d27_fwd(double fm, double f0, double fp, double & t0, double & t1, double & t2) {
  const double s = fm + fp;
  t0 = s + f0;
  t1 = fp - fm;
  t2 = (2.0/3.0)*s - (1.0/3.0)*f0;
}
// transposed transform with the 1-d weights w(0) = 2/3, w(+-1) = 1/6 folded in
__device__ __forceinline__ void d27_bwd(double t0, double t1, double t2, double & fm, double & f0, double & fp) {
  const double e = t0 + (2.0/3.0)*t2;
  fm = (1.0/6.0)*(e - t1);
  fp = (1.0/6.0)*(e + t1);
  f0 = (2.0/3.0)*(t0 - (1.0/3.0)*t2);
}

template <bool PULL, bool WRAP>
__global__ void __launch_bounds__(TPB_MAX)
collide_d3q27_sep_kernel(const Lb200Geom g, const Lb200CollideDev cp,
			 const double * __restrict__ fsrc, double * __restrict__ fdst,
			 const double * __restrict__ hforce, const char * __restrict__ status,
			 double * __restrict__ rho_out, double * __restrict__ u_out) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;

  // offsets of the site one step down / up each axis (through the periodic boundary in halo-free steps)
  int om[3], op[3];
  om[0] = (WRAP && g.wrap[0] && ic == 1)       ?  (g.nl[0] - 1)*g.xs : -g.xs;
  op[0] = (WRAP && g.wrap[0] && ic == g.nl[0]) ? -(g.nl[0] - 1)*g.xs :  g.xs;
  om[1] = (WRAP && g.wrap[1] && jc == 1)       ?  (g.nl[1] - 1)*g.ys : -g.ys;
  op[1] = (WRAP && g.wrap[1] && jc == g.nl[1]) ? -(g.nl[1] - 1)*g.ys :  g.ys;
  om[2] = (WRAP && g.wrap[2] && kc == 1)       ?  (g.nl[2] - 1) : -1;
  op[2] = (WRAP && g.wrap[2] && kc == g.nl[2]) ? -(g.nl[2] - 1) :  1;

  double f[3][3][3];                     // f[i+1][j+1][k+1], velocity (i, j, k)
#pragma unroll
  for (int i = -1; i <= 1; i++)
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
      for (int k = -1; k <= 1; k++) {
	// population (i, j, k) arrives from the site at -(i, j, k)
	const int off = PULL ? ((i > 0 ? om[0] : i < 0 ? op[0] : 0) + (j > 0 ? om[1] : j < 0 ? op[1] : 0) + (k > 0 ? om[2] : k < 0 ? op[2] : 0)) : 0;
	f[i + 1][j + 1][k + 1] = fsrc[d27_p(i, j, k)*ns + (index + off)];
      }

  const bool to_lo = (ic == 1 && g.peer_f_lo != nullptr), to_hi = (ic == g.nl[0] && g.peer_f_hi != nullptr);

  if (status != nullptr && status[index] != 0) {
#pragma unroll
    for (int i = -1; i <= 1; i++)
#pragma unroll
      for (int j = -1; j <= 1; j++)
#pragma unroll
	for (int k = -1; k <= 1; k++) {
	  const int p = d27_p(i, j, k);
	  const double v = f[i + 1][j + 1][k + 1];
	  if (PULL) fdst[p*ns + index] = v;
	  if (to_lo && i < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = v;
	  if (to_hi && i > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = v;
	}
    return;
  }

  // forward: z, then y, then x
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) d27_fwd(f[i][j][0], f[i][j][1], f[i][j][2], f[i][j][0], f[i][j][1], f[i][j][2]);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) d27_fwd(f[i][0][k], f[i][1][k], f[i][2][k], f[i][0][k], f[i][1][k], f[i][2][k]);
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int k = 0; k < 3; k++) d27_fwd(f[0][j][k], f[1][j][k], f[2][j][k], f[0][j][k], f[1][j][k], f[2][j][k]);

  double mode[27];
#pragma unroll
  for (int m = 0; m < 27; m++) mode[m] = (double) D27_MODE[m][3]*f[D27_MODE[m][0]][D27_MODE[m][1]][D27_MODE[m][2]];

  double force[3], u[3], rho;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) force[ia] = cp.fg[ia] + (hforce ? hforce[ia*ns + index] : 0.0);
  relax_hydro(mode, force, cp, rho, u);
#pragma unroll
  for (int m = 10; m < 27; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);

  // back: N_m k_m mode_m into the (a, b, c) cube, then x, y, z with the weights folded in
  const double nn[3] = {1.0, 3.0, 4.5};                  // 1/n_a, n = (1, 1/3, 2/9)
#pragma unroll
  for (int m = 0; m < 27; m++) {
    const int a = D27_MODE[m][0], b = D27_MODE[m][1], c = D27_MODE[m][2];
    f[a][b][c] = mode[m]*(nn[a]*nn[b]*nn[c]/(double) D27_MODE[m][3]);
  }
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int k = 0; k < 3; k++) d27_bwd(f[0][j][k], f[1][j][k], f[2][j][k], f[0][j][k], f[1][j][k], f[2][j][k]);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) d27_bwd(f[i][0][k], f[i][1][k], f[i][2][k], f[i][0][k], f[i][1][k], f[i][2][k]);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) d27_bwd(f[i][j][0], f[i][j][1], f[i][j][2], f[i][j][0], f[i][j][1], f[i][j][2]);

#pragma unroll
  for (int i = -1; i <= 1; i++)
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
      for (int k = -1; k <= 1; k++) {
	const int p = d27_p(i, j, k);
	const double v = f[i + 1][j + 1][k + 1];
	store_f<PULL>(fdst + p*ns + index, v);
	if (to_lo && i < 0) g.peer_f_lo[p*ns + (size_t) index + (size_t) g.nl[0]*g.xs] = v;
	if (to_hi && i > 0) g.peer_f_hi[p*ns + (size_t) index - (size_t) g.nl[0]*g.xs] = v;
      }

  rho_out[index] = rho;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];
  if (to_lo && g.peer_u_lo != nullptr) g.peer_u_lo[(size_t) index + (size_t) g.nl[0]*g.xs] = u[0];
  if (to_hi && g.peer_u_hi != nullptr) g.peer_u_hi[(size_t) index - (size_t) g.nl[0]*g.xs] = u[0];
}

// the separable form assumes the velocity ordering above: check the model tables once
static bool d27_ordering_ok(const Lb200ModelDev & mh) {
  for (int i = -1; i <= 1; i++)
    for (int j = -1; j <= 1; j++)
      for (int k = -1; k <= 1; k++) {
	const int p = d27_p(i, j, k);
	if (mh.cv[p][0] != i || mh.cv[p][1] != j || mh.cv[p][2] != k) return false;
      }
  return true;
}
#endif

// host copy of the model tables of a velocity set (deterministic per nvel), fetched once from the device copy
static const Lb200ModelDev * host_model(const Lb200ModelDev * md_dev, int nvel) {
  static Lb200ModelDev cache[3];
  static bool have[3] = {false, false, false};
  const int slot = (nvel == 15) ? 0 : (nvel == 19) ? 1 : 2;
  if (!have[slot]) {
    if (cudaMemcpy(&cache[slot], md_dev, sizeof(Lb200ModelDev), cudaMemcpyDeviceToHost) != cudaSuccess) return nullptr;
    have[slot] = true;
  }
  return &cache[slot];
}

int launch_collide(cudaStream_t st, const Lb200Geom & g, const Lb200CollideDev & cp,
		   const Lb200ModelDev * md, int nvel, int pull, const double * fsrc,
		   double * fdst, const double * force, const char * status,
		   double * rho, double * u) {
  dim3 blk;
  static const int ctpb = tuned_flag("LB200_COLLIDE_TPB", 256);
  block_shape_n(g.nl[2], (ctpb == 64 || ctpb == 128 || ctpb == 256) ? ctpb : 256, blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);

  if (nvel == 19 && md == nullptr) {
    if (g.xcnt > 0) grd.z = g.xcnt;                 // slab pipeline: planes xoff+1 .. xoff+xcnt
    static const int stream_stores = tuned_flag("LB200_STCS", 1);
    const bool wrap = pull && (g.wrap[0] || g.wrap[1] || g.wrap[2]);
#define LB200_GO(P, G, F, M) do { if (wrap && P) collide_d3q19_kernel<P, G, F, M, true, P><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u); \
    else if (stream_stores && P) collide_d3q19_kernel<P, G, F, M, true, false><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u); \
    else collide_d3q19_kernel<P, G, F, M, false, false><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u); } while (0)
#define LB200_SEL_M(P, G, F) do { if (status) LB200_GO(P, G, F, true); else LB200_GO(P, G, F, false); } while (0)
#define LB200_SEL_F(P, G) do { if (force) LB200_SEL_M(P, G, true); else LB200_SEL_M(P, G, false); } while (0)
#define LB200_SEL_G(P) do { if (cp.ghost) LB200_SEL_F(P, true); else LB200_SEL_F(P, false); } while (0)
    if (pull) LB200_SEL_G(true); else LB200_SEL_G(false);
#undef LB200_GO
#undef LB200_SEL_M
#undef LB200_SEL_F
#undef LB200_SEL_G
  }
  else if ((nvel == 15 || nvel == 27) && tuned_flag("LB200_MODEL_KERNEL", 1) && host_model(md, nvel) != nullptr) {
    const bool wrap = pull && (g.wrap[0] || g.wrap[1] || g.wrap[2]);
    const Lb200ModelDev & mh = *host_model(md, nvel);
    dim3 blk2;
    block_shape_n(g.nl[2], 128, blk2);
    dim3 grd2((g.nl[2] + blk2.x - 1)/blk2.x, (g.nl[1] + blk2.y - 1)/blk2.y, g.nl[0]);
#ifndef LB200_STRICT
    static const int sep27 = tuned_flag("LB200_D3Q27_SEPARABLE", 1);
    if (nvel == 27 && sep27 && d27_ordering_ok(mh)) {
      if (wrap)      collide_d3q27_sep_kernel<true, true><<<grd2, blk2, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u);
      else if (pull) collide_d3q27_sep_kernel<true, false><<<grd2, blk2, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u);
      else           collide_d3q27_sep_kernel<false, false><<<grd2, blk2, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u);
      return 1;
    }
#endif
#define LB200_GO(N) do { \
    if (wrap)      collide_model_kernel<N, true, true><<<grd2, blk2, 0, st>>>(g, cp, mh, fsrc, fdst, force, status, rho, u); \
    else if (pull) collide_model_kernel<N, true, false><<<grd2, blk2, 0, st>>>(g, cp, mh, fsrc, fdst, force, status, rho, u); \
    else           collide_model_kernel<N, false, false><<<grd2, blk2, 0, st>>>(g, cp, mh, fsrc, fdst, force, status, rho, u); } while (0)
    if (nvel == 15) LB200_GO(15); else LB200_GO(27);
#undef LB200_GO
  }
  else {
    const bool wrap = pull && (g.wrap[0] || g.wrap[1] || g.wrap[2]);
    if (wrap)      collide_generic_kernel<true, true><<<grd, blk, 0, st>>>(g, cp, md, fsrc, fdst, force, status, rho, u);
    else if (pull) collide_generic_kernel<true, false><<<grd, blk, 0, st>>>(g, cp, md, fsrc, fdst, force, status, rho, u);
    else           collide_generic_kernel<false, false><<<grd, blk, 0, st>>>(g, cp, md, fsrc, fdst, force, status, rho, u);
  }
  return 1;
}

// ---------------------------------------------------------------------------------------------
// symmetric_lb: two distributions (reference `free_energy symmetric_lb`).
//   phi_lb_to_field (src/phi_lb_coupler.c:39-96): phi = sum_p g_p, here optionally of the PULLED
//     populations, i.e. of the state a pending lb_propagation would produce;
//   lb_collision_binary -> lb_collision_mrt2_site (src/collision.c:604-1013): single-fluid collision with
//     the thermodynamic stress fe_symm_str_v (src/symmetric.c:371-416) in the equilibrium stress, then the
//     order-parameter distribution rebuilt from (phi, j_phi relaxed at rtau2 = 2/(1 + 2M), S_phi) by
//     d3q19_mode2f_phi (src/collision.c:2856-3135: only the non-zero terms, literal constants) or the
//     generic loop (:974-1008).  Interior sites, no status test, hydro->rho is not written (as the reference).
// ---------------------------------------------------------------------------------------------

template <bool PULL>
__global__ void __launch_bounds__(TPB_MAX)
phi_from_g_kernel(const Lb200Geom g, const Lb200ModelDev * __restrict__ md,
		  const double * __restrict__ f, double * __restrict__ phi) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  const int nvel = md->nvel;
  double phi0 = 0.0;
  for (int p = 0; p < nvel; p++) {
    const int off = PULL ? (md->cv[p][0]*g.xs + md->cv[p][1]*g.ys + md->cv[p][2]) : 0;
    phi0 += f[(size_t) (nvel + p)*ns + (index - off)];
  }
  phi[index] = phi0;
}

int launch_phi_from_g(cudaStream_t st, const Lb200Geom & g, const Lb200ModelDev * md, int pull,
		      const double * f, double * phi) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
  if (pull) phi_from_g_kernel<true><<<grd, blk, 0, st>>>(g, md, f, phi);
  else      phi_from_g_kernel<false><<<grd, blk, 0, st>>>(g, md, f, phi);
  return 1;
}

__global__ void __launch_bounds__(TPB_MAX)
phi_to_g_kernel(const Lb200Geom g, int nvel, const double * __restrict__ phi, double * __restrict__ f) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  f[(size_t) nvel*ns + index] = phi[index];
  for (int p = 1; p < nvel; p++) f[(size_t) (nvel + p)*ns + index] = 0.0;
}

int launch_phi_to_g(cudaStream_t st, const Lb200Geom & g, int nvel, const double * phi, double * f) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
  phi_to_g_kernel<<<grd, blk, 0, st>>>(g, nvel, phi, f);
  return 1;
}

// hydrodynamic relaxation of the binary collision: equilibrium stress rho u u + P^th
__device__ __forceinline__ void relax_hydro_binary(double * __restrict__ mode, const double force[3],
						   const Lb200CollideDev & cp, const double sth[3][3],
						   double u[3]) {
  const double r3 = 1.0/3.0;
  const double rho = mode[0];
  const double rrho = 1.0/rho;
  for (int ia = 0; ia < 3; ia++) u[ia] = rrho*(mode[1 + ia] + 0.5*force[ia]);

  double s[3][3], seq[3][3];
  s[0][0] = mode[4]; s[0][1] = mode[5]; s[0][2] = mode[6];
  s[1][1] = mode[7]; s[1][2] = mode[8]; s[2][2] = mode[9];

  double tr_s = 0.0, tr_seq = 0.0;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = ia; ib < 3; ib++) seq[ia][ib] = rho*u[ia]*u[ib] + sth[ia][ib];
    tr_s   += s[ia][ia];
    tr_seq += seq[ia][ia];
  }
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    s[ia][ia]   -= r3*tr_s;
    seq[ia][ia] -= r3*tr_seq;
  }
  tr_s = tr_s - cp.rtau_bulk*(tr_s - tr_seq);

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = ia; ib < 3; ib++) {
      s[ia][ib] -= cp.rtau*(s[ia][ib] - seq[ia][ib]);
      if (ia == ib) s[ia][ib] += r3*tr_s;
      s[ia][ib] += cp.tmr*(u[ia]*force[ib] + force[ia]*u[ib]);
    }
  }
  for (int ia = 0; ia < 3; ia++) mode[1 + ia] += force[ia];
  mode[4] = s[0][0]; mode[5] = s[0][1]; mode[6] = s[0][2];
  mode[7] = s[1][1]; mode[8] = s[1][2]; mode[9] = s[2][2];
}

template <bool PULL, bool GHOST, bool UNROLLED19>
__global__ void __launch_bounds__(TPB)
collide_binary_kernel(const Lb200Geom g, const Lb200CollideDev cp, const Lb200SymmDev sp,
		      const Lb200ModelDev * __restrict__ md,
		      const double * __restrict__ fsrc, double * __restrict__ fdst,
		      const double * __restrict__ hforce, const double * __restrict__ phi_,
		      const double * __restrict__ grad, const double * __restrict__ delsq_,
		      double * __restrict__ u_out) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  const int nvel = UNROLLED19 ? 19 : md->nvel;

  double force[3], u[3], gr[3], sth[3][3];
  const double phi = phi_[index];
  const double delsq = delsq_[index];
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    force[ia] = cp.fg[ia] + (hforce ? hforce[ia*ns + index] : 0.0);
    gr[ia] = grad[ia*ns + index];
  }
  {
    const double p0 = 0.5*sp.a*phi*phi + 0.75*sp.b*phi*phi*phi*phi - sp.kappa*phi*delsq
      - 0.5*sp.kappa*(gr[0]*gr[0] + gr[1]*gr[1] + gr[2]*gr[2]);
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
#pragma unroll
      for (int ib = 0; ib < 3; ib++) sth[ia][ib] = p0*((ia == ib) ? 1.0 : 0.0) + sp.kappa*gr[ia]*gr[ib];
    }
  }

  // ---- density distribution ----
  if (UNROLLED19) {
    double f[19], mode[19];
#pragma unroll
    for (int p = 0; p < 19; p++) {
      const int off = PULL ? (CV19[p][0]*g.xs + CV19[p][1]*g.ys + CV19[p][2]) : 0;
      f[p] = fsrc[p*ns + (index - off)];
    }
    d3q19_f2mode<GHOST>(f, mode);
    relax_hydro_binary(mode, force, cp, sth, u);
    if (GHOST) {
#pragma unroll
      for (int m = 10; m < 19; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);
    }
    d3q19_mode2f<GHOST>(mode, f);
#pragma unroll
    for (int p = 0; p < 19; p++) fdst[p*ns + index] = f[p];
  }
  else {
    double f[27], mode[27];
    for (int p = 0; p < nvel; p++) {
      const int off = PULL ? (md->cv[p][0]*g.xs + md->cv[p][1]*g.ys + md->cv[p][2]) : 0;
      f[p] = fsrc[p*ns + (index - off)];
    }
    for (int m = 0; m < nvel; m++) {
      double s = 0.0;
      for (int p = 0; p < nvel; p++) s += md->ma[m][p]*f[p];
      mode[m] = s;
    }
    relax_hydro_binary(mode, force, cp, sth, u);
    for (int m = 10; m < nvel; m++) mode[m] = mode[m] - cp.rtau_ghost[m]*(mode[m] - 0.0);
    for (int p = 0; p < nvel; p++) {
      double s = 0.0;
      for (int m = 0; m < nvel; m++) s += md->mi[p][m]*mode[m];
      fdst[p*ns + index] = s;
    }
  }
#pragma unroll
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];

  // ---- order-parameter distribution ----
  const double mu = sp.a*phi + sp.b*phi*phi*phi - sp.kappa*delsq;
  const double rtau2 = sp.rtau2;
  double jphi[3] = {0.0, 0.0, 0.0};
  double sphi[3][3];

  if (UNROLLED19) {
    double gp[19];
#pragma unroll
    for (int p = 0; p < 19; p++) {
      const int off = PULL ? (CV19[p][0]*g.xs + CV19[p][1]*g.ys + CV19[p][2]) : 0;
      gp[p] = fsrc[(19 + p)*ns + (index - off)];
    }
#pragma unroll
    for (int p = 1; p < 19; p++) {
#pragma unroll
      for (int ia = 0; ia < 3; ia++) {
	if (CV19[p][ia] > 0) jphi[ia] += gp[p];
	if (CV19[p][ia] < 0) jphi[ia] += -gp[p];
      }
    }
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
#pragma unroll
      for (int ib = 0; ib < 3; ib++) sphi[ia][ib] = phi*u[ia]*u[ib] + mu*((ia == ib) ? 1.0 : 0.0);
      jphi[ia] = jphi[ia] - rtau2*(jphi[ia] - phi*u[ia]);
    }
    const double q23 = 6.6666666666666663e-01, q13 = -3.3333333333333331e-01;
#pragma unroll
    for (int p = 0; p < 19; p++) {
      double jdotc = 0.0, sphidotq = 0.0;
#pragma unroll
      for (int ia = 0; ia < 3; ia++) {
	if (CV19[p][ia] > 0) jdotc += jphi[ia];
	if (CV19[p][ia] < 0) jdotc -= jphi[ia];
      }
#pragma unroll
      for (int ia = 0; ia < 3; ia++) {
#pragma unroll
	for (int ib = 0; ib < 3; ib++) {
	  const int cc = CV19[p][ia]*CV19[p][ib];
	  if (ia == ib) sphidotq += sphi[ia][ib]*(cc ? q23 : q13);
	  else if (cc > 0) sphidotq += sphi[ia][ib]*1.0;
	  else if (cc < 0) sphidotq += sphi[ia][ib]*-1.0;
	}
      }
      const double w = (p == 0) ? (12.0/36.0) : ((CV19[p][0]*CV19[p][0] + CV19[p][1]*CV19[p][1] + CV19[p][2]*CV19[p][2] == 1) ? (2.0/36.0) : (1.0/36.0));
      double v = w*(jdotc*3.0 + sphidotq*(9.0/2.0));
      if (p == 0) v = v + phi;
      fdst[(19 + p)*ns + index] = v;
    }
  }
  else {
    const double cs2 = (1.0/3.0);
    double gp[27];
    for (int p = 0; p < nvel; p++) {
      const int off = PULL ? (md->cv[p][0]*g.xs + md->cv[p][1]*g.ys + md->cv[p][2]) : 0;
      gp[p] = fsrc[(size_t) (nvel + p)*ns + (index - off)];
    }
    for (int p = 1; p < nvel; p++)
      for (int ia = 0; ia < 3; ia++) jphi[ia] += md->cv[p][ia]*gp[p];
    for (int ia = 0; ia < 3; ia++) {
      for (int ib = 0; ib < 3; ib++) sphi[ia][ib] = phi*u[ia]*u[ib] + mu*((ia == ib) ? 1.0 : 0.0);
      jphi[ia] = jphi[ia] - rtau2*(jphi[ia] - phi*u[ia]);
    }
    for (int p = 0; p < nvel; p++) {
      const int dp0 = (p == 0);
      double jdotc = 0.0, sphidotq = 0.0;
      for (int ia = 0; ia < 3; ia++) {
	jdotc += jphi[ia]*md->cv[p][ia];
	for (int ib = 0; ib < 3; ib++) {
	  sphidotq += sphi[ia][ib]*(md->cv[p][ia]*md->cv[p][ib] - cs2*((ia == ib) ? 1.0 : 0.0));
	}
      }
      fdst[(size_t) (nvel + p)*ns + index] = md->wv[p]*(jdotc*3.0 + sphidotq*4.5) + phi*dp0;
    }
  }
}

int launch_collide_binary(cudaStream_t st, const Lb200Geom & g, const Lb200CollideDev & cp,
			  const Lb200SymmDev & sp, const Lb200ModelDev * md, int unrolled19, int pull,
			  const double * fsrc, double * fdst, const double * force, const double * phi,
			  const double * grad, const double * delsq, double * u) {
  dim3 blk;
  block_shape_n(g.nl[2], TPB, blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
#define LB200_GO(P, G, U) collide_binary_kernel<P, G, U><<<grd, blk, 0, st>>>(g, cp, sp, md, fsrc, fdst, force, phi, grad, delsq, u)
  if (unrolled19) {
    if (pull) { if (cp.ghost) LB200_GO(true, true, true); else LB200_GO(true, false, true); }
    else      { if (cp.ghost) LB200_GO(false, true, true); else LB200_GO(false, false, true); }
  }
  else {
    if (pull) LB200_GO(true, true, false); else LB200_GO(false, true, false);
  }
#undef LB200_GO
  return 1;
}

// ---------------------------------------------------------------------------------------------
// lb_propagation as a stand-alone sweep: reference src/propagation.c:153-200.
// x in [1,N], every y,z of the allocation; y/z halo sites copy themselves.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB_MAX)
propagate_kernel(const Lb200Geom g, const Lb200ModelDev * __restrict__ md, int nvel, int ndist,
		 const double * __restrict__ f, double * __restrict__ fprime) {

  const int k0 = blockIdx.x*blockDim.x + threadIdx.x;       // 0 .. nall[2]-1
  const int j0 = blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (k0 >= g.nall[2] || j0 >= g.nall[1]) return;

  const int index = ((ic + g.nh - 1)*g.nall[1] + j0)*g.nall[2] + k0;
  const int jc = j0 - g.nh + 1;
  const int kc = k0 - g.nh + 1;
  const int mask = (jc >= 1 && jc <= g.nl[1] && kc >= 1 && kc <= g.nl[2]);
  const size_t ns = (size_t) g.nsites;

  for (int n = 0; n < ndist; n++) {
    for (int p = 0; p < nvel; p++) {
      const int off = mask*(md->cv[p][0]*g.xs + md->cv[p][1]*g.ys + md->cv[p][2]);
      fprime[(size_t) (n*nvel + p)*ns + index] = f[(size_t) (n*nvel + p)*ns + (index - off)];
    }
  }
}

int launch_propagate(cudaStream_t st, const Lb200Geom & g, const Lb200ModelDev * md, int nvel,
		     int ndist, const double * f, double * fprime) {
  dim3 blk;
  block_shape(g.nall[2], blk);
  dim3 grd((g.nall[2] + blk.x - 1)/blk.x, (g.nall[1] + blk.y - 1)/blk.y, g.nl[0]);
  propagate_kernel<<<grd, blk, 0, st>>>(g, md, nvel, ndist, f, fprime);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Halo shell.  Replaces the 26 pack kernels + MPI messages + 26 unpack kernels of
// lb_halo (src/lb_data.c:924-1114, 1183-1210, 1317-1477) and field_halo (src/field.c:1093-1251,
// 1329-1355, 1412-1531): every halo site within `depth` of the interior copies from the interior
// site it is the periodic image of.  Across a non-periodic boundary there is no neighbour
// (src/lb_data.c:1160-1172) and the reference unpacks zeros.  With x-slab decomposition the
// x-images live on the neighbouring GPUs; their boundary planes are staged in xlo / xhi
// (depth planes each, full y-z extent) before this kernel runs.
// Reduced distribution halo: only populations with c_p . m = |m|^2 travel in direction m
// (src/lb_data.c:1224-1239).
// ---------------------------------------------------------------------------------------------

template <bool REDUCED>
__global__ void __launch_bounds__(TPB)
halo_shell_kernel(const Lb200Geom g, const Lb200ModelDev * __restrict__ md, int ncomp, int d,
		  double * __restrict__ data, const double * __restrict__ local_src, const double * __restrict__ xlo,
		  const double * __restrict__ xhi, long long nx_slab, long long ny_slab,
		  long long nz_slab) {

  long long t = (long long) blockIdx.x*blockDim.x + threadIdx.x;
  const int ey = g.nl[1] + 2*d;     // extended extents
  const int ez = g.nl[2] + 2*d;
  int ic, jc, kc;

  if (t < 2*nx_slab) {
    // x slabs: i in [1-d,0] or [N+1,N+d], all extended j,k
    const int hi = (t >= nx_slab);
    if (hi) t -= nx_slab;
    kc = (int) (t % ez) + 1 - d;  t /= ez;
    jc = (int) (t % ey) + 1 - d;  t /= ey;
    ic = hi ? (g.nl[0] + 1 + (int) t) : ((int) t + 1 - d);
  }
  else if ((t -= 2*nx_slab) < 2*ny_slab) {
    const int hi = (t >= ny_slab);
    if (hi) t -= ny_slab;
    kc = (int) (t % ez) + 1 - d;  t /= ez;
    const int jj = (int) (t % d);  t /= d;
    jc = hi ? (g.nl[1] + 1 + jj) : (jj + 1 - d);
    ic = (int) t + 1;
  }
  else if ((t -= 2*ny_slab) < 2*nz_slab) {
    const int hi = (t >= nz_slab);
    if (hi) t -= nz_slab;
    const int kk = (int) (t % d);  t /= d;
    kc = hi ? (g.nl[2] + 1 + kk) : (kk + 1 - d);
    jc = (int) (t % g.nl[1]) + 1;  t /= g.nl[1];
    ic = (int) t + 1;
  }
  else {
    return;
  }

  // direction the arriving message travelled, m = -(offset of this halo site)
  const int mx = (ic < 1) ? 1 : (ic > g.nl[0] ? -1 : 0);
  const int my = (jc < 1) ? 1 : (jc > g.nl[1] ? -1 : 0);
  const int mz = (kc < 1) ? 1 : (kc > g.nl[2] ? -1 : 0);

  // No neighbour in direction -m (non-periodic boundary): the reference unpacks its never-written,
  // calloc'ed receive buffer there, i.e. zeros arrive (src/lb_data.c:1010-1011, src/field.c:1178-1187).
  const bool absent = (my != 0 && !g.per[1]) || (mz != 0 && !g.per[2]) || (mx > 0 && !g.has_lo)
    || (mx < 0 && !g.has_hi);

  const int sj = jc + my*g.nl[1];
  const int sk = kc + mz*g.nl[2];
  const size_t ns = (size_t) g.nsites;
  const int dst = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);

  const double * src;
  size_t sstride;
  size_t sidx;

  if (absent) {
    src = data;
    sstride = 0;
    sidx = 0;
  }
  else if (mx != 0 && g.remote_x) {
    // staging: [comp][d planes][nall_y][nall_z]; plane q of xlo = neighbour's i = N-d+1+q
    const int q = (mx > 0) ? (ic + d - 1) : (ic - g.nl[0] - 1);
    src = (mx > 0) ? xlo : xhi;
    sstride = (size_t) d*g.xs;
    sidx = (size_t) q*g.xs + (size_t) (sj + g.nh - 1)*g.nall[2] + (sk + g.nh - 1);
  }
  else {
    // the site one period away.  On a lattice thinner than the swap depth that site is itself a halo site, and what the
    // reference delivers is its content from BEFORE the swap (every send buffer is packed before any is unpacked,
    // src/field.c:1412-1531): local_src is then a snapshot of the array, otherwise the array itself
    const int si = ic + mx*g.nl[0];
    src = local_src;
    sstride = ns;
    sidx = (size_t) ((si + g.nh - 1)*g.nall[1] + (sj + g.nh - 1))*g.nall[2] + (sk + g.nh - 1);
  }

  if (REDUCED) {
    const int mm = mx*mx + my*my + mz*mz;
    for (int c = 0; c < ncomp; c++) {
      const int p = c % md->nvel;
      const int dot = mx*md->cv[p][0] + my*md->cv[p][1] + mz*md->cv[p][2];
      if (dot == mm) data[c*ns + dst] = absent ? 0.0 : src[c*sstride + sidx];
    }
  }
  else {
    for (int c = 0; c < ncomp; c++) data[c*ns + dst] = absent ? 0.0 : src[c*sstride + sidx];
  }
}

int launch_halo(cudaStream_t st, const Lb200Geom & g, const Lb200ModelDev * md, int ncomp,
		int depth, int reduced, double * data, const double * xlo, const double * xhi, const double * snapshot) {
  const double * local_src = snapshot ? snapshot : data;
  const long long ey = g.nl[1] + 2*depth;
  const long long ez = g.nl[2] + 2*depth;
  const long long nx_slab = (long long) depth*ey*ez;
  const long long ny_slab = (long long) g.nl[0]*depth*ez;
  const long long nz_slab = (long long) g.nl[0]*g.nl[1]*depth;
  const long long total = 2*(nx_slab + ny_slab + nz_slab);
  const int nblk = (int) ((total + TPB - 1)/TPB);
  if (reduced) halo_shell_kernel<true><<<nblk, TPB, 0, st>>>(g, md, ncomp, depth, data, local_src, xlo, xhi, nx_slab, ny_slab, nz_slab);
  else         halo_shell_kernel<false><<<nblk, TPB, 0, st>>>(g, md, ncomp, depth, data, local_src, xlo, xhi, nx_slab, ny_slab, nz_slab);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// 27-point gradient, reference src/gradient_3d_27pt_fluid.c:219-363, on [1-ne, N+ne]^3 with
// ne = nhalo - 1 (:91-95).  Summation order as written there.
// ---------------------------------------------------------------------------------------------

// One thread per (j,k) column of the extended region, marching GRAD_XC planes along x with the three
// 3x3 planes of the stencil held in registers: 9 new loads per site instead of 27 (the first version
// of this kernel was L1-bound at 72 % l1tex throughput, profiles/r01_ncu_full_step_256_baseline.md).
// The (j,k) plane is flattened so that no block is left with a sliver of the 258-wide rows.
constexpr int GRAD_XC = 16;

__global__ void __launch_bounds__(TPB)
grad27_kernel(const Lb200Geom g, const int ne, const double * __restrict__ field, double * __restrict__ grad,
	      double * __restrict__ delsq) {
  const int ey = g.nl[1] + 2*ne, ez = g.nl[2] + 2*ne;
  const int q = blockIdx.x*blockDim.x + threadIdx.x;
  if (q >= ey*ez) return;
  const int jc = 1 - ne + q/ez;
  const int kc = 1 - ne + q%ez;
  const int ic0 = 1 - ne + blockIdx.y*GRAD_XC;
  int ic1 = ic0 + GRAD_XC - 1;
  if (ic1 > g.nl[0] + ne) ic1 = g.nl[0] + ne;

  const int ys = g.ys;
  const size_t ns = (size_t) g.nsites;
  const double r9 = (1.0/9.0);
  int index = ((ic0 + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);

  // planes x-1 (m_), x (c_), x+1 (p_); second letter y offset, third z offset (m = -1, 0, p = +1)
  double m_mm, m_m0, m_mp, m_0m, m_00, m_0p, m_pm, m_p0, m_pp;
  double c_mm, c_m0, c_mp, c_0m, c_00, c_0p, c_pm, c_p0, c_pp;
  double p_mm, p_m0, p_mp, p_0m, p_00, p_0p, p_pm, p_p0, p_pp;

#define LB200_LOAD_PLANE(P, base) \
  P##_mm = field[(base)-ys-1]; P##_m0 = field[(base)-ys]; P##_mp = field[(base)-ys+1]; \
  P##_0m = field[(base)   -1]; P##_00 = field[(base)   ]; P##_0p = field[(base)   +1]; \
  P##_pm = field[(base)+ys-1]; P##_p0 = field[(base)+ys]; P##_pp = field[(base)+ys+1]

  // software pipeline: the plane two steps ahead (n_) is in flight while the current site is summed
  double n_mm, n_m0, n_mp, n_0m, n_00, n_0p, n_pm, n_p0, n_pp;
  LB200_LOAD_PLANE(m, index - g.xs);
  LB200_LOAD_PLANE(c, index);
  LB200_LOAD_PLANE(p, index + g.xs);

  for (int ic = ic0; ic <= ic1; ic++) {
    if (ic < ic1) { LB200_LOAD_PLANE(n, index + 2*g.xs); }

    grad[0*ns + index] = 0.5*r9*
      (+ p_mm - m_mm + p_m0 - m_m0 + p_mp - m_mp
       + p_0m - m_0m + p_00 - m_00 + p_0p - m_0p
       + p_pm - m_pm + p_p0 - m_p0 + p_pp - m_pp);
    grad[1*ns + index] = 0.5*r9*
      (+ m_pm - m_mm + m_p0 - m_m0 + m_pp - m_mp
       + c_pm - c_mm + c_p0 - c_m0 + c_pp - c_mp
       + p_pm - p_mm + p_p0 - p_m0 + p_pp - p_mp);
    grad[2*ns + index] = 0.5*r9*
      (+ m_mp - m_mm + m_0p - m_0m + m_pp - m_pm
       + c_mp - c_mm + c_0p - c_0m + c_pp - c_pm
       + p_mp - p_mm + p_0p - p_0m + p_pp - p_pm);
    delsq[index] = r9*
      (+ m_mm + m_m0 + m_mp + m_0m + m_00 + m_0p + m_pm + m_p0 + m_pp
       + c_mm + c_m0 + c_mp + c_0m        + c_0p + c_pm + c_p0 + c_pp
       + p_mm + p_m0 + p_mp + p_0m + p_00 + p_0p + p_pm + p_p0 + p_pp
       - 26.0*c_00);
    index += g.xs;

    m_mm = c_mm; m_m0 = c_m0; m_mp = c_mp; m_0m = c_0m; m_00 = c_00; m_0p = c_0p; m_pm = c_pm; m_p0 = c_p0; m_pp = c_pp;
    c_mm = p_mm; c_m0 = p_m0; c_mp = p_mp; c_0m = p_0m; c_00 = p_00; c_0p = p_0p; c_pm = p_pm; c_p0 = p_p0; c_pp = p_pp;
    if (ic < ic1) {
      p_mm = n_mm; p_m0 = n_m0; p_mp = n_mp; p_0m = n_0m; p_00 = n_00; p_0p = n_0p; p_pm = n_pm; p_p0 = n_p0; p_pp = n_pp;
    }
  }
#undef LB200_LOAD_PLANE
}

// ne: the operator is applied on [1-ne, N+ne]^3 (nhalo - 1 for grad/delsq of phi, src/gradient_3d_27pt_fluid.c:91-95;
// nhalo - 2 for the same operator applied to delsq, grad_3d_27pt_fluid_d4 :112-134)
int launch_grad27(cudaStream_t st, const Lb200Geom & g, int ne, const double * phi, double * grad,
		  double * delsq) {
  const int ex = g.nl[0] + 2*ne, ey = g.nl[1] + 2*ne, ez = g.nl[2] + 2*ne;
  dim3 grd((ey*ez + TPB - 1)/TPB, (ex + GRAD_XC - 1)/GRAD_XC, 1);
  grad27_kernel<<<grd, TPB, 0, st>>>(g, ne, phi, grad, delsq);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Chemical stress of the symmetric free energy at one site (only the column used by a face in
// direction b is needed from a neighbour): reference src/symmetric.c:371-416.
//   P_ab = p0 d_ab + kappa g_a g_b,  p0 = A/2 phi^2 + 3B/4 phi^4 - kappa phi delsq - kappa/2 |g|^2
// and the force  F_a = - d_b P_ab  with the face averages and the accumulation order
// +x, -x, +y, -y, +z, -z of reference src/phi_force_colloid.c:315-465.  The reference stores P
// for every site (72 B/site written, 7 x 72 B gathered); here P is recomputed on the fly.
// ---------------------------------------------------------------------------------------------

struct SiteFE {
  double phi, delsq, gx, gy, gz;
};

__device__ __forceinline__ SiteFE load_fe(const double * __restrict__ phi,
					   const double * __restrict__ grad,
					   const double * __restrict__ delsq, size_t ns, int idx) {
  SiteFE s;
  s.phi = phi[idx];
  s.delsq = delsq[idx];
  s.gx = grad[0*ns + idx];
  s.gy = grad[1*ns + idx];
  s.gz = grad[2*ns + idx];
  return s;
}

__device__ __forceinline__ double symm_p0(const Lb200SymmDev & sp, const SiteFE & s) {
  return 0.5*sp.a*s.phi*s.phi + 0.75*sp.b*s.phi*s.phi*s.phi*s.phi - sp.kappa*s.phi*s.delsq
    - 0.5*sp.kappa*(s.gx*s.gx + s.gy*s.gy + s.gz*s.gz);
}

// column b of P at a site: P[a][b], a = 0..2
template <int B>
__device__ __forceinline__ void symm_pcol(const Lb200SymmDev & sp, const SiteFE & s, double p[3]) {
  const double p0 = symm_p0(sp, s);
  const double gb = (B == 0) ? s.gx : (B == 1) ? s.gy : s.gz;
  const double d0 = (B == 0), d1 = (B == 1), d2 = (B == 2);
  p[0] = p0*d0 + sp.kappa*s.gx*gb;
  p[1] = p0*d1 + sp.kappa*s.gy*gb;
  p[2] = p0*d2 + sp.kappa*s.gz*gb;
}

__device__ __forceinline__ double symm_mu(const Lb200SymmDev & sp, double phi, double delsq) {
  return sp.a*phi + sp.b*phi*phi*phi - sp.kappa*delsq;
}

__device__ __forceinline__ void site_force(const Lb200SymmDev & sp, const SiteFE & s0,
					   const SiteFE & xp, const SiteFE & xm, const SiteFE & yp,
					   const SiteFE & ym, const SiteFE & zp, const SiteFE & zm,
					   double fo[3]) {
  double p0c[3], p1[3];
  symm_pcol<0>(sp, s0, p0c);
  symm_pcol<0>(sp, xp, p1);
  for (int a = 0; a < 3; a++) fo[a] = -0.5*(p1[a] + p0c[a]);
  symm_pcol<0>(sp, xm, p1);
  for (int a = 0; a < 3; a++) fo[a] += 0.5*(p1[a] + p0c[a]);
  symm_pcol<1>(sp, s0, p0c);
  symm_pcol<1>(sp, yp, p1);
  for (int a = 0; a < 3; a++) fo[a] -= 0.5*(p1[a] + p0c[a]);
  symm_pcol<1>(sp, ym, p1);
  for (int a = 0; a < 3; a++) fo[a] += 0.5*(p1[a] + p0c[a]);
  symm_pcol<2>(sp, s0, p0c);
  symm_pcol<2>(sp, zp, p1);
  for (int a = 0; a < 3; a++) fo[a] -= 0.5*(p1[a] + p0c[a]);
  symm_pcol<2>(sp, zm, p1);
  for (int a = 0; a < 3; a++) fo[a] += 0.5*(p1[a] + p0c[a]);
}

// ---------------------------------------------------------------------------------------------
// The two halves of phi_force_calculation as separate operators, for callers that use them
// directly: pth_stress_compute (src/phi_force_stress.c:171-284: P_ab stored for x in [0, N+1] and
// EVERY y, z of the allocation) and pth_force_fluid_driver (src/phi_force_colloid.c:274-465:
// force -= divergence of the stored P).  lb200_phi_force_calculation / lb200_step never store P.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB_MAX)
stress_symm_kernel(const Lb200Geom g, const Lb200SymmDev sp, const double * __restrict__ phi,
		   const double * __restrict__ grad, const double * __restrict__ delsq,
		   double * __restrict__ str) {
  const int k0 = blockIdx.x*blockDim.x + threadIdx.x;       // 0 .. nall[2]-1
  const int j0 = blockIdx.y*blockDim.y + threadIdx.y;
  const int i0 = blockIdx.z + g.nh - 1;                     // ic = 0 .. N+1
  if (k0 >= g.nall[2] || j0 >= g.nall[1]) return;
  const int index = (i0*g.nall[1] + j0)*g.nall[2] + k0;
  const size_t ns = (size_t) g.nsites;
  SiteFE s = load_fe(phi, grad, delsq, ns, index);
  const double p0 = symm_p0(sp, s);
  const double gr[3] = {s.gx, s.gy, s.gz};
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      str[(size_t) (ia*3 + ib)*ns + index] = p0*((ia == ib) ? 1.0 : 0.0) + sp.kappa*gr[ia]*gr[ib];
    }
  }
}

template <bool ACCUM>
__global__ void __launch_bounds__(TPB_MAX)
force_from_stress_kernel(const Lb200Geom g, const double * __restrict__ str, double * __restrict__ force) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  const int off[6] = {+g.xs, -g.xs, +g.ys, -g.ys, +1, -1};
  double fo[3];
#pragma unroll
  for (int d = 0; d < 6; d++) {
    const int ib = d/2;
    const int index1 = index + off[d];
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
      const double p1 = str[(size_t) (ia*3 + ib)*ns + index1];
      const double p0 = str[(size_t) (ia*3 + ib)*ns + index];
      if (d == 0)          fo[ia]  = -0.5*(p1 + p0);
      else if (d % 2 == 1) fo[ia] += 0.5*(p1 + p0);
      else                 fo[ia] -= 0.5*(p1 + p0);
    }
  }
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    if (ACCUM) force[ia*ns + index] += fo[ia];
    else       force[ia*ns + index] = fo[ia];
  }
}

int launch_stress(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, const double * phi,
		  const double * grad, const double * delsq, double * str) {
  dim3 blk;
  block_shape(g.nall[2], blk);
  dim3 grd((g.nall[2] + blk.x - 1)/blk.x, (g.nall[1] + blk.y - 1)/blk.y, g.nl[0] + 2);
  stress_symm_kernel<<<grd, blk, 0, st>>>(g, sp, phi, grad, delsq, str);
  return 1;
}

int launch_force_from_stress(cudaStream_t st, const Lb200Geom & g, int accumulate, const double * str,
			     double * force) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
  if (accumulate) force_from_stress_kernel<true><<<grd, blk, 0, st>>>(g, str, force);
  else            force_from_stress_kernel<false><<<grd, blk, 0, st>>>(g, str, force);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Cahn-Hilliard face fluxes at one site, fused: advective part (upwind order 1/2/3, reference
// src/advection.c:538-629, 770-893, 946-1141), - M (mu1 - mu0) (src/phi_cahn_hilliard.c:350-404),
// - M grad mu_ext (:1373-1397), no-normal-flux mask (src/advection_bcs.c:80-130), and the forward
// Euler update (src/phi_cahn_hilliard.c:1018-1049).  The reference keeps four flux arrays and
// five kernels; here the six face fluxes of a site are formed in registers.
// ---------------------------------------------------------------------------------------------

// Advective flux through the face between a "lo" site and the "hi" site one step up an axis, with the
// four phi values along that axis already in registers: pm1 = phi(lo-1), plo, phi_, pp1 = phi(hi+1).
// WEST = false: the face is owned by its lo site (fe, fy, fz: upwind test "u < 0");
// WEST = true : the x face owned by its hi site (fw: upwind test "u > 0").  The two only differ in
// which side a velocity of exactly zero picks (the flux is zero either way).
template <int ORDER, bool WEST>
__device__ __forceinline__ double adv_face(double u_lo, double u_hi, double pm1, double plo,
					   double phi_, double pp1) {
  if (ORDER == 1) {
    const double uf = 0.5*(u_lo + u_hi);
    const double up = WEST ? ((uf > 0.0) ? plo : phi_) : ((uf < 0.0) ? phi_ : plo);
    return uf*up;
  }
  else if (ORDER == 2) {
    return 0.5*(u_lo + u_hi)*1.0*0.5*(plo + phi_);
  }
  else if (ORDER == 4) {
    // four-point central interpolation, reference advection_le_4th (src/advection.c:1153-1262)
    const double a1 = (1.0/16.0);
    const double a2 = (9.0/16.0);
    const double uf = 0.5*(u_lo + u_hi);
    return uf*(- a1*pm1 + a2*plo + a2*phi_ - a1*pp1);
  }
  else {
    const double a1 = -0.213933;
    const double a2 =  0.927865;
    const double a3 =  0.286067;
    const double uf = 0.5*(u_lo + u_hi);
    const bool from_lo = WEST ? (uf > 0.0) : !(uf < 0.0);       // flow from lo to hi
    const double fd1 = from_lo ? pm1 : pp1;
    const double fd2 = from_lo ? plo : phi_;
    const double fd3 = from_lo ? phi_ : plo;
    return uf*(a1*fd1 + a2*fd2 + a3*fd3);
  }
}

// One thread per interior site.  Every operand of the site (13-point phi star, 7-point delsq / grad
// stars, 9 face velocities) is loaded up front into registers -- about 50 independent loads in flight
// per thread -- and everything after that is arithmetic: the first version of this kernel chased the
// upwind site with a dependent load and sat at 32 % DRAM throughput (profiles/r01_*baseline.md).
template <bool DO_FORCE, bool DO_CH, bool ACCUM, int ORDER, bool HAS_MAP, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
force_ch_kernel(const Lb200Geom g, const Lb200SymmDev sp, const double * __restrict__ phi,
		const double * __restrict__ grad, const double * __restrict__ delsq,
		const double * __restrict__ u, const char * __restrict__ status,
		double * __restrict__ force, double * __restrict__ phinew) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const int s = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  const int xs = g.xs, ys = g.ys;

  // ---- loads ----
  const double ph_c = phi[s];
  const double ph_xm = phi[s - xs], ph_xp = phi[s + xs];
  const double ph_ym = phi[s - ys], ph_yp = phi[s + ys];
  const double ph_zm = phi[s - 1],  ph_zp = phi[s + 1];
  const double d_c = delsq[s];
  const double d_xm = delsq[s - xs], d_xp = delsq[s + xs];
  const double d_ym = delsq[s - ys], d_yp = delsq[s + ys];
  const double d_zm = delsq[s - 1],  d_zp = delsq[s + 1];

  double ph_xm2 = 0.0, ph_xp2 = 0.0, ph_ym2 = 0.0, ph_yp2 = 0.0, ph_zm2 = 0.0, ph_zp2 = 0.0;
  double ux_c = 0.0, ux_xm = 0.0, ux_xp = 0.0, uy_c = 0.0, uy_ym = 0.0, uy_yp = 0.0;
  double uz_c = 0.0, uz_zm = 0.0, uz_zp = 0.0;
  if (DO_CH) {
    if (ORDER >= 3) {
      ph_xm2 = phi[s - 2*xs]; ph_xp2 = phi[s + 2*xs];
      ph_ym2 = phi[s - 2*ys]; ph_yp2 = phi[s + 2*ys];
      ph_zm2 = phi[s - 2];    ph_zp2 = phi[s + 2];
    }
    ux_c = u[0*ns + s]; ux_xm = u[0*ns + s - xs]; ux_xp = u[0*ns + s + xs];
    uy_c = u[1*ns + s]; uy_ym = u[1*ns + s - ys]; uy_yp = u[1*ns + s + ys];
    uz_c = u[2*ns + s]; uz_zm = u[2*ns + s - 1];  uz_zp = u[2*ns + s + 1];
  }

  if (DO_FORCE) {
    SiteFE s0, xp, xm, yp, ym, zp, zm;
#define LB200_FE(S, PH, D, IDX) \
    S.phi = PH; S.delsq = D; S.gx = grad[0*ns + (IDX)]; S.gy = grad[1*ns + (IDX)]; S.gz = grad[2*ns + (IDX)]
    LB200_FE(s0, ph_c, d_c, s);
    LB200_FE(xp, ph_xp, d_xp, s + xs);
    LB200_FE(xm, ph_xm, d_xm, s - xs);
    LB200_FE(yp, ph_yp, d_yp, s + ys);
    LB200_FE(ym, ph_ym, d_ym, s - ys);
    LB200_FE(zp, ph_zp, d_zp, s + 1);
    LB200_FE(zm, ph_zm, d_zm, s - 1);
#undef LB200_FE
    double fo[3];
    site_force(sp, s0, xp, xm, yp, ym, zp, zm, fo);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (ACCUM) force[a*ns + s] += fo[a];
      else       force[a*ns + s] = fo[a];
    }
  }

  if (DO_CH) {
    const double M = sp.mobility;
    const double mu0 = symm_mu(sp, ph_c, d_c);

    double mk = 1.0, mkxm = 1.0, mkxp = 1.0, mkyp = 1.0, mkym = 1.0, mkzp = 1.0, mkzm = 1.0;
    if (HAS_MAP) {
      mk   = (status[s] == 0);
      mkxm = (status[s - xs] == 0); mkxp = (status[s + xs] == 0);
      mkym = (status[s - ys] == 0); mkyp = (status[s + ys] == 0);
      mkzm = (status[s - 1] == 0);  mkzp = (status[s + 1] == 0);
    }

    // west face (s-x, s), owned by s
    double fw = adv_face<ORDER, true>(ux_xm, ux_c, ph_xm2, ph_xm, ph_c, ph_xp);
    fw -= M*(mu0 - symm_mu(sp, ph_xm, d_xm));
    fw -= M*sp.gm[0];
    if (HAS_MAP) fw *= mk*mkxm;
    // east face (s, s+x)
    double fe = adv_face<ORDER, false>(ux_c, ux_xp, ph_xm, ph_c, ph_xp, ph_xp2);
    fe -= M*(symm_mu(sp, ph_xp, d_xp) - mu0);
    fe -= M*sp.gm[0];
    if (HAS_MAP) fe *= mk*mkxp;
    // y face (s, s+y), owned by s
    double fy = adv_face<ORDER, false>(uy_c, uy_yp, ph_ym, ph_c, ph_yp, ph_yp2);
    fy -= M*(symm_mu(sp, ph_yp, d_yp) - mu0);
    fy -= M*sp.gm[1];
    if (HAS_MAP) fy *= mk*mkyp;
    // y face (s-y, s), owned by s-y
    double fym = adv_face<ORDER, false>(uy_ym, uy_c, ph_ym2, ph_ym, ph_c, ph_yp);
    fym -= M*(mu0 - symm_mu(sp, ph_ym, d_ym));
    fym -= M*sp.gm[1];
    if (HAS_MAP) fym *= mkym*mk;
    // z faces
    double fz = adv_face<ORDER, false>(uz_c, uz_zp, ph_zm, ph_c, ph_zp, ph_zp2);
    fz -= M*(symm_mu(sp, ph_zp, d_zp) - mu0);
    fz -= M*sp.gm[2];
    if (HAS_MAP) fz *= mk*mkzp;
    double fzm = adv_face<ORDER, false>(uz_zm, uz_c, ph_zm2, ph_zm, ph_c, ph_zp);
    fzm -= M*(mu0 - symm_mu(sp, ph_zm, d_zm));
    fzm -= M*sp.gm[2];
    if (HAS_MAP) fzm *= mkzm*mk;

    if (sp.csum != nullptr) {
      // phi_ch_csum_kernel (src/phi_cahn_hilliard.c:1181-1215) with kahan_add_double (src/util_sum.c:30-40), the
      // compensation carried from step to step in csum
      double sum = ph_c, cs = sp.csum[s];
      const double val[6] = {-fe, fw, -fy, fym, -sp.wz*fz, sp.wz*fzm};
#pragma unroll
      for (int n = 0; n < 6; n++) {
	const double y = val[n] + cs;
	const double t = sum + y;
	cs = y - (t - sum);
	sum = t;
      }
      sp.csum[s] = cs;
      phinew[s] = sum;
    }
    else {
      double ph = ph_c;
      ph -= (+ fe - fw + fy - fym + sp.wz*fz - sp.wz*fzm);
      phinew[s] = ph;
    }
  }
}

template <bool DO_FORCE, bool DO_CH>
int launch_force_ch_t(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, int accumulate,
		      const double * phi, const double * grad, const double * delsq,
		      const double * u, const char * status, double * force, double * phinew) {
  dim3 blk;
  block_shape_n(g.nl[2], TPB, blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
  static const int minb = tuned_flag("LB200_FCH_MINB", 4);
#define LB200_GO(A, O, M) do { \
    if (minb == 6)      force_ch_kernel<DO_FORCE, DO_CH, A, O, M, 6><<<grd, blk, 0, st>>>(g, sp, phi, grad, delsq, u, status, force, phinew); \
    else if (minb == 5) force_ch_kernel<DO_FORCE, DO_CH, A, O, M, 5><<<grd, blk, 0, st>>>(g, sp, phi, grad, delsq, u, status, force, phinew); \
    else if (minb == 3) force_ch_kernel<DO_FORCE, DO_CH, A, O, M, 3><<<grd, blk, 0, st>>>(g, sp, phi, grad, delsq, u, status, force, phinew); \
    else                force_ch_kernel<DO_FORCE, DO_CH, A, O, M, 4><<<grd, blk, 0, st>>>(g, sp, phi, grad, delsq, u, status, force, phinew); } while (0)
#define LB200_SEL_M(A, O) do { if (status) LB200_GO(A, O, true); else LB200_GO(A, O, false); } while (0)
#define LB200_SEL_O(A) do { if (sp.order == 1) LB200_SEL_M(A, 1); else if (sp.order == 2) LB200_SEL_M(A, 2); else if (sp.order == 4) LB200_SEL_M(A, 4); else LB200_SEL_M(A, 3); } while (0)
  if (accumulate) LB200_SEL_O(true); else LB200_SEL_O(false);
#undef LB200_GO
#undef LB200_SEL_M
#undef LB200_SEL_O
  return 1;
}

// fe_force_method phi_gradmu (src/phi_force.c:110-121): force += -phi grad mu (phi_grad_mu_fluid_kernel, src/phi_grad_mu.c:272-340),
// then force += -phi grad_mu_ext when that is not zero (phi_grad_mu_external_kernel, :352-384) -- two additions, as two kernels
// adding to hydro->force make them.  mu of the neighbours from phi and delsq (fe_symm_mu, src/symmetric.c:307-319).
__global__ void __launch_bounds__(TPB)
phi_gradmu_kernel(const Lb200Geom g, const Lb200SymmDev sp, const int accumulate, const double * __restrict__ phi,
		  const double * __restrict__ delsq, double * __restrict__ force) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int s = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const size_t ns = (size_t) g.nsites;
  const int off[3] = {g.xs, g.ys, 1};
  const bool ext = (sp.gm[0] != 0.0 || sp.gm[1] != 0.0 || sp.gm[2] != 0.0);
  const double phi0 = phi[s];
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    const double mum1 = symm_mu(sp, phi[s - off[ia]], delsq[s - off[ia]]);
    const double mup1 = symm_mu(sp, phi[s + off[ia]], delsq[s + off[ia]]);
    double f = 0.0;
    f += -phi0*0.5*(mup1 - mum1);
    double out = accumulate ? force[ia*ns + s] + f : 0.0 + f;
    if (ext) out += -phi0*sp.gm[ia];
    force[ia*ns + s] = out;
  }
}

int launch_phi_force(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, int accumulate,
		     const double * phi, const double * grad, const double * delsq, double * force) {
  if (sp.force_method == 1) {
    dim3 blk;
    block_shape_n(g.nl[2], TPB, blk);
    dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
    phi_gradmu_kernel<<<grd, blk, 0, st>>>(g, sp, accumulate, phi, delsq, force);
    return 1;
  }
  return launch_force_ch_t<true, false>(st, g, sp, accumulate, phi, grad, delsq, nullptr, nullptr,
					force, nullptr);
}

int launch_cahn_hilliard(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp,
			 const double * phi, const double * delsq, const double * u,
			 const char * status, double * phinew) {
  return launch_force_ch_t<false, true>(st, g, sp, 0, phi, nullptr, delsq, u, status, nullptr,
					phinew);
}

int launch_force_ch(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, int accumulate,
		    const double * phi, const double * grad, const double * delsq, const double * u,
		    const char * status, double * force, double * phinew) {
  return launch_force_ch_t<true, true>(st, g, sp, accumulate, phi, grad, delsq, u, status, force,
				       phinew);
}

// ---------------------------------------------------------------------------------------------
// The phi sector of a time step in one sweep: field_grad_compute + phi_force_calculation +
// phi_cahn_hilliard (reference src/gradient_3d_27pt_fluid.c:219-363, src/symmetric.c:371-416,
// src/phi_force_colloid.c:315-465, src/advection.c:946-1141, src/phi_cahn_hilliard.c:350-404,
// 1018-1049, 1373-1397), same operations in the same order as the separate kernels above.
//
// A CTA owns a (PS_BY-2) x (PS_BZ-2) tile of (j,k) columns plus a one-site apron and marches along
// x.  Per plane every thread (apron included) forms grad/delsq of ITS column from three phi planes
// staged in shared memory, derives p0 and mu ONCE per site (the separate kernels recompute them at
// all 7 stencil points), and publishes {p0, grad, mu, u_y, u_z} in shared memory for its y/z
// neighbours; the x neighbours are the thread's own previous / next plane, kept in registers.
// Global traffic per site: phi 8 + u 24 read, grad 24 + delsq 8 + force 24 + phi' 8 written
// (vs 136 B for gradient + force/CH kernels), 4 global loads per thread per plane, all prefetched
// one plane ahead, one __syncthreads per plane.
// ---------------------------------------------------------------------------------------------

constexpr int PS_BZ = 32;                 // threads along z (one warp)
#ifndef LB200_PS_BY
#define LB200_PS_BY 16
#endif
constexpr int PS_BY = LB200_PS_BY;        // threads along y
constexpr int PS_NT = PS_BZ*PS_BY;        // 512
constexpr int PS_TZ = PS_BZ - 2;          // interior columns per tile
constexpr int PS_TY = PS_BY - 2;
constexpr int PS_PZ = PS_BZ + 2;          // phi tile: block + one more ring
constexpr int PS_PY = PS_BY + 2;
constexpr int PS_PN = PS_PZ*PS_PY;        // 612

struct PsShared {
  double phi[4][PS_PN];                   // ring of phi planes
  double g[2][5][PS_NT];                  // p0, gx, gy, gz, mu of plane i (double buffered)
  double u[2][2][PS_NT];                  // u_y, u_z of plane i
};

// coordinate of the site that holds the value of (possibly halo) coordinate j: itself, or, when the
// dimension is read through the periodic boundary (g.wrap), the interior site it is the image of
__device__ __forceinline__ int ps_wrap(int j, int n, int w) {
  return w ? (j < 1 ? j + n : (j > n ? j - n : j)) : j;
}

[[maybe_unused]] __device__ __forceinline__ void ps_pcol(const Lb200SymmDev & sp, int B, double p0, double gx, double gy,
					double gz, double p[3]) {
  const double gb = (B == 0) ? gx : (B == 1) ? gy : gz;
  const double d0 = (B == 0), d1 = (B == 1), d2 = (B == 2);
  p[0] = p0*d0 + sp.kappa*gx*gb;
  p[1] = p0*d1 + sp.kappa*gy*gb;
  p[2] = p0*d2 + sp.kappa*gz*gb;
}

template <int ORDER>
__global__ void __launch_bounds__(PS_NT, 1)
phi_sector_kernel(const Lb200Geom g, const Lb200SymmDev sp, int xc, const double * __restrict__ phi,
		  const double * __restrict__ u, double * __restrict__ grad,
		  double * __restrict__ delsq, double * __restrict__ force,
		  double * __restrict__ phinew) {

  extern __shared__ __align__(16) unsigned char ps_smem_raw[];
  PsShared & sm = *reinterpret_cast<PsShared *>(ps_smem_raw);

  const int tz = threadIdx.x, ty = threadIdx.y;
  const int tid = ty*PS_BZ + tz;
  const int kbase = blockIdx.x*PS_TZ;              // thread column (j,k) = (jbase + ty, kbase + tz)
  const int jbase = blockIdx.y*PS_TY;
  const int kc = kbase + tz, jc = jbase + ty;
  const int i0 = 1 + g.xoff + blockIdx.z*xc;
  const int i1 = min(i0 + xc - 1, g.xcnt > 0 ? g.xoff + g.xcnt : g.nl[0]);
  const int nh = g.nh;
  const size_t ns = (size_t) g.nsites;
  const int xs = g.xs, ys = g.ys;

  const bool valid_g = (jc <= g.nl[1] + 1) && (kc <= g.nl[2] + 1);
  const bool inner = (ty >= 1 && ty <= PS_TY && tz >= 1 && tz <= PS_TZ);
  const bool out_site = inner && jc <= g.nl[1] && kc <= g.nl[2];
  const bool own_g = valid_g && ((ty >= 1 && ty <= PS_TY) || (ty == 0 && jc == 0))
    && ((tz >= 1 && tz <= PS_TZ) || (tz == 0 && kc == 0)) && !g.skip_diag;

  // clamp the column used for loads so that inactive threads stay inside the allocation
  const int jl = ps_wrap(min(jc, g.nl[1] + 1), g.nl[1], g.wrap[1]);
  const int kl = ps_wrap(min(kc, g.nl[2] + 1), g.nl[2], g.wrap[2]);
  const int col = (jl + nh - 1)*ys + (kl + nh - 1);       // + (i + nh - 1)*xs

  // cooperative phi plane load: element e of the (PS_PY x PS_PZ) tile <-> (jbase-1+r, kbase-1+c)
  const int e0 = tid, e1 = tid + PS_NT;
  const int r0 = e0/PS_PZ, c0 = e0%PS_PZ;
  const int r1 = e1/PS_PZ, c1 = e1%PS_PZ;
  const int pj0 = ps_wrap(min(jbase - 1 + r0, g.nl[1] + nh), g.nl[1], g.wrap[1]);
  const int pk0 = ps_wrap(min(kbase - 1 + c0, g.nl[2] + nh), g.nl[2], g.wrap[2]);
  const int pj1 = ps_wrap(min(jbase - 1 + r1, g.nl[1] + nh), g.nl[1], g.wrap[1]);
  const int pk1 = ps_wrap(min(kbase - 1 + c1, g.nl[2] + nh), g.nl[2], g.wrap[2]);
  const int pcol0 = (pj0 + nh - 1)*ys + (pk0 + nh - 1);
  const int pcol1 = (pj1 + nh - 1)*ys + (pk1 + nh - 1);
  const bool has_e1 = (e1 < PS_PN);

  const int istart = i0 - 2;

  // prologue: planes istart, istart+1, istart+2 into the ring
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int ip = istart + d;
    const int xo = (ps_wrap(ip, g.nl[0], g.wrap[0]) + nh - 1)*xs;
    sm.phi[(ip + 4) & 3][e0] = phi[xo + pcol0];
    if (has_e1) sm.phi[(ip + 4) & 3][e1] = phi[xo + pcol1];
  }
  double uxc = u[0*ns + (ps_wrap(istart, g.nl[0], g.wrap[0]) + nh - 1)*xs + col];
  double uxp = u[0*ns + (ps_wrap(istart + 1, g.nl[0], g.wrap[0]) + nh - 1)*xs + col];
  double uxm = 0.0;
  __syncthreads();

  // own-column history
  double gm_p0 = 0.0, gm_x = 0.0, gm_y = 0.0, gm_z = 0.0, gm_mu = 0.0;      // plane i-1
  double gc_p0 = 0.0, gc_x = 0.0, gc_y = 0.0, gc_z = 0.0, gc_mu = 0.0;      // plane i
  double phim1 = 0.0, phim2 = 0.0;                                        // phi(i-1), phi(i-2), own column

  const int pc = (ty + 1)*PS_PZ + (tz + 1);        // own position in the phi tile

  for (int i = istart; i <= i1; i++) {

    // ---- 1. prefetch for the next plane-steps (consumed after the arithmetic below) ----
    double pf0 = 0.0, pf1 = 0.0, uxn = 0.0, uyn = 0.0, uzn = 0.0;
    if (i < i1) {
      const int xo3 = (ps_wrap(i + 3, g.nl[0], g.wrap[0]) + nh - 1)*xs;
      const int xo2 = (ps_wrap(i + 2, g.nl[0], g.wrap[0]) + nh - 1)*xs;
      const int xo1 = (ps_wrap(i + 1, g.nl[0], g.wrap[0]) + nh - 1)*xs;
      pf0 = phi[xo3 + pcol0];
      if (has_e1) pf1 = phi[xo3 + pcol1];
      uxn = u[0*ns + xo2 + col];
      uyn = u[1*ns + xo1 + col];
      uzn = u[2*ns + xo1 + col];
    }

    // ---- 2. gradient of plane i+1 at the own column, from phi planes i, i+1, i+2 ----
    const double * __restrict__ fm = sm.phi[(i + 4) & 3];
    const double * __restrict__ fc = sm.phi[(i + 5) & 3];
    const double * __restrict__ fp = sm.phi[(i + 6) & 3];
    double gp_p0 = 0.0, gp_x = 0.0, gp_y = 0.0, gp_z = 0.0, gp_mu = 0.0;
    if (valid_g) {
      const double r9 = (1.0/9.0);
      const double m_mm = fm[pc-PS_PZ-1], m_m0 = fm[pc-PS_PZ], m_mp = fm[pc-PS_PZ+1];
      const double m_0m = fm[pc      -1], m_00 = fm[pc      ], m_0p = fm[pc      +1];
      const double m_pm = fm[pc+PS_PZ-1], m_p0 = fm[pc+PS_PZ], m_pp = fm[pc+PS_PZ+1];
      const double c_mm = fc[pc-PS_PZ-1], c_m0 = fc[pc-PS_PZ], c_mp = fc[pc-PS_PZ+1];
      const double c_0m = fc[pc      -1], c_00 = fc[pc      ], c_0p = fc[pc      +1];
      const double c_pm = fc[pc+PS_PZ-1], c_p0 = fc[pc+PS_PZ], c_pp = fc[pc+PS_PZ+1];
      const double p_mm = fp[pc-PS_PZ-1], p_m0 = fp[pc-PS_PZ], p_mp = fp[pc-PS_PZ+1];
      const double p_0m = fp[pc      -1], p_00 = fp[pc      ], p_0p = fp[pc      +1];
      const double p_pm = fp[pc+PS_PZ-1], p_p0 = fp[pc+PS_PZ], p_pp = fp[pc+PS_PZ+1];

      gp_x = 0.5*r9*
	(+ p_mm - m_mm + p_m0 - m_m0 + p_mp - m_mp
	 + p_0m - m_0m + p_00 - m_00 + p_0p - m_0p
	 + p_pm - m_pm + p_p0 - m_p0 + p_pp - m_pp);
      gp_y = 0.5*r9*
	(+ m_pm - m_mm + m_p0 - m_m0 + m_pp - m_mp
	 + c_pm - c_mm + c_p0 - c_m0 + c_pp - c_mp
	 + p_pm - p_mm + p_p0 - p_m0 + p_pp - p_mp);
      gp_z = 0.5*r9*
	(+ m_mp - m_mm + m_0p - m_0m + m_pp - m_pm
	 + c_mp - c_mm + c_0p - c_0m + c_pp - c_pm
	 + p_mp - p_mm + p_0p - p_0m + p_pp - p_pm);
      const double dsq = r9*
	(+ m_mm + m_m0 + m_mp + m_0m + m_00 + m_0p + m_pm + m_p0 + m_pp
	 + c_mm + c_m0 + c_mp + c_0m        + c_0p + c_pm + c_p0 + c_pp
	 + p_mm + p_m0 + p_mp + p_0m + p_00 + p_0p + p_pm + p_p0 + p_pp
	 - 26.0*c_00);

      SiteFE sf;
      sf.phi = c_00; sf.delsq = dsq; sf.gx = gp_x; sf.gy = gp_y; sf.gz = gp_z;
      gp_p0 = symm_p0(sp, sf);
      gp_mu = symm_mu(sp, c_00, dsq);

      const int ig = i + 1;
      const bool own_x = (ig >= i0 && ig <= i1) || (ig == 0 && i0 == 1) || (ig == g.nl[0] + 1 && i1 == g.nl[0]);
      if (own_g && own_x) {
	const int sidx = (ig + nh - 1)*xs + (jc + nh - 1)*ys + (kc + nh - 1);
	grad[0*ns + sidx] = gp_x;
	grad[1*ns + sidx] = gp_y;
	grad[2*ns + sidx] = gp_z;
	delsq[sidx] = dsq;
      }
    }
    {
      double (* gb)[PS_NT] = sm.g[(i + 1) & 1];
      gb[0][tid] = gp_p0; gb[1][tid] = gp_x; gb[2][tid] = gp_y; gb[3][tid] = gp_z; gb[4][tid] = gp_mu;
    }

    // ---- 3. force and Cahn-Hilliard update of plane i ----
    const double ph_c = fm[pc];
    if (i >= i0 && out_site) {
      const double (* gb)[PS_NT] = sm.g[i & 1];
      const double (* ub)[PS_NT] = sm.u[i & 1];
      const int s = (i + nh - 1)*xs + (jc + nh - 1)*ys + (kc + nh - 1);
      const int typ = tid + PS_BZ, tym = tid - PS_BZ, tzp = tid + 1, tzm = tid - 1;

      // force = - div P, accumulation order +x, -x, +y, -y, +z, -z
      double fo[3], p0c[3], p1[3];
      ps_pcol(sp, 0, gc_p0, gc_x, gc_y, gc_z, p0c);
      ps_pcol(sp, 0, gp_p0, gp_x, gp_y, gp_z, p1);
#pragma unroll
      for (int a = 0; a < 3; a++) fo[a] = -0.5*(p1[a] + p0c[a]);
      ps_pcol(sp, 0, gm_p0, gm_x, gm_y, gm_z, p1);
#pragma unroll
      for (int a = 0; a < 3; a++) fo[a] += 0.5*(p1[a] + p0c[a]);
      ps_pcol(sp, 1, gc_p0, gc_x, gc_y, gc_z, p0c);
      ps_pcol(sp, 1, gb[0][typ], gb[1][typ], gb[2][typ], gb[3][typ], p1);
#pragma unroll
      for (int a = 0; a < 3; a++) fo[a] -= 0.5*(p1[a] + p0c[a]);
      ps_pcol(sp, 1, gb[0][tym], gb[1][tym], gb[2][tym], gb[3][tym], p1);
#pragma unroll
      for (int a = 0; a < 3; a++) fo[a] += 0.5*(p1[a] + p0c[a]);
      ps_pcol(sp, 2, gc_p0, gc_x, gc_y, gc_z, p0c);
      ps_pcol(sp, 2, gb[0][tzp], gb[1][tzp], gb[2][tzp], gb[3][tzp], p1);
#pragma unroll
      for (int a = 0; a < 3; a++) fo[a] -= 0.5*(p1[a] + p0c[a]);
      ps_pcol(sp, 2, gb[0][tzm], gb[1][tzm], gb[2][tzm], gb[3][tzm], p1);
#pragma unroll
      for (int a = 0; a < 3; a++) fo[a] += 0.5*(p1[a] + p0c[a]);
#pragma unroll
      for (int a = 0; a < 3; a++) force[a*ns + s] = fo[a];

      // Cahn-Hilliard: six face fluxes in registers, forward Euler
      const double M = sp.mobility;
      const double mu0 = gc_mu;
      const double ph_xm = phim1, ph_xm2 = phim2, ph_xp = fc[pc], ph_xp2 = fp[pc];
      const double ph_ym = fm[pc - PS_PZ], ph_yp = fm[pc + PS_PZ];
      const double ph_zm = fm[pc - 1], ph_zp = fm[pc + 1];
      double ph_ym2 = 0.0, ph_yp2 = 0.0, ph_zm2 = 0.0, ph_zp2 = 0.0;
      if (ORDER == 3) {
	ph_ym2 = fm[pc - 2*PS_PZ]; ph_yp2 = fm[pc + 2*PS_PZ];
	ph_zm2 = fm[pc - 2];       ph_zp2 = fm[pc + 2];
      }
      const double uy_c = ub[0][tid], uy_ym = ub[0][tym], uy_yp = ub[0][typ];
      const double uz_c = ub[1][tid], uz_zm = ub[1][tzm], uz_zp = ub[1][tzp];

      double fw = adv_face<ORDER, true>(uxm, uxc, ph_xm2, ph_xm, ph_c, ph_xp);
      fw -= M*(mu0 - gm_mu);
      fw -= M*sp.gm[0];
      double fe = adv_face<ORDER, false>(uxc, uxp, ph_xm, ph_c, ph_xp, ph_xp2);
      fe -= M*(gp_mu - mu0);
      fe -= M*sp.gm[0];
      double fy = adv_face<ORDER, false>(uy_c, uy_yp, ph_ym, ph_c, ph_yp, ph_yp2);
      fy -= M*(gb[4][typ] - mu0);
      fy -= M*sp.gm[1];
      double fym = adv_face<ORDER, false>(uy_ym, uy_c, ph_ym2, ph_ym, ph_c, ph_yp);
      fym -= M*(mu0 - gb[4][tym]);
      fym -= M*sp.gm[1];
      double fz = adv_face<ORDER, false>(uz_c, uz_zp, ph_zm, ph_c, ph_zp, ph_zp2);
      fz -= M*(gb[4][tzp] - mu0);
      fz -= M*sp.gm[2];
      double fzm = adv_face<ORDER, false>(uz_zm, uz_c, ph_zm2, ph_zm, ph_c, ph_zp);
      fzm -= M*(mu0 - gb[4][tzm]);
      fzm -= M*sp.gm[2];

      double ph = ph_c;
      ph -= (+ fe - fw + fy - fym + sp.wz*fz - sp.wz*fzm);
      phinew[s] = ph;
      // the planes the neighbour GPUs' next phi sector reads, straight into their halo planes
      if (g.peer_phi_lo != nullptr && i <= nh) g.peer_phi_lo[(size_t) s + (size_t) g.nl[0]*xs] = ph;
      if (g.peer_phi_hi != nullptr && i > g.nl[0] - nh) g.peer_phi_hi[(size_t) s - (size_t) g.nl[0]*xs] = ph;
    }

    // ---- 4. rotate the own-column history, publish the prefetched plane ----
    gm_p0 = gc_p0; gm_x = gc_x; gm_y = gc_y; gm_z = gc_z; gm_mu = gc_mu;
    gc_p0 = gp_p0; gc_x = gp_x; gc_y = gp_y; gc_z = gp_z; gc_mu = gp_mu;
    phim2 = phim1; phim1 = ph_c;
    uxm = uxc; uxc = uxp; uxp = uxn;
    if (i < i1) {
      sm.phi[(i + 7) & 3][e0] = pf0;                 // plane i+3 -> slot of plane i-1
      if (has_e1) sm.phi[(i + 7) & 3][e1] = pf1;
      sm.u[(i + 1) & 1][0][tid] = uyn;
      sm.u[(i + 1) & 1][1][tid] = uzn;
    }
    __syncthreads();
  }
}

#ifndef LB200_STRICT
// ---------------------------------------------------------------------------------------------
// The same sweep for the FAST arithmetic mode (results within the stated FP64 tolerance of the
// reference, not bit-identical).  The exact kernel above is bound by instruction issue (ncu: 51 %
// issue-active, 46 % FP64 pipe at 25 % occupancy, ~300 FP64 operations and ~670 instructions per
// site), so this version cuts the operation count, which the reference's summation order forbids in
// strict mode:
//  * 27-point stencil from per-plane partial sums: B = 3x3 box sum, Cy / Cz = central differences of
//    the row / column sums, kept in registers for three planes:
//        d_x = (B(i+1) - B(i-1))/18,  d_y = (Cy(i-1) + Cy(i) + Cy(i+1))/18,  d_z likewise,
//        delsq = (B(i-1) + B(i) + B(i+1) - 27 phi)/9                      (26 operations, 9 loads);
//  * every site forms its stress tensor P once and publishes the five entries its y/z neighbours
//    need; F_a = 1/2 sum_b [P_ab(-b) - P_ab(+b)]  (the centre-site terms of the reference's face
//    averages cancel);
//  * every face flux is computed once, by the site on its low side, and shared with the site on its
//    high side (x: carried in registers along the march; y, z: through shared memory), so the phi
//    update of plane n-1 is completed one plane-step after its fluxes;
//  * the march is unrolled by the period of the shared-memory rings (4) for the steady-state planes
//    of a chunk, so every shared-memory address is a per-thread base plus an immediate, and all the
//    range tests of the pipeline fill / drain live in a generic step used for the first 3 and last
//    4-7 planes only.
// One __syncthreads per plane; all shared buffers are double buffered.
// ---------------------------------------------------------------------------------------------

constexpr int PF_RING = 6;                // phi planes in flight: n .. n+2 read, n+3 landed, n+4 in flight
struct PfShared {
  double phi[PF_RING][PS_PN];             // ring of phi planes, slot = (plane - first plane) % 6
  double g[2][6][PS_NT];                  // Pxy, Pyy, Pyz, Pxz, Pzz, mu of a plane
  double u[3][2][PS_NT];                  // u_y, u_z, slot = plane % 3
  double ux[3][PS_NT];                    // u_x, slot = plane % 3
  double fl[2][2][PS_NT];                 // y and z face fluxes (face between the site and site+1)
};

// global -> shared without a register stop-over (LDGSTS): the prefetch distance of the march is two
// plane-steps, more than the registers of this kernel could hold
__device__ __forceinline__ void pf_cp_async8(double * smem_dst, const double * gsrc) {
  const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void pf_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pf_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

struct PfRegs {                            // own-column history carried along the march (plane n = current)
  double Bm, Cym, Czm, Bc, Cyc, Czc;      // plane sums of planes n, n+1
  double gm_xx, gm_xy, gm_xz;             // P_xa of plane n-1
  double gc_xx, gc_xy, gc_xz, gc_mu;      // P_xa, mu of plane n
  double phim1;                           // phi(n-1)
  double fxm1, fxm2;                      // x-face fluxes (n-1 | n), (n-2 | n-1)
  double fy_prev, fz_prev;                // own y / z face fluxes of plane n-1
  double uxc;                             // u_x(n)
};

struct PfK {                               // per-thread / per-CTA constants
  int pc, tid, col, scol, pcol0, pcol1, e0, e1;
  bool has_e1, valid_g, out_site, face_row, own_g;
  int xs, ns, nh, nlx, wx, i0, i1;
  double M, kappa, a, b, mg0, mg1, mg2, wz;
  double * peer_lo, * peer_hi;            // neighbour GPUs' phi' arrays (nullptr: none)
};

// partial sums of one phi plane at the own column: B, Cy, Cz
__device__ __forceinline__ void pf_plane_sums(const double * __restrict__ q, int pc, double & B,
					      double & Cy, double & Cz) {
  const double mm = q[pc-PS_PZ-1], m0 = q[pc-PS_PZ], mp = q[pc-PS_PZ+1];
  const double zm = q[pc      -1], z0 = q[pc      ], zp = q[pc      +1];
  const double pm = q[pc+PS_PZ-1], p0 = q[pc+PS_PZ], pp = q[pc+PS_PZ+1];
  const double am = (mm + m0) + mp;
  const double a0 = (zm + z0) + zp;
  const double ap = (pm + p0) + pp;
  B  = (am + a0) + ap;
  Cy = ap - am;
  Cz = ((mp - mm) + (zp - zm)) + (pp - pm);
}

// One plane-step of the march at plane n.  PH >= 0: steady state, PH = (n - first plane) % 6 known at
// compile time, every stage active, no periodic wrap in x (planes n+1 .. n+4 inside the chunk).
// PH < 0: generic (pipeline fill and drain).
template <int ORDER, int PH>
__device__ __forceinline__ void pf_step(PfShared & sm, PfRegs & r, const PfK & k, const int n,
					const double * __restrict__ phi, const double * __restrict__ u,
					double * __restrict__ grad, double * __restrict__ delsq,
					double * __restrict__ force, double * __restrict__ phinew) {
  constexpr bool STEADY = (PH >= 0);
  const int q = STEADY ? PH : (n - (k.i0 - 2)) % 6;        // phase: plane n relative to the first plane
  const int tid = k.tid, pc = k.pc;
  const int typ = tid + PS_BZ, tym = tid - PS_BZ, tzp = tid + 1, tzm = tid - 1;
  const double r9 = (1.0/9.0), r18 = 0.5*(1.0/9.0);

  const bool do_grad = STEADY || (n <= k.i1);
  const bool do_fx   = STEADY || (n >= k.i0 - 1 && n <= k.i1);
  const bool do_full = STEADY || (n >= k.i0 && n <= k.i1);
  const bool do_upd  = STEADY || (n >= k.i0 + 1);

  // ---- 1. asynchronous prefetch, two plane-steps ahead: phi(n+4), u_x(n+3), u_y / u_z (n+2) ----
  {
    int xo2, xo3, xo4;
    if (STEADY) {
      xo2 = (n + 1 + k.nh)*k.xs; xo3 = xo2 + k.xs; xo4 = xo3 + k.xs;
    }
    else {
      xo2 = (ps_wrap(n + 2, k.nlx, k.wx) + k.nh - 1)*k.xs;
      xo3 = (ps_wrap(n + 3, k.nlx, k.wx) + k.nh - 1)*k.xs;
      xo4 = (ps_wrap(n + 4, k.nlx, k.wx) + k.nh - 1)*k.xs;
    }
    if (STEADY || n + 4 <= k.i1 + 2) {
      pf_cp_async8(&sm.phi[(q + 4) % 6][k.e0], phi + xo4 + k.pcol0);
      if (k.has_e1) pf_cp_async8(&sm.phi[(q + 4) % 6][k.e1], phi + xo4 + k.pcol1);
    }
    if (STEADY || n + 3 <= k.i1 + 1) pf_cp_async8(&sm.ux[q % 3][tid], u + xo3 + k.col);
    if (STEADY || n + 2 <= k.i1) {
      pf_cp_async8(&sm.u[(q + 2) % 3][0][tid], u + k.ns + xo2 + k.col);
      pf_cp_async8(&sm.u[(q + 2) % 3][1][tid], u + 2*k.ns + xo2 + k.col);
    }
    pf_cp_async_commit();
  }

  const double * __restrict__ fm = sm.phi[q];
  const double * __restrict__ fc = sm.phi[(q + 1) % 6];
  const double * __restrict__ fp = sm.phi[(q + 2) % 6];

  // ---- 2. gradient, chemical potential and stress of plane n+1 at the own column ----
  double gp_xx = 0.0, gp_xy = 0.0, gp_xz = 0.0, gp_mu = 0.0;
  double Bp = 0.0, Cyp = 0.0, Czp = 0.0;
  if (do_grad) {
    pf_plane_sums(fp, pc, Bp, Cyp, Czp);
    const double phc = fc[pc];
    const double gx = r18*(Bp - r.Bm);
    const double gy = r18*((r.Cym + r.Cyc) + Cyp);
    const double gz = r18*((r.Czm + r.Czc) + Czp);
    const double dsq = r9*(((r.Bm + r.Bc) + Bp) - 27.0*phc);

    const int ig = n + 1;
    const bool own_x = STEADY || (ig >= k.i0 && ig <= k.i1) || (ig == 0 && k.i0 == 1) || (ig == k.nlx + 1 && k.i1 == k.nlx);
    if (k.own_g && own_x) {
      const int sidx = (ig + k.nh - 1)*k.xs + k.scol;
      grad[sidx] = gx;
      grad[k.ns + sidx] = gy;
      grad[2*k.ns + sidx] = gz;
      delsq[sidx] = dsq;
    }

    const double ph2 = phc*phc;
    const double p0 = ph2*(0.5*k.a + 0.75*k.b*ph2) - k.kappa*(phc*dsq + 0.5*((gx*gx + gy*gy) + gz*gz));
    gp_mu = phc*(k.a + k.b*ph2) - k.kappa*dsq;
    const double kgx = k.kappa*gx, kgy = k.kappa*gy, kgz = k.kappa*gz;
    gp_xx = p0 + kgx*gx; gp_xy = kgx*gy; gp_xz = kgx*gz;
    double (* gb)[PS_NT] = sm.g[(q + 1) & 1];
    gb[0][tid] = gp_xy;
    gb[1][tid] = p0 + kgy*gy;
    gb[2][tid] = kgy*gz;
    gb[3][tid] = gp_xz;
    gb[4][tid] = p0 + kgz*gz;
    gb[5][tid] = gp_mu;
  }

  // ---- 3. plane n: x-face flux (n | n+1), force, y/z face fluxes ----
  // (rows 0 .. PS_TY own faces; lane PS_BZ-1 of those rows computes harmless values that nobody reads)
  const double ph_c = fm[pc];
  const double uxp = sm.ux[(q + 1) % 3][tid];               // u_x(n+1)
  double fx = 0.0, fy = 0.0, fz = 0.0;
  if (STEADY || (do_fx && k.face_row)) {          // steady state: branch-free, one basic block for the scheduler
    fx = adv_face<ORDER, false>(r.uxc, uxp, r.phim1, ph_c, fc[pc], fp[pc]) - k.M*(gp_mu - r.gc_mu) - k.mg0;

    if (do_full) {
      const double (* gb)[PS_NT] = sm.g[q & 1];
      const double (* ub)[PS_NT] = sm.u[q % 3];

      if (k.out_site) {
	const int s = (n + k.nh - 1)*k.xs + k.scol;
	force[s]          = 0.5*(((r.gm_xx - gp_xx) + (gb[0][tym] - gb[0][typ])) + (gb[3][tzm] - gb[3][tzp]));
	force[k.ns + s]   = 0.5*(((r.gm_xy - gp_xy) + (gb[1][tym] - gb[1][typ])) + (gb[2][tzm] - gb[2][tzp]));
	force[2*k.ns + s] = 0.5*(((r.gm_xz - gp_xz) + (gb[2][tym] - gb[2][typ])) + (gb[4][tzm] - gb[4][tzp]));
      }

      double ph_yp2 = 0.0, ph_zp2 = 0.0;
      if (ORDER == 3) { ph_yp2 = fm[pc + 2*PS_PZ]; ph_zp2 = fm[pc + 2]; }
      fy = adv_face<ORDER, false>(ub[0][tid], ub[0][typ], fm[pc - PS_PZ], ph_c, fm[pc + PS_PZ], ph_yp2)
	- k.M*(gb[5][typ] - r.gc_mu) - k.mg1;
      fz = adv_face<ORDER, false>(ub[1][tid], ub[1][tzp], fm[pc - 1], ph_c, fm[pc + 1], ph_zp2)
	- k.M*(gb[5][tzp] - r.gc_mu) - k.mg2;
      sm.fl[q & 1][0][tid] = fy;
      sm.fl[q & 1][1][tid] = fz;
    }
  }

  // ---- 4. phi update of plane n-1, whose y/z face fluxes were published one plane-step ago ----
  if (do_upd && k.out_site) {
    const double (* fl)[PS_NT] = sm.fl[(q + 1) & 1];
    const int s = (n - 1 + k.nh - 1)*k.xs + k.scol;
    const double phn = r.phim1 - (((r.fxm1 - r.fxm2) + (r.fy_prev - fl[0][tym])) + k.wz*(r.fz_prev - fl[1][tzm]));
    phinew[s] = phn;
    // the planes the neighbour GPUs' next phi sector reads, straight into their halo planes
    if (k.peer_lo != nullptr && n - 1 <= k.nh) k.peer_lo[(size_t) s + (size_t) k.nlx*k.xs] = phn;
    if (k.peer_hi != nullptr && n - 1 > k.nlx - k.nh) k.peer_hi[(size_t) s - (size_t) k.nlx*k.xs] = phn;
  }

  // ---- 5. rotate the own-column history; the prefetch issued ONE step ago must have landed ----
  r.Bm = r.Bc; r.Cym = r.Cyc; r.Czm = r.Czc;
  r.Bc = Bp; r.Cyc = Cyp; r.Czc = Czp;
  r.gm_xx = r.gc_xx; r.gm_xy = r.gc_xy; r.gm_xz = r.gc_xz;
  r.gc_xx = gp_xx; r.gc_xy = gp_xy; r.gc_xz = gp_xz;
  r.gc_mu = gp_mu;
  r.phim1 = ph_c;
  r.fxm2 = r.fxm1; r.fxm1 = fx;
  r.fy_prev = fy; r.fz_prev = fz;
  r.uxc = uxp;
  pf_cp_async_wait<1>();
  __syncthreads();
}

template <int ORDER>
__global__ void __launch_bounds__(PS_NT, 512/PS_NT)
phi_sector_fast_kernel(const Lb200Geom g, const Lb200SymmDev sp, int xc,
		       const double * __restrict__ phi, const double * __restrict__ u,
		       double * __restrict__ grad, double * __restrict__ delsq,
		       double * __restrict__ force, double * __restrict__ phinew) {

  extern __shared__ __align__(16) unsigned char ps_smem_raw[];
  PfShared & sm = *reinterpret_cast<PfShared *>(ps_smem_raw);

  const int tz = threadIdx.x, ty = threadIdx.y;
  const int kbase = blockIdx.x*PS_TZ;              // thread column (j,k) = (jbase + ty, kbase + tz)
  const int jbase = blockIdx.y*PS_TY;
  const int kc = kbase + tz, jc = jbase + ty;
  const int nh = g.nh, ys = g.ys;

  PfK k;
  k.tid = ty*PS_BZ + tz;
  k.pc = (ty + 1)*PS_PZ + (tz + 1);                // own position in the phi tile
  k.xs = g.xs; k.ns = g.nsites; k.nh = nh; k.nlx = g.nl[0]; k.wx = g.wrap[0];
  k.i0 = 1 + g.xoff + blockIdx.z*xc;
  k.i1 = min(k.i0 + xc - 1, g.xcnt > 0 ? g.xoff + g.xcnt : g.nl[0]);
  k.M = sp.mobility; k.kappa = sp.kappa; k.a = sp.a; k.b = sp.b; k.wz = sp.wz;
  k.peer_lo = g.peer_phi_lo; k.peer_hi = g.peer_phi_hi;
  k.mg0 = sp.mobility*sp.gm[0]; k.mg1 = sp.mobility*sp.gm[1]; k.mg2 = sp.mobility*sp.gm[2];

  const bool inner = (ty >= 1 && ty <= PS_TY && tz >= 1 && tz <= PS_TZ);
  k.valid_g = (jc <= g.nl[1] + 1) && (kc <= g.nl[2] + 1);
  k.out_site = inner && jc <= g.nl[1] && kc <= g.nl[2];
  k.face_row = (ty <= PS_TY);                      // rows that own faces towards j+1 / k+1
  k.own_g = k.valid_g && ((ty >= 1 && ty <= PS_TY) || (ty == 0 && jc == 0))
    && ((tz >= 1 && tz <= PS_TZ) || (tz == 0 && kc == 0)) && !g.skip_diag;

  // column used for loads: clamped inside the allocation, through the periodic boundary if wrapping
  const int jl = ps_wrap(min(jc, g.nl[1] + 1), g.nl[1], g.wrap[1]);
  const int kl = ps_wrap(min(kc, g.nl[2] + 1), g.nl[2], g.wrap[2]);
  k.col = (jl + nh - 1)*ys + (kl + nh - 1);
  k.scol = (jc + nh - 1)*ys + (kc + nh - 1);       // column of the stores (never wrapped)

  // cooperative phi plane load: element e of the (PS_PY x PS_PZ) tile <-> (jbase-1+r, kbase-1+c)
  k.e0 = k.tid; k.e1 = k.tid + PS_NT;
  {
    const int r0 = k.e0/PS_PZ, c0 = k.e0%PS_PZ;
    const int r1 = k.e1/PS_PZ, c1 = k.e1%PS_PZ;
    const int pj0 = ps_wrap(min(jbase - 1 + r0, g.nl[1] + nh), g.nl[1], g.wrap[1]);
    const int pk0 = ps_wrap(min(kbase - 1 + c0, g.nl[2] + nh), g.nl[2], g.wrap[2]);
    const int pj1 = ps_wrap(min(jbase - 1 + r1, g.nl[1] + nh), g.nl[1], g.wrap[1]);
    const int pk1 = ps_wrap(min(kbase - 1 + c1, g.nl[2] + nh), g.nl[2], g.wrap[2]);
    k.pcol0 = (pj0 + nh - 1)*ys + (pk0 + nh - 1);
    k.pcol1 = (pj1 + nh - 1)*ys + (pk1 + nh - 1);
    k.has_e1 = (k.e1 < PS_PN);
  }

  const int istart = k.i0 - 2;

  // prologue: phi planes istart .. istart+3 -> ring slots 0 .. 3; u_x(istart+1), u_x(istart+2) -> slots 1, 2;
  // u_y, u_z (istart+1) -> slot 1.  (The first plane-step prefetches phi(istart+4), u_x(istart+3), u_y/u_z(istart+2).)
#pragma unroll
  for (int d = 0; d < 4; d++) {
    const int xo = (ps_wrap(istart + d, k.nlx, k.wx) + nh - 1)*k.xs;
    pf_cp_async8(&sm.phi[d][k.e0], phi + xo + k.pcol0);
    if (k.has_e1) pf_cp_async8(&sm.phi[d][k.e1], phi + xo + k.pcol1);
    if (d == 1 || d == 2) pf_cp_async8(&sm.ux[d][k.tid], u + xo + k.col);
    if (d == 1) {
      pf_cp_async8(&sm.u[1][0][k.tid], u + k.ns + xo + k.col);
      pf_cp_async8(&sm.u[1][1][k.tid], u + 2*k.ns + xo + k.col);
    }
  }
  pf_cp_async_commit();
  pf_cp_async_wait<0>();
  __syncthreads();

  PfRegs r;
  r.uxc = 0.0;                                     // u_x(n): not used before n = i0 - 1
  pf_plane_sums(sm.phi[0], k.pc, r.Bm, r.Cym, r.Czm);
  pf_plane_sums(sm.phi[1], k.pc, r.Bc, r.Cyc, r.Czc);
  r.gm_xx = r.gm_xy = r.gm_xz = 0.0;
  r.gc_xx = r.gc_xy = r.gc_xz = r.gc_mu = 0.0;
  r.phim1 = 0.0;
  r.fxm1 = r.fxm2 = r.fy_prev = r.fz_prev = 0.0;

  int n = istart;
  // pipeline fill: planes i0-2, i0-1, i0 (phases 0, 1, 2)
  for (; n <= k.i0 && n <= k.i1 + 1; n++) pf_step<ORDER, -1>(sm, r, k, n, phi, u, grad, delsq, force, phinew);
  // steady state: n >= i0+1 and n <= i1-4, phases 3, 4, 5, 0, 1, 2
  for (; n + 5 <= k.i1 - 4; n += 6) {
    pf_step<ORDER, 3>(sm, r, k, n,     phi, u, grad, delsq, force, phinew);
    pf_step<ORDER, 4>(sm, r, k, n + 1, phi, u, grad, delsq, force, phinew);
    pf_step<ORDER, 5>(sm, r, k, n + 2, phi, u, grad, delsq, force, phinew);
    pf_step<ORDER, 0>(sm, r, k, n + 3, phi, u, grad, delsq, force, phinew);
    pf_step<ORDER, 1>(sm, r, k, n + 4, phi, u, grad, delsq, force, phinew);
    pf_step<ORDER, 2>(sm, r, k, n + 5, phi, u, grad, delsq, force, phinew);
  }
  // remaining planes and pipeline drain
  for (; n <= k.i1 + 1; n++) pf_step<ORDER, -1>(sm, r, k, n, phi, u, grad, delsq, force, phinew);
}
#endif

// Planes per x chunk: every chunk pays PS_XPRO extra plane-steps of pipeline fill, and the chunks are
// scheduled in rounds of (SMs x resident CTAs), so pick the chunk count that minimises
// rounds x (planes per chunk + fill).  256^3 on 148 SMs: 6 chunks of 43 planes = 1026 CTAs = 6.9 rounds.
static int ps_pick_xc(const void * kernel, int nt, size_t smem, int tiles, int nx, int fill, int balance = 0) {
  // per device: one process may hold contexts on several GPUs
  static int nsm_dev[LB200_MAX_DEVICES] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < LB200_MAX_DEVICES) ? dev : 0;
  if (nsm_dev[dev] == 0) {
    cudaDeviceGetAttribute(&nsm_dev[dev], cudaDevAttrMultiProcessorCount, dev);
    if (nsm_dev[dev] <= 0) nsm_dev[dev] = 148;
  }
  const int nsm = nsm_dev[dev];
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const long long slots = (long long) nsm*per_sm;
  const char * e = getenv("LB200_PS_XC");
  if (e && atoi(e) > 0) return atoi(e);
  long long best = -1;
  int best_xc = nx;
  for (int nchunk = 1; nchunk <= nx; nchunk++) {
    const int xc = (nx + nchunk - 1)/nchunk;
    if (xc < 8 && nchunk > 1) break;
    const long long rounds = ((long long) tiles*((nx + xc - 1)/xc) + slots - 1)/slots;
    long long cost = rounds*(xc + fill)*100;
    // balance: SMs do not run at one speed (distance to the L2 slices, neighbours' traffic); with only two or three long
    // CTAs per SM the slowest SM sets the time (measured at 256^3: + 15 % with 2 rounds, + 8 % with 4, ~0 with 10),
    // with many short ones the block scheduler evens it out
    if (balance) cost += cost*40/(100*rounds);
    if (best < 0 || cost < best) { best = cost; best_xc = xc; }
  }
  return best_xc;
}

int launch_phi_sector(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, const double * phi,
		      const double * u, double * grad, double * delsq, double * force, double * phinew) {
  dim3 blk(PS_BZ, PS_BY, 1);
  const int gz = (g.nl[2] + 1 + PS_TZ - 1)/PS_TZ, gy = (g.nl[1] + 1 + PS_TY - 1)/PS_TY;
#ifdef LB200_STRICT
#define LB200_PS_KERNEL phi_sector_kernel
  const size_t smem = sizeof(PsShared);
  const int fill = 3;
#else
#define LB200_PS_KERNEL phi_sector_fast_kernel
  static const int smem_pad = tuned_flag("LB200_PS_SMEM_PAD", 0);     // tuning: extra dynamic shared memory (limits CTAs per SM)
  const size_t smem = sizeof(PfShared) + (size_t) smem_pad;
  const int fill = 6;       // 4 extra plane-steps + the slower generic steps of fill and drain
#endif
  // the opt-in to > 48 kB of dynamic shared memory is a per-device function attribute
  static bool configured[LB200_MAX_DEVICES] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= LB200_MAX_DEVICES || !configured[dev]) {
    cudaFuncSetAttribute(LB200_PS_KERNEL<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(LB200_PS_KERNEL<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(LB200_PS_KERNEL<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (dev >= 0 && dev < LB200_MAX_DEVICES) configured[dev] = true;
  }
  const int nx = (g.xcnt > 0) ? g.xcnt : g.nl[0];
  const int xc = (g.xchunk > 0) ? g.xchunk : ps_pick_xc((const void *) LB200_PS_KERNEL<3>, PS_NT, smem, gz*gy, nx, fill);
  dim3 grd(gz, gy, (nx + xc - 1)/xc);
  if (sp.order == 1)      LB200_PS_KERNEL<1><<<grd, blk, smem, st>>>(g, sp, xc, phi, u, grad, delsq, force, phinew);
  else if (sp.order == 2) LB200_PS_KERNEL<2><<<grd, blk, smem, st>>>(g, sp, xc, phi, u, grad, delsq, force, phinew);
  else                    LB200_PS_KERNEL<3><<<grd, blk, smem, st>>>(g, sp, xc, phi, u, grad, delsq, force, phinew);
#undef LB200_PS_KERNEL
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU flags of the peer-store exchange.  The signal runs in stream order after the producing
// kernel, whose peer stores are therefore complete and visible system-wide; the consumer waits with a
// stream memory operation (cuStreamWaitValue32) or, where that is unavailable, with this polling kernel.
// ---------------------------------------------------------------------------------------------

__global__ void signal_kernel(unsigned int * a, unsigned int * b, unsigned int value) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (a != nullptr) *((volatile unsigned int *) a) = value;
    if (b != nullptr) *((volatile unsigned int *) b) = value;
    __threadfence_system();
  }
}

// two flag pairs at one point of the stream (the one-kernel step produces the phi planes and the f / u planes together)
__global__ void signal2_kernel(unsigned int * a, unsigned int * b, unsigned int vab, unsigned int * c, unsigned int * d, unsigned int vcd) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (a != nullptr) *((volatile unsigned int *) a) = vab;
    if (b != nullptr) *((volatile unsigned int *) b) = vab;
    __threadfence_system();                        // whoever sees the second pair's value also sees the first pair's
    if (c != nullptr) *((volatile unsigned int *) c) = vcd;
    if (d != nullptr) *((volatile unsigned int *) d) = vcd;
    __threadfence_system();
  }
}

__global__ void spin_wait_kernel(const unsigned int * flag, unsigned int value, long long timeout_ns, int * err) {
  if (threadIdx.x == 0) {
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    // (int) difference: the counters wrap around after 2^32 exchanges
    while ((int) (*((volatile const unsigned int *) flag) - value) < 0) {
      long long t1;
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      // a neighbour that never signals: fail loudly (the stream errors out, lb200_sync returns LB200_ECUDA)
      // instead of letting the consumer kernel run on stale halo planes
      if (t1 - t0 > timeout_ns) { *err = 1; __threadfence_system(); __trap(); }
    }
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------------
// cahn_hilliard_options_conserve 2 (PHI_CONSERVE_GLOBAL_SUBTRACT): after the forward step the sum of phi over the
// fluid sites is brought back to its initial value by subtracting (sum - sum0)/nfluid everywhere
// (phi_ch_subtract_sum_phi_after_forward_step, src/phi_cahn_hilliard.c:1102-1169, kernels :1225-1319).  The
// reference's own summation order depends on its thread count; here the order is fixed (so results are reproducible
// run to run and independent of the GPU count up to the last bit of the sum) and the sum is compensated
// (Neumaier), as the reference's statistics code compensates the initial sum (src/cahn_hilliard_stats.c:150-330).
//   partial[2b], partial[2b+1] = (sum, compensation) of block b; partial[2 nblk] = fluid sites of the lattice
// ---------------------------------------------------------------------------------------------

constexpr int PSUM_BLOCKS = 256;
constexpr int PSUM_TPB = 256;

__device__ __forceinline__ void neumaier_add(double & sum, double & c, double v) {
  const double t = sum + v;
  c += (fabs(sum) >= fabs(v)) ? ((sum - t) + v) : ((v - t) + sum);
  sum = t;
}

__global__ void __launch_bounds__(PSUM_TPB)
phi_sum_partial_kernel(const Lb200Geom g, const double * __restrict__ phi, const char * __restrict__ status,
		       double * __restrict__ partial) {
  __shared__ double ssum[PSUM_TPB], scmp[PSUM_TPB];
  __shared__ int scnt[PSUM_TPB];
  const long long nint = (long long) g.nl[0]*g.nl[1]*g.nl[2];
  double sum = 0.0, cmp = 0.0;
  int cnt = 0;
  // interior sites in (i, j, k) order, dealt round-robin to the threads of the grid
  for (long long q = (long long) blockIdx.x*PSUM_TPB + threadIdx.x; q < nint; q += (long long) PSUM_BLOCKS*PSUM_TPB) {
    const int kc = 1 + (int) (q % g.nl[2]);
    const int jc = 1 + (int) ((q/g.nl[2]) % g.nl[1]);
    const int ic = 1 + (int) (q/((long long) g.nl[2]*g.nl[1]));
    const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
    if (status == nullptr || status[index] == 0) { neumaier_add(sum, cmp, phi[index]); cnt += 1; }
  }
  ssum[threadIdx.x] = sum; scmp[threadIdx.x] = cmp; scnt[threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, c = 0.0;
    int n = 0;
    for (int t = 0; t < PSUM_TPB; t++) { neumaier_add(s, c, ssum[t]); c += scmp[t]; n += scnt[t]; }
    partial[2*blockIdx.x] = s; partial[2*blockIdx.x + 1] = c;
    atomicAdd(reinterpret_cast<int *>(partial + 2*PSUM_BLOCKS), n);          // integer: order does not matter
  }
}

// result[0] = compensated sum of the partials, result[1] = fluid sites (as a double); one thread, fixed order
__global__ void phi_sum_final_kernel(double * __restrict__ partial, double * __restrict__ result) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0, c = 0.0;
  for (int b = 0; b < PSUM_BLOCKS; b++) { neumaier_add(s, c, partial[2*b]); c += partial[2*b + 1]; }
  result[0] = s + c;
  result[1] = (double) *reinterpret_cast<int *>(partial + 2*PSUM_BLOCKS);
  *reinterpret_cast<int *>(partial + 2*PSUM_BLOCKS) = 0;
}

// all[2r], all[2r+1] = (sum, nfluid) of rank r -> total[0], total[1], summed in rank order on every GPU
__global__ void phi_sum_ranks_kernel(const double * __restrict__ all, int nranks, double * __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0, c = 0.0, n = 0.0;
  for (int r = 0; r < nranks; r++) { neumaier_add(s, c, all[2*r]); n += all[2*r + 1]; }
  total[0] = s + c; total[1] = n;
}

__global__ void __launch_bounds__(TPB_MAX)
phi_subtract_kernel(const Lb200Geom g, const double * __restrict__ total, double phi0, const char * __restrict__ status,
		    double * __restrict__ phi) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  if (status != nullptr && status[index] != 0) return;
  // phi -= (correct->phi - correct->phi0)/correct->nfluid, src/phi_cahn_hilliard.c:1312
  phi[index] = phi[index] - (total[0] - phi0)/total[1];
}

int launch_phi_sum(cudaStream_t st, const Lb200Geom & g, const double * phi, const char * status, double * partial, double * result) {
  phi_sum_partial_kernel<<<PSUM_BLOCKS, PSUM_TPB, 0, st>>>(g, phi, status, partial);
  phi_sum_final_kernel<<<1, 32, 0, st>>>(partial, result);
  return 2;
}

int launch_phi_sum_ranks(cudaStream_t st, const double * all, int nranks, double * total) {
  phi_sum_ranks_kernel<<<1, 32, 0, st>>>(all, nranks, total);
  return 1;
}

int launch_phi_subtract(cudaStream_t st, const Lb200Geom & g, const double * total, double phi0, const char * status, double * phi) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
  phi_subtract_kernel<<<grd, blk, 0, st>>>(g, total, phi0, status, phi);
  return 1;
}

int launch_signal(cudaStream_t st, unsigned int * a, unsigned int * b, unsigned int value) {
  signal_kernel<<<1, 32, 0, st>>>(a, b, value);
  return 1;
}

int launch_signal2(cudaStream_t st, unsigned int * a, unsigned int * b, unsigned int vab, unsigned int * c, unsigned int * d, unsigned int vcd) {
  signal2_kernel<<<1, 32, 0, st>>>(a, b, vab, c, d, vcd);
  return 1;
}

int launch_spin_wait(cudaStream_t st, const unsigned int * flag, unsigned int value, int timeout_ms, int * err) {
  spin_wait_kernel<<<1, 32, 0, st>>>(flag, value, (long long) timeout_ms*1000000LL, err);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// zero everything outside the interior (materialises "logically zero" halos of force / u)
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB_MAX)
zero_outside_kernel(const Lb200Geom g, int ncomp, double * __restrict__ data) {
  const int k0 = blockIdx.x*blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y*blockDim.y + threadIdx.y;
  const int i0 = blockIdx.z;
  if (k0 >= g.nall[2] || j0 >= g.nall[1]) return;
  const int ic = i0 - g.nh + 1, jc = j0 - g.nh + 1, kc = k0 - g.nh + 1;
  const bool inside = (ic >= 1 && ic <= g.nl[0] && jc >= 1 && jc <= g.nl[1] && kc >= 1 && kc <= g.nl[2]);
  if (inside) return;
  const size_t idx = ((size_t) i0*g.nall[1] + j0)*g.nall[2] + k0;
  for (int c = 0; c < ncomp; c++) data[(size_t) c*g.nsites + idx] = 0.0;
}

int launch_zero_outside(cudaStream_t st, const Lb200Geom & g, int ncomp, double * data) {
  dim3 blk;
  block_shape(g.nall[2], blk);
  dim3 grd((g.nall[2] + blk.x - 1)/blk.x, (g.nall[1] + blk.y - 1)/blk.y, g.nall[0]);
  zero_outside_kernel<<<grd, blk, 0, st>>>(g, ncomp, data);
  return 1;
}

#include "lb200_fused.cuh"
#include "lb200_fused_ws.cuh"
#include "lb200_le.cuh"
#include "lb200_lc.cuh"

}  // anonymous namespace
}  // namespace lb200_fast / lb200_strict

#ifdef LB200_STRICT
using namespace lb200_strict;
#else
using namespace lb200_fast;
#endif

const Lb200Kernels LB200_TABLE = {
  launch_collide,
  launch_propagate,
  launch_halo,
  launch_grad27,
  launch_phi_force,
  launch_cahn_hilliard,
  launch_force_ch,
  launch_phi_sector,
  launch_zero_outside,
  launch_phi_from_g,
  launch_phi_to_g,
  launch_collide_binary,
  launch_stress,
  launch_force_from_stress,
  launch_signal,
  launch_spin_wait,
  launch_le_interp,
  launch_le_grad_planes,
  launch_le_force_prep,
  launch_le_ch_prep,
  launch_le_prep_both,
  launch_le_force_ch,
  launch_le_lb_bc,
  launch_grad7,
  launch_lc_stress,
  launch_lc_force_be,
  launch_f_convert,
  launch_collide_f32,
  launch_step_fused,
  launch_phi_sum,
  launch_phi_sum_ranks,
  launch_phi_subtract,
  launch_le_yz_images,
  launch_le_interp_both,
  launch_signal2,
  PSUM_BLOCKS,
};
