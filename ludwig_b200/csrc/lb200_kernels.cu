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

__host__ inline void block_shape(int nz, dim3 & blk) {
  int bx = ((nz + 31)/32)*32;
  if (bx > TPB) bx = TPB;
  blk = dim3(bx, TPB/bx, 1);
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
template <bool PULL, bool GHOST, bool HAS_FORCE, bool HAS_MAP>
__global__ void __launch_bounds__(TPB)
collide_d3q19_kernel(const Lb200Geom g, const Lb200CollideDev cp,
		     const double * __restrict__ fsrc, double * __restrict__ fdst,
		     const double * __restrict__ hforce, const char * __restrict__ status,
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

#pragma unroll
  for (int p = 0; p < 19; p++) {
    const int off = PULL ? (CV19[p][0]*g.xs + CV19[p][1]*g.ys + CV19[p][2]) : 0;
    f[p] = fsrc[p*ns + (index - off)];
  }

  if (HAS_MAP) {
    if (status[index] != 0) {
      // non-fluid site: propagation still moves the populations; no collision, no rho/u
      if (PULL) {
#pragma unroll
	for (int p = 0; p < 19; p++) fdst[p*ns + index] = f[p];
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
  for (int p = 0; p < 19; p++) fdst[p*ns + index] = f[p];

  rho_out[index] = rho;
#pragma unroll
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];
}

// Generic velocity set (D3Q15, D3Q27; also D3Q19 with the model matrices instead of the coded
// constants): reference src/collision.c:335-342, 541-551.
template <bool PULL>
__global__ void __launch_bounds__(TPB)
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

  for (int p = 0; p < nvel; p++) {
    const int off = PULL ? (md->cv[p][0]*g.xs + md->cv[p][1]*g.ys + md->cv[p][2]) : 0;
    f[p] = fsrc[p*ns + (index - off)];
  }

  if (status != nullptr && status[index] != 0) {
    if (PULL) {
      for (int p = 0; p < nvel; p++) fdst[p*ns + index] = f[p];
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

  for (int p = 0; p < nvel; p++) {
    double s = 0.0;
    for (int m = 0; m < nvel; m++) s += md->mi[p][m]*mode[m];
    fdst[p*ns + index] = s;
  }

  rho_out[index] = rho;
  for (int ia = 0; ia < 3; ia++) u_out[ia*ns + index] = u[ia];
}

int launch_collide(cudaStream_t st, const Lb200Geom & g, const Lb200CollideDev & cp,
		   const Lb200ModelDev * md, int nvel, int pull, const double * fsrc,
		   double * fdst, const double * force, const char * status,
		   double * rho, double * u) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);

  if (nvel == 19 && md == nullptr) {
#define LB200_GO(P, G, F, M) collide_d3q19_kernel<P, G, F, M><<<grd, blk, 0, st>>>(g, cp, fsrc, fdst, force, status, rho, u)
#define LB200_SEL_M(P, G, F) do { if (status) LB200_GO(P, G, F, true); else LB200_GO(P, G, F, false); } while (0)
#define LB200_SEL_F(P, G) do { if (force) LB200_SEL_M(P, G, true); else LB200_SEL_M(P, G, false); } while (0)
#define LB200_SEL_G(P) do { if (cp.ghost) LB200_SEL_F(P, true); else LB200_SEL_F(P, false); } while (0)
    if (pull) LB200_SEL_G(true); else LB200_SEL_G(false);
#undef LB200_GO
#undef LB200_SEL_M
#undef LB200_SEL_F
#undef LB200_SEL_G
  }
  else {
    if (pull) collide_generic_kernel<true><<<grd, blk, 0, st>>>(g, cp, md, fsrc, fdst, force, status, rho, u);
    else      collide_generic_kernel<false><<<grd, blk, 0, st>>>(g, cp, md, fsrc, fdst, force, status, rho, u);
  }
  return 1;
}

// ---------------------------------------------------------------------------------------------
// lb_propagation as a stand-alone sweep: reference src/propagation.c:153-200.
// x in [1,N], every y,z of the allocation; y/z halo sites copy themselves.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB)
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
		  double * __restrict__ data, const double * __restrict__ xlo,
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
    const int si = ic + mx*g.nl[0];
    src = data;
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
		int depth, int reduced, double * data, const double * xlo, const double * xhi) {
  const long long ey = g.nl[1] + 2*depth;
  const long long ez = g.nl[2] + 2*depth;
  const long long nx_slab = (long long) depth*ey*ez;
  const long long ny_slab = (long long) g.nl[0]*depth*ez;
  const long long nz_slab = (long long) g.nl[0]*g.nl[1]*depth;
  const long long total = 2*(nx_slab + ny_slab + nz_slab);
  const int nblk = (int) ((total + TPB - 1)/TPB);
  if (reduced) halo_shell_kernel<true><<<nblk, TPB, 0, st>>>(g, md, ncomp, depth, data, xlo, xhi, nx_slab, ny_slab, nz_slab);
  else         halo_shell_kernel<false><<<nblk, TPB, 0, st>>>(g, md, ncomp, depth, data, xlo, xhi, nx_slab, ny_slab, nz_slab);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// 27-point gradient, reference src/gradient_3d_27pt_fluid.c:219-363, on [1-ne, N+ne]^3 with
// ne = nhalo - 1 (:91-95).  Summation order as written there.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB)
grad27_kernel(const Lb200Geom g, const double * __restrict__ field, double * __restrict__ grad,
	      double * __restrict__ delsq) {
  const int ne = g.nh - 1;
  const int kc = 1 - ne + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 - ne + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 - ne + blockIdx.z;
  if (kc > g.nl[2] + ne || jc > g.nl[1] + ne) return;

  const int ys = g.ys;
  const int index = ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
  const int indexm1 = index - g.xs;
  const int indexp1 = index + g.xs;
  const size_t ns = (size_t) g.nsites;
  const double r9 = (1.0/9.0);

  // the 27 values, named [x][y][z] with 0 = -1, 1 = 0, 2 = +1
  const double m_mm = field[indexm1-ys-1], m_m0 = field[indexm1-ys], m_mp = field[indexm1-ys+1];
  const double m_0m = field[indexm1   -1], m_00 = field[indexm1   ], m_0p = field[indexm1   +1];
  const double m_pm = field[indexm1+ys-1], m_p0 = field[indexm1+ys], m_pp = field[indexm1+ys+1];
  const double c_mm = field[index  -ys-1], c_m0 = field[index  -ys], c_mp = field[index  -ys+1];
  const double c_0m = field[index     -1], c_00 = field[index     ], c_0p = field[index     +1];
  const double c_pm = field[index  +ys-1], c_p0 = field[index  +ys], c_pp = field[index  +ys+1];
  const double p_mm = field[indexp1-ys-1], p_m0 = field[indexp1-ys], p_mp = field[indexp1-ys+1];
  const double p_0m = field[indexp1   -1], p_00 = field[indexp1   ], p_0p = field[indexp1   +1];
  const double p_pm = field[indexp1+ys-1], p_p0 = field[indexp1+ys], p_pp = field[indexp1+ys+1];

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
}

int launch_grad27(cudaStream_t st, const Lb200Geom & g, const double * phi, double * grad,
		  double * delsq) {
  const int ne = g.nh - 1;
  const int ex = g.nl[0] + 2*ne, ey = g.nl[1] + 2*ne, ez = g.nl[2] + 2*ne;
  dim3 blk;
  block_shape(ez, blk);
  dim3 grd((ez + blk.x - 1)/blk.x, (ey + blk.y - 1)/blk.y, ex);
  grad27_kernel<<<grd, blk, 0, st>>>(g, phi, grad, delsq);
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
// Cahn-Hilliard face fluxes at one site, fused: advective part (upwind order 1/2/3, reference
// src/advection.c:538-629, 770-893, 946-1141), - M (mu1 - mu0) (src/phi_cahn_hilliard.c:350-404),
// - M grad mu_ext (:1373-1397), no-normal-flux mask (src/advection_bcs.c:80-130), and the forward
// Euler update (src/phi_cahn_hilliard.c:1018-1049).  The reference keeps four flux arrays and
// five kernels; here the six face fluxes of a site are formed in registers.
// ---------------------------------------------------------------------------------------------

// flux through the face between site s and s + str ("east"-like face of s), component velocity
// u0 = u_a(s), u1 = u_a(s + str)
template <int ORDER>
__device__ __forceinline__ double adv_hi(const double * __restrict__ phi, int s, int str,
					 double u0, double u1) {
  if (ORDER == 1) {
    const double uf = 0.5*(u0 + u1);
    const int idx = (uf < 0.0) ? s + str : s;
    return uf*phi[idx];
  }
  else if (ORDER == 2) {
    return 0.5*(u0 + u1)*1.0*0.5*(phi[s] + phi[s + str]);
  }
  else {
    const double a1 = -0.213933;
    const double a2 =  0.927865;
    const double a3 =  0.286067;
    const double uf = 0.5*(u0 + u1);
    double fd1, fd2, fd3;
    if (uf < 0.0) { fd1 = phi[s + 2*str]; fd2 = phi[s + str]; fd3 = phi[s]; }
    else          { fd1 = phi[s - str];   fd2 = phi[s];       fd3 = phi[s + str]; }
    return uf*(a1*fd1 + a2*fd2 + a3*fd3);
  }
}

// flux through the face between s - str and s as computed AT s ("west" face, x only)
template <int ORDER>
__device__ __forceinline__ double adv_west(const double * __restrict__ phi, int s, int str,
					   double u0, double u1) {
  if (ORDER == 1) {
    const double uf = 0.5*(u0 + u1);
    const int idx = (uf > 0.0) ? s - str : s;
    return uf*phi[idx];
  }
  else if (ORDER == 2) {
    return 0.5*(u0 + u1)*1.0*0.5*(phi[s - str] + phi[s]);
  }
  else {
    const double a1 = -0.213933;
    const double a2 =  0.927865;
    const double a3 =  0.286067;
    const double uf = 0.5*(u0 + u1);
    double fd1, fd2, fd3;
    if (uf > 0.0) { fd1 = phi[s - 2*str]; fd2 = phi[s - str]; fd3 = phi[s]; }
    else          { fd1 = phi[s + str];   fd2 = phi[s];       fd3 = phi[s - str]; }
    return uf*(a1*fd1 + a2*fd2 + a3*fd3);
  }
}

template <int ORDER, bool HAS_MAP>
__device__ __forceinline__ double site_phi_update(const Lb200Geom & g, const Lb200SymmDev & sp,
						  const double * __restrict__ phi,
						  const double * __restrict__ delsq,
						  const double * __restrict__ u,
						  const char * __restrict__ status, int s) {
  const size_t ns = (size_t) g.nsites;
  const int xs = g.xs, ys = g.ys;
  const double M = sp.mobility;

  const double mu0 = symm_mu(sp, phi[s], delsq[s]);
  const double ux0 = u[0*ns + s], uy0 = u[1*ns + s], uz0 = u[2*ns + s];

  double mk = 1.0, mkxm = 1.0, mkxp = 1.0, mkyp = 1.0, mkym = 1.0, mkzp = 1.0, mkzm = 1.0;
  if (HAS_MAP) {
    mk   = (status[s] == 0);
    mkxm = (status[s - xs] == 0); mkxp = (status[s + xs] == 0);
    mkym = (status[s - ys] == 0); mkyp = (status[s + ys] == 0);
    mkzm = (status[s - 1] == 0);  mkzp = (status[s + 1] == 0);
  }

  // west (computed at s): u1 = u_x(s - x)
  double fw = adv_west<ORDER>(phi, s, xs, ux0, u[0*ns + s - xs]);
  fw -= M*(mu0 - symm_mu(sp, phi[s - xs], delsq[s - xs]));
  fw -= M*sp.gm[0];
  if (HAS_MAP) fw *= mk*mkxm;
  // east
  double fe = adv_hi<ORDER>(phi, s, xs, ux0, u[0*ns + s + xs]);
  fe -= M*(symm_mu(sp, phi[s + xs], delsq[s + xs]) - mu0);
  fe -= M*sp.gm[0];
  if (HAS_MAP) fe *= mk*mkxp;
  // y: face (s, s+y) computed at s
  double fy = adv_hi<ORDER>(phi, s, ys, uy0, u[1*ns + s + ys]);
  fy -= M*(symm_mu(sp, phi[s + ys], delsq[s + ys]) - mu0);
  fy -= M*sp.gm[1];
  if (HAS_MAP) fy *= mk*mkyp;
  // y: face (s-y, s) computed at s - y
  double fym = adv_hi<ORDER>(phi, s - ys, ys, u[1*ns + s - ys], uy0);
  fym -= M*(mu0 - symm_mu(sp, phi[s - ys], delsq[s - ys]));
  fym -= M*sp.gm[1];
  if (HAS_MAP) fym *= mkym*mk;
  // z
  double fz = adv_hi<ORDER>(phi, s, 1, uz0, u[2*ns + s + 1]);
  fz -= M*(symm_mu(sp, phi[s + 1], delsq[s + 1]) - mu0);
  fz -= M*sp.gm[2];
  if (HAS_MAP) fz *= mk*mkzp;
  double fzm = adv_hi<ORDER>(phi, s - 1, 1, u[2*ns + s - 1], uz0);
  fzm -= M*(mu0 - symm_mu(sp, phi[s - 1], delsq[s - 1]));
  fzm -= M*sp.gm[2];
  if (HAS_MAP) fzm *= mkzm*mk;

  double ph = phi[s];
  ph -= (+ fe - fw + fy - fym + sp.wz*fz - sp.wz*fzm);
  return ph;
}

template <bool DO_FORCE, bool DO_CH, bool ACCUM, int ORDER, bool HAS_MAP>
__global__ void __launch_bounds__(TPB)
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

  if (DO_FORCE) {
    double fo[3];
    const SiteFE s0 = load_fe(phi, grad, delsq, ns, s);
    const SiteFE xp = load_fe(phi, grad, delsq, ns, s + g.xs);
    const SiteFE xm = load_fe(phi, grad, delsq, ns, s - g.xs);
    const SiteFE yp = load_fe(phi, grad, delsq, ns, s + g.ys);
    const SiteFE ym = load_fe(phi, grad, delsq, ns, s - g.ys);
    const SiteFE zp = load_fe(phi, grad, delsq, ns, s + 1);
    const SiteFE zm = load_fe(phi, grad, delsq, ns, s - 1);
    site_force(sp, s0, xp, xm, yp, ym, zp, zm, fo);
    for (int a = 0; a < 3; a++) {
      if (ACCUM) force[a*ns + s] += fo[a];
      else       force[a*ns + s] = fo[a];
    }
  }

  if (DO_CH) {
    phinew[s] = site_phi_update<ORDER, HAS_MAP>(g, sp, phi, delsq, u, status, s);
  }
}

template <bool DO_FORCE, bool DO_CH>
int launch_force_ch_t(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, int accumulate,
		      const double * phi, const double * grad, const double * delsq,
		      const double * u, const char * status, double * force, double * phinew) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
#define LB200_GO(A, O, M) force_ch_kernel<DO_FORCE, DO_CH, A, O, M><<<grd, blk, 0, st>>>(g, sp, phi, grad, delsq, u, status, force, phinew)
#define LB200_SEL_M(A, O) do { if (status) LB200_GO(A, O, true); else LB200_GO(A, O, false); } while (0)
#define LB200_SEL_O(A) do { if (sp.order == 1) LB200_SEL_M(A, 1); else if (sp.order == 2) LB200_SEL_M(A, 2); else LB200_SEL_M(A, 3); } while (0)
  if (accumulate) LB200_SEL_O(true); else LB200_SEL_O(false);
#undef LB200_GO
#undef LB200_SEL_M
#undef LB200_SEL_O
  return 1;
}

int launch_phi_force(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, int accumulate,
		     const double * phi, const double * grad, const double * delsq, double * force) {
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
// zero everything outside the interior (materialises "logically zero" halos of force / u)
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB)
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
  launch_zero_outside,
};
