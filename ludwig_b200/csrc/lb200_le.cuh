// lb200_le.cuh -- Lees-Edwards sliding periodic planes, included by lb200_kernels.cu
// inside its anonymous namespace (so it is built twice: fast and strict).
//
// The reference does every one of these steps on the HOST in its GPU build (whole-array device->host and
// host->device copies around serial loops: src/field.c:443-446, src/hydro.c:379-382, src/phi_force.c:95-97,
// src/phi_cahn_hilliard.c:641-645, src/gradient_3d_27pt_fluid.c:416-419).  Here they are kernels on the planes
// concerned; nothing leaves the device.
//
// Geometry (reference src/leesedwards.c): plane p lies between x = loc[p] and loc[p] + 1.  Field arrays carry
// nxbuf = 2*nhalo*nplane extra x-planes after the high x halo; buffer plane ib (x coordinate N + nhalo + 1 + ib)
// holds the real plane ibuff_to_real(ib) displaced along y by -/+ uy t and interpolated.  A stencil that
// crosses a plane reads the buffer plane instead of the real neighbour (lees_edw_ic_to_buff).

constexpr int LE_RED_NT = 1024;

// lees_edw_ic_to_buff, src/leesedwards.c:1030-1065: x coordinate of the neighbour di planes from ic
__device__ __forceinline__ int le_x(const Lb200LeDev & le, const Lb200Geom & g, int ic, int di) {
  if (le.nplane > 0) {
    int p = ic/le.xblock;
    p = max(0, min(p, le.nplane - 1));
    const int nh = g.nh;
    int ip = le.loc[p] - (nh - 1);
    if (di > 0 && ic >= ip && ic < ip + nh && ic + di >= ip + nh) return g.nl[0] + (1 + 2*p)*nh + (ic - ip + 1) + di;
    ip = le.loc[p] + 1;
    if (di < 0 && ic >= ip && ic < ip + nh && ic + di < ip) return g.nl[0] + (2 + 2*p)*nh + (ic - ip + 1) + di;
  }
  int x = ic + di;
  if (g.wrap[0]) { if (x < 1) x += g.nl[0]; else if (x > g.nl[0]) x -= g.nl[0]; }
  return x;
}

// Halo-free steps (Lb200Geom::wrap): the periodic images in y / z are read from the interior sites they mirror
// (select arithmetic, no branches: whether ptxas turns `if / else if` into SEL or into divergent branches depends on
// unrelated details of the kernel -- the liquid-crystal sweeps lost 5-11 % when it chose branches)
__device__ __forceinline__ int lb200_wrap1(int i, int n, int wrap) {
  const int w = wrap ? n : 0;
  return i + ((i < 1) ? w : 0) - ((i > n) ? w : 0);
}
__device__ __forceinline__ int le_wy(const Lb200Geom & g, int j) { return lb200_wrap1(j, g.nl[1], g.wrap[1]); }
__device__ __forceinline__ int le_wz(const Lb200Geom & g, int k) { return lb200_wrap1(k, g.nl[2], g.wrap[2]); }

__device__ __forceinline__ int le_index(const Lb200Geom & g, int ic, int jc, int kc) {
  return ((ic + g.nh - 1)*g.nall[1] + (jc + g.nh - 1))*g.nall[2] + (kc + g.nh - 1);
}

// ---------------------------------------------------------------------------------------------
// Buffer planes of a field.  CUBIC: field_leesedwards, 4-point Lagrange (src/field.c:460-505);
// !CUBIC: hydro_lees_edwards, linear + velocity jump (src/hydro.c:386-430).  The per-sign integer
// displacement and weights are formed on the host exactly as the reference does (fmod, floor).
// One thread per (jc, kc) of the allocation and buffer plane.
// ---------------------------------------------------------------------------------------------

template <bool CUBIC>
__global__ void __launch_bounds__(TPB_MAX)
le_interp_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200LeInterp ip, int ncomp, int zext,
		 double * __restrict__ data) {
  const int kc = 1 - zext + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 - g.nh + blockIdx.y*blockDim.y + threadIdx.y;
  const int ib = blockIdx.z;
  if (kc > g.nl[2] + zext || jc > g.nl[1] + g.nh) return;
  const int nh = g.nh;
  const int p = ib/(2*nh);
  const int r = ib % (2*nh);
  const int ic = le.loc[p] - (nh - 1) + r;                 // lees_edw_ibuff_to_real
  const int sgn = (r < nh) ? 0 : 1;                        // lees_edw_buffer_duy: -1 / +1
  const int jdy = ip.jdy[sgn];
  const int ny = g.nl[1];
  const size_t ns = (size_t) g.nsites;
  const int dst = le_index(g, g.nl[0] + nh + 1 + ib, jc, kc);
  const int ks = le_wz(g, kc);                             // source column (its z halo may be stale in halo-free steps)

  if (CUBIC) {
    const int j0 = 1 + (jc - jdy - 3 + 2*ny) % ny;
    const int j1 = 1 + j0 % ny;
    const int j2 = 1 + j1 % ny;
    const int j3 = 1 + j2 % ny;
    const double w0 = ip.w[sgn][0], w1 = ip.w[sgn][1], w2 = ip.w[sgn][2], w3 = ip.w[sgn][3];
    for (int n = 0; n < ncomp; n++) {
      const double * d = data + n*ns;
      data[n*ns + dst] = - w0*d[le_index(g, ic, j0, ks)] + w1*d[le_index(g, ic, j1, ks)]
	- w2*d[le_index(g, ic, j2, ks)] + w3*d[le_index(g, ic, j3, ks)];
    }
  }
  else {
    const int j1 = 1 + (jc - jdy - 2 + 2*ny) % ny;
    const int j2 = 1 + j1 % ny;
    const double fr = ip.w[sgn][0], omfr = ip.w[sgn][1];
    for (int n = 0; n < ncomp; n++) {
      const double * d = data + n*ns;
      const double ule = (n == 1) ? le.uy*(sgn ? 1 : -1) : 0.0;
      data[n*ns + dst] = ule + d[le_index(g, ic, j1, ks)]*fr + d[le_index(g, ic, j2, ks)]*omfr;
    }
  }
}

// field_leesedwards(phi) and hydro_lees_edwards(u) in one launch (the patch chain of the one-kernel step): same
// expressions as le_interp_kernel<true> / <false>
__global__ void __launch_bounds__(TPB_MAX)
le_interp_both_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200LeInterp ipc, const Lb200LeInterp ipl,
		      int zext, double * __restrict__ phi, double * __restrict__ u) {
  const int kc = 1 - zext + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 - g.nh + blockIdx.y*blockDim.y + threadIdx.y;
  const int ib = blockIdx.z;
  if (kc > g.nl[2] + zext || jc > g.nl[1] + g.nh) return;
  const int nh = g.nh;
  const int p = ib/(2*nh);
  const int r = ib % (2*nh);
  const int ic = le.loc[p] - (nh - 1) + r;
  const int sgn = (r < nh) ? 0 : 1;
  const int ny = g.nl[1];
  const size_t ns = (size_t) g.nsites;
  const int dst = le_index(g, g.nl[0] + nh + 1 + ib, jc, kc);
  const int ks = le_wz(g, kc);
  {
    const int jdy = ipc.jdy[sgn];
    const int j0 = 1 + (jc - jdy - 3 + 2*ny) % ny;
    const int j1 = 1 + j0 % ny;
    const int j2 = 1 + j1 % ny;
    const int j3 = 1 + j2 % ny;
    const double w0 = ipc.w[sgn][0], w1 = ipc.w[sgn][1], w2 = ipc.w[sgn][2], w3 = ipc.w[sgn][3];
    phi[dst] = - w0*phi[le_index(g, ic, j0, ks)] + w1*phi[le_index(g, ic, j1, ks)]
      - w2*phi[le_index(g, ic, j2, ks)] + w3*phi[le_index(g, ic, j3, ks)];
  }
  {
    const int jdy = ipl.jdy[sgn];
    const int j1 = 1 + (jc - jdy - 2 + 2*ny) % ny;
    const int j2 = 1 + j1 % ny;
    const double fr = ipl.w[sgn][0], omfr = ipl.w[sgn][1];
    const int s1 = le_index(g, ic, j1, ks), s2 = le_index(g, ic, j2, ks);
#pragma unroll
    for (int n = 0; n < 3; n++) {
      const double * d = u + n*ns;
      const double ule = (n == 1) ? le.uy*(sgn ? 1 : -1) : 0.0;
      u[n*ns + dst] = ule + d[s1]*fr + d[s2]*omfr;
    }
  }
}

int launch_le_interp_both(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200LeInterp & ipc,
			  const Lb200LeInterp & ipl, int zext, double * phi, double * u) {
  if (le.nplane == 0) return 0;
  dim3 blk;
  block_shape(g.nl[2] + 2*zext, blk);
  dim3 grd((g.nl[2] + 2*zext + blk.x - 1)/blk.x, (g.nall[1] + blk.y - 1)/blk.y, 2*g.nh*le.nplane);
  le_interp_both_kernel<<<grd, blk, 0, st>>>(g, le, ipc, ipl, zext, phi, u);
  return 1;
}

int launch_le_interp(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200LeInterp & ip,
		     int cubic, int ncomp, int zext, double * data) {
  if (le.nplane == 0) return 0;
  dim3 blk;
  block_shape(g.nl[2] + 2*zext, blk);
  dim3 grd((g.nl[2] + 2*zext + blk.x - 1)/blk.x, (g.nall[1] + blk.y - 1)/blk.y, 2*g.nh*le.nplane);
  if (cubic) le_interp_kernel<true><<<grd, blk, 0, st>>>(g, le, ip, ncomp, zext, data);
  else       le_interp_kernel<false><<<grd, blk, 0, st>>>(g, le, ip, ncomp, zext, data);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// 27-point gradient at a list of x-planes given as (x-1, x, x+1) coordinate triples: the two real planes
// next to each Lees-Edwards plane (their far x-neighbour is a buffer plane, src/gradient_3d_27pt_fluid.c:
// 250-253) and the buffer planes themselves (grad_3d_27pt_fluid_le, :375-651).  Same summation order as
// grad27_kernel.
// ---------------------------------------------------------------------------------------------

template <bool SEVEN>
__global__ void __launch_bounds__(TPB)
le_grad_planes_kernel(const Lb200Geom g, const int ne, const int * __restrict__ trip,
		      const double * __restrict__ field, double * __restrict__ grad, double * __restrict__ delsq) {
  const int ey = g.nl[1] + 2*ne, ez = g.nl[2] + 2*ne;
  const int q = blockIdx.x*blockDim.x + threadIdx.x;
  if (q >= ey*ez) return;
  const int jc = 1 - ne + q/ez;
  const int kc = 1 - ne + q%ez;
  const int xm = trip[3*blockIdx.y + 0], xc = trip[3*blockIdx.y + 1], xp = trip[3*blockIdx.y + 2];
  const size_t ns = (size_t) g.nsites;
  const double r9 = (1.0/9.0);
  const int index = le_index(g, xc, jc, kc);
  const int jm = le_wy(g, jc - 1), jp = le_wy(g, jc + 1), km = le_wz(g, kc - 1), kp = le_wz(g, kc + 1);
  if (SEVEN) {
    // fd_gradient_calculation 3d_7pt_fluid: grad_3d_7pt_fluid_kernel_v / grad_3d_7pt_fluid_le (src/gradient_3d_7pt_fluid.c:231-300,
    // 317-440), same expressions as grad7_kernel
    const double f0 = field[index];
    const double fxm = field[le_index(g, xm, jc, kc)], fxp = field[le_index(g, xp, jc, kc)];
    const double fym = field[le_index(g, xc, jm, kc)], fyp = field[le_index(g, xc, jp, kc)];
    const double fzm = field[le_index(g, xc, jc, km)], fzp = field[le_index(g, xc, jc, kp)];
    grad[0*ns + index] = 0.5*(fxp - fxm);
    grad[1*ns + index] = 0.5*(fyp - fym);
    grad[2*ns + index] = 0.5*(fzp - fzm);
    delsq[index] = fxp + fxm + fyp + fym + fzp + fzm - 6.0*f0;
    return;
  }

  double m_mm, m_m0, m_mp, m_0m, m_00, m_0p, m_pm, m_p0, m_pp;
  double c_mm, c_m0, c_mp, c_0m, c_00, c_0p, c_pm, c_p0, c_pp;
  double p_mm, p_m0, p_mp, p_0m, p_00, p_0p, p_pm, p_p0, p_pp;
#define LB200_LOAD_PLANE(P, X) \
  P##_mm = field[le_index(g, X, jm, km)]; P##_m0 = field[le_index(g, X, jm, kc)]; P##_mp = field[le_index(g, X, jm, kp)]; \
  P##_0m = field[le_index(g, X, jc, km)]; P##_00 = field[le_index(g, X, jc, kc)]; P##_0p = field[le_index(g, X, jc, kp)]; \
  P##_pm = field[le_index(g, X, jp, km)]; P##_p0 = field[le_index(g, X, jp, kc)]; P##_pp = field[le_index(g, X, jp, kp)]
  LB200_LOAD_PLANE(m, xm);
  LB200_LOAD_PLANE(c, xc);
  LB200_LOAD_PLANE(p, xp);
#undef LB200_LOAD_PLANE

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

// ne < 0: the 7-point stencil (3d_7pt_fluid), extension -ne - 1
int launch_le_grad_planes(cudaStream_t st, const Lb200Geom & g, int ne_, int ntrip, const int * trip,
			  const double * phi, double * grad, double * delsq) {
  if (ntrip == 0) return 0;
  const bool seven = (ne_ < 0);
  const int ne = seven ? -ne_ - 1 : ne_;
  const int ey = g.nl[1] + 2*ne, ez = g.nl[2] + 2*ne;
  dim3 grd((ey*ez + TPB - 1)/TPB, ntrip, 1);
  if (seven) le_grad_planes_kernel<true><<<grd, TPB, 0, st>>>(g, ne, trip, phi, grad, delsq);
  else       le_grad_planes_kernel<false><<<grd, TPB, 0, st>>>(g, ne, trip, phi, grad, delsq);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Flux form of the thermodynamic force, phi_force_flux (src/phi_force.c:289-345): face fluxes
// 1/2 [P(i) + P(i')]_{a,face} (:360-440), the per-plane correction that makes the integrated east flux below a
// plane equal the integrated west flux above it (phi_force_flux_fix_local, :595-673), divergence (:452-495).
//   le_force_term_kernel : per (plane, j, k) the summand  - fluxe(loc) + fluxw(loc + 1)   -> term[p][a][j][k]
//   le_force_sum_kernel  : fcor[p][a] = sum_{j,k} term   (strict: the reference's sequential j, k order;
//                          fast: fixed-shape tree, deterministic)
//   le_force_ch_kernel   : the divergence with the corrected plane fluxes
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void le_pcol_x(const Lb200SymmDev & sp, const double * __restrict__ phi,
					  const double * __restrict__ grad, const double * __restrict__ delsq,
					  size_t ns, int idx, double p[3]) {
  const SiteFE s = load_fe(phi, grad, delsq, ns, idx);
  symm_pcol<0>(sp, s, p);
}

__global__ void __launch_bounds__(TPB_MAX)
le_force_term_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200SymmDev sp,
		     const double * __restrict__ phi, const double * __restrict__ grad,
		     const double * __restrict__ delsq, double * __restrict__ term) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int p = blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const size_t ns = (size_t) g.nsites;
  const int ic = le.loc[p];
  double p0[3], p1[3], fluxe[3], fluxw[3];
  le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, ic, jc, kc), p0);
  le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, le_x(le, g, ic, +1), jc, kc), p1);
  for (int a = 0; a < 3; a++) fluxe[a] = 0.5*(p1[a] + p0[a]);
  le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, ic + 1, jc, kc), p0);
  le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, le_x(le, g, ic + 1, -1), jc, kc), p1);
  for (int a = 0; a < 3; a++) fluxw[a] = 0.5*(p1[a] + p0[a]);
  const size_t nyz = (size_t) g.nl[1]*g.nl[2];
  const size_t q = (size_t) (jc - 1)*g.nl[2] + (kc - 1);
  for (int a = 0; a < 3; a++) term[(p*3 + a)*nyz + q] = - fluxe[a] + fluxw[a];
}

__global__ void __launch_bounds__(LE_RED_NT)
le_force_sum_kernel(int nyz, const double * __restrict__ term, double * __restrict__ fcor) {
  const double * t = term + (size_t) blockIdx.x*nyz;       // blockIdx.x = p*3 + a
#ifdef LB200_STRICT
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nyz; i++) s += t[i];
    fcor[blockIdx.x] = s;
  }
#else
  __shared__ double sh[LE_RED_NT];
  double s = 0.0;
  for (int i = threadIdx.x; i < nyz; i += LE_RED_NT) s += t[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = LE_RED_NT/2; w > 0; w >>= 1) {
    if ((int) threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) fcor[blockIdx.x] = sh[0];
#endif
}

// raw (uncorrected) Cahn-Hilliard flux through an x face of site (ic, jc, kc): advective + diffusive +
// external, masked -- what the reference holds in flux->fw / flux->fe before phi_ch_le_fix_fluxes
template <int ORDER, bool WEST>
__device__ __forceinline__ double le_ch_xflux(const Lb200Geom & g, const Lb200LeDev & le, const Lb200SymmDev & sp,
					      const double * __restrict__ phi, const double * __restrict__ delsq,
					      const double * __restrict__ u, const char * __restrict__ status,
					      int ic, int jc, int kc) {
  const int s = le_index(g, ic, jc, kc);
  const int sm1 = le_index(g, le_x(le, g, ic, -1), jc, kc);
  const int sp1 = le_index(g, le_x(le, g, ic, +1), jc, kc);
  const double ph_c = phi[s], ph_xm = phi[sm1], ph_xp = phi[sp1];
  const double mu0 = symm_mu(sp, ph_c, delsq[s]);
  double fx;
  if (WEST) {
    const double ph_xm2 = (ORDER >= 3) ? phi[le_index(g, le_x(le, g, ic, -2), jc, kc)] : 0.0;
    fx = adv_face<ORDER, true>(u[sm1], u[s], ph_xm2, ph_xm, ph_c, ph_xp);
    fx -= sp.mobility*(mu0 - symm_mu(sp, ph_xm, delsq[sm1]));
    fx -= sp.mobility*sp.gm[0];
    if (status) fx *= (double) (status[s] == 0)*(double) (status[s - g.xs] == 0);
  }
  else {
    const double ph_xp2 = (ORDER >= 3) ? phi[le_index(g, le_x(le, g, ic, +2), jc, kc)] : 0.0;
    fx = adv_face<ORDER, false>(u[s], u[sp1], ph_xm, ph_c, ph_xp, ph_xp2);
    fx -= sp.mobility*(symm_mu(sp, ph_xp, delsq[sp1]) - mu0);
    fx -= sp.mobility*sp.gm[0];
    if (status) fx *= (double) (status[s] == 0)*(double) (status[s + g.xs] == 0);
  }
  return fx;
}

// raw fe(loc, j, k) -> chx[p][0][j][k], raw fw(loc + 1, j, k) -> chx[p][1][j][k]
template <int ORDER>
__global__ void __launch_bounds__(TPB_MAX)
le_ch_xflux_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200SymmDev sp,
		   const double * __restrict__ phi, const double * __restrict__ delsq,
		   const double * __restrict__ u, const char * __restrict__ status, double * __restrict__ chx) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int p = blockIdx.z/2, side = blockIdx.z % 2;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const size_t nyz = (size_t) g.nl[1]*g.nl[2];
  const size_t q = (size_t) (jc - 1)*g.nl[2] + (kc - 1);
  double fx;
  if (side == 0) fx = le_ch_xflux<ORDER, false>(g, le, sp, phi, delsq, u, status, le.loc[p], jc, kc);
  else           fx = le_ch_xflux<ORDER, true>(g, le, sp, phi, delsq, u, status, le.loc[p] + 1, jc, kc);
  chx[(size_t) (2*p + side)*nyz + q] = fx;
}

#ifndef LB200_STRICT
// Fast build, force AND Cahn-Hilliard wanted: the summands of the force correction, their per-block sums and the raw x-face
// fluxes either side of every plane in ONE launch, then one small launch that adds the block sums in a fixed order
// (le_force_term_kernel + le_force_sum_kernel + le_ch_xflux_kernel: 22 us of three dependent launches -> 2 launches).
constexpr int LE_PREP_NT = 128;
template <int ORDER>
__global__ void __launch_bounds__(LE_PREP_NT)
le_prep_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200SymmDev sp,
	       const double * __restrict__ phi, const double * __restrict__ grad, const double * __restrict__ delsq,
	       const double * __restrict__ u, const char * __restrict__ status, double * __restrict__ partial,
	       double * __restrict__ chx) {
  __shared__ double sh[3][LE_PREP_NT];
  const int nyz = g.nl[1]*g.nl[2];
  const int q = blockIdx.x*LE_PREP_NT + threadIdx.x;
  const int p = blockIdx.y;
  double t[3] = {0.0, 0.0, 0.0};
  if (q < nyz) {
    const int jc = 1 + q/g.nl[2], kc = 1 + q % g.nl[2];
    const size_t ns = (size_t) g.nsites;
    const int ic = le.loc[p];
    double p0[3], p1[3], fluxe[3], fluxw[3];
    le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, ic, jc, kc), p0);
    le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, le_x(le, g, ic, +1), jc, kc), p1);
    for (int a = 0; a < 3; a++) fluxe[a] = 0.5*(p1[a] + p0[a]);
    le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, ic + 1, jc, kc), p0);
    le_pcol_x(sp, phi, grad, delsq, ns, le_index(g, le_x(le, g, ic + 1, -1), jc, kc), p1);
    for (int a = 0; a < 3; a++) fluxw[a] = 0.5*(p1[a] + p0[a]);
    for (int a = 0; a < 3; a++) t[a] = - fluxe[a] + fluxw[a];
    chx[(size_t) (2*p + 0)*nyz + q] = le_ch_xflux<ORDER, false>(g, le, sp, phi, delsq, u, status, ic, jc, kc);
    chx[(size_t) (2*p + 1)*nyz + q] = le_ch_xflux<ORDER, true>(g, le, sp, phi, delsq, u, status, ic + 1, jc, kc);
  }
#pragma unroll
  for (int a = 0; a < 3; a++) sh[a][threadIdx.x] = t[a];
  __syncthreads();
  for (int w = LE_PREP_NT/2; w > 0; w >>= 1) {
    if ((int) threadIdx.x < w) {
#pragma unroll
      for (int a = 0; a < 3; a++) sh[a][threadIdx.x] += sh[a][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[(size_t) (p*3 + threadIdx.x)*gridDim.x + blockIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(LE_RED_NT)
le_prep_sum_kernel(int nblk, const double * __restrict__ partial, double * __restrict__ fcor) {
  __shared__ double sh[LE_RED_NT];
  const double * t = partial + (size_t) blockIdx.x*nblk;       // blockIdx.x = p*3 + a
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += LE_RED_NT) s += t[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = LE_RED_NT/2; w > 0; w >>= 1) {
    if ((int) threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) fcor[blockIdx.x] = sh[0];
}
#endif

// both preparations; term: at least 3*nplane*ceil(Ny*Nz/128) doubles (the 3*nplane*Ny*Nz summand array of the separate form is)
int launch_le_prep_both(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200SymmDev & sp,
			const double * phi, const double * grad, const double * delsq, const double * u, const char * status,
			double * term, double * fcor, double * chx);

// One thread per interior site of the x-planes in xlist (nullptr: every plane).  Force: flux form with the
// plane correction.  Cahn-Hilliard: as force_ch_kernel with the x-neighbours through the buffer planes and,
// next to a plane, the x-face flux averaged with the interpolated flux of the other side
// (phi_ch_le_fix_fluxes, src/phi_cahn_hilliard.c:648-735).
template <bool DO_FORCE, bool DO_CH, int ORDER>
__global__ void __launch_bounds__(TPB)
le_force_ch_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200SymmDev sp, const Lb200LeFix fix,
		   const int * __restrict__ xlist, const int accumulate,
		   const double * __restrict__ phi, const double * __restrict__ grad,
		   const double * __restrict__ delsq, const double * __restrict__ u,
		   const char * __restrict__ status, const double * __restrict__ fcor,
		   const double * __restrict__ chx, double * __restrict__ force, double * __restrict__ phinew) {

  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = xlist ? xlist[blockIdx.z] : 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;

  const size_t ns = (size_t) g.nsites;
  const int s = le_index(g, ic, jc, kc);
  const int sxm = le_index(g, le_x(le, g, ic, -1), jc, kc);
  const int sxp = le_index(g, le_x(le, g, ic, +1), jc, kc);
  const int sym = le_index(g, ic, le_wy(g, jc - 1), kc), syp = le_index(g, ic, le_wy(g, jc + 1), kc);
  const int szm = le_index(g, ic, jc, le_wz(g, kc - 1)), szp = le_index(g, ic, jc, le_wz(g, kc + 1));

  // which plane (if any) this site touches: below (ic == loc) or above (ic == loc + 1)
  int pl = -1, side = -1;
  for (int p = 0; p < le.nplane; p++) {
    if (ic == le.loc[p])     { pl = p; side = 0; }
    if (ic == le.loc[p] + 1) { pl = p; side = 1; }
  }

  if (DO_FORCE) {
    const SiteFE s0 = load_fe(phi, grad, delsq, ns, s);
    const SiteFE xm = load_fe(phi, grad, delsq, ns, sxm);
    const SiteFE xp = load_fe(phi, grad, delsq, ns, sxp);
    const SiteFE ym = load_fe(phi, grad, delsq, ns, sym);
    const SiteFE yp = load_fe(phi, grad, delsq, ns, syp);
    const SiteFE zm = load_fe(phi, grad, delsq, ns, szm);
    const SiteFE zp = load_fe(phi, grad, delsq, ns, szp);
    double p0[3], p1[3], fluxe[3], fluxw[3], fluxy[3], fluxym[3], fluxz[3], fluxzm[3];
    symm_pcol<0>(sp, s0, p0);
    symm_pcol<0>(sp, xm, p1);
    for (int a = 0; a < 3; a++) fluxw[a] = 0.5*(p1[a] + p0[a]);
    symm_pcol<0>(sp, xp, p1);
    for (int a = 0; a < 3; a++) fluxe[a] = 0.5*(p1[a] + p0[a]);
    symm_pcol<1>(sp, s0, p0);
    symm_pcol<1>(sp, yp, p1);
    for (int a = 0; a < 3; a++) fluxy[a] = 0.5*(p1[a] + p0[a]);
    symm_pcol<1>(sp, ym, p1);
    for (int a = 0; a < 3; a++) fluxym[a] = 0.5*(p0[a] + p1[a]);
    symm_pcol<2>(sp, s0, p0);
    symm_pcol<2>(sp, zp, p1);
    for (int a = 0; a < 3; a++) fluxz[a] = 0.5*(p1[a] + p0[a]);
    symm_pcol<2>(sp, zm, p1);
    for (int a = 0; a < 3; a++) fluxzm[a] = 0.5*(p0[a] + p1[a]);
    if (side == 0) { for (int a = 0; a < 3; a++) fluxe[a] += fix.ra*fcor[3*pl + a]; }
    if (side == 1) { for (int a = 0; a < 3; a++) fluxw[a] -= fix.ra*fcor[3*pl + a]; }
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double fo = -(+ fluxe[a] - fluxw[a] + fluxy[a] - fluxym[a] + fluxz[a] - fluxzm[a]);
      if (accumulate) force[a*ns + s] += fo;
      else            force[a*ns + s] = fo;
    }
  }

  if (DO_CH) {
    const double M = sp.mobility;
    const double ph_c = phi[s];
    const double ph_ym = phi[sym], ph_yp = phi[syp], ph_zm = phi[szm], ph_zp = phi[szp];
    const double d_c = delsq[s];
    const double mu0 = symm_mu(sp, ph_c, d_c);
    double ph_ym2 = 0.0, ph_yp2 = 0.0, ph_zm2 = 0.0, ph_zp2 = 0.0;
    if (ORDER >= 3) {
      ph_ym2 = phi[le_index(g, ic, le_wy(g, jc - 2), kc)]; ph_yp2 = phi[le_index(g, ic, le_wy(g, jc + 2), kc)];
      ph_zm2 = phi[le_index(g, ic, jc, le_wz(g, kc - 2))]; ph_zp2 = phi[le_index(g, ic, jc, le_wz(g, kc + 2))];
    }
    const double uy_c = u[1*ns + s], uy_ym = u[1*ns + sym], uy_yp = u[1*ns + syp];
    const double uz_c = u[2*ns + s], uz_zm = u[2*ns + szm], uz_zp = u[2*ns + szp];
    double mk = 1.0, mkyp = 1.0, mkym = 1.0, mkzp = 1.0, mkzm = 1.0;
    if (status) {
      mk = (status[s] == 0);
      mkym = (status[sym] == 0); mkyp = (status[syp] == 0);
      mkzm = (status[szm] == 0); mkzp = (status[szp] == 0);
    }

    double fw = le_ch_xflux<ORDER, true>(g, le, sp, phi, delsq, u, status, ic, jc, kc);
    double fe = le_ch_xflux<ORDER, false>(g, le, sp, phi, delsq, u, status, ic, jc, kc);
    if (side >= 0) {
      const int ny = g.nl[1];
      const size_t nyz = (size_t) ny*g.nl[2];
      // below the plane: fe <- 1/2 (fe + fw of the site above, displaced by +dy); above: fw <- 1/2 (fw + fe below, -dy)
      const int jdy = fix.jdy[side];
      const double fr = fix.fr[side];
      const int j1 = 1 + (jc - jdy - 2 + 2*ny) % ny;
      const int j2 = 1 + j1 % ny;
      const double * other = chx + (size_t) (2*pl + (1 - side))*nyz;
      const double b = other[(size_t) (j1 - 1)*g.nl[2] + (kc - 1)]*fr + other[(size_t) (j2 - 1)*g.nl[2] + (kc - 1)]*(1.0 - fr);
      if (side == 0) fe = 0.5*(fe + b);
      else           fw = 0.5*(fw + b);
    }

    double fy = adv_face<ORDER, false>(uy_c, uy_yp, ph_ym, ph_c, ph_yp, ph_yp2);
    fy -= M*(symm_mu(sp, ph_yp, delsq[syp]) - mu0);
    fy -= M*sp.gm[1];
    if (status) fy *= mk*mkyp;
    double fym = adv_face<ORDER, false>(uy_ym, uy_c, ph_ym2, ph_ym, ph_c, ph_yp);
    fym -= M*(mu0 - symm_mu(sp, ph_ym, delsq[sym]));
    fym -= M*sp.gm[1];
    if (status) fym *= mkym*mk;
    double fz = adv_face<ORDER, false>(uz_c, uz_zp, ph_zm, ph_c, ph_zp, ph_zp2);
    fz -= M*(symm_mu(sp, ph_zp, delsq[szp]) - mu0);
    fz -= M*sp.gm[2];
    if (status) fz *= mk*mkzp;
    double fzm = adv_face<ORDER, false>(uz_zm, uz_c, ph_zm2, ph_zm, ph_c, ph_zp);
    fzm -= M*(mu0 - symm_mu(sp, ph_zm, delsq[szm]));
    fzm -= M*sp.gm[2];
    if (status) fzm *= mkzm*mk;

    if (sp.csum != nullptr) {
      // cahn_hilliard_options_conserve 1: phi_ch_csum_kernel (src/phi_cahn_hilliard.c:1181-1215), as in force_ch_kernel
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

// phi_force_calculation with planes: fcor, then the divergence (do_ch = 0) -- or force + Cahn-Hilliard in one
// sweep.  nx planes from xlist (nullptr: all nl[0]).  term: 3*nplane*Ny*Nz scratch, chx: 2*nplane*Ny*Nz.
int launch_le_force_prep(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200SymmDev & sp,
			 const double * phi, const double * grad, const double * delsq, double * term, double * fcor) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, le.nplane);
  le_force_term_kernel<<<grd, blk, 0, st>>>(g, le, sp, phi, grad, delsq, term);
  le_force_sum_kernel<<<3*le.nplane, LE_RED_NT, 0, st>>>(g.nl[1]*g.nl[2], term, fcor);
  return 2;
}

int launch_le_ch_prep(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200SymmDev & sp,
		      const double * phi, const double * delsq, const double * u, const char * status, double * chx) {
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, 2*le.nplane);
  if (sp.order == 1)      le_ch_xflux_kernel<1><<<grd, blk, 0, st>>>(g, le, sp, phi, delsq, u, status, chx);
  else if (sp.order == 2) le_ch_xflux_kernel<2><<<grd, blk, 0, st>>>(g, le, sp, phi, delsq, u, status, chx);
  else if (sp.order == 4) le_ch_xflux_kernel<4><<<grd, blk, 0, st>>>(g, le, sp, phi, delsq, u, status, chx);
  else                    le_ch_xflux_kernel<3><<<grd, blk, 0, st>>>(g, le, sp, phi, delsq, u, status, chx);
  return 1;
}

int launch_le_prep_both(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200SymmDev & sp,
			const double * phi, const double * grad, const double * delsq, const double * u, const char * status,
			double * term, double * fcor, double * chx) {
#ifdef LB200_STRICT
  // the reference's sequential sum over (j, k): the separate kernels
  return launch_le_force_prep(st, g, le, sp, phi, grad, delsq, term, fcor) + launch_le_ch_prep(st, g, le, sp, phi, delsq, u, status, chx);
#else
  const int nyz = g.nl[1]*g.nl[2];
  const int nblk = (nyz + LE_PREP_NT - 1)/LE_PREP_NT;
  dim3 grd(nblk, le.nplane, 1);
  if (sp.order == 1)      le_prep_kernel<1><<<grd, LE_PREP_NT, 0, st>>>(g, le, sp, phi, grad, delsq, u, status, term, chx);
  else if (sp.order == 2) le_prep_kernel<2><<<grd, LE_PREP_NT, 0, st>>>(g, le, sp, phi, grad, delsq, u, status, term, chx);
  else if (sp.order == 4) le_prep_kernel<4><<<grd, LE_PREP_NT, 0, st>>>(g, le, sp, phi, grad, delsq, u, status, term, chx);
  else                    le_prep_kernel<3><<<grd, LE_PREP_NT, 0, st>>>(g, le, sp, phi, grad, delsq, u, status, term, chx);
  le_prep_sum_kernel<<<3*le.nplane, LE_RED_NT, 0, st>>>(nblk, term, fcor);
  return 2;
#endif
}

int launch_le_force_ch(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200SymmDev & sp,
		       const Lb200LeFix & fix, int nx, const int * xlist, int do_force, int do_ch, int accumulate,
		       const double * phi, const double * grad, const double * delsq, const double * u,
		       const char * status, const double * fcor, const double * chx, double * force, double * phinew) {
  dim3 blk;
  block_shape_n(g.nl[2], TPB, blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, nx);
#define LB200_GO(F, C, O) le_force_ch_kernel<F, C, O><<<grd, blk, 0, st>>>(g, le, sp, fix, xlist, accumulate, phi, grad, delsq, u, status, fcor, chx, force, phinew)
#define LB200_SEL_O(F, C) do { if (sp.order == 1) LB200_GO(F, C, 1); else if (sp.order == 2) LB200_GO(F, C, 2); else if (sp.order == 4) LB200_GO(F, C, 4); else LB200_GO(F, C, 3); } while (0)
  if (do_force && do_ch) LB200_SEL_O(true, true);
  else if (do_force)     LB200_GO(true, false, 1);
  else                   LB200_SEL_O(false, true);
#undef LB200_GO
#undef LB200_SEL_O
  return 1;
}

// ---------------------------------------------------------------------------------------------
// lb_data_apply_le_boundary_conditions (src/model_le.c:78-180): the populations about to cross a plane
// (c_x = +1 on x = loc, c_x = -1 on x = loc + 1) are re-projected with the velocity jump -/+ uy
// (le_reproject, :264-345), displaced by the integer part of the plane displacement (:358-400) and linearly
// interpolated with its fractional part (:584-640).  Two kernels on 2*nplane x-planes; sbuf holds the
// re-projected values [ix][n][ip][j][k] (the reference's send buffer; its receive buffer is the same data
// shifted by whole sites, which the second kernel does by indexing).
// ---------------------------------------------------------------------------------------------

// D3Q19: the same statements with the velocity set as a compile-time table and every loop unrolled (the generic form below
// reads cv from global memory inside data-dependent loops: 13 us for two planes of 256^2 sites; this one 2-3x less)
__global__ void __launch_bounds__(TPB_MAX)
le_lb_reproject19_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200ModelDev * __restrict__ md,
			 int ndist, const double * __restrict__ f, double * __restrict__ sbuf) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ix = blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int iplane = ix/2, iside = ix % 2;
  const int cx = 1 - 2*iside;
  const int ic = iside + le.loc[iplane];
  const int index = le_index(g, ic, jc, kc);
  const size_t ns = (size_t) g.nsites;
  const size_t nyz = (size_t) g.nl[1]*g.nl[2];
  const size_t q = (size_t) (jc - 1)*g.nl[2] + (kc - 1);
  constexpr int nvel = 19;
  const double cs2 = (1.0/3.0);
  const double rcs2 = 1.0/cs2;
  double du[3] = {0.0, 0.0, 0.0};
  du[1] = le.uy;
  du[1] = -1.0*cx*du[1];

  for (int n = 0; n < ndist; n++) {
    double fl[nvel];
#pragma unroll
    for (int p = 0; p < nvel; p++) fl[p] = f[(size_t) (n*nvel + p)*ns + index];
    double rho = 0.0;
    double gv[3] = {0.0, 0.0, 0.0};
    double ds[3][3];
#pragma unroll
    for (int p = 0; p < nvel; p++) {
      rho += fl[p];
#pragma unroll
      for (int ia = 0; ia < 3; ia++) gv[ia] += CV19[p][ia]*fl[p];
    }
#pragma unroll
    for (int ia = 0; ia < 3; ia++)
#pragma unroll
      for (int ib = 0; ib < 3; ib++)
	ds[ia][ib] = (gv[ia]*du[ib] + du[ia]*gv[ib] + rho*du[ia]*du[ib]);
    int ip = 0;
#pragma unroll
    for (int p = 1; p < nvel; p++) {
      if (CV19[p][0] != cx) continue;
      const double udotc = du[1]*CV19[p][1];
      double sdotq = 0.0;
#pragma unroll
      for (int ia = 0; ia < 3; ia++) {
#pragma unroll
	for (int ib = 0; ib < 3; ib++) {
	  const double dab = cs2*(ia == ib);
	  const double qab = (CV19[p][ia]*CV19[p][ib] - dab);
	  sdotq += ds[ia][ib]*qab;
	}
      }
      double fp = fl[p];
      fp += md->wv[p]*(rho*udotc*rcs2 + 0.5*sdotq*rcs2*rcs2);
      sbuf[((size_t) (ix*ndist + n)*le.nprop + ip)*nyz + q] = fp;
      ip++;
    }
  }
}

__global__ void __launch_bounds__(TPB_MAX)
le_lb_reproject_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200ModelDev * __restrict__ md,
		       int ndist, const double * __restrict__ f, double * __restrict__ sbuf) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ix = blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int iplane = ix/2, iside = ix % 2;
  const int cx = 1 - 2*iside;
  const int ic = iside + le.loc[iplane];
  const int index = le_index(g, ic, jc, kc);
  const size_t ns = (size_t) g.nsites;
  const size_t nyz = (size_t) g.nl[1]*g.nl[2];
  const size_t q = (size_t) (jc - 1)*g.nl[2] + (kc - 1);
  const int nvel = md->nvel;
  const double cs2 = (1.0/3.0);
  const double rcs2 = 1.0/cs2;
  double du[3] = {0.0, 0.0, 0.0};
  du[1] = le.uy;
  du[1] = -1.0*cx*du[1];

  for (int n = 0; n < ndist; n++) {
    double rho = 0.0;
    double gv[3] = {0.0, 0.0, 0.0};
    double ds[3][3];
    for (int p = 0; p < nvel; p++) {               // (each accumulator sees the reference's order of additions)
      const double fp = f[(size_t) (n*nvel + p)*ns + index];
      rho += fp;
      for (int ia = 0; ia < 3; ia++) gv[ia] += md->cv[p][ia]*fp;
    }
    for (int ia = 0; ia < 3; ia++)
      for (int ib = 0; ib < 3; ib++)
	ds[ia][ib] = (gv[ia]*du[ib] + du[ia]*gv[ib] + rho*du[ia]*du[ib]);
    int ip = 0;
    for (int p = 1; p < nvel; p++) {
      if (md->cv[p][0] != cx) continue;
      const double udotc = du[1]*md->cv[p][1];
      double sdotq = 0.0;
      for (int ia = 0; ia < 3; ia++) {
	for (int ib = 0; ib < 3; ib++) {
	  const double dab = cs2*(ia == ib);
	  const double qab = (md->cv[p][ia]*md->cv[p][ib] - dab);
	  sdotq += ds[ia][ib]*qab;
	}
      }
      double fp = f[(size_t) (n*nvel + p)*ns + index];
      fp += md->wv[p]*(rho*udotc*rcs2 + 0.5*sdotq*rcs2*rcs2);
      sbuf[((size_t) (ix*ndist + n)*le.nprop + ip)*nyz + q] = fp;
      ip++;
    }
  }
}

__global__ void __launch_bounds__(TPB_MAX)
le_lb_interp_kernel(const Lb200Geom g, const __grid_constant__ Lb200LeDev le, const Lb200LeFix fix,
		    const Lb200ModelDev * __restrict__ md, int ndist, const double * __restrict__ sbuf,
		    double * __restrict__ f) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ix = blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const int iplane = ix/2, iside = ix % 2;
  const int cx = 1 - 2*iside;
  const int ic = iside + le.loc[iplane];
  const int index = le_index(g, ic, jc, kc);
  const size_t ns = (size_t) g.nsites;
  const int ny = g.nl[1];
  const size_t nyz = (size_t) ny*g.nl[2];
  const int nvel = md->nvel;
  const int dj = fix.jdy[iside];
  const double fr = fix.fr[iside];
  // receive-buffer rows jc and jc + 1 are send-buffer rows js(jc), js(jc + 1) (displace kernel, :380-383)
  const int js0 = 1 + (jc + dj - 1 + 2*ny) % ny;
  const int js1 = 1 + (jc + 1 + dj - 1 + 2*ny) % ny;
  const size_t q0 = (size_t) (js0 - 1)*g.nl[2] + (kc - 1);
  const size_t q1 = (size_t) (js1 - 1)*g.nl[2] + (kc - 1);
  for (int n = 0; n < ndist; n++) {
    int ip = 0;
    for (int p = 1; p < nvel; p++) {
      if (md->cv[p][0] != cx) continue;
      const double * sb = sbuf + ((size_t) (ix*ndist + n)*le.nprop + ip)*nyz;
      f[(size_t) (n*nvel + p)*ns + index] = (1.0 - fr)*sb[q0] + fr*sb[q1];
      ip++;
    }
  }
}

int launch_le_lb_bc(cudaStream_t st, const Lb200Geom & g, const Lb200LeDev & le, const Lb200LeFix & fix,
		    const Lb200ModelDev * md, int ndist, double * f, double * sbuf) {
  if (le.nplane == 0) return 0;
  dim3 blk;
  block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, 2*le.nplane);
  if (le.nvel == 19) le_lb_reproject19_kernel<<<grd, blk, 0, st>>>(g, le, md, ndist, f, sbuf);
  else               le_lb_reproject_kernel<<<grd, blk, 0, st>>>(g, le, md, ndist, f, sbuf);
  le_lb_interp_kernel<<<grd, blk, 0, st>>>(g, le, fix, md, ndist, sbuf, f);
  return 2;
}

// ---------------------------------------------------------------------------------------------
// y / z periodic images, depth d, of ncomp components on a list of x-planes.  The one-kernel step (lb200_fused*.cuh) keeps
// the images of f', phi' and u up to date from the warps that produce them; the planes next to a Lees-Edwards plane
// are produced again by the patch kernels, which write interior sites only: this brings their images up to date.
// One thread per shell site (2d full rows of the allocation's cross-section, then 2d sites of each interior row).
// ---------------------------------------------------------------------------------------------

struct Lb200Images { double * data[3]; int ncomp[3]; int depth[3]; };

__global__ void __launch_bounds__(TPB)
le_yz_images_kernel(const Lb200Geom g, const int * __restrict__ xlist, const Lb200Images im) {
  // grid layer -> (array, component)
  int n = blockIdx.z, a = 0;
  if (n >= im.ncomp[0]) { n -= im.ncomp[0]; a = 1; if (n >= im.ncomp[1]) { n -= im.ncomp[1]; a = 2; } }
  const int d = im.depth[a];
  double * __restrict__ data = im.data[a];
  const int ny = g.nl[1], nz = g.nl[2];
  const int ez = nz + 2*d;
  const int nshell = 2*d*ez + 2*d*ny;
  int q = blockIdx.x*blockDim.x + threadIdx.x;
  if (q >= nshell) return;
  int jc, kc;
  if (q < 2*d*ez) {
    const int r = q/ez;
    jc = (r < d) ? 1 - d + r : ny + 1 + (r - d);
    kc = 1 - d + q % ez;
  }
  else {
    q -= 2*d*ez;
    const int t = q % (2*d);
    jc = 1 + q/(2*d);
    kc = (t < d) ? 1 - d + t : nz + 1 + (t - d);
  }
  const int ic = xlist[blockIdx.y];
  const int dst = le_index(g, ic, jc, kc);
  const int src = le_index(g, ic, lb200_wrap1(jc, ny, 1), lb200_wrap1(kc, nz, 1));
  const size_t ns = (size_t) g.nsites;
  data[n*ns + dst] = data[n*ns + src];
}

// up to three arrays (data == nullptr: none) in one launch
int launch_le_yz_images(cudaStream_t st, const Lb200Geom & g, int nx, const int * xlist,
			double * d0, int ncomp0, int depth0, double * d1, int ncomp1, int depth1, double * d2, int ncomp2, int depth2) {
  if (nx == 0) return 0;
  Lb200Images im;
  im.data[0] = d0; im.ncomp[0] = d0 ? ncomp0 : 0; im.depth[0] = depth0;
  im.data[1] = d1; im.ncomp[1] = d1 ? ncomp1 : 0; im.depth[1] = depth1;
  im.data[2] = d2; im.ncomp[2] = d2 ? ncomp2 : 0; im.depth[2] = depth2;
  int dmax = 0, nl = 0;
  for (int a = 0; a < 3; a++) if (im.ncomp[a] > 0) { dmax = max(dmax, im.depth[a]); nl += im.ncomp[a]; }
  if (nl == 0 || dmax <= 0) return 0;
  const int nshell = 2*dmax*(g.nl[2] + 2*dmax) + 2*dmax*g.nl[1];
  dim3 grd((nshell + TPB - 1)/TPB, nx, nl);
  le_yz_images_kernel<<<grd, TPB, 0, st>>>(g, xlist, im);
  return 1;
}
