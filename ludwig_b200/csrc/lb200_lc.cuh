// lb200_lc.cuh -- liquid crystal (Landau-de Gennes Q tensor, `free_energy lc_blue_phase`) coupled to the LB fluid
// (BASELINE config 4), included by lb200_kernels.cu inside its anonymous namespace (built twice:
// fast and strict).  Redshift 1, no activity, no noise, all-fluid lattices.
//
// Reference per step (src/ludwig.c:579-586, 716-719, 769-779): field_halo(q); grad_3d_7pt_fluid (15 + 5 arrays
// written); pth_stress_compute with fe_lc_stress_v (reads q, 15 gradients, 5 laplacians; writes 9); the stress
// divergence; hydro_u_halo; advection_x (20 flux arrays); beris_edw_h_driver (5 written); beris_edw_kernel_v
// (reads everything again): about 1.1 kB/site of array traffic besides the collision.
// Here: two sweeps.  `lc_stress_kernel` forms the gradients of the five Q components from the 7-point star in
// registers and writes only the stress (40 B read + 72 B written per site); `lc_force_be_kernel` gathers the
// stress columns for the force and, for the Beris-Edwards update, rebuilds gradients and molecular field from the
// 13-point star of Q it needs anyway for the advective fluxes (176 B read + 64 B written).  The 7-point gradient
// is also available on its own (`grad7_kernel`) for callers that want the arrays (free-energy statistics).
//
// Operation order = the reference's vectorised forms (fe_lc_compute_h_v, fe_lc_compute_fed_v,
// fe_lc_compute_stress_v, src/blue_phase.c:1908-2775; beris_edw_kernel_v, src/blue_phase_beris_edwards.c:538-850),
// including their use of kappa1 = kappa0 in the molecular field and the free-energy density.

#ifndef LC_BE_MINB
#define LC_BE_MINB 4
#endif
#ifndef LC_ST_MINB
#define LC_ST_MINB 4
#endif
__device__ constexpr int LC_D[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
__device__ constexpr int LC_E[3][3][3] = {{{0, 0, 0}, {0, 0, 1}, {0, -1, 0}},
					   {{0, 0, -1}, {0, 0, 0}, {1, 0, 0}},
					   {{0, 1, 0}, {-1, 0, 0}, {0, 0, 0}}};

__host__ inline void lc_block_shape(int nz, dim3 & blk);

// site index of (ic + dx, jc + dy, kc + dz), through the periodic boundary where the geometry says "wrap"
__device__ __forceinline__ int lc_nbr(const Lb200Geom & g, int ic, int jc, int kc) {
  return le_index(g, lb200_wrap1(ic, g.nl[0], g.wrap[0]), le_wy(g, jc), le_wz(g, kc));
}

// expand the five stored components (XX, XY, XZ, YY, YZ) of a traceless symmetric tensor
__device__ __forceinline__ void lc_expand5(const double c[5], double t[3][3]) {
  t[0][0] = c[0]; t[0][1] = c[1]; t[0][2] = c[2];
  t[1][0] = c[1]; t[1][1] = c[3]; t[1][2] = c[4];
  t[2][0] = c[2]; t[2][1] = c[4]; t[2][2] = 0.0 - c[0] - c[3];
}

// q, d_a q, laplacian(q) at a site from its 7-point star (grad_3d_7pt_fluid_kernel_v, src/gradient_3d_7pt_fluid.c:231-300)
// g2d: fd_gradient_calculation 2d_5pt_fluid (src/gradient_2d_5pt_fluid.c:108-172) -- a host loop over the plane kc = 1 in the
// reference: no z terms there, and gradient arrays that stay zero at kc != 1 (the stress on the z halo sites reads zeros)
__device__ __forceinline__ void lc_load_star(const Lb200Geom & g, const double * __restrict__ qf, int ic, int jc, int kc,
					     double q[3][3], double dq[3][3][3], double dsq[3][3], int g2d = 0) {
  const size_t ns = (size_t) g.nsites;
  const int s = lc_nbr(g, ic, jc, kc);
  const int sxm = lc_nbr(g, ic - 1, jc, kc), sxp = lc_nbr(g, ic + 1, jc, kc);
  const int sym = lc_nbr(g, ic, jc - 1, kc), syp = lc_nbr(g, ic, jc + 1, kc);
  const int szm = lc_nbr(g, ic, jc, kc - 1), szp = lc_nbr(g, ic, jc, kc + 1);
  double c[5], gx[5], gy[5], gz[5], d2[5];
#pragma unroll
  for (int n = 0; n < 5; n++) {
    const double * f = qf + n*ns;
    const double f0 = f[s], fxm = f[sxm], fxp = f[sxp], fym = f[sym], fyp = f[syp], fzm = f[szm], fzp = f[szp];
    c[n] = f0;
    gx[n] = 0.5*(fxp - fxm);
    gy[n] = 0.5*(fyp - fym);
    gz[n] = 0.5*(fzp - fzm);
    d2[n] = fxp + fxm + fyp + fym + fzp + fzm - 6.0*f0;
    if (g2d) {
      gz[n] = 0.0;
      d2[n] = fxp + fxm + fyp + fym - 4.0*f0;
      if (kc != 1) { gx[n] = 0.0; gy[n] = 0.0; d2[n] = 0.0; }
    }
  }
  lc_expand5(c, q);
  lc_expand5(gx, dq[0]);
  lc_expand5(gy, dq[1]);
  lc_expand5(gz, dq[2]);
  lc_expand5(d2, dsq);
}

// fe_lc_compute_h_v, src/blue_phase.c:2094-2270
__device__ __forceinline__ void lc_compute_h(const Lb200LcDev & p, const double q[3][3], const double dq[3][3][3],
					     const double dsq[3][3], double h[3][3]) {
  const double r3 = (1.0/3.0);
  const double q0 = p.rredshift*p.q0;
  const double kappa0 = p.redshift*p.redshift*p.kappa0;
  const double kappa1 = kappa0;
  const double gamma = p.gamma;
  double q2 = 0.0, edq = 0.0, e2 = 0.0;

#pragma unroll
  for (int ia = 0; ia < 3; ia++)
#pragma unroll
    for (int ib = 0; ib < 3; ib++) q2 += q[ia][ib]*q[ia][ib];

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      double sum = 0.0;
#pragma unroll
      for (int ic = 0; ic < 3; ic++) sum += q[ia][ic]*q[ib][ic];
      if (LC_D[ia][ib]) h[ia][ib] = - p.a0*(1.0 - r3*gamma)*q[ia][ib] + p.a0*gamma*(sum - r3*q2) - p.a0*gamma*q2*q[ia][ib];
      else              h[ia][ib] = - p.a0*(1.0 - r3*gamma)*q[ia][ib] + p.a0*gamma*(sum - 0.0) - p.a0*gamma*q2*q[ia][ib];
    }
  }

#pragma unroll
  for (int ib = 0; ib < 3; ib++)
#pragma unroll
    for (int ic = 0; ic < 3; ic++)
#pragma unroll
      for (int ia = 0; ia < 3; ia++) {
	if (LC_E[ib][ic][ia] > 0) edq += dq[ib][ic][ia];
	if (LC_E[ib][ic][ia] < 0) edq += -dq[ib][ic][ia];
      }

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      double sum = 0.0;
#pragma unroll
      for (int ic = 0; ic < 3; ic++) {
#pragma unroll
	for (int id = 0; id < 3; id++) {
	  const int ea = LC_E[ia][ic][id], eb = LC_E[ib][ic][id];
	  const double ta = (ea > 0) ? dq[ic][ib][id] : -dq[ic][ib][id];
	  const double tb = (eb > 0) ? dq[ic][ia][id] : -dq[ic][ia][id];
	  if (ea != 0 && eb != 0) sum += ta + tb;
	  else if (ea != 0) sum += ta;
	  else if (eb != 0) sum += tb;
	}
      }
      if (ia == ib) h[ia][ib] += kappa0*dsq[ia][ib] - 2.0*kappa1*q0*sum + 4.0*r3*kappa1*q0*edq - 4.0*kappa1*q0*q0*q[ia][ib];
      else          h[ia][ib] += kappa0*dsq[ia][ib] - 2.0*kappa1*q0*sum - 4.0*kappa1*q0*q0*q[ia][ib];
    }
  }

#pragma unroll
  for (int ia = 0; ia < 3; ia++) { const double ea = p.e0[ia]*1.0; e2 += ea*ea; }
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    const double ea = p.e0[ia]*1.0;
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      const double eb = p.e0[ib]*1.0;
      if (ia == ib) h[ia][ib] += p.epsilon*(ea*eb - r3*e2);
      else          h[ia][ib] += p.epsilon*(ea*eb - 0.0);
    }
  }
}

// fe_lc_compute_fed_v, src/blue_phase.c:1908-2075
__device__ __forceinline__ double lc_compute_fed(const Lb200LcDev & p, const double q[3][3], const double dq[3][3][3]) {
  const double r3 = 1.0/3.0;
  const double q0 = p.rredshift*p.q0;
  const double kappa0 = p.redshift*p.redshift*p.kappa0;
  const double kappa1 = kappa0;
  double q2 = 0.0, q3 = 0.0, dq0 = 0.0, dq1 = 0.0, efield = 0.0;

#pragma unroll
  for (int ia = 0; ia < 3; ia++)
#pragma unroll
    for (int ib = 0; ib < 3; ib++) q2 += q[ia][ib]*q[ia][ib];
#pragma unroll
  for (int ia = 0; ia < 3; ia++)
#pragma unroll
    for (int ib = 0; ib < 3; ib++)
#pragma unroll
      for (int ic = 0; ic < 3; ic++) q3 += q[ia][ib]*q[ib][ic]*q[ia][ic];

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    double sum = 0.0;
#pragma unroll
    for (int ib = 0; ib < 3; ib++) sum += dq[ib][ia][ib];
    dq0 += sum*sum;
  }
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      double sum = 0.0;
#pragma unroll
      for (int ic = 0; ic < 3; ic++) {
#pragma unroll
	for (int id = 0; id < 3; id++) {
	  if (LC_E[ia][ic][id] > 0) sum += dq[ic][ib][id];
	  if (LC_E[ia][ic][id] < 0) sum -= dq[ic][ib][id];
	}
      }
      sum += 2.0*q0*q[ia][ib];
      dq1 += sum*sum;
    }
  }
#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
    const double ea = p.e0[ia]*1.0;
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      const double eb = p.e0[ib]*1.0;
      efield += ea*q[ia][ib]*eb;
    }
  }
  return 0.5*p.a0*(1.0 - r3*p.gamma)*q2 - r3*p.a0*p.gamma*q3 + 0.25*p.a0*p.gamma*q2*q2
    + 0.5*kappa0*dq0 + 0.5*kappa1*dq1 - p.epsilon*efield;
}

// fe_lc_compute_stress_v, src/blue_phase.c:2279-2775 (the unrolled form of fe_lc_compute_stress, :827-925)
__device__ __forceinline__ void lc_compute_stress(const Lb200LcDev & p, const double q[3][3], const double dq[3][3][3],
						  const double h[3][3], double s[3][3]) {
  const double r3 = (1.0/3.0);
  const double q0 = p.q0*p.rredshift;
  const double kappa0 = p.kappa0*p.redshift*p.redshift;
  const double kappa1 = p.kappa1*p.redshift*p.redshift;
  const double xi = p.xi;
  double qh = 0.0;
  double p0 = lc_compute_fed(p, q, dq);
  p0 = 0.0 - p0;

#pragma unroll
  for (int ia = 0; ia < 3; ia++)
#pragma unroll
    for (int ib = 0; ib < 3; ib++) qh += q[ia][ib]*h[ia][ib];

#pragma unroll
  for (int ia = 0; ia < 3; ia++) {
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      double sth;
      if (ia == ib) sth = 2.0*xi*(q[ia][ib] + r3)*qh - p0;
      else          sth = 2.0*xi*(q[ia][ib])*qh;
#pragma unroll
      for (int ic = 0; ic < 3; ic++) {
	const double qb = (ib == ic) ? (q[ib][ic] + r3) : (q[ib][ic]);
	const double qa = (ia == ic) ? (q[ia][ic] + r3) : (q[ia][ic]);
	sth += -xi*h[ia][ic]*qb - xi*qa*h[ib][ic];
      }
#pragma unroll
      for (int ic = 0; ic < 3; ic++) {
#pragma unroll
	for (int id = 0; id < 3; id++) {
	  sth += - kappa0*dq[ia][ib][ic]*dq[id][ic][id] - kappa1*dq[ia][ic][id]*dq[ib][ic][id]
	    + kappa1*dq[ia][ic][id]*dq[ic][ib][id];
	  if (ib != ic) {
	    const int ie = 3 - ib - ic;
	    if (LC_E[ib][ic][ie] > 0) sth -= 2.0*kappa1*q0*dq[ia][ic][id]*q[id][ie];
	    else                      sth += 2.0*kappa1*q0*dq[ia][ic][id]*q[id][ie];
	  }
	}
      }
#pragma unroll
      for (int ic = 0; ic < 3; ic++) sth += q[ia][ic]*h[ib][ic] - h[ia][ic]*q[ib][ic];
      s[ia][ib] = -sth;
    }
  }
}


#ifndef LB200_STRICT
// ---------------------------------------------------------------------------------------------
// Fast-mode forms of the same quantities (re-associated; within 1e-12 of the reference order, tested): the
// reference's unrolled stress spends ~650 FP64 operations on the 81-term gradient contraction alone, which makes
// the kernel FP64-issue bound on this part (ncu: FP64 pipe 73 % active, DRAM 16 %).  Shared sub-expressions:
//   QQ = Q Q (symmetric), div_c = d_d Q_cd, C_ab = e_acd d_c Q_bd, M = Q H,
//   W_bcd = d_c Q_bd - d_b Q_cd - 2 q0 e_bce Q_de   (zero for b == c)
//   h_ab  = c1 Q_ab + a0 g (QQ_ab - q2/3 d_ab) + k0 lap Q_ab - 2 k0 q0 (C_ab + C_ba) + 4/3 k0 q0 (e:dQ) d_ab - 4 k0 q0^2 Q_ab + field
//   fed   = bulk(q2, QQ:Q) + k0/2 |div|^2 + k0/2 |C + 2 q0 Q|^2 - field
//   -s_ab = 2 xi (Q_ab + d_ab/3) Q:H - p0 d_ab - xi (M_ab + M_ba + 2/3 h_ab) + (M_ab - M_ba)
//           - k0 d_a Q_bc div_c + k1 d_a Q_cd W_bcd
// ---------------------------------------------------------------------------------------------
struct LcShared {
  double q2, qq[3][3], div[3], cc[3][3];
};

__device__ __forceinline__ void lc_h_fast(const Lb200LcDev & p, const double q[3][3], const double dq[3][3][3],
					  const double dsq[3][3], double h[3][3], LcShared & sh) {
  const double r3 = (1.0/3.0);
  const double q0 = p.q0*p.rredshift, k0 = p.kappa0*p.redshift*p.redshift;
  double q2 = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) q2 += q[a][b]*q[a][b];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = a; b < 3; b++) {
      const double v = q[a][0]*q[b][0] + q[a][1]*q[b][1] + q[a][2]*q[b][2];
      sh.qq[a][b] = v; sh.qq[b][a] = v;
    }
#pragma unroll
  for (int c = 0; c < 3; c++) sh.div[c] = dq[0][c][0] + dq[1][c][1] + dq[2][c][2];
#pragma unroll
  for (int b = 0; b < 3; b++) {
    sh.cc[0][b] = dq[1][b][2] - dq[2][b][1];
    sh.cc[1][b] = dq[2][b][0] - dq[0][b][2];
    sh.cc[2][b] = dq[0][b][1] - dq[1][b][0];
  }
  sh.q2 = q2;
  // e_bca d_b Q_ca
  const double edq = (dq[0][1][2] - dq[0][2][1]) + (dq[1][2][0] - dq[1][0][2]) + (dq[2][0][1] - dq[2][1][0]);
  const double ag = p.a0*p.gamma;
  const double c1 = -(p.a0*(1.0 - r3*p.gamma) + ag*q2) - 4.0*k0*q0*q0;
  const double dia = -ag*r3*q2 + 4.0*r3*k0*q0*edq;
  double e2 = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++) e2 += p.e0[a]*p.e0[a];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = a; b < 3; b++) {
      double v = c1*q[a][b] + ag*sh.qq[a][b] + k0*dsq[a][b] - 2.0*k0*q0*(sh.cc[a][b] + sh.cc[b][a]) + p.epsilon*p.e0[a]*p.e0[b];
      if (a == b) v += dia - p.epsilon*r3*e2;
      h[a][b] = v; h[b][a] = v;
    }
}

__device__ __forceinline__ void lc_stress_fast(const Lb200LcDev & p, const double q[3][3], const double dq[3][3][3],
						const double h[3][3], const LcShared & sh, double s[3][3]) {
  const double r3 = (1.0/3.0);
  const double q0 = p.q0*p.rredshift, k0 = p.kappa0*p.redshift*p.redshift, k1 = p.kappa1*p.redshift*p.redshift, xi = p.xi;
  // free-energy density (kappa1 = kappa0 there, as in the reference's vectorised form)
  double q3 = 0.0, dq0 = 0.0, dq1 = 0.0, efield = 0.0, qh = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    dq0 += sh.div[a]*sh.div[a];
#pragma unroll
    for (int b = 0; b < 3; b++) {
      q3 += sh.qq[a][b]*q[a][b];
      const double t = sh.cc[a][b] + 2.0*q0*q[a][b];
      dq1 += t*t;
      efield += p.e0[a]*q[a][b]*p.e0[b];
      qh += q[a][b]*h[a][b];
    }
  }
  const double fed = 0.5*p.a0*(1.0 - r3*p.gamma)*sh.q2 - r3*p.a0*p.gamma*q3 + 0.25*p.a0*p.gamma*sh.q2*sh.q2
    + 0.5*k0*dq0 + 0.5*k0*dq1 - p.epsilon*efield;
  const double p0 = -fed;

  double m[3][3];                                         // M = Q H
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) m[a][b] = q[a][0]*h[0][b] + q[a][1]*h[1][b] + q[a][2]*h[2][b];

#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double sth = 2.0*xi*(q[a][b] + (a == b ? r3 : 0.0))*qh - (a == b ? p0 : 0.0)
	- xi*(m[a][b] + m[b][a] + 2.0*r3*h[a][b]) + (m[a][b] - m[b][a]);
      double g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) {
	g0 += dq[a][b][c]*sh.div[c];
	if (c != b) {
	  const int e = 3 - b - c;
	  const double sg = (LC_E[b][c][e] > 0) ? 2.0*q0 : -2.0*q0;
#pragma unroll
	  for (int d = 0; d < 3; d++) g1 += dq[a][c][d]*(dq[c][b][d] - dq[b][c][d] - sg*q[d][e]);
	}
      }
      sth += -k0*g0 + k1*g1;
      s[a][b] = -sth;
    }
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// 7-point gradient of nf components as arrays: grad[(n*3 + a)*ns + i], delsq[n*ns + i] on [1-ne, N+ne]^3
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB_MAX)
grad7_kernel(const Lb200Geom g, int ne, int nf, int g2d, const double * __restrict__ field, double * __restrict__ grad,
	     double * __restrict__ delsq) {
  const int kc = 1 - ne + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 - ne + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 - ne + blockIdx.z;
  if (kc > g.nl[2] + ne || jc > g.nl[1] + ne) return;
  if (g2d && kc != 1) return;                      // 2d_5pt_fluid: the plane kc = 1 only
  const size_t ns = (size_t) g.nsites;
  const int s = le_index(g, ic, jc, kc);
  const int sxm = lc_nbr(g, ic - 1, jc, kc), sxp = lc_nbr(g, ic + 1, jc, kc);
  const int sym = lc_nbr(g, ic, jc - 1, kc), syp = lc_nbr(g, ic, jc + 1, kc);
  const int szm = lc_nbr(g, ic, jc, kc - 1), szp = lc_nbr(g, ic, jc, kc + 1);
  for (int n = 0; n < nf; n++) {
    const double * f = field + n*ns;
    const double f0 = f[s], fxm = f[sxm], fxp = f[sxp], fym = f[sym], fyp = f[syp], fzm = f[szm], fzp = f[szp];
    grad[(size_t) (n*3 + 0)*ns + s] = 0.5*(fxp - fxm);
    grad[(size_t) (n*3 + 1)*ns + s] = 0.5*(fyp - fym);
    grad[(size_t) (n*3 + 2)*ns + s] = g2d ? 0.0 : 0.5*(fzp - fzm);
    delsq[n*ns + s] = g2d ? fxp + fxm + fyp + fym - 4.0*f0 : fxp + fxm + fyp + fym + fzp + fzm - 6.0*f0;
  }
}

// nf < 0: -nf components with the 2d_5pt_fluid stencil (the plane kc = 1 only)
int launch_grad7(cudaStream_t st, const Lb200Geom & g, int ne, int nf_, const double * field, double * grad, double * delsq) {
  const int g2d = (nf_ < 0), nf = g2d ? -nf_ : nf_;
  dim3 blk;
  block_shape(g.nl[2] + 2*ne, blk);
  dim3 grd((g.nl[2] + 2*ne + blk.x - 1)/blk.x, (g.nl[1] + 2*ne + blk.y - 1)/blk.y, g.nl[0] + 2*ne);
  grad7_kernel<<<grd, blk, 0, st>>>(g, ne, nf, g2d, field, grad, delsq);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// pth_stress_compute with fe_lc_stress_v (src/phi_force_stress.c:171-284, src/blue_phase.c:1737-1800): the
// stress on [1-ne, N+ne]^3 (ne = 1 when the divergence reads halo sites, 0 in halo-free steps)
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB, LC_ST_MINB)
lc_stress_kernel(const Lb200Geom g, const __grid_constant__ Lb200LcDev p, int nex, int ne, const double * __restrict__ qf,
		 double * __restrict__ str) {
  const int kc = 1 - ne + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 - ne + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 - nex + blockIdx.z;
  if (kc > g.nl[2] + ne || jc > g.nl[1] + ne) return;
  const size_t ns = (size_t) g.nsites;
  double q[3][3], dq[3][3][3], dsq[3][3], h[3][3], s[3][3];
  lc_load_star(g, qf, ic, jc, kc, q, dq, dsq, p.g2d);
#ifdef LB200_STRICT
  lc_compute_h(p, q, dq, dsq, h);
  lc_compute_stress(p, q, dq, h, s);
#else
  LcShared sh;
  lc_h_fast(p, q, dq, dsq, h, sh);
  lc_stress_fast(p, q, dq, h, sh, s);
#endif
  if (p.is_active) {
    // fe_lc_compute_stress_active (src/blue_phase.c:934-972) with zeta2 = 0, added as fe_lc_stress_v does (:1825-1845)
#pragma unroll
    for (int ia = 0; ia < 3; ia++)
#pragma unroll
      for (int ib = 0; ib < 3; ib++) {
	double sa = p.zeta0*LC_D[ia][ib] - p.zeta1*q[ia][ib];
	sa = -sa;
	s[ia][ib] += sa;
      }
  }
  const int idx = le_index(g, ic, jc, kc);
#pragma unroll
  for (int ia = 0; ia < 3; ia++)
#pragma unroll
    for (int ib = 0; ib < 3; ib++) str[(size_t) (ia*3 + ib)*ns + idx] = s[ia][ib];
}

// nex / ne: extension of the swept region beyond the interior in x / in y and z (x-slabs with in-kernel y, z images
// need the stress on the x planes 0 and N + 1 only)
int launch_lc_stress(cudaStream_t st, const Lb200Geom & g, const Lb200LcDev & p, int nex, int ne, const double * q, double * str) {
  dim3 blk;
  lc_block_shape(g.nl[2] + 2*ne, blk);
  dim3 grd((g.nl[2] + 2*ne + blk.x - 1)/blk.x, (g.nl[1] + 2*ne + blk.y - 1)/blk.y, g.nl[0] + 2*nex);
  lc_stress_kernel<<<grd, blk, 0, st>>>(g, p, nex, ne, q, str);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Force = - div(stored stress) (pth_force_fluid_kernel_v, src/phi_force_colloid.c:315-465) and / or the
// Beris-Edwards update (beris_edw_update: advection_x + beris_edw_h_driver + beris_edw_kernel_v,
// src/blue_phase_beris_edwards.c:266-296, 538-850) of one interior site, in one sweep.
// ---------------------------------------------------------------------------------------------

template <bool DO_FORCE, bool DO_BE, int ORDER>
__global__ void __launch_bounds__(TPB, LC_BE_MINB)
lc_force_be_kernel(const Lb200Geom g, const __grid_constant__ Lb200LcDev p, const int accumulate,
		   const double * __restrict__ qf, const double * __restrict__ str, const double * __restrict__ u,
		   double * __restrict__ force, double * __restrict__ qnew) {
  const int kc = 1 + blockIdx.x*blockDim.x + threadIdx.x;
  const int jc = 1 + blockIdx.y*blockDim.y + threadIdx.y;
  const int ic = 1 + blockIdx.z;
  if (kc > g.nl[2] || jc > g.nl[1]) return;
  const size_t ns = (size_t) g.nsites;
  const int s = le_index(g, ic, jc, kc);
  const int sxm = lc_nbr(g, ic - 1, jc, kc), sxp = lc_nbr(g, ic + 1, jc, kc);
  const int sym = lc_nbr(g, ic, jc - 1, kc), syp = lc_nbr(g, ic, jc + 1, kc);
  const int szm = lc_nbr(g, ic, jc, kc - 1), szp = lc_nbr(g, ic, jc, kc + 1);

  if (DO_FORCE) {
    // F_a = - sum_b 1/2 [(P_ab(+b) + P_ab) - (P_ab(-b) + P_ab)], accumulated +x, -x, +y, -y, +z, -z
    const int nb[6] = {sxp, sxm, syp, sym, szp, szm};
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
      double fo = 0.0;
#pragma unroll
      for (int d = 0; d < 6; d++) {
	const int ib = d/2;
	const double p1 = str[(size_t) (ia*3 + ib)*ns + nb[d]];
	const double p0 = str[(size_t) (ia*3 + ib)*ns + s];
	if (d == 0)          fo  = -0.5*(p1 + p0);
	else if (d % 2 == 1) fo += 0.5*(p1 + p0);
	else                 fo -= 0.5*(p1 + p0);
      }
      if (accumulate) force[ia*ns + s] += fo;
      else            force[ia*ns + s] = fo;
    }
  }

#ifdef LB200_STRICT
  if (DO_BE) {
    const double r3 = (1.0/3.0);
    const double dt = 1.0;
    double q[3][3], dq[3][3][3], dsq[3][3], h[3][3];
    lc_load_star(g, qf, ic, jc, kc, q, dq, dsq, p.g2d);
#ifdef LB200_STRICT
    lc_compute_h(p, q, dq, dsq, h);
#else
    { LcShared sh; lc_h_fast(p, q, dq, dsq, h, sh); }
#endif

    // velocity gradient tensor w[a][b] = d_b u_a
    double w[3][3], d[3][3], omega[3][3], sm[3][3];
    const int nbp[3] = {sxp, syp, szp}, nbm[3] = {sxm, sym, szm};
#pragma unroll
    for (int ib = 0; ib < 3; ib++)
#pragma unroll
      for (int ia = 0; ia < 3; ia++) w[ia][ib] = 0.5*(u[ia*ns + nbp[ib]] - u[ia*ns + nbm[ib]]);
    const double tr = r3*(w[0][0] + w[1][1] + w[2][2]);
    w[0][0] -= tr; w[1][1] -= tr; w[2][2] -= tr;
    double trace_qw = 0.0;
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
#pragma unroll
      for (int ib = 0; ib < 3; ib++) {
	trace_qw += q[ia][ib]*w[ib][ia];
	d[ia][ib] = 0.5*(w[ia][ib] + w[ib][ia]);
	omega[ia][ib] = 0.5*(w[ia][ib] - w[ib][ia]);
      }
    }
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
#pragma unroll
      for (int ib = 0; ib < 3; ib++) {
	if (ia > ib) continue;                   // only XX, XY, XZ, YY, YZ (and ZZ, unused) are needed
	const double qd = LC_D[ia][ib] ? (q[ia][ib] + r3) : (q[ia][ib] + 0.0);
	double sab = -2.0*p.xi*qd*trace_qw;
#pragma unroll
	for (int id = 0; id < 3; id++) {
	  const double q1 = LC_D[id][ib] ? (q[id][ib] + r3) : (q[id][ib] + 0.0);
	  const double q2 = LC_D[ia][id] ? (q[ia][id] + r3) : (q[ia][id] + 0.0);
	  sab += (p.xi*d[ia][id] + omega[ia][id])*q1 + q2*(p.xi*d[id][ib] - omega[id][ib]);
	}
	sm[ia][ib] = sab;
      }
    }

    // advective fluxes of the five components through the six faces, in registers
    const double ux_c = u[0*ns + s], ux_m = u[0*ns + sxm], ux_p = u[0*ns + sxp];
    const double uy_c = u[1*ns + s], uy_m = u[1*ns + sym], uy_p = u[1*ns + syp];
    const double uz_c = u[2*ns + s], uz_m = u[2*ns + szm], uz_p = u[2*ns + szp];
    int sxm2 = s, sxp2 = s, sym2 = s, syp2 = s, szm2 = s, szp2 = s;
    if (ORDER >= 3) {
      sxm2 = lc_nbr(g, ic - 2, jc, kc); sxp2 = lc_nbr(g, ic + 2, jc, kc);
      sym2 = lc_nbr(g, ic, jc - 2, kc); syp2 = lc_nbr(g, ic, jc + 2, kc);
      szm2 = lc_nbr(g, ic, jc, kc - 2); szp2 = lc_nbr(g, ic, jc, kc + 2);
    }
    const int ca[5] = {0, 0, 0, 1, 1}, cb[5] = {0, 1, 2, 1, 2};
#pragma unroll
    for (int n = 0; n < 5; n++) {
      const double * f = qf + n*ns;
      const double f0 = f[s];
      const double fxm = f[sxm], fxp = f[sxp], fym = f[sym], fyp = f[syp], fzm = f[szm], fzp = f[szp];
      double fxm2 = 0.0, fxp2 = 0.0, fym2 = 0.0, fyp2 = 0.0, fzm2 = 0.0, fzp2 = 0.0;
      if (ORDER >= 3) { fxm2 = f[sxm2]; fxp2 = f[sxp2]; fym2 = f[sym2]; fyp2 = f[syp2]; fzm2 = f[szm2]; fzp2 = f[szp2]; }
      const double fw  = adv_face<ORDER, true>(ux_m, ux_c, fxm2, fxm, f0, fxp);
      const double fe  = adv_face<ORDER, false>(ux_c, ux_p, fxm, f0, fxp, fxp2);
      const double fy  = adv_face<ORDER, false>(uy_c, uy_p, fym, f0, fyp, fyp2);
      const double fyl = adv_face<ORDER, false>(uy_m, uy_c, fym2, fym, f0, fyp);
      const double fz  = adv_face<ORDER, false>(uz_c, uz_p, fzm, f0, fzp, fzp2);
      const double fzl = adv_face<ORDER, false>(uz_m, uz_c, fzm2, fzm, f0, fzp);
      double qn = q[ca[n]][cb[n]];
      qn += dt*(sm[ca[n]][cb[n]] + 0.0 + p.Gamma*h[ca[n]][cb[n]] - fe + fw - fy + fyl - fz + fzl);
      qnew[n*ns + s] = qn;
    }
  }
#else
  if (DO_BE) {
    // Fast mode: one pass over the 13-point star of each component gives Q, its gradients and Laplacian AND the
    // divergence of its six advective face fluxes (the strict form above loads the 7-point star twice and keeps the
    // reference's left-to-right sum of the six fluxes); loads of a component are independent of the arithmetic of
    // the previous one, so many are in flight.
    const double r3 = (1.0/3.0);
    const int nbp[3] = {sxp, syp, szp}, nbm[3] = {sxm, sym, szm};
    double w[3][3];
#pragma unroll
    for (int ib = 0; ib < 3; ib++)
#pragma unroll
      for (int ia = 0; ia < 3; ia++) w[ia][ib] = 0.5*(u[ia*ns + nbp[ib]] - u[ia*ns + nbm[ib]]);
    const double ux_c = u[0*ns + s], ux_m = u[0*ns + sxm], ux_p = u[0*ns + sxp];
    const double uy_c = u[1*ns + s], uy_m = u[1*ns + sym], uy_p = u[1*ns + syp];
    const double uz_c = u[2*ns + s], uz_m = u[2*ns + szm], uz_p = u[2*ns + szp];
    int sxm2 = s, sxp2 = s, sym2 = s, syp2 = s, szm2 = s, szp2 = s;
    if (ORDER >= 3) {
      sxm2 = lc_nbr(g, ic - 2, jc, kc); sxp2 = lc_nbr(g, ic + 2, jc, kc);
      sym2 = lc_nbr(g, ic, jc - 2, kc); syp2 = lc_nbr(g, ic, jc + 2, kc);
      szm2 = lc_nbr(g, ic, jc, kc - 2); szp2 = lc_nbr(g, ic, jc, kc + 2);
    }
    double c[5], gx[5], gy[5], gz[5], d2[5], adv[5];
#pragma unroll
    for (int n = 0; n < 5; n++) {
      const double * f = qf + n*ns;
      const double f0 = f[s];
      const double fxm = f[sxm], fxp = f[sxp], fym = f[sym], fyp = f[syp], fzm = f[szm], fzp = f[szp];
      double fxm2 = 0.0, fxp2 = 0.0, fym2 = 0.0, fyp2 = 0.0, fzm2 = 0.0, fzp2 = 0.0;
      if (ORDER >= 3) { fxm2 = f[sxm2]; fxp2 = f[sxp2]; fym2 = f[sym2]; fyp2 = f[syp2]; fzm2 = f[szm2]; fzp2 = f[szp2]; }
      c[n] = f0;
      gx[n] = 0.5*(fxp - fxm);
      gy[n] = 0.5*(fyp - fym);
      gz[n] = p.g2d ? 0.0 : 0.5*(fzp - fzm);
      d2[n] = p.g2d ? fxp + fxm + fyp + fym - 4.0*f0 : fxp + fxm + fyp + fym + fzp + fzm - 6.0*f0;
      const double fw  = adv_face<ORDER, true>(ux_m, ux_c, fxm2, fxm, f0, fxp);
      const double fe  = adv_face<ORDER, false>(ux_c, ux_p, fxm, f0, fxp, fxp2);
      const double fy  = adv_face<ORDER, false>(uy_c, uy_p, fym, f0, fyp, fyp2);
      const double fyl = adv_face<ORDER, false>(uy_m, uy_c, fym2, fym, f0, fyp);
      const double fz  = adv_face<ORDER, false>(uz_c, uz_p, fzm, f0, fzp, fzp2);
      const double fzl = adv_face<ORDER, false>(uz_m, uz_c, fzm2, fzm, f0, fzp);
      adv[n] = (fw - fe) + (fyl - fy) + (fzl - fz);
    }
    double q[3][3], dq[3][3][3], dsq[3][3], h[3][3];
    lc_expand5(c, q);
    lc_expand5(gx, dq[0]);
    lc_expand5(gy, dq[1]);
    lc_expand5(gz, dq[2]);
    lc_expand5(d2, dsq);
    { LcShared sh; lc_h_fast(p, q, dq, dsq, h, sh); }

    const double tr = r3*(w[0][0] + w[1][1] + w[2][2]);
    w[0][0] -= tr; w[1][1] -= tr; w[2][2] -= tr;
    double d[3][3], omega[3][3], qs[3][3];
    double trace_qw = 0.0;
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
#pragma unroll
      for (int ib = 0; ib < 3; ib++) {
	trace_qw += q[ia][ib]*w[ib][ia];
	d[ia][ib] = 0.5*(w[ia][ib] + w[ib][ia]);
	omega[ia][ib] = 0.5*(w[ia][ib] - w[ib][ia]);
	qs[ia][ib] = q[ia][ib] + (ia == ib ? r3 : 0.0);
      }
    }
    const int ca[5] = {0, 0, 0, 1, 1}, cb[5] = {0, 1, 2, 1, 2};
#pragma unroll
    for (int n = 0; n < 5; n++) {
      const int ia = ca[n], ib = cb[n];
      double sab = -2.0*p.xi*qs[ia][ib]*trace_qw;
#pragma unroll
      for (int id = 0; id < 3; id++) {
	sab += (p.xi*d[ia][id] + omega[ia][id])*qs[id][ib] + qs[ia][id]*(p.xi*d[id][ib] - omega[id][ib]);
      }
      qnew[n*ns + s] = c[n] + (sab + p.Gamma*h[ia][ib] + adv[n]);
    }
  }
#endif
}

// block shape of the two LC sweeps: bx threads along z, TPB/bx rows of y (y neighbours of a row then hit in L1);
// LB200_LC_BX (32 / 64 / 128) overrides for tuning runs
__host__ inline void lc_block_shape(int nz, dim3 & blk) {
  static const int bx_env = tuned_flag("LB200_LC_BX", 32);
  int bx = ((nz + 31)/32)*32;
  if (bx > bx_env) bx = bx_env;
  if (bx > TPB) bx = TPB;
  blk = dim3(bx, TPB/bx, 1);
}

int launch_lc_force_be(cudaStream_t st, const Lb200Geom & g, const Lb200LcDev & p, int do_force, int do_be, int accumulate,
		       const double * q, const double * str, const double * u, double * force, double * qnew) {
  dim3 blk;
  lc_block_shape(g.nl[2], blk);
  dim3 grd((g.nl[2] + blk.x - 1)/blk.x, (g.nl[1] + blk.y - 1)/blk.y, g.nl[0]);
#define LB200_GO(F, B, O) lc_force_be_kernel<F, B, O><<<grd, blk, 0, st>>>(g, p, accumulate, q, str, u, force, qnew)
#define LB200_SEL_O(F, B) do { if (p.order == 1) LB200_GO(F, B, 1); else if (p.order == 2) LB200_GO(F, B, 2); else if (p.order == 4) LB200_GO(F, B, 4); else LB200_GO(F, B, 3); } while (0)
  if (do_force && do_be) LB200_SEL_O(true, true);
  else if (do_force)     LB200_GO(true, false, 1);
  else                   LB200_SEL_O(false, true);
#undef LB200_GO
#undef LB200_SEL_O
  return 1;
}
