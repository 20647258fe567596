// lb200_fused_ws.cuh -- the one-kernel binary-fluid step with WARP-SPECIALISED roles (fast arithmetic mode).
//
// Same sweep, data flow, TMA staging and producer-side halos as step_fused_kernel (lb200_fused.cuh), but the two halves of a
// plane-step are done by different warps of the CTA, concurrently:
//   producer warps (one per tile row incl. the two apron rows)  phi sector of plane n: gradient / stress of n+1, force and face
//                                                               fluxes of n, phi update of n-1; the force of the tile's sites
//                                                               goes to a double-buffered shared-memory block
//   consumer warps (one per interior tile row)                  pull-stream + collision of plane m <= n with that force
// linked by mbarriers (force full / empty), each group with its own named barrier -- no CTA-wide __syncthreads in the march.
// Why: in step_fused_kernel every warp does both halves one after the other; with 164 registers an SM holds 10 warps, 2.5 per
// scheduler, and the schedulers idle half the time (profiles/r02_ncu_step_fused.md).  The phi sector is shared-memory-pipe work,
// the collision FP64-pipe work: as separate warps neither carries the other's registers (18 warps x 112 registers), 4.5 warps
// per scheduler are busy, and the two kinds of work overlap all the time instead of half the time.
// (setmaxnreg was tried first -- 12 producer warps at 120-128 registers, 8 consumer warps at 64: it deadlocks, because the
// increase can only be served from what the decrease of the SAME CTA released, 8 x 32 x 32 registers < 12 x 32 x 24.)

constexpr int FW_BY = 10;                            // tile rows incl. apron (producer warps 0 .. 9)
constexpr int FW_NPROD = 10;                         // producer warps = tile rows
constexpr int FW_NCONS = 8;                          // consumer warps = interior rows
constexpr int FW_NT = 32*(FW_NPROD + FW_NCONS + 1);  // + one loader warp (TMA of the populations): 608 threads
constexpr int FW_NF = 3;                             // force blocks in flight between the two groups

struct FwShared {
  FuShared<FW_BY> s;
  double F[FW_NF][3][FW_NCONS*32];                   // force of the tile's sites, ring over planes
  unsigned long long ffull[FW_NF], fempty[FW_NF];    // mbarriers of the force hand-over
  unsigned long long sempty[FU_NSTAGE];              // mbarriers: every collision warp has taken its populations out of a stage
};

__device__ __forceinline__ void fw_mbar_arrive(unsigned int bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
template <int ID, int N> __device__ __forceinline__ void fw_bar_sync() { asm volatile("bar.sync %0, %1;\n" :: "n"(ID), "n"(N) : "memory"); }

#define FW_AD(member) (k.smb + (unsigned int) offsetof(FwShared, member))

template <int ORDER, bool GHOST>
__global__ void __launch_bounds__(FW_NT, 1)
step_fused_ws_kernel(const __grid_constant__ CUtensorMap fmap, const __grid_constant__ CUtensorMap phimap,
		     const __grid_constant__ CUtensorMap umap, const Lb200Geom g, const Lb200SymmDev sp,
		     const Lb200CollideDev cp, int xc, int le_x0, int le_xb,
		     const double * __restrict__ phi, const double * __restrict__ u,
		     const double * __restrict__ fsrc, double * __restrict__ fdst,
		     double * __restrict__ grad, double * __restrict__ delsq,
		     double * __restrict__ force, double * __restrict__ phinew,
		     double * __restrict__ rho_out, double * __restrict__ u_out) {
  constexpr int BY = FW_BY;
  using G = FuGeo<BY>;
  extern __shared__ __align__(1024) unsigned char fu_smem_raw[];
  FwShared & sw = *reinterpret_cast<FwShared *>(fu_smem_raw);
  FuShared<BY> & sm = sw.s;

  const int warp = threadIdx.x >> 5, tz = threadIdx.x & 31;
  const bool producer = (warp < FW_NPROD), loader = (warp == FW_NPROD + FW_NCONS);
  // tile row of this warp: producers 0 .. BY-1, consumers the interior rows 1 .. TY (the loader warp: no row)
  const int ty = producer ? warp : (loader ? 1 : warp - FW_NPROD + 1);
  const int fmode = 0, skew = 0;
  const int kbase = blockIdx.x*G::TZ;              // thread column (j,k) = (jbase + ty, kbase + tz)
  const int jbase = blockIdx.y*G::TY;
  const int kc = kbase + tz, jc = jbase + ty;
  const int nh = g.nh, ys = g.ys;

  FuK k;
  k.tid = ty*G::BZ + tz;
  k.pc = (ty + 1)*G::PZ + (tz + 1);                // own position in the phi tile
  k.xs = g.xs; k.ns = (size_t) g.nsites; k.nh = nh; k.nlx = g.nl[0]; k.wx = g.wrap[0];
  k.i0 = 1 + g.xoff + blockIdx.z*xc;
  k.i1 = min(k.i0 + xc - 1, g.xcnt > 0 ? g.xoff + g.xcnt : g.nl[0]);
  k.M = sp.mobility; k.kappa = sp.kappa; k.a = sp.a; k.b = sp.b; k.wz = sp.wz;
  k.peer_lo = g.peer_phi_lo; k.peer_hi = g.peer_phi_hi;
  k.peer_f_lo = g.peer_f_lo; k.peer_f_hi = g.peer_f_hi; k.peer_u_lo = g.peer_u_lo; k.peer_u_hi = g.peer_u_hi;
  k.mg0 = sp.mobility*sp.gm[0]; k.mg1 = sp.mobility*sp.gm[1]; k.mg2 = sp.mobility*sp.gm[2];
  k.skip_diag = (g.skip_diag != 0);
  k.fmode = fmode;
  k.odd = (skew != 0) && ((ty & 1) != 0);
  k.has_sites = (kbase + 1 <= g.nl[2]) && (jbase + 1 <= g.nl[1]);      // the tile owns lattice sites (CTA-uniform)

  const bool inner = (ty >= 1 && ty <= G::TY && tz >= 1 && tz <= G::TZ);
  k.valid_g = (jc <= g.nl[1] + 1) && (kc <= g.nl[2] + 1);
  k.out_site = inner && jc <= g.nl[1] && kc <= g.nl[2];
  k.face_row = (ty <= G::TY);                      // rows that own faces towards j+1 / k+1
  k.own_g = k.valid_g && ((ty >= 1 && ty <= G::TY) || (ty == 0 && jc == 0))
    && ((tz >= 1 && tz <= G::TZ) || (tz == 0 && kc == 0)) && !g.skip_diag;

  // column used for loads: clamped inside the allocation, through the periodic boundary if wrapping
  const int jl = ps_wrap(min(jc, g.nl[1] + 1), g.nl[1], g.wrap[1]);
  const int kl = ps_wrap(min(kc, g.nl[2] + 1), g.nl[2], g.wrap[2]);
  k.col = (jl + nh - 1)*ys + (kl + nh - 1);
  k.scol = (jc + nh - 1)*ys + (kc + nh - 1);       // column of the stores (never wrapped)

  // pull offsets of the own site in y / z (direct loads, tuning mode 1)
  k.oym = (g.wrap[1] && jc == 1)       ?  (g.nl[1] - 1)*ys : -ys;
  k.oyp = (g.wrap[1] && jc == g.nl[1]) ? -(g.nl[1] - 1)*ys :  ys;
  k.ozm = (g.wrap[2] && kc == 1)       ?  (g.nl[2] - 1) : -1;
  k.ozp = (g.wrap[2] && kc == g.nl[2]) ? -(g.nl[2] - 1) :  1;
  k.frow = min(max(ty - 1, 0), G::TY - 1);
  k.flane = tz;                                    // staged row element e <-> array z index kbase + e; own site: tz + 1

  // TMA boxes: origin (array z index kbase: even, so the run is 16-byte aligned; array row of the first interior row)
  k.kbase = kbase; k.jrow0 = jbase + 1 + nh - 1;
  // periodic images of the own site in y / z
  k.sy = (g.wrap[1] && jc == 1) ? 1 : ((g.wrap[1] && jc == g.nl[1]) ? -1 : 0);
  k.sz = (g.wrap[2] && kc == 1) ? 1 : ((g.wrap[2] && kc == g.nl[2]) ? -1 : 0);
  if (g.nl[1] == 1) k.sy = 0;                      // (a one-site dimension: the fused step is not used, see the launcher)
  k.imy = k.sy*g.nl[1]*ys; k.imz = k.sz*g.nl[2];
  k.py = (g.wrap[1] && jc <= nh) ? 1 : ((g.wrap[1] && jc > g.nl[1] - nh) ? -1 : 0);
  k.pz = (g.wrap[2] && kc <= nh) ? 1 : ((g.wrap[2] && kc > g.nl[2] - nh) ? -1 : 0);
  if (!k.out_site) { k.sy = 0; k.sz = 0; k.py = 0; k.pz = 0; }
  k.tu = ty*FU_ROW + tz + 1;
  k.smb = fu_smem_u32(&sm);
  k.imy_unit = g.nl[1]*ys; k.imz_unit = g.nl[2];


  const int istart = k.i0 - 2;
  k.smb = fu_smem_u32(&sw);                        // (FuShared is the first member)

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < FU_NSTAGE; s++) { fu_mbar_init(FU_AD(full) + 8u*s, 1u); fu_mbar_init(FW_AD(sempty) + 8u*s, (unsigned int) FW_NCONS); }
#pragma unroll
    for (int s = 0; s < FU_RING; s++) fu_mbar_init(FU_AD(pl) + 8u*s, 1u);
#pragma unroll
    for (int s = 0; s < FW_NF; s++) { fu_mbar_init(FW_AD(ffull) + 8u*s, 1u); fu_mbar_init(FW_AD(fempty) + 8u*s, (unsigned int) FW_NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();                                 // the only CTA-wide barrier: the mbarriers are initialised

  if (producer) {
    // ================================ phi sector ================================

  // array coordinates of the staged planes: phi rows jbase-1 .. jbase+BY (array row jbase + nh - 2), u rows jbase .. jbase+BY-1
  const int prow = jbase + nh - 2, urow = jbase + nh - 1;
  const bool issuer = (k.tid == 32*(BY - 1));      // lane 0 of the last warp (an apron row: no collision work)

  // prologue: phi planes istart .. istart+3 -> ring slots 0 .. 3; u_x(istart+1), u_x(istart+2) -> slots 1, 2;
  // u_y, u_z (istart+1) -> slot 1.  (The first plane-step prefetches phi(istart+4), u_x(istart+3), u_y/u_z(istart+2).)
  // They use barrier pl[5], whose first regular use is five plane-steps away.
  if (issuer) {
    fu_mbar_expect(FU_AD(pl) + 40u, (unsigned int) (8*(4*G::PY*FU_ROW + 4*BY*FU_ROW)));
#pragma unroll
    for (int d = 0; d < 4; d++) {
      const int xp = ps_wrap(istart + d, k.nlx, k.wx) + nh - 1;
      fu_tma_box3(FU_AD(phi) + 8u*G::PSLOT*d, &phimap, k.kbase, prow, xp, FU_AD(pl) + 40u);
      if (d == 1 || d == 2) fu_tma_box(FU_AD(ux) + 8u*G::USLOT*d, &umap, k.kbase, urow, xp, 0, FU_AD(pl) + 40u);
      if (d == 1) {
	fu_tma_box(FU_AD(u) + 8u*G::USLOT*2, &umap, k.kbase, urow, xp, 1, FU_AD(pl) + 40u);
	fu_tma_box(FU_AD(u) + 8u*G::USLOT*3, &umap, k.kbase, urow, xp, 2, FU_AD(pl) + 40u);
      }
    }
  }
  fu_mbar_wait(FU_AD(pl) + 40u, 0u);

#ifndef LB200_STRICT
  PfRegs r;
  r.uxc = 0.0;                                     // u_x(n): not used before n = i0 - 1
  fu_plane_sums<G::PZ>(sm.phi[0], k.pc, r.Bm, r.Cym, r.Czm);
  fu_plane_sums<G::PZ>(sm.phi[1], k.pc, r.Bc, r.Cyc, r.Czc);
  r.gm_xx = r.gm_xy = r.gm_xz = 0.0;
  r.gc_xx = r.gc_xy = r.gc_xz = r.gc_mu = 0.0;
  r.phim1 = 0.0;
  r.fxm1 = r.fxm2 = r.fy_prev = r.fz_prev = 0.0;
#else
  // strict arithmetic (the reference's operation order, phi_sector_kernel): own-column history of the march
  double gm_p0 = 0.0, gm_x = 0.0, gm_y = 0.0, gm_z = 0.0, gm_mu = 0.0;      // plane n-1
  double gc_p0 = 0.0, gc_x = 0.0, gc_y = 0.0, gc_z = 0.0, gc_mu = 0.0;      // plane n
  double phim1 = 0.0, phim2 = 0.0;                                        // phi(n-1), phi(n-2), own column
  double uxm = 0.0, uxc = 0.0;                                            // u_x(n-1), u_x(n)
#endif

  const int tid = k.tid, pc = k.pc;
  const int typ = tid + G::BZ, tym = tid - G::BZ, tzp = tid + 1, tzm = tid - 1;
  const double r9 = (1.0/9.0), r18 = 0.5*(1.0/9.0);
  const int fidx = (ty - 1)*32 + tz;                // own slot of the force block (interior rows)

  // Lees-Edwards planes (le_xb > 0: one every le_xb x-planes; le_x0 = the plane below the first one, minus one): the 2*nhalo
  // x-planes whose stencils cross a plane are produced by the patch kernels, concurrently -- nothing is stored for them here
  int ler = (le_xb > 0) ? (istart - 1 - le_x0 + 2*le_xb) % le_xb : 0;      // (plane n-1) relative to the patched window
  int q = 0;                                       // phase: (plane n - first plane) % 6
  for (int n = istart; n <= k.i1 + 1; n++) {
    const bool le_skip = (le_xb > 0) && (ler < 4);
    if (++ler == le_xb) ler = 0;
    const bool do_grad = (n <= k.i1);
    const bool do_fx   = (n >= k.i0 - 1 && n <= k.i1);
    const bool do_full = (n >= k.i0 && n <= k.i1);
    const bool do_upd  = (n >= k.i0 + 1);
#ifdef LB200_STRICT
    (void) le_skip; (void) do_fx; (void) do_upd; (void) r18;   // (planes: the fast build only; the strict march updates plane n itself)
#endif
    const int q1 = (q + 1 >= 6) ? q - 5 : q + 1;
    const int q2 = (q + 2 >= 6) ? q - 4 : q + 2;
    const int q4 = (q + 4 >= 6) ? q - 2 : q + 4;
    const int u0 = (q >= 3) ? q - 3 : q;           // q % 3
    const int u1 = (u0 + 1 >= 3) ? u0 - 2 : u0 + 1;
    const int u2 = (u0 + 2 >= 3) ? u0 - 1 : u0 + 2;

    // ---- 1. asynchronous prefetch (TMA): phi(n+4), u_x(n+3), u_y / u_z (n+2) ----
    {
      if (issuer) {
	const bool do_phi = (n + 4 <= k.i1 + 2), do_ux = (n + 3 <= k.i1 + 1), do_uyz = (n + 2 <= k.i1);
	// (the prologue's use of pl[5] was phase 0 of that barrier: the regular uses start one phase later)
	const unsigned int plq = FU_AD(pl) + 8u*q;
	fu_mbar_expect(plq, (unsigned int) (8*((do_phi ? G::PY*FU_ROW : 0) + (do_ux ? BY*FU_ROW : 0) + (do_uyz ? 2*BY*FU_ROW : 0))));
	if (do_phi) fu_tma_box3(FU_AD(phi) + 8u*G::PSLOT*q4, &phimap, k.kbase, prow, ps_wrap(n + 4, k.nlx, k.wx) + k.nh - 1, plq);
	if (do_ux) fu_tma_box(FU_AD(ux) + 8u*G::USLOT*u0, &umap, k.kbase, urow, ps_wrap(n + 3, k.nlx, k.wx) + k.nh - 1, 0, plq);
	if (do_uyz) {
	  const int xp = ps_wrap(n + 2, k.nlx, k.wx) + k.nh - 1;
	  fu_tma_box(FU_AD(u) + 8u*G::USLOT*(2*u2), &umap, k.kbase, urow, xp, 1, plq);
	  fu_tma_box(FU_AD(u) + 8u*G::USLOT*(2*u2 + 1), &umap, k.kbase, urow, xp, 2, plq);
	}
      }

    }

#ifndef LB200_STRICT
    const double * __restrict__ fm = sm.phi[q];
    const double * __restrict__ fc = sm.phi[q1];
    const double * __restrict__ fp = sm.phi[q2];

    // ---- 2. gradient, chemical potential and stress of plane n+1 at the own column ----
    double gp_xx = 0.0, gp_xy = 0.0, gp_xz = 0.0, gp_mu = 0.0;
    double Bp = 0.0, Cyp = 0.0, Czp = 0.0;
    if (do_grad) {
      fu_plane_sums<G::PZ>(fp, pc, Bp, Cyp, Czp);
      const double phc = fc[pc];
      const double gx = r18*(Bp - r.Bm);
      const double gy = r18*((r.Cym + r.Cyc) + Cyp);
      const double gz = r18*((r.Czm + r.Czc) + Czp);
      const double dsq = r9*(((r.Bm + r.Bc) + Bp) - 27.0*phc);

      const int ig = n + 1;
      const bool own_x = (ig >= k.i0 && ig <= k.i1) || (ig == 0 && k.i0 == 1) || (ig == k.nlx + 1 && k.i1 == k.nlx);
      if (k.own_g && own_x) {
	const size_t sidx = (size_t) ((ig + k.nh - 1)*k.xs + k.scol);
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
      double (* gb)[G::NT] = sm.g[q1 & 1];
      gb[0][tid] = gp_xy;
      gb[1][tid] = p0 + kgy*gy;
      gb[2][tid] = kgy*gz;
      gb[3][tid] = gp_xz;
      gb[4][tid] = p0 + kgz*gz;
      gb[5][tid] = gp_mu;
    }

    // ---- 3. plane n: x-face flux (n | n+1), force, y/z face fluxes ----
    const double ph_c = fm[pc];
    const double uxp = sm.ux[u1][k.tu];              // u_x(n+1)
    double fx = 0.0, fy = 0.0, fz = 0.0;
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    if (do_fx && k.face_row) {
      fx = adv_face<ORDER, false>(r.uxc, uxp, r.phim1, ph_c, fc[pc], fp[pc]) - k.M*(gp_mu - r.gc_mu) - k.mg0;

      if (do_full) {
	const double (* gb)[G::NT] = sm.g[q & 1];
	const double (* ub)[G::USLOT] = sm.u[u0];
	const int tu = k.tu;

	if (k.out_site) {
	  F0 = 0.5*(((r.gm_xx - gp_xx) + (gb[0][tym] - gb[0][typ])) + (gb[3][tzm] - gb[3][tzp]));
	  F1 = 0.5*(((r.gm_xy - gp_xy) + (gb[1][tym] - gb[1][typ])) + (gb[2][tzm] - gb[2][tzp]));
	  F2 = 0.5*(((r.gm_xz - gp_xz) + (gb[2][tym] - gb[2][typ])) + (gb[4][tzm] - gb[4][tzp]));
	}

	double ph_yp2 = 0.0, ph_zp2 = 0.0;
	if (ORDER == 3) { ph_yp2 = fm[pc + 2*G::PZ]; ph_zp2 = fm[pc + 2]; }
	fy = adv_face<ORDER, false>(ub[0][tu], ub[0][tu + FU_ROW], fm[pc - G::PZ], ph_c, fm[pc + G::PZ], ph_yp2)
	  - k.M*(gb[5][typ] - r.gc_mu) - k.mg1;
	fz = adv_face<ORDER, false>(ub[1][tu], ub[1][tu + 1], fm[pc - 1], ph_c, fm[pc + 1], ph_zp2)
	  - k.M*(gb[5][tzp] - r.gc_mu) - k.mg2;
	sm.fl[q & 1][0][tid] = fy;
	sm.fl[q & 1][1][tid] = fz;
      }
    }

    // ---- 4. phi update of plane n-1, whose y/z face fluxes were published one plane-step ago ----
    if (do_upd && k.out_site && !le_skip) {
      const double (* fl)[G::NT] = sm.fl[q1 & 1];
      const int s = (n - 1 + k.nh - 1)*k.xs + k.scol;
      const double phn = r.phim1 - (((r.fxm1 - r.fxm2) + (r.fy_prev - fl[0][tym])) + k.wz*(r.fz_prev - fl[1][tzm]));
      phinew[s] = phn;
      // periodic images within nhalo of a y / z boundary (read by the next step's TMA boxes of phi)
      const int ipy = k.py*k.imy_unit, ipz = k.pz*k.imz_unit;
      if (k.py != 0) phinew[s + ipy] = phn;
      if (k.pz != 0) phinew[s + ipz] = phn;
      if (k.py != 0 && k.pz != 0) phinew[s + ipy + ipz] = phn;
      // the planes the neighbour GPUs' next phi sector reads, straight into their halo planes (with their y / z images)
      if (k.peer_lo != nullptr && n - 1 <= k.nh) {
	double * pl = k.peer_lo + ((size_t) s + (size_t) k.nlx*k.xs);
	pl[0] = phn;
	if (k.py != 0) pl[ipy] = phn;
	if (k.pz != 0) pl[ipz] = phn;
	if (k.py != 0 && k.pz != 0) pl[ipy + ipz] = phn;
      }
      if (k.peer_hi != nullptr && n - 1 > k.nlx - k.nh) {
	double * ph = k.peer_hi + ((size_t) s - (size_t) k.nlx*k.xs);
	ph[0] = phn;
	if (k.py != 0) ph[ipy] = phn;
	if (k.pz != 0) ph[ipz] = phn;
	if (k.py != 0 && k.pz != 0) ph[ipy + ipz] = phn;
      }
    }

    // ---- 5. rotate the own-column history ----
    r.Bm = r.Bc; r.Cym = r.Cyc; r.Czm = r.Czc;
    r.Bc = Bp; r.Cyc = Cyp; r.Czc = Czp;
    r.gm_xx = r.gc_xx; r.gm_xy = r.gc_xy; r.gm_xz = r.gc_xz;
    r.gc_xx = gp_xx; r.gc_xy = gp_xy; r.gc_xz = gp_xz;
    r.gc_mu = gp_mu;
    r.phim1 = ph_c;
    r.fxm2 = r.fxm1; r.fxm1 = fx;
    r.fy_prev = fy; r.fz_prev = fz;
    r.uxc = uxp;


#else
    const double * __restrict__ fm = sm.phi[q];
    const double * __restrict__ fc = sm.phi[q1];
    const double * __restrict__ fp = sm.phi[q2];
    const double uxp = sm.ux[u1][k.tu];              // u_x(n+1)
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;

    // ---- 2. gradient of plane n+1 at the own column, from phi planes n, n+1, n+2: the reference's summation order
    //         (src/gradient_3d_27pt_fluid.c:268-358), then p0 and mu of that site (src/symmetric.c:307-319, 371-416) ----
    double gp_p0 = 0.0, gp_x = 0.0, gp_y = 0.0, gp_z = 0.0, gp_mu = 0.0;
    if (k.valid_g && do_grad) {
      constexpr int PZ = G::PZ;
      const double m_mm = fm[pc-PZ-1], m_m0 = fm[pc-PZ], m_mp = fm[pc-PZ+1];
      const double m_0m = fm[pc   -1], m_00 = fm[pc   ], m_0p = fm[pc   +1];
      const double m_pm = fm[pc+PZ-1], m_p0 = fm[pc+PZ], m_pp = fm[pc+PZ+1];
      const double c_mm = fc[pc-PZ-1], c_m0 = fc[pc-PZ], c_mp = fc[pc-PZ+1];
      const double c_0m = fc[pc   -1], c_00 = fc[pc   ], c_0p = fc[pc   +1];
      const double c_pm = fc[pc+PZ-1], c_p0 = fc[pc+PZ], c_pp = fc[pc+PZ+1];
      const double p_mm = fp[pc-PZ-1], p_m0 = fp[pc-PZ], p_mp = fp[pc-PZ+1];
      const double p_0m = fp[pc   -1], p_00 = fp[pc   ], p_0p = fp[pc   +1];
      const double p_pm = fp[pc+PZ-1], p_p0 = fp[pc+PZ], p_pp = fp[pc+PZ+1];

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

      const int ig = n + 1;
      const bool own_x = (ig >= k.i0 && ig <= k.i1) || (ig == 0 && k.i0 == 1) || (ig == k.nlx + 1 && k.i1 == k.nlx);
      if (k.own_g && own_x) {
	const size_t sidx = (size_t) ((ig + k.nh - 1)*k.xs + k.scol);
	grad[sidx] = gp_x;
	grad[k.ns + sidx] = gp_y;
	grad[2*k.ns + sidx] = gp_z;
	delsq[sidx] = dsq;
      }
    }
    {
      double (* gb)[G::NT] = sm.g[q1 & 1];
      gb[0][tid] = gp_p0; gb[1][tid] = gp_x; gb[2][tid] = gp_y; gb[3][tid] = gp_z; gb[4][tid] = gp_mu;
    }

    // ---- 3. force and Cahn-Hilliard update of plane n (src/phi_force_colloid.c:315-465, src/advection.c:946-1141,
    //         src/phi_cahn_hilliard.c:350-404, 1018-1049, 1373-1397): the reference's operations in the reference's order ----
    const double ph_c = fm[pc];
    if (do_full && k.out_site) {
      const double (* gb)[G::NT] = sm.g[q & 1];
      const double (* ub)[G::USLOT] = sm.u[u0];
      const int tu = k.tu;
      const int s = (n + k.nh - 1)*k.xs + k.scol;

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
      F0 = fo[0]; F1 = fo[1]; F2 = fo[2];

      // Cahn-Hilliard: six face fluxes in registers, forward Euler
      const double M = sp.mobility;
      const double mu0 = gc_mu;
      const double ph_xm = phim1, ph_xm2 = phim2, ph_xp = fc[pc], ph_xp2 = fp[pc];
      const double ph_ym = fm[pc - G::PZ], ph_yp = fm[pc + G::PZ];
      const double ph_zm = fm[pc - 1], ph_zp = fm[pc + 1];
      double ph_ym2 = 0.0, ph_yp2 = 0.0, ph_zm2 = 0.0, ph_zp2 = 0.0;
      if (ORDER == 3) {
	ph_ym2 = fm[pc - 2*G::PZ]; ph_yp2 = fm[pc + 2*G::PZ];
	ph_zm2 = fm[pc - 2];       ph_zp2 = fm[pc + 2];
      }
      const double uy_c = ub[0][tu], uy_ym = ub[0][tu - FU_ROW], uy_yp = ub[0][tu + FU_ROW];
      const double uz_c = ub[1][tu], uz_zm = ub[1][tu - 1], uz_zp = ub[1][tu + 1];

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

      double phn = ph_c;
      phn -= (+ fe - fw + fy - fym + sp.wz*fz - sp.wz*fzm);
      phinew[s] = phn;
      // periodic images within nhalo of a y / z boundary (read by the next step's TMA boxes of phi)
      const int ipy = k.py*k.imy_unit, ipz = k.pz*k.imz_unit;
      if (k.py != 0) phinew[s + ipy] = phn;
      if (k.pz != 0) phinew[s + ipz] = phn;
      if (k.py != 0 && k.pz != 0) phinew[s + ipy + ipz] = phn;
      // the planes the neighbour GPUs' next phi sector reads, straight into their halo planes (with their y / z images)
      if (k.peer_lo != nullptr && n <= k.nh) {
	double * pl = k.peer_lo + ((size_t) s + (size_t) k.nlx*k.xs);
	pl[0] = phn;
	if (k.py != 0) pl[ipy] = phn;
	if (k.pz != 0) pl[ipz] = phn;
	if (k.py != 0 && k.pz != 0) pl[ipy + ipz] = phn;
      }
      if (k.peer_hi != nullptr && n > k.nlx - k.nh) {
	double * ph = k.peer_hi + ((size_t) s - (size_t) k.nlx*k.xs);
	ph[0] = phn;
	if (k.py != 0) ph[ipy] = phn;
	if (k.pz != 0) ph[ipz] = phn;
	if (k.py != 0 && k.pz != 0) ph[ipy + ipz] = phn;
      }
    }

    // ---- 4. rotate the own-column history ----
    gm_p0 = gc_p0; gm_x = gc_x; gm_y = gc_y; gm_z = gc_z; gm_mu = gc_mu;
    gc_p0 = gp_p0; gc_x = gp_x; gc_y = gp_y; gc_z = gp_z; gc_mu = gp_mu;
    phim2 = phim1; phim1 = ph_c;
    uxm = uxc; uxc = uxp;

#endif
    // ---- 6. hand the force of plane n to the collision warps ----
    if (do_full && k.has_sites) {
      const int fs = (n - k.i0) % FW_NF;
      // the slot is free once the collision warps have taken the force of plane n - FW_NF out of it
      if (n - k.i0 >= FW_NF) fu_mbar_wait(FW_AD(fempty) + 8u*fs, (unsigned int) (((n - k.i0)/FW_NF - 1) & 1));
      if (ty >= 1 && ty <= G::TY) { sw.F[fs][0][fidx] = F0; sw.F[fs][1][fidx] = F1; sw.F[fs][2][fidx] = F2; }
    }

    // the phi / u planes issued ONE plane-step ago have landed (every producer thread reads them after the barrier)
    if (n > istart) {
      const int qp = (q == 0) ? 5 : q - 1;
      const int uses = (n - 1 - istart)/6 + (qp == 5 ? 1 : 0);
      fu_mbar_wait(FU_AD(pl) + 8u*qp, (unsigned int) (uses & 1));
    }
    fw_bar_sync<1, 32*FW_BY>();                    // the producer warps' plane-step barrier
    if (do_full && k.has_sites && threadIdx.x == 0) fw_mbar_arrive(FW_AD(ffull) + 8u*((n - k.i0) % FW_NF));    // the force of plane n is complete
    q = q1;
  }
  }
  else if (loader) {
    // ================================ TMA of the populations ================================
    // plane m into stage (m - i0) % 3 as soon as every collision warp has taken plane m - 3 out of it
    if (tz == 0 && k.has_sites) {
      for (int m = k.i0; m <= k.i1; m++) {
	const int st = (m - k.i0) % FU_NSTAGE;
	const unsigned int fst = FU_AD(full) + 8u*st;
	if (m - k.i0 >= FU_NSTAGE) fu_mbar_wait(FW_AD(sempty) + 8u*st, (unsigned int) (((m - k.i0)/FU_NSTAGE - 1) & 1));
	fu_mbar_expect(fst, (unsigned int) (19*G::TY*FU_ROW*8));
#pragma unroll
	for (int p = 0; p < 19; p++) {
	  // population p arrives from the site at -c_p: rows j - c_y of plane m - c_x
	  fu_tma_box(FU_AD(f) + 8u*G::FBLK*(19*st + p), &fmap, k.kbase, k.jrow0 - CV19[p][1], ps_wrap(m - CV19[p][0], k.nlx, k.wx) + k.nh - 1, p, fst);
	}
      }
    }
  }
  else {
    // ================================ pull-stream + collision ================================
    const int fidx = (ty - 1)*32 + tz;
    int ler = (le_xb > 0) ? (k.i0 - le_x0 + 2*le_xb) % le_xb : 0;
    for (int m = k.i0; m <= k.i1 && k.has_sites; m++) {
      const bool le_skip = (le_xb > 0) && (ler < 4);     // a plane next to a Lees-Edwards plane: collided by the patch kernels
      if (++ler == le_xb) ler = 0;
      const int fs = (m - k.i0) % FW_NF;
      // the populations of the own site (every lane reads its slot, so that the stage can be released by the warp) ...
      double f[19];
      fu_collide_read<BY>(sm, k, m, f);
      // ... and the force of plane m (written by the producer warps one barrier ago)
      fu_mbar_wait(FW_AD(ffull) + 8u*fs, (unsigned int) (((m - k.i0)/FW_NF) & 1));
      const double F0 = sw.F[fs][0][fidx], F1 = sw.F[fs][1][fidx], F2 = sw.F[fs][2][fidx];
      __syncwarp();
      if (tz == 0) { fw_mbar_arrive(FW_AD(sempty) + 8u*((m - k.i0) % FU_NSTAGE)); fw_mbar_arrive(FW_AD(fempty) + 8u*fs); }
      if (k.out_site && !le_skip) fu_collide<GHOST, BY, true>(sm, k, cp, m, F0, F1, F2, fsrc, fdst, force, rho_out, u_out, f);
    }
  }
}

#undef FW_AD

int launch_step_fused_ws(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, const Lb200CollideDev & cp,
			 const double * phi, const double * u, const double * fsrc, double * fdst, double * grad,
			 double * delsq, double * force, double * phinew, double * rho, double * u_out, int le_x0, int le_xb) {
  constexpr int BY = FW_BY;
  using G = FuGeo<BY>;
  const int ext = g.skip_diag ? 0 : 1;
  const int gz = (g.nl[2] + ext + G::TZ - 1)/G::TZ, gy = (g.nl[1] + ext + G::TY - 1)/G::TY;
  const size_t smem = sizeof(FwShared);
  static bool configured[LB200_MAX_DEVICES] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= LB200_MAX_DEVICES || !configured[dev]) {
#define LB200_FW_ATTR(O, GH) cudaFuncSetAttribute(step_fused_ws_kernel<O, GH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)
    LB200_FW_ATTR(1, false); LB200_FW_ATTR(2, false); LB200_FW_ATTR(3, false);
    LB200_FW_ATTR(1, true); LB200_FW_ATTR(2, true); LB200_FW_ATTR(3, true);
#undef LB200_FW_ATTR
    if (dev >= 0 && dev < LB200_MAX_DEVICES) configured[dev] = true;
  }
  CUtensorMap fmap, phimap, umap;
  if (!fu_tensor_map(fsrc, g, 19, FU_ROW, G::TY, &fmap) || !fu_tensor_map(phi, g, 0, FU_ROW, G::PY, &phimap)
      || !fu_tensor_map(u, g, 3, FU_ROW, BY, &umap)) return 0;
  const int nx = (g.xcnt > 0) ? g.xcnt : g.nl[0];
  // x-chunks of ~24 planes: measured at 256^3 (1.36 ms with chunks of 64, 1.22 with 32, 1.20 with 16-26, 1.24 with 37) -- the
  // pipeline fill of a chunk is hidden behind the collision warps of the previous CTA's tail, and many short CTAs keep the SMs
  // evenly loaded (LB200_PS_XC / g.xchunk override)
  int xc = g.xchunk;
  if (xc <= 0) {
    const char * e = getenv("LB200_PS_XC");
    static const int target = tuned_flag("LB200_FUSED_XC", 24);
    if (e && atoi(e) > 0) xc = atoi(e);
    else {
      const int nchunk = (nx + target/2)/target > 0 ? (nx + target/2)/target : 1;
      xc = (nx + nchunk - 1)/nchunk;
    }
  }
  dim3 grd(gz, gy, (nx + xc - 1)/xc);
#define LB200_FW_GO(O, GH) step_fused_ws_kernel<O, GH><<<grd, FW_NT, smem, st>>>(fmap, phimap, umap, g, sp, cp, xc, le_x0, le_xb, phi, u, fsrc, fdst, grad, delsq, force, phinew, rho, u_out)
  if (cp.ghost) {
    if (sp.order == 1) LB200_FW_GO(1, true); else if (sp.order == 2) LB200_FW_GO(2, true); else LB200_FW_GO(3, true);
  }
  else {
    if (sp.order == 1) LB200_FW_GO(1, false); else if (sp.order == 2) LB200_FW_GO(2, false); else LB200_FW_GO(3, false);
  }
#undef LB200_FW_GO
  return 1;
}


