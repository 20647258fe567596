// lb200_fused.cuh -- one kernel per binary-fluid time step (fast arithmetic mode).
//
// The phi sector (27-point gradient, chemical stress, force divergence, Cahn-Hilliard update: the march of
// phi_sector_fast_kernel) and the pull-stream + MRT collision of the SAME plane in ONE sweep, so that the body force
// never exists in memory: the thread that has just formed F(n,j,k) in registers also collides the 19 populations
// pulled to site (n,j,k) and stores f', u (reference: src/ludwig.c:528-860 = field_grad_compute,
// phi_force_calculation, phi_cahn_hilliard, lb_collide, lb_halo, lb_propagation; the arithmetic is that of
// pf_step and collide_d3q19_kernel, whose device functions are shared).
//
// HBM traffic per site and step: f 152 read + 152 written, phi 8 + 8, u 24 + 24 = 368 B instead of 416 B for
// the two-kernel step (force 24 written + 24 read) and 496 B for the reference's minimal three sweeps.
//
// How the two halves share an SM:
//  * The populations arrive by TMA tensor copies (cp.async.bulk.tensor.4d, SASS UTMALDG) -- one box per population and
//    plane-step: TY source rows x 34 doubles (the 16-byte aligned z run that contains the pulled range) of plane n - c_x,
//    rows j - c_y.  No register, no LSU instruction and no L1 line is held while ~40 kB per SM are in flight; a thread
//    reads its 19 values back with conflict-free LDS.64 at element (lane + 1 - c_z).  Three stages, one mbarrier
//    (complete_tx) each: the boxes of plane n+1 are issued at the start of plane-step n and first read 1.5 plane-steps
//    later.  A box cannot wrap around the lattice, so the kernel that PRODUCES f' also stores, for sites on a y / z
//    boundary, the few populations its periodic images will be pulled from into the halo rows / columns (5 of 19 per
//    face image, 1 per edge image), and likewise into the neighbour GPUs' halo planes: the next step's boxes read valid
//    halos without a halo sweep.  (The first such step after anything else wrote f runs one lb_halo.)
//  * The collision is FP64-pipe work, the phi sector shared-memory-pipe work.  Warps of even rows run
//    [phi(n); collide(n)], warps of odd rows [collide(n-1); phi(n)] inside the same barrier interval, so at any time
//    half of the warps are in each kind of work.
//
// u is double-buffered by the caller (the phi sector reads u(t-1) of neighbouring sites while the collision of
// other CTAs writes u(t)); phi / phinew and f / fprime are double-buffered anyway.

constexpr int FU_RING = 6;                   // phi planes in flight: n .. n+2 read, n+3 landed, n+4 in flight
constexpr int FU_NSTAGE = 3;
constexpr int FU_ROW = 34;                   // doubles per staged row: pulled z range [kbase, kbase+31] inside an aligned run

template <int BY> struct FuGeo {
  static constexpr int BZ = 32;
  static constexpr int NT = BZ*BY;
  static constexpr int TZ = BZ - 2;          // interior columns per tile
  static constexpr int TY = BY - 2;
  static constexpr int PZ = BZ + 2;          // phi tile: block + one more ring
  static constexpr int PY = BY + 2;
  static constexpr int PN = PZ*PY;
  static constexpr int FBLK = ((TY*FU_ROW + 15)/16)*16;   // doubles per staged population (128-byte aligned blocks)
  static constexpr int PSLOT = ((PN + 15)/16)*16;         // doubles per staged phi plane [PY][34]
  static constexpr int USLOT = ((BY*FU_ROW + 15)/16)*16;  // doubles per staged velocity-component plane [BY][34]
};

template <int BY> struct FuShared {
  double f[FU_NSTAGE][19][FuGeo<BY>::FBLK];  // staged boxes [TY rows][34] of the pulled populations
  double phi[FU_RING][FuGeo<BY>::PSLOT];     // ring of phi planes [PY rows][34], slot = (plane - first plane) % 6
  double u[3][2][FuGeo<BY>::USLOT];          // u_y, u_z planes [BY rows][34], slot = plane % 3
  double ux[3][FuGeo<BY>::USLOT];            // u_x, slot = plane % 3
  double g[2][6][FuGeo<BY>::NT];             // Pxy, Pyy, Pyz, Pxz, Pzz, mu of a plane
  double fl[2][2][FuGeo<BY>::NT];            // y and z face fluxes (face between the site and site+1)
  unsigned long long full[FU_NSTAGE];        // mbarriers: the populations of a stage have landed
  unsigned long long pl[FU_RING];            // mbarriers: the phi / u planes issued at phase q have landed
};

struct FuK {                                 // per-thread / per-CTA constants
  int pc, tid, col, scol;
  bool valid_g, out_site, face_row, own_g, skip_diag, odd, has_sites;
  int fmode;                                 // tuning experiments: 0 bulk copies, 1 direct loads, 2 no loads (timing only)
  int xs, nh, nlx, wx, i0, i1;
  size_t ns;
  int oym, oyp, ozm, ozp;                    // offsets of the y / z neighbours (through the periodic boundary), direct loads
  int frow, flane;                           // interior row (0 .. TY-1) and lane of the own site in the staged rows
  int kbase, jrow0;                          // box origin: array z index / array row index of the first interior row
  int sy, sz;                                // +1: the site's image lies one period up (site on the low boundary), -1: down, 0: none
  int imy, imz;                              // element offsets of those images
  int py, pz;                                // the same for the nhalo-deep images of phi (sites within nhalo of a boundary)
  int tu;                                    // own element of a staged velocity plane
  unsigned int smb;                          // shared-window address of the FuShared block
  int imy_unit, imz_unit;                    // one period in y / z (elements)
  double M, kappa, a, b, mg0, mg1, mg2, wz;
  double * peer_lo, * peer_hi;               // neighbour GPUs' phi' arrays (nullptr: none)
  double * peer_f_lo, * peer_f_hi, * peer_u_lo, * peer_u_hi;
};

// c_x + 1 (a = 0) or c_y + 1 (a = 1) of the 19 velocities, two bits each
__device__ constexpr unsigned long long fu_cbits(int a) {
  unsigned long long b = 0;
  for (int p = 0; p < 19; p++) b |= (unsigned long long) (CV19[p][a] + 1) << (2*p);
  return b;
}

__device__ __forceinline__ unsigned int fu_smem_u32(const void * p) { return (unsigned int) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void fu_mbar_init(unsigned int bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fu_mbar_expect(unsigned int bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fu_mbar_wait(unsigned int bar, unsigned int parity) {
  asm volatile("{\n"
	       ".reg .pred p;\n"
	       "FU_WAIT_%=:\n"
	       "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	       "@p bra FU_DONE_%=;\n"
	       "bra FU_WAIT_%=;\n"
	       "FU_DONE_%=:\n"
	       "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// global -> shared TMA box load: coordinates (z, y, x, population) in elements of the 4-d tensor map of f
__device__ __forceinline__ void fu_tma_box(unsigned int dst, const CUtensorMap * map, int c0, int c1, int c2, int c3,
					   unsigned int bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
	       :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

__device__ __forceinline__ void fu_tma_box3(unsigned int dst, const CUtensorMap * map, int c0, int c1, int c2, unsigned int bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
	       :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <int PZ>
__device__ __forceinline__ void fu_plane_sums(const double * __restrict__ q, int pc, double & B,
					      double & Cy, double & Cz) {
  const double mm = q[pc-PZ-1], m0 = q[pc-PZ], mp = q[pc-PZ+1];
  const double zm = q[pc   -1], z0 = q[pc   ], zp = q[pc   +1];
  const double pm = q[pc+PZ-1], p0 = q[pc+PZ], pp = q[pc+PZ+1];
  const double am = (mm + m0) + mp;
  const double a0 = (zm + z0) + zp;
  const double ap = (pm + p0) + pp;
  B  = (am + a0) + ap;
  Cy = ap - am;
  Cz = ((mp - mm) + (zp - zm)) + (pp - pm);
}

// shared-window addresses of the staging areas and their barriers
#define FU_AD(member) (k.smb + (unsigned int) offsetof(FuShared<BY>, member))

// pull-stream + collision of the own site of plane m with the force (F0, F1, F2) (lb_collide, src/collision.c:253-593)
// the 19 populations pulled to the own site of plane m, out of the staged boxes (waits for the stage to land)
template <int BY>
__device__ __forceinline__ void fu_collide_read(FuShared<BY> & sm, const FuK & k, const int m, double (&f)[19]) {
  const int st = (m - k.i0) % FU_NSTAGE;
  fu_mbar_wait(FU_AD(full) + 8u*st, (unsigned int) (((m - k.i0)/FU_NSTAGE) & 1));
  const double * __restrict__ fb = &sm.f[st][0][k.frow*FU_ROW + k.flane];
#pragma unroll
  for (int p = 0; p < 19; p++) f[p] = fb[p*FuGeo<BY>::FBLK + 1 - CV19[p][2]];
}

template <bool GHOST, int BY, bool HAVE_F = false>
__device__ __forceinline__ void fu_collide(FuShared<BY> & sm, const FuK & k, const Lb200CollideDev & cp, const int m,
					   const double F0, const double F1, const double F2,
					   const double * __restrict__ fsrc, double * __restrict__ fdst,
					   double * __restrict__ force, double * __restrict__ rho_out,
					   double * __restrict__ u_out, const double * fin = nullptr) {
  const size_t s = (size_t) ((m + k.nh - 1)*k.xs + k.scol);
  double f[19], mode[19], fo[3], uu[3], rho;

  if (HAVE_F) {
    // (the caller has taken the populations out of the stage already: fu_collide_read)
#pragma unroll
    for (int p = 0; p < 19; p++) f[p] = fin[p];
  }
  else if (k.fmode == 1) {
    const int oxm = (k.wx && m == 1)     ?  (k.nlx - 1)*k.xs : -k.xs;
    const int oxp = (k.wx && m == k.nlx) ? -(k.nlx - 1)*k.xs :  k.xs;
#pragma unroll
    for (int p = 0; p < 19; p++) {
      const int off = (CV19[p][0] > 0 ? oxm : CV19[p][0] < 0 ? oxp : 0)
	+ (CV19[p][1] > 0 ? k.oym : CV19[p][1] < 0 ? k.oyp : 0)
	+ (CV19[p][2] > 0 ? k.ozm : CV19[p][2] < 0 ? k.ozp : 0);
      f[p] = fsrc[p*k.ns + s + off];
    }
  }
  else {
    const int st = (m - k.i0) % FU_NSTAGE;
    if (k.fmode == 0) fu_mbar_wait(FU_AD(full) + 8u*st, (unsigned int) (((m - k.i0)/FU_NSTAGE) & 1));
    const double * __restrict__ fb = &sm.f[st][0][k.frow*FU_ROW + k.flane];
#pragma unroll
    for (int p = 0; p < 19; p++) f[p] = fb[p*FuGeo<BY>::FBLK + 1 - CV19[p][2]];
  }
  fo[0] = cp.fg[0] + F0; fo[1] = cp.fg[1] + F1; fo[2] = cp.fg[2] + F2;

  d3q19_f2mode<GHOST>(f, mode);
  relax_hydro(mode, fo, cp, rho, uu);
  if (GHOST) {
#pragma unroll
    for (int i = 10; i < 19; i++) mode[i] = mode[i] - cp.rtau_ghost[i]*(mode[i] - 0.0);
  }
  d3q19_mode2f<GHOST>(mode, f);

  {
    // one 64-bit BYTE-pointer increment per population (2 instructions) instead of index -> address chains (4)
    // (the add is opaque to the compiler, which otherwise rebuilds every address from a running index: 4 instructions)
    const unsigned long long nsb = (unsigned long long) k.ns*sizeof(double);
    unsigned long long pf = (unsigned long long) (fdst + s);
#pragma unroll
    for (int p = 0; p < 19; p++) {
      __stcs(reinterpret_cast<double *>(pf), f[p]);
      asm("add.u64 %0, %0, %1;" : "+l"(pf) : "l"(nsb));
    }
    unsigned long long pu = (unsigned long long) (u_out + s);
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
      *reinterpret_cast<double *>(pu) = uu[ia];
      asm("add.u64 %0, %0, %1;" : "+l"(pu) : "l"(nsb));
    }
  }
  if (!k.skip_diag) {
    rho_out[s] = rho;
    force[s] = F0; force[k.ns + s] = F1; force[2*k.ns + s] = F2;
  }

  // Periodic images in y / z (the next step's TMA boxes read the halo rows / columns, they cannot wrap): a site on
  // the low boundary has its image one period up, where the populations with c = -1 in that dimension are pulled from.
  if ((k.sy | k.sz) != 0) {
#pragma unroll
    for (int ia = 0; ia < 3; ia++) {
      if (k.sy != 0) u_out[ia*k.ns + s + k.imy] = uu[ia];
      if (k.sz != 0) u_out[ia*k.ns + s + k.imz] = uu[ia];
      if (k.sy != 0 && k.sz != 0) u_out[ia*k.ns + s + k.imy + k.imz] = uu[ia];
    }
#pragma unroll
    for (int p = 0; p < 19; p++) {
      const bool my = (CV19[p][1] != 0) && (CV19[p][1] == -k.sy);
      const bool mz = (CV19[p][2] != 0) && (CV19[p][2] == -k.sz);
      if (CV19[p][1] != 0 && my) fdst[p*k.ns + s + k.imy] = f[p];
      if (CV19[p][2] != 0 && mz) fdst[p*k.ns + s + k.imz] = f[p];
      if (CV19[p][1] != 0 && CV19[p][2] != 0 && my && mz) fdst[p*k.ns + s + k.imy + k.imz] = f[p];
    }
  }

  // boundary planes straight into the neighbour GPUs' halo planes (see collide_d3q19_kernel), with their y / z images
  if (m == 1 && k.peer_f_lo != nullptr) {
    const size_t dst = s + (size_t) k.nlx*k.xs;
#pragma unroll
    for (int p = 0; p < 19; p++) {
      if (CV19[p][0] < 0) {
	k.peer_f_lo[p*k.ns + dst] = f[p];
	if (CV19[p][1] != 0 && CV19[p][1] == -k.sy) k.peer_f_lo[p*k.ns + dst + k.imy] = f[p];
	if (CV19[p][2] != 0 && CV19[p][2] == -k.sz) k.peer_f_lo[p*k.ns + dst + k.imz] = f[p];
      }
    }
    if (k.peer_u_lo != nullptr) k.peer_u_lo[dst] = uu[0];
  }
  if (m == k.nlx && k.peer_f_hi != nullptr) {
    const size_t dst = s - (size_t) k.nlx*k.xs;
#pragma unroll
    for (int p = 0; p < 19; p++) {
      if (CV19[p][0] > 0) {
	k.peer_f_hi[p*k.ns + dst] = f[p];
	if (CV19[p][1] != 0 && CV19[p][1] == -k.sy) k.peer_f_hi[p*k.ns + dst + k.imy] = f[p];
	if (CV19[p][2] != 0 && CV19[p][2] == -k.sz) k.peer_f_hi[p*k.ns + dst + k.imz] = f[p];
      }
    }
    if (k.peer_u_hi != nullptr) k.peer_u_hi[dst] = uu[0];
  }
}

#ifndef LB200_STRICT
template <int ORDER, bool GHOST, int BY>
__global__ void __launch_bounds__(FuGeo<BY>::NT, 1)
step_fused_kernel(const __grid_constant__ CUtensorMap fmap, const __grid_constant__ CUtensorMap phimap,
		  const __grid_constant__ CUtensorMap umap, const Lb200Geom g, const Lb200SymmDev sp,
		  const Lb200CollideDev cp, int xc, int fmode, int skew,
		  const double * __restrict__ phi, const double * __restrict__ u,
		  const double * __restrict__ fsrc, double * __restrict__ fdst,
		  double * __restrict__ grad, double * __restrict__ delsq,
		  double * __restrict__ force, double * __restrict__ phinew,
		  double * __restrict__ rho_out, double * __restrict__ u_out) {
  using G = FuGeo<BY>;
  extern __shared__ __align__(1024) unsigned char fu_smem_raw[];
  FuShared<BY> & sm = *reinterpret_cast<FuShared<BY> *>(fu_smem_raw);

  const int tz = threadIdx.x, ty = threadIdx.y;
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

  // tuning: de-synchronise the CTAs of a round (skew >> 1 = nanoseconds per CTA index step, 16 phases)
  if ((skew >> 1) > 0) __nanosleep((unsigned int) (((blockIdx.x + 3*blockIdx.y + 7*blockIdx.z) & 15)*(skew >> 1)));

  if (k.tid == 0) {
#pragma unroll
    for (int s = 0; s < FU_NSTAGE; s++) fu_mbar_init(FU_AD(full) + 8u*s, (unsigned int) BY);
#pragma unroll
    for (int s = 0; s < FU_RING; s++) fu_mbar_init(FU_AD(pl) + 8u*s, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  __syncthreads();                                 // the barriers are initialised

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

  PfRegs r;
  r.uxc = 0.0;                                     // u_x(n): not used before n = i0 - 1
  fu_plane_sums<G::PZ>(sm.phi[0], k.pc, r.Bm, r.Cym, r.Czm);
  fu_plane_sums<G::PZ>(sm.phi[1], k.pc, r.Bc, r.Cyc, r.Czc);
  r.gm_xx = r.gm_xy = r.gm_xz = 0.0;
  r.gc_xx = r.gc_xy = r.gc_xz = r.gc_mu = 0.0;
  r.phim1 = 0.0;
  r.fxm1 = r.fxm2 = r.fy_prev = r.fz_prev = 0.0;

  const int tid = k.tid, pc = k.pc;
  const int typ = tid + G::BZ, tym = tid - G::BZ, tzp = tid + 1, tzm = tid - 1;
  const double r9 = (1.0/9.0), r18 = 0.5*(1.0/9.0);
  double Fs0 = 0.0, Fs1 = 0.0, Fs2 = 0.0;          // force of the previous plane (rows that collide one plane-step late)

  int q = 0;                                       // phase: (plane n - first plane) % 6
  for (int n = istart; n <= k.i1 + 1; n++) {
    const bool do_grad = (n <= k.i1);
    const bool do_fx   = (n >= k.i0 - 1 && n <= k.i1);
    const bool do_full = (n >= k.i0 && n <= k.i1);
    const bool do_upd  = (n >= k.i0 + 1);
    const int q1 = (q + 1 >= 6) ? q - 5 : q + 1;
    const int q2 = (q + 2 >= 6) ? q - 4 : q + 2;
    const int q4 = (q + 4 >= 6) ? q - 2 : q + 4;
    const int u0 = (q >= 3) ? q - 3 : q;           // q % 3
    const int u1 = (u0 + 1 >= 3) ? u0 - 2 : u0 + 1;
    const int u2 = (u0 + 2 >= 3) ? u0 - 1 : u0 + 2;

    // ---- 1. asynchronous prefetch (TMA): phi(n+4), u_x(n+3), u_y / u_z (n+2); the populations of plane n+1 ----
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

      const int m = n + 1;
      if (k.fmode == 0 && tz == 0 && k.has_sites && m >= k.i0 && m <= k.i1) {
	// population p arrives from the site at -c_p (lb_propagation, src/propagation.c:153-200): one box per population,
	// TY rows from row j - c_y of plane m - c_x; lane 0 of warp w issues populations w and w + BY
	const int st = (m - k.i0) % FU_NSTAGE;
	const unsigned int fst = FU_AD(full) + 8u*st;
	fu_mbar_expect(fst, (unsigned int) ((ty + BY < 19 ? 2 : 1)*G::TY*FU_ROW*8));
#pragma unroll
	for (int pp = 0; pp < 2; pp++) {
	  const int p = ty + pp*BY;
	  if (p < 19) {
	    const int cx = (int) ((fu_cbits(0) >> (2*p)) & 3ull) - 1, cy = (int) ((fu_cbits(1) >> (2*p)) & 3ull) - 1;
	    fu_tma_box(FU_AD(f) + 8u*G::FBLK*(19*st + p), &fmap, k.kbase, k.jrow0 - cy, ps_wrap(m - cx, k.nlx, k.wx) + k.nh - 1, p, fst);
	  }
	}
      }
    }

    // ---- rows that collide one plane-step late: plane n-1 first (FP64 work next to the other rows' phi sector) ----
    if (k.odd && k.out_site && n - 1 >= k.i0 && n - 1 <= k.i1) {
      fu_collide<GHOST, BY>(sm, k, cp, n - 1, Fs0, Fs1, Fs2, fsrc, fdst, force, rho_out, u_out);
    }

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
    if (do_upd && k.out_site) {
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

    // ---- 6. the other rows: collision of plane n with the force of stage 3 ----
    if (!k.odd && do_full && k.out_site) {
      fu_collide<GHOST, BY>(sm, k, cp, n, F0, F1, F2, fsrc, fdst, force, rho_out, u_out);
    }
    Fs0 = F0; Fs1 = F1; Fs2 = F2;

    // the phi / u planes issued ONE plane-step ago have landed (every thread reads them after the barrier)
    if (n > istart) {
      const int qp = (q == 0) ? 5 : q - 1;
      // use count of pl[qp]: the prologue used pl[5] once before the march started
      const int uses = (n - 1 - istart)/6 + (qp == 5 ? 1 : 0);
      fu_mbar_wait(FU_AD(pl) + 8u*qp, (unsigned int) (uses & 1));
    }
    __syncthreads();
    q = q1;
  }
}

#endif   // the every-warp-does-both kernel: fast arithmetic mode only

// Tensor maps of the distribution arrays (rank 4: z, y, x, population; box 34 x TY x 1 x 1), created on first use
// through the driver entry point and cached per (array, geometry, box).
typedef CUresult (*fu_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
				 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
				 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct FuMapEntry { const void * ptr; int nall[3]; int nsites; int ncomp, bz, by; CUtensorMap map; };
// ncomp = 0: a scalar field (rank 3: z, y, x); else rank 4 with the component as the slowest dimension
static bool fu_tensor_map(const double * a, const Lb200Geom & g, int ncomp, int bz, int by, CUtensorMap * out) {
  static std::mutex mtx;
  static std::vector<FuMapEntry *> cache;
  static fu_encode_fn encode = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lock(mtx);
  for (FuMapEntry * e : cache) {
    if (e->ptr == (const void *) a && e->nall[0] == g.nall[0] && e->nall[1] == g.nall[1] && e->nall[2] == g.nall[2]
	&& e->nsites == g.nsites && e->ncomp == ncomp && e->bz == bz && e->by == by) { *out = e->map; return true; }
  }
  if (!tried) {
    tried = true;
    void * fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess
	&& q == cudaDriverEntryPointSuccess) encode = (fu_encode_fn) fn;
    else cudaGetLastError();
  }
  if (encode == nullptr) return false;
  FuMapEntry * e = new FuMapEntry;
  e->ptr = a; e->nall[0] = g.nall[0]; e->nall[1] = g.nall[1]; e->nall[2] = g.nall[2]; e->nsites = g.nsites;
  e->ncomp = ncomp; e->bz = bz; e->by = by;
  const cuuint32_t rank = ncomp > 0 ? 4 : 3;
  const cuuint64_t dims[4] = {(cuuint64_t) g.nall[2], (cuuint64_t) g.nall[1], (cuuint64_t) g.nall[0], (cuuint64_t) (ncomp > 0 ? ncomp : 1)};
  const cuuint64_t strides[3] = {(cuuint64_t) g.ys*8, (cuuint64_t) g.xs*8, (cuuint64_t) g.nsites*8};
  const cuuint32_t box[4] = {(cuuint32_t) bz, (cuuint32_t) by, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (encode(&e->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, (void *) a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    delete e;
    return false;
  }
  if (cache.size() >= 64) { delete cache.front(); cache.erase(cache.begin()); }      // contexts come and go
  cache.push_back(e);
  *out = e->map;
  return true;
}

#ifndef LB200_STRICT
template <int BY>
int launch_step_fused_by(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, const Lb200CollideDev & cp,
			 const double * phi, const double * u, const double * fsrc, double * fdst, double * grad,
			 double * delsq, double * force, double * phinew, double * rho, double * u_out) {
  using G = FuGeo<BY>;
  dim3 blk(G::BZ, BY, 1);
  // tiles: the steps that store grad / delsq (on [0, N+1]^3) need the tile that holds column N+1, the others only the sites
  const int ext = g.skip_diag ? 0 : 1;
  const int gz = (g.nl[2] + ext + G::TZ - 1)/G::TZ, gy = (g.nl[1] + ext + G::TY - 1)/G::TY;
  const size_t smem = sizeof(FuShared<BY>);
  static bool configured[LB200_MAX_DEVICES] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= LB200_MAX_DEVICES || !configured[dev]) {
#define LB200_FU_ATTR(O, GH) cudaFuncSetAttribute(step_fused_kernel<O, GH, BY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)
    LB200_FU_ATTR(1, false); LB200_FU_ATTR(2, false); LB200_FU_ATTR(3, false);
    LB200_FU_ATTR(1, true); LB200_FU_ATTR(2, true); LB200_FU_ATTR(3, true);
#undef LB200_FU_ATTR
    if (dev >= 0 && dev < LB200_MAX_DEVICES) configured[dev] = true;
  }
  CUtensorMap fmap, phimap, umap;
  if (!fu_tensor_map(fsrc, g, 19, FU_ROW, G::TY, &fmap) || !fu_tensor_map(phi, g, 0, FU_ROW, G::PY, &phimap)
      || !fu_tensor_map(u, g, 3, FU_ROW, BY, &umap)) return 0;
  const int nx = (g.xcnt > 0) ? g.xcnt : g.nl[0];
  const int xc = (g.xchunk > 0) ? g.xchunk
    : ps_pick_xc((const void *) step_fused_kernel<3, false, BY>, G::NT, smem, gz*gy, nx, 4, 1);
  dim3 grd(gz, gy, (nx + xc - 1)/xc);
  static const int fmode = tuned_flag("LB200_FUSED_FMODE", 0);
  static const int skew = tuned_flag("LB200_FUSED_SKEW", 1);
#define LB200_FU_GO(O, GH) step_fused_kernel<O, GH, BY><<<grd, blk, smem, st>>>(fmap, phimap, umap, g, sp, cp, xc, fmode, skew, phi, u, fsrc, fdst, grad, delsq, force, phinew, rho, u_out)
  if (cp.ghost) {
    if (sp.order == 1) LB200_FU_GO(1, true); else if (sp.order == 2) LB200_FU_GO(2, true); else LB200_FU_GO(3, true);
  }
  else {
    if (sp.order == 1) LB200_FU_GO(1, false); else if (sp.order == 2) LB200_FU_GO(2, false); else LB200_FU_GO(3, false);
  }
#undef LB200_FU_GO
  return 1;
}

#endif

int launch_step_fused_ws(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, const Lb200CollideDev & cp,
			 const double * phi, const double * u, const double * fsrc, double * fdst, double * grad,
			 double * delsq, double * force, double * phinew, double * rho, double * u_out, int le_x0, int le_xb);

// 1: launched; 0: this configuration has no one-kernel step (the caller runs the two-kernel step).
// le_xb > 0: Lees-Edwards planes every le_xb x-planes, the first between x = le_x0 + 1 and le_x0 + 2 -- the sweep stores
// nothing (phinew, f', u, rho, force) for x-planes le_x0 .. le_x0 + 3 (+ m le_xb): the caller's patch kernels produce them
int launch_step_fused(cudaStream_t st, const Lb200Geom & g, const Lb200SymmDev & sp, const Lb200CollideDev & cp,
		      const double * phi, const double * u, const double * fsrc, double * fdst, double * grad,
		      double * delsq, double * force, double * phinew, double * rho, double * u_out, int le_x0, int le_xb) {
  if (sp.order < 1 || sp.order > 3 || sp.csum != nullptr) return 0;
  // the staged source rows start at even array indices: 16-byte aligned only if every row does
  if (g.nh != 2 || (g.nall[2] & 1) || (g.nsites & 1) || (g.nl[2] & 1) || g.nl[1] < 2 || g.nl[2] < 2) return 0;
#ifndef LB200_STRICT
  static const int ws = tuned_flag("LB200_FUSED_WS", 1);      // warp-specialised roles (lb200_fused_ws.cuh); 0: every warp does both halves
  if (!ws && le_xb <= 0) {
    static const int by = tuned_flag("LB200_FUSED_BY", 10);
    if (by == 8) return launch_step_fused_by<8>(st, g, sp, cp, phi, u, fsrc, fdst, grad, delsq, force, phinew, rho, u_out);
    return launch_step_fused_by<10>(st, g, sp, cp, phi, u, fsrc, fdst, grad, delsq, force, phinew, rho, u_out);
  }
#endif
  return launch_step_fused_ws(st, g, sp, cp, phi, u, fsrc, fdst, grad, delsq, force, phinew, rho, u_out, le_x0, le_xb);
}
