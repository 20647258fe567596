// lb200_kernels.h -- launcher table shared by the two builds of lb200_kernels.cu
// (fast: FMA contraction on; strict: -fmad=false, bit-identical to the reference CPU build).
#pragma once

#include <cuda_runtime.h>

struct Lb200Geom {
  int nl[3];        // local extent
  int nh;           // halo width of the allocation
  int nall[3];      // nl + 2 nh
  int xs, ys;       // strides (zs = 1)
  int nsites;
  int per[3];       // periodic (global)
  int has_lo;       // an x-neighbour exists on the low / high side (periodic or interior slab)
  int has_hi;
  int remote_x;     // 1: x-neighbours are other GPUs, their planes arrive in staging buffers
  int wrap[3];      // 1: kernels read the periodic images of this dimension from the interior (no halo needed)
  // Peer stores (x-slabs on several GPUs, lb200_step): the neighbour GPUs' arrays, mapped into this process
  // (cudaIpc over NVLink).  A kernel that produces a boundary plane also stores it in the halo plane of the
  // neighbour that will read it: collide -> populations with c_x = -1 / +1 and u_x of planes 1 / N into the
  // low / high neighbour's destination buffers; phi sector -> the new phi of planes 1..nh / N-nh+1..N.
  // nullptr: no peer store.
  double * peer_f_lo, * peer_f_hi;
  double * peer_u_lo, * peer_u_hi;
  double * peer_phi_lo, * peer_phi_hi;
  // x sub-range of one launch (the slab pipeline of lb200_step, collide_d3q19 and the phi-sector kernels only):
  // planes xoff+1 .. xoff+xcnt; xcnt == 0: all of 1 .. nl[0].  xchunk > 0: planes per phi-sector CTA (0: chosen
  // from the SM count).
  int xoff, xcnt, xchunk;
  // 1: an intermediate step of a multi-step lb200_step call -- the arrays that only the caller reads are not
  // stored (hydro->rho by collide_d3q19*, grad / delsq by the phi-sector kernels, which keep them in registers);
  // the last step of the call stores them, so they hold what the reference's arrays hold when the call returns
  int skip_diag;
};

// Lees-Edwards planes (reference src/leesedwards.c).  Field arrays of an LE context carry 2*nh*nplane buffer
// x-planes after the high x halo (Lb200Geom::nsites counts them; nall[0] does not).
constexpr int LB200_LE_MAXPLANES = 16;
struct Lb200LeDev {
  int nplane;           // planes in this slab (0: none)
  int xblock;           // nl[0]/nplane
  int nprop;            // populations with c_x = +1 (= those with c_x = -1)
  int nvel;             // velocity set of the context (19: the plane kernels use the compile-time table)
  int loc[LB200_LE_MAXPLANES];   // plane p lies between local x = loc[p] and loc[p] + 1
  double uy;            // plane speed
};
// displacement of the buffer planes, formed on the host as the reference does (fmod, floor):
// index 0: buffers with velocity jump -uy, 1: +uy.  Cubic (phi): w = the four Lagrange weights;
// linear (u): w[0] = fr, w[1] = 1 - fr
struct Lb200LeInterp {
  int jdy[2];
  double w[2][4];
};
// index 0: the side below a plane (c_x = +1 populations / east face flux), 1: above
struct Lb200LeFix {
  int jdy[2];
  double fr[2];
  double ra;            // 0.5/(Ly Lz), force flux correction
};

// liquid crystal (fe_lc_param_t + beris_edw_param_t, src/blue_phase.h:52-75, src/blue_phase_beris_edwards.h:30-37)
struct Lb200LcDev {
  double a0, q0, gamma, kappa0, kappa1, xi;
  double Gamma;         // rotational diffusion constant
  double epsilon;       // dielectric anisotropy, already divided by 12 pi
  double e0[3];
  int order;            // advection order 1..3
  int is_active;        // active stress zeta0 d_ab - zeta1 Q_ab (lc_activity)
  double zeta0, zeta1;
  double redshift, rredshift;   // static redshift and its reciprocal (1.0/redshift, formed on the host as fe_lc_redshift_set does)
  int g2d;              // fd_gradient_calculation 2d_5pt_fluid (lattices with one plane in z)
};

struct Lb200CollideDev {
  double fg[3];         // force_global
  double rtau;          // 1/tau_shear
  double rtau_bulk;
  double tmr;           // 2.0 - rtau
  double rtau_ghost[27];
  int ghost;            // 0: every ghost mode relaxes at rate 1 (M10): ghosts are not projected
};

struct Lb200SymmDev {
  double a, b, kappa, mobility;
  double gm[3];
  int order;
  double wz;            // 0 if nlocal[Z] == 1
  double rtau2;         // 2/(1 + 2 mobility): relaxation of the order-parameter flux (symmetric_lb)
  double * csum;        // cahn_hilliard_options_conserve 1: per-site Kahan compensation (pch->csum); nullptr: plain update
  int force_method;     // fe_force_method: 0 = stress_divergence, 1 = phi_gradmu (phi_force only: the sweeps that fuse the force
                        // with other operators are the stress-divergence form)
};

struct Lb200ModelDev {       // generic (non-unrolled) model tables
  int nvel;
  signed char cv[27][3];
  double wv[27];
  double ma[27][27];
  double mi[27][27];
};

struct Lb200Kernels {
  // fused pull-stream + collide (pull = 1: fdst <- collide(pull(fsrc)));  pull = 0: in place on fsrc
  // force == nullptr: force field is identically zero.  status == nullptr: all fluid.
  int (*collide)(cudaStream_t, const Lb200Geom &, const Lb200CollideDev &, const Lb200ModelDev *,
		 int nvel, int pull, const double * fsrc, double * fdst, const double * force,
		 const char * status, double * rho, double * u);
  // reference propagation: fprime <- pull(f), x in [1,N], y/z halo self copy
  int (*propagate)(cudaStream_t, const Lb200Geom &, const Lb200ModelDev *, int nvel, int ndist,
		   const double * f, double * fprime);
  // halo shell of depth d for ncomp components; reduced != 0 only for distributions (needs cv)
  // snapshot != nullptr (lattices thinner than the swap depth): local sources are read from this copy of data taken
  // before the swap, as the reference's pack-everything-then-unpack order delivers them
  int (*halo)(cudaStream_t, const Lb200Geom &, const Lb200ModelDev *, int ncomp, int depth,
	      int reduced, double * data, const double * xlo, const double * xhi, const double * snapshot);
  int (*grad27)(cudaStream_t, const Lb200Geom &, int ne, const double * phi, double * grad, double * delsq);
  // force = [force +] -div P(phi, grad, delsq)   (accumulate = 0: plain store)
  int (*phi_force)(cudaStream_t, const Lb200Geom &, const Lb200SymmDev &, int accumulate,
		   const double * phi, const double * grad, const double * delsq, double * force);
  // phinew(interior) = phi - div(flux(phi, delsq, u)); status may be nullptr
  int (*cahn_hilliard)(cudaStream_t, const Lb200Geom &, const Lb200SymmDev &, const double * phi,
		       const double * delsq, const double * u, const char * status, double * phinew);
  // both of the above in one sweep
  int (*force_ch)(cudaStream_t, const Lb200Geom &, const Lb200SymmDev &, int accumulate,
		  const double * phi, const double * grad, const double * delsq, const double * u,
		  const char * status, double * force, double * phinew);
  // the whole phi sector of one time step in ONE sweep (all-fluid lattices): 27-point gradient (stored
  // on [0,N+1]^3), chemical stress + force divergence, Cahn-Hilliard fluxes + update
  int (*phi_sector)(cudaStream_t, const Lb200Geom &, const Lb200SymmDev &, const double * phi,
		    const double * u, double * grad, double * delsq, double * force, double * phinew);
  // zero everything outside the interior (ncomp components)
  int (*zero_outside)(cudaStream_t, const Lb200Geom &, int ncomp, double * data);
  // symmetric_lb (ndist = 2): phi = sum_p g_p (of the pulled populations if pull), g = (phi, 0, ..., 0),
  // and the two-distribution collision (pull = 1: fdst <- collide(pull(fsrc)), else in place, fdst == fsrc)
  int (*phi_from_g)(cudaStream_t, const Lb200Geom &, const Lb200ModelDev *, int pull, const double * f,
		    double * phi);
  int (*phi_to_g)(cudaStream_t, const Lb200Geom &, int nvel, const double * phi, double * f);
  int (*collide_binary)(cudaStream_t, const Lb200Geom &, const Lb200CollideDev &, const Lb200SymmDev &,
			const Lb200ModelDev *, int unrolled19, int pull, const double * fsrc, double * fdst,
			const double * force, const double * phi, const double * grad, const double * delsq,
			double * u);
  // pth_stress_compute / pth_force_fluid_driver as separate operators (P stored: 9 x nsites)
  int (*stress)(cudaStream_t, const Lb200Geom &, const Lb200SymmDev &, const double * phi, const double * grad,
		const double * delsq, double * str);
  int (*force_from_stress)(cudaStream_t, const Lb200Geom &, int accumulate, const double * str, double * force);
  // cross-GPU flags of the peer-store exchange: *a = *b = value after everything earlier in the stream
  // (either pointer may be nullptr); wait until *flag >= value (fallback when stream memory operations are
  // not available), giving up after ~timeout_ms with *err = 1
  int (*signal)(cudaStream_t, unsigned int * a, unsigned int * b, unsigned int value);
  int (*spin_wait)(cudaStream_t, const unsigned int * flag, unsigned int value, int timeout_ms, int * err);
  // Lees-Edwards planes (lb200_le.cuh): buffer planes of a field (cubic: phi, linear + jump: u), the 27-point
  // gradient on (x-1, x, x+1) plane triples, the flux-form force with its per-plane correction, the
  // Cahn-Hilliard x-face fluxes either side of a plane, force and/or Cahn-Hilliard on a list of x-planes,
  // and the re-projection + displacement + interpolation of the plane-crossing populations
  int (*le_interp)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200LeInterp &, int cubic,
		   int ncomp, int zext, double * data);
  int (*le_grad_planes)(cudaStream_t, const Lb200Geom &, int ne, int ntrip, const int * trip, const double * phi,
			double * grad, double * delsq);
  int (*le_force_prep)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200SymmDev &, const double * phi,
		       const double * grad, const double * delsq, double * term, double * fcor);
  int (*le_ch_prep)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200SymmDev &, const double * phi,
		    const double * delsq, const double * u, const char * status, double * chx);
  // le_force_prep + le_ch_prep (fast build: one sweep over the planes + a fixed-order sum of block partials)
  int (*le_prep_both)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200SymmDev &, const double * phi,
		      const double * grad, const double * delsq, const double * u, const char * status, double * term,
		      double * fcor, double * chx);
  int (*le_force_ch)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200SymmDev &, const Lb200LeFix &,
		     int nx, const int * xlist, int do_force, int do_ch, int accumulate, const double * phi,
		     const double * grad, const double * delsq, const double * u, const char * status,
		     const double * fcor, const double * chx, double * force, double * phinew);
  int (*le_lb_bc)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200LeFix &, const Lb200ModelDev *,
		  int ndist, double * f, double * sbuf);
  // liquid crystal (lb200_lc.cuh): 7-point gradient arrays of nf components on [1-ne, N+ne]^3; the stress from the
  // 7-point star of Q; force from the stored stress and / or the Beris-Edwards update in one sweep
  // (nf < 0: -nf components with the 2d_5pt_fluid stencil, on the plane kc = 1 only)
  int (*grad7)(cudaStream_t, const Lb200Geom &, int ne, int nf, const double * field, double * grad, double * delsq);
  int (*lc_stress)(cudaStream_t, const Lb200Geom &, const Lb200LcDev &, int nex, int ne, const double * q, double * str);
  int (*lc_force_be)(cudaStream_t, const Lb200Geom &, const Lb200LcDev &, int do_force, int do_be, int accumulate,
		     const double * q, const double * str, const double * u, double * force, double * qnew);
  // FP32 storage of the D3Q19 distributions inside lb200_step (LB200_KNOB_F32): the arrays hold float(f_p - w_p),
  // arithmetic stays FP64.  to_f32 != 0: f32 <- f64, else f64 <- f32 (every site of the allocation);
  // collide_f32: pull-stream (periodic images from the interior) + collision, f32 in / f32 out
  int (*f_convert)(cudaStream_t, const Lb200Geom &, int to_f32, double * f64, float * f32);
  int (*collide_f32)(cudaStream_t, const Lb200Geom &, const Lb200CollideDev &, const float * fsrc, float * fdst,
		     const double * force, double * rho, double * u);
  // the whole binary-fluid time step in ONE sweep (lb200_fused.cuh): phi sector + pull-stream + collision of the
  // same plane, the force never stored (force / rho / grad / delsq arrays written only when !skip_diag).  u_in and
  // u_out must be different buffers.  Returns 0 without launching when this build has no such kernel (strict
  // mode, compensated Cahn-Hilliard update): the caller then runs phi_sector + collide.  le_xb > 0 (Lees-Edwards planes
  // every le_xb x-planes): nothing is stored for x-planes le_x0 .. le_x0 + 3 (+ m le_xb), the patch kernels produce them.
  int (*step_fused)(cudaStream_t, const Lb200Geom &, const Lb200SymmDev &, const Lb200CollideDev &,
		    const double * phi, const double * u_in, const double * fsrc, double * fdst, double * grad,
		    double * delsq, double * force, double * phinew, double * rho, double * u_out, int le_x0, int le_xb);
  // cahn_hilliard_options_conserve 2: result[0] = compensated sum of phi over the fluid interior sites of this GPU in a
  // fixed order, result[1] = their number; partial: 2*psum_blocks + 1 doubles of scratch (zero before the first call).
  // phi_sum_ranks: all[2r], all[2r+1] of every rank -> total[0..1] in rank order.  phi_subtract: phi -= (total[0] - phi0)/total[1]
  int (*phi_sum)(cudaStream_t, const Lb200Geom &, const double * phi, const char * status, double * partial, double * result);
  int (*phi_sum_ranks)(cudaStream_t, const double * all, int nranks, double * total);
  int (*phi_subtract)(cudaStream_t, const Lb200Geom &, const double * total, double phi0, const char * status, double * phi);
  // y / z periodic images of up to three arrays on a list of x-planes (the one-kernel step with Lees-Edwards
  // planes: the patched planes' images)
  int (*le_yz_images)(cudaStream_t, const Lb200Geom &, int nx, const int * xlist, double * d0, int ncomp0, int depth0,
		      double * d1, int ncomp1, int depth1, double * d2, int ncomp2, int depth2);
  // field_leesedwards(phi) + hydro_lees_edwards(u) in one launch
  int (*le_interp_both)(cudaStream_t, const Lb200Geom &, const Lb200LeDev &, const Lb200LeInterp & cubic,
			const Lb200LeInterp & linear, int zext, double * phi, double * u);
  // signal() for two flag pairs with their own values, one launch
  int (*signal2)(cudaStream_t, unsigned int * a, unsigned int * b, unsigned int vab, unsigned int * c, unsigned int * d, unsigned int vcd);
  int psum_blocks;
};

extern const Lb200Kernels lb200_kernels_fast;
extern const Lb200Kernels lb200_kernels_strict;
