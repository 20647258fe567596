/*
 * test_host_api.c -- unit tests of the hot path THROUGH LUDWIG'S OWN HOST FUNCTION NAMES
 * (include/ludwig_host.h, implemented on libludwig_b200.so), written in the style of the
 * reference's unit tests and reproducing their known-answer checks:
 *
 *   test_lb_prop_source   : tests/unit/test_prop.c:164-258 (every population has moved exactly one
 *                           site from its expected global source after halo + propagation; exact)
 *   test_lb_halo_fill     : tests/unit/test_lb_data.c:672-758 (halo known-answer fill / check)
 *   test_field_halo       : tests/unit/test_field.c:744-800 (field halo known answer)
 *   test_collide_*, test_binary_step : the calls of src/ludwig.c:528-860 in the reference's order,
 *                           compared with the CPU oracle (oracle/lb_oracle.c; LB200_MATH=strict => exact)
 *
 * Build and run (GPU box): see tests/test_host_c.py.  Exit code 0 and "PASS" lines on success.
 */

#include <assert.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ludwig_host.h"
#include "lb_oracle.h"

#define test_assert(x) do { if (!(x)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #x); exit(1); } } while (0)

static pe_t * pe = NULL;

/* reference tests/unit/test_prop.c:164-258 */
static void test_lb_prop_source(int nvel, lb_halo_enum_t halo) {
  cs_t * cs = NULL;
  lb_t * lb = NULL;
  int ntotal[3] = {8, 6, 10};
  int nlocal[3];
  lb_data_options_t opts = lb_data_options_ndim_nvel_ndist(3, nvel, 1);

  cs_create(pe, &cs);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  opts.halo = halo;
  lb_data_create(pe, cs, &opts, &lb);

  /* f_p(r) = unique tag of the global site */
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	double tag = 1.0*(ntotal[Y]*ntotal[Z]*ic + ntotal[Z]*jc + kc);
	for (int p = 0; p < nvel; p++) lb_f_set(lb, index, p, 0, tag + 0.001*p);
      }
  lb_memcpy(lb, tdpMemcpyHostToDevice);
  lb_halo(lb);
  lb_propagation(lb);
  lb_memcpy(lb, tdpMemcpyDeviceToHost);

  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	for (int p = 0; p < nvel; p++) {
	  double f_actual, f_expect;
	  int isource = ic - lb->model.cv[p][X];
	  int jsource = jc - lb->model.cv[p][Y];
	  int ksource = kc - lb->model.cv[p][Z];
	  if (isource == 0) isource += ntotal[X];
	  if (isource == ntotal[X] + 1) isource = 1;
	  if (jsource == 0) jsource += ntotal[Y];
	  if (jsource == ntotal[Y] + 1) jsource = 1;
	  if (ksource == 0) ksource += ntotal[Z];
	  if (ksource == ntotal[Z] + 1) ksource = 1;
	  f_expect = 1.0*(ntotal[Y]*ntotal[Z]*isource + ntotal[Z]*jsource + ksource) + 0.001*p;
	  lb_f(lb, index, p, 0, &f_actual);
	  test_assert(fabs(f_actual - f_expect) < DBL_EPSILON);
	}
      }
  lb_free(lb);
  cs_free(cs);
  printf("PASS test_lb_prop_source nvel=%d halo=%d\n", nvel, (int) halo);
}

/* reference tests/unit/test_lb_data.c:672-758: every halo site holds its periodic image */
static void test_lb_halo_fill(void) {
  cs_t * cs = NULL;
  lb_t * lb = NULL;
  int ntotal[3] = {6, 7, 5};
  int nlocal[3];
  lb_data_options_t opts = lb_data_options_ndim_nvel_ndist(3, 19, 1);

  cs_create(pe, &cs);
  cs_nhalo_set(cs, 2);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  lb_data_create(pe, cs, &opts, &lb);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	for (int p = 0; p < 19; p++)
	  lb_f_set(lb, cs_index(cs, ic, jc, kc), p, 0, 1.0*(100*ic + 10*jc + kc) + 0.01*p);
  lb_memcpy(lb, tdpMemcpyHostToDevice);
  lb_halo(lb);
  lb_memcpy(lb, tdpMemcpyDeviceToHost);
  for (int ic = 0; ic <= nlocal[X] + 1; ic++)
    for (int jc = 0; jc <= nlocal[Y] + 1; jc++)
      for (int kc = 0; kc <= nlocal[Z] + 1; kc++) {
	int is = (ic == 0) ? nlocal[X] : (ic == nlocal[X] + 1) ? 1 : ic;
	int js = (jc == 0) ? nlocal[Y] : (jc == nlocal[Y] + 1) ? 1 : jc;
	int ks = (kc == 0) ? nlocal[Z] : (kc == nlocal[Z] + 1) ? 1 : kc;
	for (int p = 0; p < 19; p++) {
	  double f;
	  lb_f(lb, cs_index(cs, ic, jc, kc), p, 0, &f);
	  test_assert(f == 1.0*(100*is + 10*js + ks) + 0.01*p);
	}
      }
  lb_free(lb);
  cs_free(cs);
  printf("PASS test_lb_halo_fill\n");
}

/* reference tests/unit/test_field.c:744-800 */
static void test_field_halo(void) {
  cs_t * cs = NULL;
  lees_edw_t * le = NULL;
  field_t * phi = NULL;
  hydro_t * hydro = NULL;
  lees_edw_options_t leopts = {0};
  int ntotal[3] = {6, 4, 8};
  int nlocal[3], nhalo = 2;
  field_options_t fopts = field_options_ndata_nhalo(1, 2);
  hydro_options_t hopts = hydro_options_default();

  cs_create(pe, &cs);
  cs_nhalo_set(cs, nhalo);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  lees_edw_create(pe, cs, &leopts, &le);
  field_create(pe, cs, le, "phi", &fopts, &phi);
  hydro_create(pe, cs, le, &hopts, &hydro);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	double u[3] = {1.0*ic, 2.0*jc, 3.0*kc};
	field_scalar_set(phi, index, 1.0*(100*ic + 10*jc + kc));
	hydro_u_set(hydro, index, u);
      }
  field_memcpy(phi, tdpMemcpyHostToDevice);
  hydro_memcpy(hydro, tdpMemcpyHostToDevice);
  field_halo(phi);
  hydro_u_halo(hydro);
  field_memcpy(phi, tdpMemcpyDeviceToHost);
  hydro_memcpy(hydro, tdpMemcpyDeviceToHost);
  for (int ic = 1 - nhalo; ic <= nlocal[X] + nhalo; ic++)
    for (int jc = 1 - nhalo; jc <= nlocal[Y] + nhalo; jc++)
      for (int kc = 1 - nhalo; kc <= nlocal[Z] + nhalo; kc++) {
	int is = (ic < 1) ? ic + nlocal[X] : (ic > nlocal[X]) ? ic - nlocal[X] : ic;
	int js = (jc < 1) ? jc + nlocal[Y] : (jc > nlocal[Y]) ? jc - nlocal[Y] : jc;
	int ks = (kc < 1) ? kc + nlocal[Z] : (kc > nlocal[Z]) ? kc - nlocal[Z] : kc;
	double p, u[3];
	field_scalar(phi, cs_index(cs, ic, jc, kc), &p);
	hydro_u(hydro, cs_index(cs, ic, jc, kc), u);
	test_assert(p == 1.0*(100*is + 10*js + ks));
	test_assert(u[X] == 1.0*is && u[Y] == 2.0*js && u[Z] == 3.0*ks);
      }
  hydro_free(hydro);
  field_free(phi);
  lees_edw_free(le);
  cs_free(cs);
  printf("PASS test_field_halo\n");
}

static double frand(unsigned int * s) { *s = 1664525u*(*s) + 1013904223u; return (*s >> 8)*(1.0/16777216.0); }

/* whole binary-fluid time steps through the reference's entry points vs the oracle */
static void test_binary_step(int order, int nsteps, int strict, int conserve) {
  cs_t * cs = NULL;
  physics_t * phys = NULL;
  lees_edw_t * le = NULL;
  lb_t * lb = NULL;
  hydro_t * hydro = NULL;
  map_t * map = NULL;
  field_t * phi = NULL;
  field_grad_t * phi_grad = NULL;
  fe_symm_t * fe = NULL;
  pth_t * pth = NULL;
  phi_ch_t * pch = NULL;
  int ntotal[3] = {12, 10, 34};
  int nlocal[3], ns;
  unsigned int seed = 12345;
  double fbody[3] = {1.0e-6, -2.0e-6, 5.0e-7};
  const double zero[3] = {0.0, 0.0, 0.0};

  cs_create(pe, &cs);
  cs_nhalo_set(cs, 2);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  cs_nsites(cs, &ns);
  physics_create(pe, &phys);
  physics_eta_shear_set(phys, 0.00625);
  physics_eta_bulk_set(phys, 0.00625);
  physics_fbody_set(phys, fbody);
  physics_mobility_set(phys, 1.25);
  { lees_edw_options_t o = {0}; lees_edw_create(pe, cs, &o, &le); }
  { lb_data_options_t o = lb_data_options_ndim_nvel_ndist(3, 19, 1); lb_data_create(pe, cs, &o, &lb); }
  { hydro_options_t o = hydro_options_default(); hydro_create(pe, cs, le, &o, &hydro); }
  { map_options_t o = map_options_default(); map_create(pe, cs, &o, &map); }
  { field_options_t o = field_options_ndata_nhalo(1, 2); field_create(pe, cs, le, "phi", &o, &phi); }
  field_grad_create(pe, phi, 2, &phi_grad);
  field_grad_set(phi_grad, grad_3d_27pt_fluid_d2, NULL);
  fe_symm_create(pe, cs, phi, phi_grad, &fe);
  { fe_symm_param_t p = {.a = -0.00625, .b = 0.00625, .kappa = 0.004}; fe_symm_param_set(fe, p); }
  { phi_ch_info_t o = {0}; o.conserve = conserve; phi_ch_create(pe, cs, le, &o, &pch); }   /* cahn_hilliard_options_conserve */
  pth_create(pe, cs, FE_FORCE_METHOD_STRESS_DIVERGENCE, &pth);
  advection_order_set(order);

  lb_init_rest_f(lb, 1.0);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	field_scalar_set(phi, cs_index(cs, ic, jc, kc), 0.1*(frand(&seed) - 0.5));

  /* oracle copies (the host layout IS the canonical layout) */
  orc_geom_t g = {{nlocal[X], nlocal[Y], nlocal[Z]}, 2, {1, 1, 1}};
  orc_model_t model;
  orc_collide_param_t ocp = {ORC_RELAX_M10, 1.0, 0.00625, 0.00625, {fbody[0], fbody[1], fbody[2]}};
  orc_symm_param_t osp = {-0.00625, 0.00625, 0.004, 1.25, {0.0, 0.0, 0.0}, order, conserve};
  double * of = malloc(sizeof(double)*19*ns), * ophi = malloc(sizeof(double)*ns);
  double * ou = calloc(3*ns, sizeof(double)), * orho = calloc(ns, sizeof(double)), * oforce = calloc(3*ns, sizeof(double));
  double * ograd = calloc(3*ns, sizeof(double)), * odelsq = calloc(ns, sizeof(double));
  orc_model_create(19, &model);
  memcpy(of, lb->f, sizeof(double)*19*ns);
  memcpy(ophi, phi->data, sizeof(double)*ns);

  /* src/ludwig.c:501-506 then the loop body :528-860 */
  map_memcpy(map, tdpMemcpyHostToDevice);
  lb_memcpy(lb, tdpMemcpyHostToDevice);
  field_memcpy(phi, tdpMemcpyHostToDevice);
  for (int n = 0; n < nsteps; n++) {
    hydro_f_zero(hydro, zero);
    field_halo(phi);
    field_grad_compute(phi_grad);
    phi_force_calculation(pe, cs, le, NULL, pth, (fe_t *) fe, map, phi, hydro);
    phi_cahn_hilliard(pch, (fe_t *) fe, phi, hydro, map, NULL);
    hydro_u_zero(hydro, zero);
    lb_collide(lb, hydro, map, NULL, (fe_t *) fe, NULL);
    lb_halo(lb);
    lb_propagation(lb);
  }
  lb_memcpy(lb, tdpMemcpyDeviceToHost);
  field_memcpy(phi, tdpMemcpyDeviceToHost);
  hydro_memcpy(hydro, tdpMemcpyDeviceToHost);
  field_grad_memcpy(phi_grad, tdpMemcpyDeviceToHost);

  orc_step(&g, &model, &ocp, &osp, 1, 0, nsteps, of, ophi, ou, orho, oforce, ograd, odelsq);

  double fed = 0.0;
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	double fe1;
	for (int p = 0; p < 19; p++) {
	  double a = lb->f[LB_ADDR(ns, 1, 19, index, 0, p)], b = of[(size_t) p*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.34 + 1e-14);
	}
	{
	  double a = phi->data[index], b = ophi[index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.05 + 1e-14);
	}
	for (int ia = 0; ia < 3; ia++) {
	  double a = hydro->u->data[addr_rank1(ns, 3, index, ia)], b = ou[(size_t) ia*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-14);
	  a = phi_grad->grad[addr_rank2(ns, 1, 3, index, 0, ia)]; b = ograd[(size_t) ia*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.05 + 1e-14);
	}
	fe_symm_fed(fe, index, &fe1);
	fed += fe1;
	{
	  /* fe_symm_str (src/symmetric.c:333-362): symmetric, trace = 3 p0 + kappa |grad phi|^2 */
	  double s[3][3], gp[3], gg = 0.0;
	  fe_symm_str(fe, index, s);
	  field_grad_scalar_grad(phi_grad, index, gp);
	  for (int ia = 0; ia < 3; ia++) gg += gp[ia]*gp[ia];
	  /* (kappa g_a) g_b and (kappa g_b) g_a round differently, as in the reference: symmetric to an ulp */
	  test_assert(fabs(s[0][1] - s[1][0]) <= 4e-16*fabs(s[0][1]) && fabs(s[0][2] - s[2][0]) <= 4e-16*fabs(s[0][2])
		      && fabs(s[1][2] - s[2][1]) <= 4e-16*fabs(s[1][2]));
	  test_assert(fabs((s[0][0] - 0.004*gp[0]*gp[0]) - (s[2][2] - 0.004*gp[2]*gp[2])) <= 1e-18 + 1e-12*fabs(s[0][0]));
	  (void) gg;
	}
      }
  printf("PASS test_binary_step order=%d nsteps=%d conserve=%d %s  (free energy %.10e)\n", order, nsteps, conserve,
	 strict ? "bit-exact" : "within tolerance", fed);

  free(of); free(ophi); free(ou); free(orho); free(oforce); free(ograd); free(odelsq);
  pth_free(pth); phi_ch_free(pch); fe_symm_free(fe); field_grad_free(phi_grad); field_free(phi);
  map_free(&map); hydro_free(hydro); lb_free(lb); lees_edw_free(le); physics_free(phys); cs_free(cs);
}

/* `free_energy symmetric_lb`: two distributions, the calls of src/ludwig.c:528-860 with ndist == 2 */
static void test_symmetric_lb(int nvel, int strict) {
  cs_t * cs = NULL;
  physics_t * phys = NULL;
  lees_edw_t * le = NULL;
  lb_t * lb = NULL;
  hydro_t * hydro = NULL;
  map_t * map = NULL;
  field_t * phi = NULL;
  field_grad_t * phi_grad = NULL;
  fe_symm_t * fe = NULL;
  int ntotal[3] = {8, 6, 34};
  int nlocal[3], ns, nsteps = 5;
  double fbody[3] = {1.0e-6, 2.0e-6, -1.0e-6};
  const double zero[3] = {0.0, 0.0, 0.0};
  unsigned int seed = 77;

  cs_create(pe, &cs);
  cs_nhalo_set(cs, 1);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  cs_nsites(cs, &ns);
  physics_create(pe, &phys);
  physics_eta_shear_set(phys, 0.00625);
  physics_eta_bulk_set(phys, 0.00625);
  physics_fbody_set(phys, fbody);
  physics_mobility_set(phys, 3.75);
  { lees_edw_options_t o = {0}; lees_edw_create(pe, cs, &o, &le); }
  { lb_data_options_t o = lb_data_options_ndim_nvel_ndist(3, nvel, 2); lb_data_create(pe, cs, &o, &lb); }
  { hydro_options_t o = hydro_options_default(); hydro_create(pe, cs, le, &o, &hydro); }
  { map_options_t o = map_options_default(); map_create(pe, cs, &o, &map); }
  { field_options_t o = field_options_ndata_nhalo(1, 1); field_create(pe, cs, le, "phi", &o, &phi); }
  field_grad_create(pe, phi, 2, &phi_grad);
  field_grad_set(phi_grad, grad_3d_27pt_fluid_d2, NULL);
  fe_symm_create(pe, cs, phi, phi_grad, &fe);
  { fe_symm_param_t p = {.a = -0.00625, .b = 0.00625, .kappa = 0.004}; fe_symm_param_set(fe, p); }

  lb_init_rest_f(lb, 1.0);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	field_scalar_set(phi, cs_index(cs, ic, jc, kc), 0.1*(frand(&seed) - 0.5));
  phi_lb_from_field(phi, lb);                        /* src/ludwig.c:402 */

  orc_geom_t g = {{nlocal[X], nlocal[Y], nlocal[Z]}, 1, {1, 1, 1}};
  orc_model_t model;
  orc_collide_param_t ocp = {ORC_RELAX_M10, 1.0, 0.00625, 0.00625, {fbody[0], fbody[1], fbody[2]}};
  orc_symm_param_t osp = {-0.00625, 0.00625, 0.004, 3.75, {0.0, 0.0, 0.0}, 1};
  double * of = malloc(sizeof(double)*2*nvel*ns), * ophi = calloc(ns, sizeof(double));
  double * ou = calloc(3*ns, sizeof(double)), * oforce = calloc(3*ns, sizeof(double));
  double * ograd = calloc(3*ns, sizeof(double)), * odelsq = calloc(ns, sizeof(double));
  orc_model_create(nvel, &model);
  memcpy(of, lb->f, sizeof(double)*2*nvel*ns);

  lb_memcpy(lb, tdpMemcpyHostToDevice);
  for (int n = 0; n < nsteps; n++) {
    hydro_f_zero(hydro, zero);
    phi_lb_to_field(phi, lb);
    field_halo(phi);
    field_grad_compute(phi_grad);
    hydro_u_zero(hydro, zero);
    lb_collide(lb, hydro, map, NULL, (fe_t *) fe, NULL);
    lb_halo(lb);
    lb_propagation(lb);
  }
  lb_memcpy(lb, tdpMemcpyDeviceToHost);
  field_memcpy(phi, tdpMemcpyDeviceToHost);
  hydro_memcpy(hydro, tdpMemcpyDeviceToHost);

  orc_step_lb2(&g, &model, &ocp, &osp, 0, nsteps, of, ophi, ou, oforce, ograd, odelsq);

  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	for (int n = 0; n < 2; n++)
	  for (int p = 0; p < nvel; p++) {
	    double a = lb->f[LB_ADDR(ns, 2, nvel, index, n, p)], b = of[(size_t) (n*nvel + p)*ns + index];
	    if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.34 + 1e-14);
	  }
	{
	  double a = phi->data[index], b = ophi[index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.05 + 1e-14);
	}
	for (int ia = 0; ia < 3; ia++) {
	  double a = hydro->u->data[addr_rank1(ns, 3, index, ia)], b = ou[(size_t) ia*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-14);
	}
      }
  printf("PASS test_symmetric_lb nvel=%d nsteps=%d %s\n", nvel, nsteps, strict ? "bit-exact" : "within tolerance");
  free(of); free(ophi); free(ou); free(oforce); free(ograd); free(odelsq);
  fe_symm_free(fe); field_grad_free(phi_grad); field_free(phi);
  map_free(&map); hydro_free(hydro); lb_free(lb); lees_edw_free(le); physics_free(phys); cs_free(cs);
}

/* Lees-Edwards sheared binary fluid through the reference's entry points (src/ludwig.c:528-860 with planes:
 * tests/regression/d3q19-short/serial-le3d-st7.inp at a smaller size) vs the Lees-Edwards oracle */
static void test_lees_edwards_step(int order, int nplanes, int nsteps, int strict) {
  cs_t * cs = NULL;
  physics_t * phys = NULL;
  lees_edw_t * le = NULL;
  lb_t * lb = NULL;
  hydro_t * hydro = NULL;
  map_t * map = NULL;
  field_t * phi = NULL;
  field_grad_t * phi_grad = NULL;
  fe_symm_t * fe = NULL;
  pth_t * pth = NULL;
  phi_ch_t * pch = NULL;
  int ntotal[3] = {16, 12, 10};
  int nlocal[3], ns, nsle;
  unsigned int seed = 777;
  const double zero[3] = {0.0, 0.0, 0.0};
  const double uy = 0.05, eta = 0.1;

  cs_create(pe, &cs);
  cs_nhalo_set(cs, 2);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  cs_nsites(cs, &ns);
  physics_create(pe, &phys);
  physics_eta_shear_set(phys, eta);
  physics_eta_bulk_set(phys, eta);
  physics_mobility_set(phys, 0.15);
  physics_control_init_time(phys, 0, nsteps);
  { lees_edw_options_t o = {.nplanes = nplanes, .type = LE_SHEAR_TYPE_STEADY, .nt0 = 0, .uy = uy}; lees_edw_create(pe, cs, &o, &le); }
  lees_edw_nsites(le, &nsle);
  test_assert(nsle == (nlocal[X] + 4 + 4*nplanes)*(nlocal[Y] + 4)*(nlocal[Z] + 4));
  test_assert(lees_edw_plane_location(le, 0) == nlocal[X]/(2*nplanes));
  test_assert(lees_edw_ic_to_buff(le, lees_edw_plane_location(le, 0), +1) == nlocal[X] + 2 + 1 + 2);
  { lb_data_options_t o = lb_data_options_ndim_nvel_ndist(3, 19, 1); lb_data_create(pe, cs, &o, &lb); }
  { hydro_options_t o = hydro_options_default(); hydro_create(pe, cs, le, &o, &hydro); }
  { map_options_t o = map_options_default(); map_create(pe, cs, &o, &map); }
  { field_options_t o = field_options_ndata_nhalo(1, 2); field_create(pe, cs, le, "phi", &o, &phi); }
  test_assert(phi->nsites == nsle && hydro->nsite == nsle && lb->nsite == ns);
  field_grad_create(pe, phi, 2, &phi_grad);
  field_grad_set(phi_grad, grad_3d_27pt_fluid_d2, NULL);
  fe_symm_create(pe, cs, phi, phi_grad, &fe);
  { fe_symm_param_t p = {.a = -0.0625, .b = 0.0625, .kappa = 0.04}; fe_symm_param_set(fe, p); }
  { phi_ch_info_t o = {0}; phi_ch_create(pe, cs, le, &o, &pch); }
  pth_create(pe, cs, FE_FORCE_METHOD_STRESS_DIVERGENCE, &pth);
  advection_order_set(order);

  lb_le_init_shear_profile(lb, le);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	field_scalar_set(phi, cs_index(cs, ic, jc, kc), 0.1*(frand(&seed) - 0.5));

  orc_geom_t g = {{nlocal[X], nlocal[Y], nlocal[Z]}, 2, {1, 1, 1}, nplanes};
  orc_le_t ole = {uy, 0.0};
  orc_model_t model;
  orc_collide_param_t ocp = {ORC_RELAX_M10, 1.0, eta, eta, {0.0, 0.0, 0.0}};
  orc_symm_param_t osp = {-0.0625, 0.0625, 0.04, 0.15, {0.0, 0.0, 0.0}, order};
  double * of = malloc(sizeof(double)*19*ns), * ophi = malloc(sizeof(double)*nsle);
  double * ou = calloc(3*nsle, sizeof(double)), * orho = calloc(nsle, sizeof(double)), * oforce = calloc(3*nsle, sizeof(double));
  double * ograd = calloc(3*nsle, sizeof(double)), * odelsq = calloc(nsle, sizeof(double));
  orc_model_create(19, &model);
  test_assert(orc_nsites(&g) == nsle && orc_nsites_lb(&g) == ns);
  /* the host initial condition is the reference's (oracle restatement of lb_le_init_shear_profile) */
  orc_le_init_shear_profile(&g, &model, &ole, 1.0, eta, of);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	for (int p = 0; p < 19; p++) test_assert(of[(size_t) p*ns + cs_index(cs, ic, jc, kc)] == lb->f[LB_ADDR(ns, 1, 19, cs_index(cs, ic, jc, kc), 0, p)]);
  memcpy(of, lb->f, sizeof(double)*19*ns);
  memcpy(ophi, phi->data, sizeof(double)*nsle);

  map_memcpy(map, tdpMemcpyHostToDevice);
  lb_memcpy(lb, tdpMemcpyHostToDevice);
  field_memcpy(phi, tdpMemcpyHostToDevice);
  while (physics_control_next_step(phys)) {
    hydro_f_zero(hydro, zero);
    field_halo(phi);
    field_grad_compute(phi_grad);
    phi_force_calculation(pe, cs, le, NULL, pth, (fe_t *) fe, map, phi, hydro);
    phi_cahn_hilliard(pch, (fe_t *) fe, phi, hydro, map, NULL);
    hydro_u_zero(hydro, zero);
    lb_collide(lb, hydro, map, NULL, (fe_t *) fe, NULL);
    lb_data_apply_le_boundary_conditions(lb, le);
    lb_halo(lb);
    lb_propagation(lb);
  }
  test_assert(physics_control_timestep(phys) == nsteps + 1);     /* the loop test advances once more, as in the reference */
  lb_memcpy(lb, tdpMemcpyDeviceToHost);
  field_memcpy(phi, tdpMemcpyDeviceToHost);
  hydro_memcpy(hydro, tdpMemcpyDeviceToHost);

  orc_le_step(&g, &model, &ocp, &osp, &ole, 0, nsteps, of, ophi, ou, orho, oforce, ograd, odelsq);

  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = lees_edw_index(le, ic, jc, kc);
	for (int p = 0; p < 19; p++) {
	  double a = lb->f[LB_ADDR(ns, 1, 19, index, 0, p)], b = of[(size_t) p*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.34 + 1e-14);
	}
	{
	  double a = phi->data[index], b = ophi[index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.05 + 1e-14);
	}
	for (int ia = 0; ia < 3; ia++) {
	  double a = hydro->u->data[addr_rank1(nsle, 3, index, ia)], b = ou[(size_t) ia*nsle + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.03 + 1e-14);
	}
      }
  printf("PASS test_lees_edwards_step order=%d planes=%d nsteps=%d %s\n", order, nplanes, nsteps,
	 strict ? "bit-exact" : "within tolerance");

  free(of); free(ophi); free(ou); free(orho); free(oforce); free(ograd); free(odelsq);
  pth_free(pth); phi_ch_free(pch); fe_symm_free(fe); field_grad_free(phi_grad); field_free(phi);
  map_free(&map); hydro_free(hydro); lb_free(lb); lees_edw_free(le); physics_free(phys); cs_free(cs);
}

/* liquid crystal (Q tensor + Beris-Edwards) through the reference's entry points (src/ludwig.c:528-860 with ludwig->q:
 * tests/regression/d3q19/pmpi08-chol-s01.inp at a smaller size) vs the liquid-crystal oracle */
static void test_liquid_crystal_step(int order, int nsteps, int strict) {
  cs_t * cs = NULL;
  physics_t * phys = NULL;
  lees_edw_t * le = NULL;
  lb_t * lb = NULL;
  hydro_t * hydro = NULL;
  map_t * map = NULL;
  field_t * q = NULL;
  field_grad_t * q_grad = NULL;
  fe_lc_t * fe = NULL;
  beris_edw_t * be = NULL;
  pth_t * pth = NULL;
  int ntotal[3] = {10, 8, 12};
  int nlocal[3], ns;
  unsigned int seed = 4321;
  const double zero[3] = {0.0, 0.0, 0.0};
  const double eta = 0.1;
  fe_lc_param_t p = {.a0 = 0.01, .q0 = 0.19635, .gamma = 3.0, .kappa0 = 0.000648456, .kappa1 = 0.0008, .xi = 0.7,
		     .redshift = 1.0, .rredshift = 1.0, .amplitude0 = 0.333333333333333, .coswt = 1.0};
  beris_edw_param_t bp = {.xi = 0.7, .gamma = 0.5};

  cs_create(pe, &cs);
  cs_nhalo_set(cs, 2);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  cs_nsites(cs, &ns);
  physics_create(pe, &phys);
  physics_eta_shear_set(phys, eta);
  physics_eta_bulk_set(phys, eta);
  { lees_edw_options_t o = {0}; lees_edw_create(pe, cs, &o, &le); }
  { lb_data_options_t o = lb_data_options_ndim_nvel_ndist(3, 19, 1); lb_data_create(pe, cs, &o, &lb); }
  { hydro_options_t o = hydro_options_nhalo(2); hydro_create(pe, cs, le, &o, &hydro); }
  { map_options_t o = map_options_default(); map_create(pe, cs, &o, &map); }
  { field_options_t o = field_options_ndata_nhalo(NQAB, 2); field_create(pe, cs, le, "q", &o, &q); }
  field_grad_create(pe, q, 2, &q_grad);
  field_grad_set(q_grad, grad_3d_7pt_fluid_d2, NULL);
  fe_lc_create(pe, cs, le, q, q_grad, &fe);
  fe_lc_param_set(fe, &p);
  beris_edw_create(pe, cs, le, &be);
  beris_edw_param_set(be, &bp);
  pth_create(pe, cs, FE_FORCE_METHOD_STRESS_DIVERGENCE, &pth);
  advection_order_set(order);

  lb_init_rest_f(lb, 1.0);
  blue_phase_twist_init(cs, &p, q, Z);
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++)
	for (int n = 0; n < NQAB; n++) q->data[addr_rank1(ns, NQAB, cs_index(cs, ic, jc, kc), n)] += 0.02*(frand(&seed) - 0.5);

  orc_geom_t g = {{nlocal[X], nlocal[Y], nlocal[Z]}, 2, {1, 1, 1}, 0};
  orc_model_t model;
  orc_collide_param_t ocp = {ORC_RELAX_M10, 1.0, eta, eta, {0.0, 0.0, 0.0}};
  orc_lc_param_t olc = {.a0 = p.a0, .q0 = p.q0, .gamma = p.gamma, .kappa0 = p.kappa0, .kappa1 = p.kappa1, .xi = p.xi,
			.Gamma = bp.gamma, .redshift = 1.0, .rredshift = 1.0};
  double * of = malloc(sizeof(double)*19*ns), * oq = malloc(sizeof(double)*NQAB*ns);
  double * ou = calloc(3*ns, sizeof(double)), * orho = calloc(ns, sizeof(double)), * oforce = calloc(3*ns, sizeof(double));
  double * ograd = calloc(15*ns, sizeof(double)), * odelsq = calloc(5*ns, sizeof(double));
  orc_model_create(19, &model);
  memcpy(of, lb->f, sizeof(double)*19*ns);
  memcpy(oq, q->data, sizeof(double)*NQAB*ns);

  map_memcpy(map, tdpMemcpyHostToDevice);
  lb_memcpy(lb, tdpMemcpyHostToDevice);
  field_memcpy(q, tdpMemcpyHostToDevice);
  for (int n = 0; n < nsteps; n++) {
    hydro_f_zero(hydro, zero);
    field_halo(q);
    field_grad_compute(q_grad);
    phi_force_calculation(pe, cs, le, NULL, pth, (fe_t *) fe, map, NULL, hydro);
    hydro_u_halo(hydro);
    beris_edw_update(be, (fe_t *) fe, q, q_grad, hydro, NULL, map, NULL);
    hydro_u_zero(hydro, zero);
    lb_collide(lb, hydro, map, NULL, (fe_t *) fe, NULL);
    lb_halo(lb);
    lb_propagation(lb);
  }
  lb_memcpy(lb, tdpMemcpyDeviceToHost);
  field_memcpy(q, tdpMemcpyDeviceToHost);
  hydro_memcpy(hydro, tdpMemcpyDeviceToHost);

  orc_lc_step(&g, &model, &ocp, &olc, order, nsteps, of, oq, ou, orho, oforce, ograd, odelsq);

  double umax = 0.0;
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	for (int p1 = 0; p1 < 19; p1++) {
	  double a = lb->f[LB_ADDR(ns, 1, 19, index, 0, p1)], b = of[(size_t) p1*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.34 + 1e-14);
	}
	for (int n = 0; n < NQAB; n++) {
	  double a = q->data[addr_rank1(ns, NQAB, index, n)], b = oq[(size_t) n*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.34 + 1e-14);
	}
	for (int ia = 0; ia < 3; ia++) {
	  double a = hydro->u->data[addr_rank1(ns, 3, index, ia)], b = ou[(size_t) ia*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-14);
	  if (fabs(b) > umax) umax = fabs(b);
	}
      }
  test_assert(umax > 1e-9);
  printf("PASS test_liquid_crystal_step order=%d nsteps=%d %s\n", order, nsteps, strict ? "bit-exact" : "within tolerance");

  free(of); free(oq); free(ou); free(orho); free(oforce); free(ograd); free(odelsq);
  pth_free(pth); beris_edw_free(be); fe_lc_free(fe); field_grad_free(q_grad); field_free(q);
  map_free(&map); hydro_free(hydro); lb_free(lb); lees_edw_free(le); physics_free(phys); cs_free(cs);
}

/* single-fluid collision + propagation steps, each relaxation scheme */
static void test_single_fluid(int nvel, lb_relaxation_enum_t nrelax, int strict) {
  cs_t * cs = NULL;
  physics_t * phys = NULL;
  lb_t * lb = NULL;
  hydro_t * hydro = NULL;
  map_t * map = NULL;
  int ntotal[3] = {8, 6, 34};
  int nlocal[3], ns, nsteps = 5;
  double fbody[3] = {1.0e-6, 2.0e-6, 3.0e-6};
  const double zero[3] = {0.0, 0.0, 0.0};
  unsigned int seed = 99;

  cs_create(pe, &cs);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  cs_nsites(cs, &ns);
  physics_create(pe, &phys);
  physics_eta_shear_set(phys, 0.05);
  physics_eta_bulk_set(phys, 0.08);
  physics_fbody_set(phys, fbody);
  { lb_data_options_t o = lb_data_options_ndim_nvel_ndist(3, nvel, 1); o.nrelax = nrelax; lb_data_create(pe, cs, &o, &lb); }
  { hydro_options_t o = hydro_options_default(); hydro_create(pe, cs, NULL, &o, &hydro); }
  { map_options_t o = map_options_default(); map_create(pe, cs, &o, &map); }
  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	double u[3] = {0.01*(frand(&seed) - 0.5), 0.01*(frand(&seed) - 0.5), 0.01*(frand(&seed) - 0.5)};
	lb_1st_moment_equilib_set(lb, cs_index(cs, ic, jc, kc), 1.0 + 0.01*frand(&seed), u);
      }
  orc_geom_t g = {{nlocal[X], nlocal[Y], nlocal[Z]}, 1, {1, 1, 1}};
  orc_model_t model;
  orc_collide_param_t ocp = {(int) nrelax, 1.0, 0.05, 0.08, {fbody[0], fbody[1], fbody[2]}};
  double * of = malloc(sizeof(double)*nvel*ns);
  double * ou = calloc(3*ns, sizeof(double)), * orho = calloc(ns, sizeof(double)), * oforce = calloc(3*ns, sizeof(double));
  orc_model_create(nvel, &model);
  memcpy(of, lb->f, sizeof(double)*nvel*ns);

  lb_memcpy(lb, tdpMemcpyHostToDevice);
  for (int n = 0; n < nsteps; n++) {
    hydro_f_zero(hydro, zero);
    hydro_u_zero(hydro, zero);
    lb_collide(lb, hydro, map, NULL, NULL, NULL);
    lb_halo(lb);
    lb_propagation(lb);
  }
  lb_memcpy(lb, tdpMemcpyDeviceToHost);
  hydro_memcpy(hydro, tdpMemcpyDeviceToHost);
  orc_step(&g, &model, &ocp, NULL, 0, 0, nsteps, of, NULL, ou, orho, oforce, NULL, NULL);

  for (int ic = 1; ic <= nlocal[X]; ic++)
    for (int jc = 1; jc <= nlocal[Y]; jc++)
      for (int kc = 1; kc <= nlocal[Z]; kc++) {
	int index = cs_index(cs, ic, jc, kc);
	double rho;
	for (int p = 0; p < nvel; p++) {
	  double a = lb->f[LB_ADDR(ns, 1, nvel, index, 0, p)], b = of[(size_t) p*ns + index];
	  if (strict) test_assert(a == b); else test_assert(fabs(a - b) <= 1e-12*0.34 + 1e-14);
	}
	hydro_rho(hydro, index, &rho);
	if (strict) test_assert(rho == orho[index]); else test_assert(fabs(rho - orho[index]) <= 1e-12);
      }
  printf("PASS test_single_fluid nvel=%d nrelax=%d %s\n", nvel, (int) nrelax, strict ? "bit-exact" : "within tolerance");
  free(of); free(ou); free(orho); free(oforce);
  map_free(&map); hydro_free(hydro); lb_free(lb); physics_free(phys); cs_free(cs);
}

int main(void) {
  const char * math = getenv("LB200_MATH");
  int strict = (math && strcmp(math, "strict") == 0);

  pe_create(MPI_COMM_WORLD, PE_QUIET, &pe);

  test_lb_prop_source(19, LB_HALO_FULL);
  test_lb_prop_source(19, LB_HALO_REDUCED);
  test_lb_prop_source(15, LB_HALO_FULL);
  test_lb_prop_source(27, LB_HALO_REDUCED);
  test_lb_halo_fill();
  test_field_halo();
  test_single_fluid(19, LB_RELAXATION_M10, strict);
  test_single_fluid(19, LB_RELAXATION_TRT, strict);
  test_single_fluid(15, LB_RELAXATION_BGK, strict);
  test_single_fluid(27, LB_RELAXATION_M10, strict);
  test_binary_step(1, 5, strict, 0);
  test_binary_step(3, 5, strict, 0);
  test_binary_step(3, 6, strict, 1);
  test_symmetric_lb(19, strict);
  test_symmetric_lb(15, strict);
  test_lees_edwards_step(3, 2, 6, strict);
  test_lees_edwards_step(1, 1, 6, strict);
  test_liquid_crystal_step(3, 6, strict);
  test_liquid_crystal_step(1, 6, strict);

  pe_free(pe);
  printf("PASS test_host_api (%s)\n", strict ? "strict" : "fast");
  return 0;
}
