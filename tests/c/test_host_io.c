/*
 * test_host_io.c -- on-disk formats through Ludwig's own host function names (include/ludwig_host.h).
 * CPU only (no device context is ever created).  Driven by tests/test_host_io.py:
 *
 *   test_host_io.exe write <nx> <ny> <nz> <nvel> <ndist> <nf> <fieldname> <le_planes> [ascii]
 *       fills lb->f and the field with the tag function below and writes dist-000000007.001-001,
 *       <fieldname>-000000007.001-001 and their metadata into the current directory;
 *   test_host_io.exe read  ... (same arguments)
 *       reads the files of the current directory (written by the REFERENCE) and checks every value against the tags.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ludwig_host.h"

static double tag(int ic, int jc, int kc, int c) { return (double) ((((ic*64 + jc)*64 + kc)*64) + c); }

int main(int argc, char ** argv) {
  pe_t * pe = NULL;
  cs_t * cs = NULL;
  lees_edw_t * le = NULL;
  lb_t * lb = NULL;
  field_t * fld = NULL;
  io_event_t ev = {0};
  if (argc < 10) { printf("usage\n"); return 2; }
  const int write = (strcmp(argv[1], "write") == 0);
  int ntotal[3] = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4])};
  const int nvel = atoi(argv[5]), ndist = atoi(argv[6]), nf = atoi(argv[7]), nplanes = atoi(argv[9]);
  int nlocal[3];

  pe_create(MPI_COMM_WORLD, PE_QUIET, &pe);
  cs_create(pe, &cs);
  cs_nhalo_set(cs, 2);
  cs_ntotal_set(cs, ntotal);
  cs_init(cs);
  cs_nlocal(cs, nlocal);
  { lees_edw_options_t o = {.nplanes = nplanes, .type = LE_SHEAR_TYPE_STEADY, .uy = 0.05}; lees_edw_create(pe, cs, &o, &le); }
  const int ascii = (argc > 10 && strcmp(argv[10], "ascii") == 0);
  { lb_data_options_t o = lb_data_options_ndim_nvel_ndist(3, nvel, ndist);
    if (ascii) o.iodata.input = o.iodata.output = io_options_with_format(IO_MODE_MPIIO, IO_RECORD_ASCII);
    lb_data_create(pe, cs, &o, &lb); }
  { field_options_t o = field_options_ndata_nhalo(nf, 2);
    if (ascii) o.iodata.input = o.iodata.output = io_options_with_format(IO_MODE_MPIIO, IO_RECORD_ASCII);
    field_create(pe, cs, le, argv[8], &o, &fld); }

  if (write) {
    for (int ic = 1; ic <= nlocal[X]; ic++)
      for (int jc = 1; jc <= nlocal[Y]; jc++)
	for (int kc = 1; kc <= nlocal[Z]; kc++) {
	  int index = cs_index(cs, ic, jc, kc);
	  for (int n = 0; n < ndist; n++)
	    for (int p = 0; p < nvel; p++) lb_f_set(lb, index, p, n, tag(ic, jc, kc, n*nvel + p));
	  for (int n = 0; n < nf; n++) fld->data[addr_rank1(fld->nsites, nf, index, n)] = 0.5 + tag(ic, jc, kc, n);
	}
    lb_io_write(lb, 7, &ev);
    field_io_write(fld, 7, &ev);
    printf("PASS wrote\n");
  }
  else {
    int nbad = 0;
    lb_io_read(lb, 7, &ev);
    field_io_read(fld, 7, &ev);
    for (int ic = 1; ic <= nlocal[X]; ic++)
      for (int jc = 1; jc <= nlocal[Y]; jc++)
	for (int kc = 1; kc <= nlocal[Z]; kc++) {
	  int index = cs_index(cs, ic, jc, kc);
	  for (int n = 0; n < ndist; n++)
	    for (int p = 0; p < nvel; p++) {
	      double f;
	      lb_f(lb, index, p, n, &f);
	      if (f != tag(ic, jc, kc, n*nvel + p)) nbad++;
	    }
	  for (int n = 0; n < nf; n++) if (fld->data[addr_rank1(fld->nsites, nf, index, n)] != 0.5 + tag(ic, jc, kc, n)) nbad++;
	}
    if (nbad) { printf("FAIL %d values differ\n", nbad); return 1; }
    printf("PASS read\n");
  }
  field_free(fld); lb_free(lb); lees_edw_free(le); cs_free(cs); pe_free(pe);
  return 0;
}
