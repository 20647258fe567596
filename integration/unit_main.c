/* unit_main.c -- runs the reference's OWN unit-test suites of the hot path (tests/unit/*.c, compiled unchanged; the
 * suite functions are declared in the reference's tests/unit/tests.h) in place of the reference's tests.c driver, which
 * runs all 96 suites, many of them about subsystems outside SURVEY 8 and some needing input files of that directory.
 * A failing check is an assert() inside the reference's test code: the process aborts. */

#include <stdio.h>
#include <string.h>

#include "pe.h"
#include "tests.h"

int main(int argc, char ** argv) {

  MPI_Init(&argc, &argv);

  /* the fixtures that pin this path in the reference's own tests */
  test_lb_prop_suite();        /* tests/unit/test_prop.c:80-258: propagation + halo, bit exact */
  test_lb_data_suite();        /* tests/unit/test_lb_data.c:299-353, 672-758: halo (full / reduced), io */
  test_field_suite();          /* tests/unit/test_field.c:425-446: field halo */
  test_field_grad_suite();
  test_hydro_suite();          /* tests/unit/test_hydro.c:132-177: u halo */
  test_lb_model_suite();
  test_lb_d3q19_suite();
  test_phi_ch_suite();
  test_fe_symmetric_suite();
  test_le_suite();

  MPI_Finalize();
  printf("unit_main: the reference's hot-path unit suites passed\n");

  return 0;
}
