"""the drop-in boundary, proven with the reference's OWN callers.

integration/Makefile (run where the reference source tree exists; outputs travel to the GPU box like oracle/_ref) links
the reference's unmodified objects -- src/ludwig.c, the run-time set-up, statistics, I/O; tests/unit/*.c -- with
integration/ludwig_b200_shim.c: every call they make to a hot-path entry point (lb_collide, lb_propagation, lb_halo,
field_halo, field_grad_compute, phi_force_calculation, phi_cahn_hilliard, hydro_*_zero, the *_memcpy family ...) is
re-routed by ld --wrap to the C-ABI of libludwig_b200.so.  Here:

  * Ludwig_b200.exe runs regression-style input files (the keys the reference's own parser reads) on the GPU, and its
    log -- the reference's own statistics code printing what it copied back from the device -- is compared with the
    log of Ludwig_soa.exe (the same objects without the shim, i.e. the reference itself, run on the host cores of the same
    box) under the rules of the reference's tests/test-diff.sh + tests/awk-fp-diff.sh: version / timer / compiler lines
    dropped, numbers equal within 1e-12 absolute;
  * unit_b200.exe runs the reference's unit-test suites of the hot path (tests/unit/test_prop.c, test_lb_data.c,
    test_field.c, test_hydro.c, ... compiled unchanged) with the library underneath: their own assert()s are the check.
"""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_ref")

# input files in the reference's own format (key value), restated here: the configuration of
# tests/regression/d3q19-short/serial-spin-fd1.inp (BASELINE config 2 at 64^3) and relatives
SPINODAL = {
    "N_cycles": "10", "size": "64_64_64", "viscosity": "0.00625", "ghost_modes": "off",
    "free_energy": "symmetric", "A": "-0.00625", "B": "0.00625", "K": "0.004", "phi0": "0.0",
    "phi_initialisation": "spinodal", "mobility": "1.25", "fd_gradient_calculation": "3d_27pt_fluid",
    "fd_advection_scheme_order": "1", "colloid_init": "no_colloids", "boundary_walls": "0_0_0",
    "periodicity": "1_1_1", "freq_statistics": "10", "config_at_end": "no", "random_seed": "8361235",
}
CASES = {
    "spinodal_order1": SPINODAL,                                                   # serial-spin-fd1
    "spinodal_order3_trt": dict(SPINODAL, size="32_48_40", fd_advection_scheme_order="3", N_cycles="20", freq_statistics="5",
                                lb_relaxation_scheme="trt", force="0.00001_0.0_-0.00002"),
    "spinodal_7pt_bgk": dict(SPINODAL, size="32_32_32", fd_gradient_calculation="3d_7pt_fluid", fd_advection_scheme_order="2",
                             lb_relaxation_scheme="bgk", N_cycles="10", freq_statistics="5"),
    "single_fluid": {"N_cycles": "20", "size": "32_32_32", "viscosity": "0.1", "free_energy": "none",
                     "distribution_initialisation": "3d_uniform_u", "distribution_uniform_u": "0.002_0.003_0.004",
                     "colloid_init": "no_colloids", "periodicity": "1_1_1", "freq_statistics": "10", "config_at_end": "no",
                     "boundary_walls": "0_0_0"},
    "lees_edwards": {"N_cycles": "10", "size": "32_32_32", "viscosity": "0.1", "free_energy": "symmetric", "A": "-0.0625",
                     "B": "0.0625", "K": "0.04", "phi0": "0.0", "phi_initialisation": "spinodal", "mobility": "0.15",
                     "fd_gradient_calculation": "3d_27pt_fluid", "fd_advection_scheme_order": "3", "colloid_init": "no_colloids",
                     "periodicity": "1_1_1", "freq_statistics": "10", "config_at_end": "no", "N_LE_plane": "2",
                     "LE_plane_vel": "0.05", "LE_init_profile": "1", "random_seed": "7361237"},   # serial-le3d-st7
    # a 2-d run: the lattice is thinner than the halo (the configuration of tests/regression/d3q19/pmpi08-le2d-fd1, 100 of its steps)
    "lees_edwards_2d": {"N_cycles": "100", "size": "64_64_1", "viscosity": "0.1", "free_energy": "symmetric", "A": "-0.0625",
                        "B": "0.0625", "K": "0.04", "phi0": "0.0", "phi_initialisation": "spinodal", "mobility": "0.15",
                        "fd_gradient_calculation": "3d_27pt_fluid", "fd_advection_scheme_order": "3", "colloid_init": "no_colloids",
                        "periodicity": "1_1_1", "freq_statistics": "50", "config_at_end": "no", "N_LE_plane": "2",
                        "LE_plane_vel": "0.05", "LE_init_profile": "1", "random_seed": "-7361237"},
    # liquid crystal: Landau-de Gennes Q tensor + Beris-Edwards (BASELINE config 4; the configuration of
    # tests/regression/d3q19/pmpi08-chol-s01 at 48 x 32 x 64)
    "cholesteric": {"N_start": "0", "N_cycles": "10", "size": "48_32_64", "reduced_halo": "no", "viscosity": "1.0",
                    "isothermal_fluctuations": "off", "free_energy": "lc_blue_phase", "fd_advection_scheme_order": "3",
                    "fd_gradient_calculation": "3d_7pt_fluid", "lc_a0": "0.01", "lc_gamma": "3.0", "lc_q0": "0.19635",
                    "lc_kappa0": "0.000648456", "lc_kappa1": "0.000648456", "lc_xi": "0.7", "lc_Gamma": "0.5",
                    "lc_q_initialisation": "twist", "lc_q_init_amplitude": "0.333333333333333", "lc_init_redshift": "1.0",
                    "lc_anchoring_method": "two", "lc_wall_anchoring": "normal", "lc_coll_anchoring": "normal",
                    "lc_anchoring_strength_colloid": "0.002593824", "colloid_init": "no_colloids", "periodicity": "1_1_1",
                    "freq_statistics": "10", "config_at_end": "no", "random_seed": "8361235"},
}
# active liquid crystal: lc_activity yes with the constants of tests/regression/d3q19-short/serial-actv-s01.inp
CASES["active_cholesteric"] = dict(CASES["cholesteric"], lc_activity="yes", lc_active_zeta0="0.33333333333333333",
                                   lc_active_zeta1="0.005", size="32_32_32")
# symmetric_lb (two distributions) with planes on a 2-d lattice: tests/regression/d3q19-short/serial-le2d-lb1.inp, 50 of its steps
CASES["lees_edwards_symmetric_lb_2d"] = {
    "N_cycles": "50", "size": "64_64_1", "viscosity": "0.1", "ghost_modes": "off", "free_energy": "symmetric_lb", "A": "-0.0625",
    "B": "0.0625", "K": "0.04", "phi0": "0.0", "phi_initialisation": "spinodal", "mobility": "0.45",
    "fd_gradient_calculation": "3d_27pt_fluid", "colloid_init": "no_colloids", "periodicity": "1_1_1", "freq_statistics": "50",
    "config_at_end": "no", "N_LE_plane": "2", "LE_plane_vel": "0.05", "LE_init_profile": "1", "random_seed": "13"}
# static redshift (lc_init_redshift != 1, no dynamic update)
CASES["cholesteric_redshift"] = dict(CASES["cholesteric"], lc_init_redshift="0.95", size="32_32_32")

# lines tests/test-diff.sh deletes before comparing
DROP = re.compile(r"call\)|calls\)|Welcome|Git commit:|Compiler:|\.\.name:|\.\.version-string:|\.\.options:|Target thread model:|"
                  r"Default threads per block|OpenMP|Note assertions|Timer|user.parameters.from|GPU INFO|SIMD vector|Start time|"
                  r"End time|Halo type|Decomposition|Local domain|Final cell list|Final cell lengths|SVN.revision")
NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")
TOLERANCE = 1.0e-12          # tests/awk-fp-diff.sh


def have(exe):
    return os.path.exists(os.path.join(BIN, exe))


def run_exe(exe, keys, cwd, env=None):
    with open(os.path.join(cwd, "input"), "w") as fh:
        for k, v in keys.items():
            fh.write(f"{k} {v}\n")
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([os.path.join(BIN, exe)], cwd=cwd, capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, (exe, r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


def filtered(log):
    return [ln.rstrip() for ln in log.splitlines() if ln.strip() and not DROP.search(ln)]


def diff_logs(a, b, tol=TOLERANCE):
    """the reference's awk-fp-diff.sh rule: same lines, tokens equal as strings or as numbers within tol (absolute)"""
    la, lb_ = filtered(a), filtered(b)
    bad = []
    if len(la) != len(lb_):
        bad.append(f"line counts differ: {len(la)} vs {len(lb_)}")
    for x, y in zip(la, lb_):
        tx, ty = x.split(), y.split()
        ok = len(tx) == len(ty)
        if ok:
            for p, q in zip(tx, ty):
                if p == q:
                    continue
                if NUM.match(p) and NUM.match(q) and abs(float(p) - float(q)) <= tol:
                    continue
                ok = False
                break
        if not ok:
            bad.append(f"< {x}\n> {y}")
    return bad


@pytest.mark.skipif(not (have("Ludwig_b200.exe") and have("Ludwig_soa.exe")), reason="integration/_ref not built (needs the reference source tree)")
@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("case", list(CASES))
def test_reference_driver_runs_on_the_library(case, math, tmp_path):
    """src/ludwig.c end to end: input parsing, initialisation, H2D, time-step loop through the library, D2H, the
    reference's own statistics -- log equal to the reference's log under the reference's own regression-diff rules"""
    keys = CASES[case]
    d_gpu, d_cpu = tmp_path / "gpu", tmp_path / "cpu"
    d_gpu.mkdir(); d_cpu.mkdir()
    got = run_exe("Ludwig_b200.exe", keys, str(d_gpu), env={"LB200_MATH": math})
    ref = run_exe("Ludwig_soa.exe", keys, str(d_cpu), env={"OMP_NUM_THREADS": "8"})
    assert "Completed cycle" in got and "Ludwig finished normally" in got
    # (fast arithmetic, 2-d symmetric_lb with planes: the total momentum printed after 50 steps is a sum of 4096 terms of
    # either sign, 3.48e-05 in all -- FMA contraction moves its 8th printed digit by one unit, 1e-12 absolute, exactly the
    # reference's tolerance; the strict mode of the same case meets 1e-12)
    tol = 2.0e-12 if (case == "lees_edwards_symmetric_lb_2d" and math == "fast") else TOLERANCE
    bad = diff_logs(ref, got, tol)
    assert not bad, "\n".join(bad[:20])


@pytest.mark.skipif(not have("unit_b200.exe"), reason="integration/_ref not built (needs the reference source tree)")
def test_reference_unit_suites_pass_on_the_library(tmp_path):
    """tests/unit/test_prop.c, test_lb_data.c, test_field.c, test_field_grad.c, test_hydro.c, test_lb_model.c, test_lb_d3q19.c,
    test_phi_ch.c, test_fe_symmetric.c, test_le.c: compiled unchanged, hot-path calls routed to the GPU library"""
    r = subprocess.run([os.path.join(BIN, "unit_b200.exe")], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "the reference's hot-path unit suites passed" in r.stdout
    for suite in ("test_prop", "test_lb_data", "test_field", "test_hydro"):
        assert re.search(rf"PASS\s+\S*{suite}\b", r.stdout), suite


# ---- the reference's own regression inputs (tests/regression/d3q19-short/serial-*.inp, restated as key / value pairs in
# tests/golden/regression_inputs_d3q19_short.json by tools/make_regression_inputs.py) -- a sample of the full sweep
# (tools/regression_sweep.py, profiles/r02_regression_sweep.md: 29 logs equal, 72 explicit refusals, 0 different)
import json

REGRESSION = json.load(open(os.path.join(ROOT, "tests", "golden", "regression_inputs_d3q19_short.json")))
IN_SCOPE = ["serial-actv-s01", "serial-le3d-st1", "serial-le3d-st7", "serial-le2d-lb1", "serial-relx-bp1", "serial-chol-fld", "serial-symm-dr1",
            "serial-dist-3du", "serial-init-bp1", "serial-spin-lb1"]
OUT_OF_SCOPE = {"serial-auto-c01": "colloids are outside this library", "serial-wall-st2": "walls are outside this library",
                "serial-elec-gc1": "porous media are outside this library", "serial-pola-r01": "this free energy is outside this library"}


def run_pairs(exe, pairs, cwd, env=None):
    with open(os.path.join(cwd, "input"), "w") as fh:
        for k, v in pairs:
            fh.write(f"{k} {v}\n")
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([os.path.join(BIN, exe)], cwd=cwd, capture_output=True, text=True, timeout=600, env=e)


@pytest.mark.skipif(not (have("Ludwig_b200.exe") and have("Ludwig_soa.exe")), reason="integration/_ref not built (needs the reference source tree)")
@pytest.mark.parametrize("case", IN_SCOPE)
def test_reference_regression_input_matches(case, tmp_path):
    d_gpu, d_cpu = tmp_path / "gpu", tmp_path / "cpu"
    d_gpu.mkdir(); d_cpu.mkdir()
    got = run_pairs("Ludwig_b200.exe", REGRESSION[case], str(d_gpu), env={"LB200_MATH": "strict"})
    ref = run_pairs("Ludwig_soa.exe", REGRESSION[case], str(d_cpu), env={"OMP_NUM_THREADS": "8"})
    assert ref.returncode == 0 and "Ludwig finished normally" in ref.stdout
    assert got.returncode == 0 and "Ludwig finished normally" in got.stdout, (got.stdout[-1500:], got.stderr[-1500:])
    bad = diff_logs(ref.stdout, got.stdout)
    assert not bad, "\n".join(bad[:20])


@pytest.mark.skipif(not have("Ludwig_b200.exe"), reason="integration/_ref not built (needs the reference source tree)")
@pytest.mark.parametrize("case", list(OUT_OF_SCOPE))
def test_reference_regression_input_outside_the_scope_is_refused(case, tmp_path):
    """walls, colloids, other free energies: an explicit message and the reference's own pe_fatal exit (its serial MPI stub's
    MPI_Abort ends the process with status 0) -- never a finished run with other numbers"""
    got = run_pairs("Ludwig_b200.exe", REGRESSION[case], str(tmp_path), env={"LB200_MATH": "strict"})
    assert "Ludwig finished normally" not in got.stdout and "Completed cycle" not in got.stdout
    assert OUT_OF_SCOPE[case] in (got.stdout + got.stderr)
