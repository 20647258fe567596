"""The reference-side binding (integration/): built here, where the reference source tree exists, by __graft_entry__.build();
the GPU tests (tests/test_gpu_reference_callers.py) run what this builds.  No compute here: only that the programs exist,
are linked against the product library and really route the hot-path entry points through the shim."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_ref")

needs_ref = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference source tree")


@needs_ref
def test_programs_are_built_and_linked_against_the_library():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "integration"), "-j8"])
    for exe in ("Ludwig_b200.exe", "unit_b200.exe", "Ludwig_soa.exe", "unit_soa.exe"):
        assert os.path.exists(os.path.join(BIN, exe)), exe
    for exe in ("Ludwig_b200.exe", "unit_b200.exe"):
        dyn = subprocess.run(["readelf", "-d", os.path.join(BIN, exe)], capture_output=True, text=True).stdout
        assert "libludwig_b200.so" in dyn, exe
        syms = subprocess.run(["nm", os.path.join(BIN, exe)], capture_output=True, text=True).stdout
        # the reference's callers reach the shim, the shim reaches the C-ABI, the reference's own bodies stay linked
        for name in ("__wrap_lb_collide", "__wrap_lb_propagation", "__wrap_lb_halo", "__wrap_field_halo", "__wrap_phi_cahn_hilliard"):
            assert name in syms, (exe, name)
        for name in ("lb200_lb_collide", "lb200_lb_propagation", "lb200_phi_cahn_hilliard", "lb200_memcpy"):
            assert f"U {name}" in syms, (exe, name)
    # the unmodified reference program does NOT depend on the library
    dyn = subprocess.run(["readelf", "-d", os.path.join(BIN, "Ludwig_soa.exe")], capture_output=True, text=True).stdout
    assert "libludwig_b200" not in dyn


@needs_ref
def test_shim_uses_the_reference_headers_not_copies():
    src = open(os.path.join(ROOT, "integration", "ludwig_b200_shim.c")).read()
    for hdr in ("lb_data.h", "field.h", "hydro.h", "phi_cahn_hilliard.h"):
        assert f'#include "{hdr}"' in src
        assert not os.path.exists(os.path.join(ROOT, "integration", hdr))      # the reference's own, found with -I at build time
    assert "struct lb_data_s" not in src and "struct field_s" not in src


@needs_ref
def test_reference_unit_suites_pass_on_the_reference_itself(tmp_path):
    """the same unit_main.c + tests/unit/*.c against the reference's own code (CPU): the suites are runnable as selected"""
    r = subprocess.run([os.path.join(BIN, "unit_soa.exe")], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "the reference's hot-path unit suites passed" in r.stdout, r.stdout[-2000:]


@needs_ref
@pytest.mark.parametrize("model", ["d3q15", "d3q27"])
def test_drivers_of_the_other_velocity_sets_are_built(model):
    """the reference fixes the velocity set at compile time: the two drivers again per set (tools/regression_sweep.py --model)"""
    out = os.path.join(BIN, model)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "integration"), "-j8", "MODEL=_%s_" % model.upper(), "OUT=" + out, "drivers"])
    for exe in ("Ludwig_b200.exe", "Ludwig_soa.exe"):
        assert os.path.exists(os.path.join(out, exe)), exe
    dyn = subprocess.run(["readelf", "-d", os.path.join(out, "Ludwig_b200.exe")], capture_output=True, text=True).stdout
    assert "libludwig_b200.so" in dyn and "$ORIGIN/../../../ludwig_b200" in dyn


@needs_ref
@pytest.mark.parametrize("sub,fixture,prefixes", [("d3q19-short", "regression_inputs_d3q19_short.json", ("serial-",)),
                                                   ("d3q27", "regression_inputs_d3q27.json", ("serial-",)),
                                                   ("d3q15", "regression_inputs_d3q15.json", ("serial-",))])
def test_regression_input_fixtures_restate_the_reference_inputs(sub, fixture, prefixes):
    """tests/golden/regression_inputs_*.json (tools/make_regression_inputs.py) hold every key / value line of the reference's
    serial-*.inp files of that suite, in file order -- nothing else, nothing missing"""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_regression_inputs as m
    have = json.load(open(os.path.join(ROOT, "tests", "golden", fixture)))
    d = os.path.join("/root/reference", "tests", "regression", sub)
    names = sorted(n[:-4] for n in os.listdir(d) if n.endswith(".inp") and n.startswith(prefixes))
    keyed = {k.split("/")[-1]: v for k, v in have.items()}
    assert sorted(keyed) == names
    for n in names:
        assert keyed[n] == m.parse(os.path.join(d, n + ".inp")), n
