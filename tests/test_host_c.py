"""The C host layer with Ludwig's own function names (include/ludwig_host.h): a C test program written like
the reference's unit tests (tests/c/test_host_api.c) is compiled with gcc against libludwig_b200.so and the
oracle, and run on the GPU in both arithmetic modes."""
import os
import subprocess

import pytest

import oracle

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
EXE = os.path.join(ROOT, "tests", "c", "test_host_api.exe")


def build_exe():
    oracle.build()
    src = os.path.join(ROOT, "tests", "c", "test_host_api.c")
    libdir = os.path.join(ROOT, "ludwig_b200")
    odir = os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-O1", "-std=gnu11", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + odir, src, "-o", EXE,
           "-L" + libdir, "-lludwig_b200", "-L" + odir, "-loracle", "-lm",
           "-Wl,-rpath," + os.path.abspath(libdir), "-Wl,-rpath," + os.path.abspath(odir)]
    subprocess.check_call(cmd)


def test_host_api_test_program_compiles_and_links():
    """CPU part: the program builds against the headers and the library exports every name it uses."""
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("math", ["strict", "fast"])
def test_host_api_on_gpu(math):
    build_exe()
    env = dict(os.environ, LB200_MATH=math)
    r = subprocess.run([EXE], capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "FAIL" not in r.stdout
    assert f"PASS test_host_api ({math})" in r.stdout
