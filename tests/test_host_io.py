"""On-disk formats: the distribution / field files and their metadata written by the C host layer
(include/ludwig_host.h: lb_io_write, field_io_write) are BYTE-IDENTICAL with the files the unmodified reference writes
for the same data, and the host layer reads the reference's files back exactly -- so either code can restart from the
other's output.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

import refharness as R

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
EXE = os.path.join(ROOT, "tests", "c", "test_host_io.exe")

pytestmark = pytest.mark.skipif(not R.available(), reason="reference library oracle/_ref not built")


def build_exe():
    libdir = os.path.join(ROOT, "ludwig_b200")
    cmd = ["gcc", "-O1", "-std=gnu11", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "test_host_io.c"),
           "-o", EXE, "-L" + libdir, "-lludwig_b200", "-lm", "-Wl,-rpath," + os.path.abspath(libdir)]
    subprocess.check_call(cmd)


def tags(n, nhalo, ncomp, offset=0.0):
    nall = tuple(m + 2 * nhalo for m in n)
    a = np.zeros((ncomp,) + nall)
    ic, jc, kc = np.meshgrid(*(np.arange(1, m + 1) for m in n), indexing="ij")
    for c in range(ncomp):
        a[c, nhalo:-nhalo, nhalo:-nhalo, nhalo:-nhalo] = offset + (((ic * 64 + jc) * 64 + kc) * 64 + c)
    return a.reshape(ncomp, -1)


@pytest.mark.parametrize("fmt", ["binary_records", "ascii_records"])
@pytest.mark.parametrize("case", ["binary", "symmlb", "lc", "le"])
def test_files_match_reference(case, fmt, tmp_path):
    """fmt ascii_records: options.iodata.{input,output}.iorformat = IO_RECORD_ASCII (the reference's default_io_format ascii)"""
    build_exe()
    ascii_ = int(fmt == "ascii_records")
    n = (6, 4, 5) if case != "le" else (16, 4, 5)
    kw = dict(nhalo=2, adv_order=1, eta_shear=0.1, io_ascii=ascii_)
    nvel, ndist, nf, name, planes = 19, 1, 1, "phi", 0
    if case == "binary":
        kw.update(have_phi=1, a=-0.1, b=0.1, kappa=0.1, mobility=0.1)
    elif case == "symmlb":
        ndist = 2
        kw.update(have_phi=1, ndist=2, a=-0.1, b=0.1, kappa=0.1, mobility=0.1)
    elif case == "lc":
        nf, name = 5, "q"
        kw.update(lc=dict(a0=0.01, q0=0.1, gamma=3.0, kappa0=0.01, kappa1=0.01, xi=0.7, Gamma=0.5))
    else:
        planes = 2
        kw.update(have_phi=1, a=-0.1, b=0.1, kappa=0.1, mobility=0.1, le_nplanes=2, le_uy=0.05)
    dref, dours = tmp_path / "ref", tmp_path / "ours"
    dref.mkdir(); dours.mkdir()
    cwd = os.getcwd()
    try:
        os.chdir(dref)
        with R.RefSim(n, **kw) as s:
            s.set(R.REF_F, tags(n, 2, ndist * nvel))
            fld = tags(n, 2, nf, 0.5)
            if s.nsites_le != s.nsites:                       # Lees-Edwards: field arrays carry buffer planes
                fld = np.concatenate([fld, np.zeros((nf, s.nsites_le - s.nsites))], axis=1)
            s.set(R.REF_Q if case == "lc" else R.REF_PHI, fld)
            assert s.io("lb_io_write", 7) == 0 and s.io("field_io_write", 7) == 0
    finally:
        os.chdir(cwd)
    args = [str(x) for x in (*n, nvel, ndist, nf, name, planes)] + (["ascii"] if ascii_ else [])
    r = subprocess.run([EXE, "write"] + args, cwd=dours, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    files = sorted(os.listdir(dref))
    assert files == sorted(os.listdir(dours)) and len(files) == 4, (files, sorted(os.listdir(dours)))
    for f in files:
        a, b = (dref / f).read_bytes(), (dours / f).read_bytes()
        assert a == b, (f, len(a), len(b), a[:300] if "meta" in f else None, b[:300] if "meta" in f else None)
    # and the host layer reads the reference's own files
    r = subprocess.run([EXE, "read"] + args, cwd=dref, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS read" in r.stdout, r.stdout + r.stderr
