"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; without a
device it refuses to create a context (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_functions(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long|void \*|const char \*|double)\s*\**\s*((?:lb200|lb|field|hydro|phi|pth|fe|cs|pe|physics|map|grad|advection|lees_edw)_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ludwig_b200 as lb
    lib = lb.load_library()
    headers = [os.path.join(ROOT, "include", h) for h in os.listdir(os.path.join(ROOT, "include")) if h.endswith(".h")]
    names = []
    for h in headers:
        names += declared_functions(h)
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import ludwig_b200 as lb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(lb.Lb200Error, match="no CUDA device"):
        lb.Lb200((8, 8, 8))


def test_product_does_not_reference_the_oracle():
    """Nothing under ludwig_b200/ or include/ may import, link or execute oracle/."""
    bad = []
    for base in ("ludwig_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in fs:
                if f.endswith((".py", ".c", ".cu", ".h", ".cuh", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle|refharness|_ref/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    lib = os.path.join(ROOT, "ludwig_b200", "libludwig_b200.so")
    if os.path.exists(lib):
        import subprocess
        out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "oracle" not in out and "ludwig_ref" not in out
