"""The generated D3Q19 projection tables (oracle/d3q19_tables.h, ludwig_b200/csrc/d3q19_proj.cuh) are
derived from the model definition; here they are (a) re-derived and compared with the committed files and
(b) where /root/reference exists, compared entry by entry with the reference's unrolled source."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_generated_files_are_current(tmp_path):
    import gen_d3q19_tables as g
    fwd, bwd, na = g.derive(quirk=True)
    h = tmp_path / "t.h"
    cu = tmp_path / "t.cuh"
    g.emit_c_header(fwd, bwd, str(h))
    g.emit_cuda(fwd, bwd, str(cu))
    assert h.read_text() == open(os.path.join(ROOT, "oracle", "d3q19_tables.h")).read()
    assert cu.read_text() == open(os.path.join(ROOT, "ludwig_b200", "csrc", "d3q19_proj.cuh")).read()
    # normalisers quoted in SURVEY.md appendix A
    assert [float(x) for x in na] == [1, 3, 3, 3, 4.5, 9, 9, 4.5, 9, 4.5, 0.75, 1.5, 1.5, 1.5, 2.25, 4.5, 4.5, 4.5, 0.5]


def test_model_matrices_orthogonal():
    from oracle import Oracle
    for nvel in (15, 19, 27):
        o = Oracle((2, 2, 2), nvel=nvel)
        ma = np.array([[o.m.ma[m][p] for p in range(nvel)] for m in range(nvel)])
        mi = np.array([[o.m.mi[p][m] for m in range(nvel)] for p in range(nvel)])
        assert np.allclose(ma @ mi, np.eye(nvel), atol=1e-14)
        assert abs(o.wv.sum() - 1.0) < 1e-15
        for p in range(1, nvel):
            assert (o.cv[nvel - p] == -o.cv[p]).all()     # used by the halo code, src/lb_data.c:990-992


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not present")
def test_tables_match_reference_source():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_d3q19_tables.py"), "--check", "/root/reference"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches: 0" in r.stdout
