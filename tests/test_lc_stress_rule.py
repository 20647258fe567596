"""The liquid-crystal stress of the reference's kernels, fe_lc_compute_stress_v (src/blue_phase.c:2279-2775), is a
500-line mechanical unrolling.  The oracle (oracle/lb_oracle_lc.c: orc_lc_compute_stress) and the CUDA kernel
(ludwig_b200/csrc/lb200_lc.cuh: lc_compute_stress) restate it as ONE rule over constant-index loops.  This test
regenerates the statement sequence from that rule and compares it, token for token, with the reference's source text
(in addition to the bit-for-bit numerical pin in tests/test_lc_oracle.py).  Needs /root/reference (build container)."""
import os
import re

import pytest

SRC = "/root/reference/src/blue_phase.c"
pytestmark = pytest.mark.skipif(not os.path.exists(SRC), reason="reference sources not present")

E = {(0, 1, 2): 1, (1, 2, 0): 1, (2, 0, 1): 1, (0, 2, 1): -1, (2, 1, 0): -1, (1, 0, 2): -1}


def rule():
    """the statement list of the rule: (op, expression) with op in '=', '+=', '-=', and the final assignments"""
    out = []
    q = lambda a, b: f"q[{a}][{b}][iv]"
    h = lambda a, b: f"h[{a}][{b}][iv]"
    dq = lambda a, b, c: f"dq[{a}][{b}][{c}][iv]"
    for ia in range(3):
        for ib in range(3):
            if ia == ib:
                out.append(("=", f"2.0*xi*({q(ia, ib)}+r3)*qh[iv]-p0[iv]"))
            else:
                out.append(("=", f"2.0*xi*({q(ia, ib)})*qh[iv]"))
            for ic in range(3):
                qb = f"({q(ib, ic)}+r3)" if ib == ic else f"({q(ib, ic)})"
                qa = f"({q(ia, ic)}+r3)" if ia == ic else f"({q(ia, ic)})"
                out.append(("+=", f"-xi*{h(ia, ic)}*{qb}-xi*{qa}*{h(ib, ic)}"))
            for ic in range(3):
                for id_ in range(3):
                    out.append(("+=", f"-kappa0*{dq(ia, ib, ic)}*{dq(id_, ic, id_)}-kappa1*{dq(ia, ic, id_)}*{dq(ib, ic, id_)}"
                                      f"+kappa1*{dq(ia, ic, id_)}*{dq(ic, ib, id_)}"))
                    if ib != ic:
                        ie = 3 - ib - ic
                        op = "-=" if E[(ib, ic, ie)] > 0 else "+="
                        out.append((op, f"2.0*kappa1*q0*{dq(ia, ic, id_)}*{q(id_, ie)}"))
            for ic in range(3):
                out.append(("+=", f"{q(ia, ic)}*{h(ib, ic)}-{h(ia, ic)}*{q(ib, ic)}"))
            out.append(("store", f"s[{'XYZ'[ia]}][{'XYZ'[ib]}][iv]=-sthtmp[iv]"))
    return out


def reference_statements():
    text = open(SRC).read()
    a = text.index("void fe_lc_compute_stress_v(")
    b = text.index("\n}\n", a)
    body = text[a:b]
    body = body[body.index("The rest is automatically generated"):]
    out = []
    for m in re.finditer(r"for_simd_v\(iv, NSIMDVL\)\s*(.*?);", body, re.S):
        st = re.sub(r"\s+", "", m.group(1))
        mm = re.match(r"sthtmp\[iv\](\+=|-=|=)(.*)", st)
        if mm:
            out.append((mm.group(1), mm.group(2)))
        else:
            out.append(("store", st))
    return out


def test_unrolled_stress_follows_the_rule():
    ours, ref = rule(), reference_statements()
    assert len(ours) == len(ref) == 9 * (1 + 3 + 9 + 6 + 3 + 1)
    for i, (x, y) in enumerate(zip(ours, ref)):
        assert x == y, (i, x, y)
