#!/usr/bin/env python3
"""Run the reference's own regression inputs (tests/golden/regression_inputs_*.json, restated by tools/make_regression_inputs.py)
through the reference's own driver linked against libludwig_b200.so (integration/_ref/Ludwig_b200.exe) and compare each log
with the log of the unmodified reference (integration/_ref/Ludwig_soa.exe) under the rules of the reference's tests/test-diff.sh
(same filter and 1e-12 tolerance as tests/test_gpu_reference_callers.py).

    python tools/regression_sweep.py reference [--threads 8]     # CPU: the reference's logs -> integration/_ref/sweep/<case>.log
    python tools/regression_sweep.py library [--math strict]      # GPU: the library's logs, the comparison, a summary

Outcome per case: MATCH (logs equal) | REFUSED (the shim or the library said "outside this library / build": an explicit refusal) |
DIFF (both ran, logs differ) | REF-FAILED (the reference itself did not run this input here, e.g. it needs a restart or colloid
file) | FAILED (the library run crashed or timed out)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
BIN = os.path.join(ROOT, "integration", "_ref")
SWEEP = os.path.join(BIN, "sweep")


def run(exe, pairs, env, timeout, bindir=None):
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "input"), "w") as fh:
            for k, v in pairs:
                fh.write(f"{k} {v}\n")
        e = dict(os.environ)
        e.update(env)
        try:
            r = subprocess.run([os.path.join(bindir or BIN, exe)], cwd=d, capture_output=True, text=True, timeout=timeout, env=e)
            return r.returncode, r.stdout, r.stderr
        except subprocess.TimeoutExpired:
            return -999, "", "timeout"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("side", choices=["reference", "library"])
    ap.add_argument("--inputs", default=os.path.join(ROOT, "tests", "golden", "regression_inputs_d3q19_short.json"))
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--math", default="strict")
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "regression_sweep.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--model", default="", help="d3q15 / d3q27: the drivers of integration/_ref/<model>/ (make MODEL=... drivers)")
    args = ap.parse_args()
    cases = json.load(open(args.inputs))
    if args.only:
        cases = {k: v for k, v in cases.items() if args.only in k}
    os.makedirs(SWEEP, exist_ok=True)
    bindir = os.path.join(BIN, args.model) if args.model else BIN
    if args.side == "reference":
        for name, pairs in cases.items():
            rc, out, err = run("Ludwig_soa.exe", pairs, {"OMP_NUM_THREADS": str(args.threads)}, args.timeout, bindir)
            ok = (rc == 0 and "Ludwig finished normally" in out)
            with open(os.path.join(SWEEP, name.replace("/", "__") + ".log"), "w") as fh:
                fh.write(out if ok else "REF-FAILED rc=%d\n%s\n%s" % (rc, out[-2000:], err[-2000:]))
            print(name, "ok" if ok else "REF-FAILED", flush=True)
        return
    from test_gpu_reference_callers import diff_logs
    results = {}
    for name, pairs in cases.items():
        path = os.path.join(SWEEP, name.replace("/", "__") + ".log")
        ref = open(path).read() if os.path.exists(path) else "REF-FAILED (no log)"
        if ref.startswith("REF-FAILED"):
            results[name] = {"outcome": "REF-FAILED", "detail": ref.splitlines()[-1][:200] if ref.splitlines() else ""}
            print(name, "REF-FAILED", flush=True)
            continue
        rc, out, err = run("Ludwig_b200.exe", pairs, {"LB200_MATH": args.math}, args.timeout, bindir)
        text = out + err
        if rc == 0 and "Ludwig finished normally" in out:
            bad = diff_logs(ref, out)
            results[name] = {"outcome": "MATCH"} if not bad else {"outcome": "DIFF", "detail": bad[:4], "ndiff": len(bad)}
        elif "libludwig_b200" in text or "outside this" in text:
            msg = [ln for ln in text.splitlines() if "libludwig_b200" in ln or "outside this" in ln]
            results[name] = {"outcome": "REFUSED", "detail": msg[-1][:200] if msg else ""}
        else:
            results[name] = {"outcome": "FAILED", "detail": (err or out)[-400:], "rc": rc}
        print(name, results[name]["outcome"], str(results[name].get("detail", ""))[:160], flush=True)
    count = {}
    for r in results.values():
        count[r["outcome"]] = count.get(r["outcome"], 0) + 1
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"math": args.math, "count": count, "cases": results}, open(args.out, "w"), indent=1, sort_keys=True)
    print(json.dumps(count))


if __name__ == "__main__":
    main()
