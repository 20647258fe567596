#!/usr/bin/env python3
"""Restate the reference's own regression inputs (tests/regression/<dir>/serial-*.inp: `key value` lines, comments dropped) as
one JSON file of {case: [[key, value], ...]} (file order kept, duplicates included) -- the fixture tools/regression_sweep.py feeds to the reference's driver linked against
libludwig_b200.so.  Run where /root/reference exists:
    python tools/make_regression_inputs.py /root/reference d3q19-short tests/golden/regression_inputs_d3q19_short.json"""
import json
import os
import sys


def parse(path):
    keys = []
    with open(path) as fh:
        for ln in fh:
            ln = ln.strip()
            if not ln or ln.startswith("#"):
                continue
            parts = ln.split(None, 1)
            if len(parts) == 2:
                keys.append([parts[0], parts[1].strip()])
    return keys


def main():
    ref, sub, out = sys.argv[1], sys.argv[2], sys.argv[3]
    d = os.path.join(ref, "tests", "regression", sub)
    cases = {}
    for name in sorted(os.listdir(d)):
        if name.startswith("serial-") and name.endswith(".inp"):
            cases[name[:-4]] = parse(os.path.join(d, name))
    with open(out, "w") as fh:                                    # one case per line
        fh.write("{\n" + ",\n".join(json.dumps(k) + ": " + json.dumps(v) for k, v in sorted(cases.items())) + "\n}\n")
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
