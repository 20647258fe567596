#!/bin/bash
# quick GPU check: parity tests of the step paths + one short bench line (per-kernel ms)
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3)
env "$@" python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('MLUPS %.0f  ms/step %.3f |' % (d['value'], d['ms_per_step']), ' '.join('%s %.3f x%d' % (a, b['ms_per_launch'], b['launches']) for a, b in k.items() if b['launches']), '| e2e %.0f' % d['e2e']['value'])
"
