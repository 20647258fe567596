#!/bin/bash
# quick per-kernel timing of variants (each in a fresh process: knobs are read once)
run() {
  echo "== $*"
  env "$@" python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('MLUPS %.0f  ms/step %.3f |' % (d['value'], d['ms_per_step']), ' '.join('%s %.3f x%d' % (a, b['ms_per_launch'], b['launches']) for a, b in k.items() if b['launches']), '| e2e %.0f' % d['e2e']['value'])
"
}
run LB200_PHI_SECTOR=1
run LB200_PHI_SECTOR=0
run LB200_PHI_SECTOR=0 LB200_SPLIT_FCH=1
