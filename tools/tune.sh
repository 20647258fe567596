#!/bin/bash
# quick per-kernel timing of variants (each in a fresh process: knobs are read once)
run() {
  echo "== $*"
  env "$@" python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('MLUPS %.0f  ms/step %.3f | collide %.3f force_ch %.3f grad %.3f halo %.3f x%d | e2e %.0f' % (d['value'], d['ms_per_step'], k['collide']['ms_per_launch'], k['force_ch']['ms_per_launch'], k['grad']['ms_per_launch'], k['halo']['ms_per_launch'], k['halo']['launches']//k['collide']['launches'], d['e2e']['value']))
"
}
run LB200_FCH_MINB=4
run LB200_FCH_MINB=3
run LB200_FCH_MINB=5
run LB200_FCH_MINB=6
run LB200_FCH_MINB=4 LB200_STCS=0
