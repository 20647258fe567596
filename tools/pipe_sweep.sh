#!/bin/bash
# slab-pipeline sweep: "S:SMS[:CHUNKS[:GREEN]]" configurations, one bench line each (0 = serial step)
out=gpurun_out/pipe_sweep.log
: > $out
for cfg in "$@"; do
  IFS=: read -r s sms ch gr <<< "$cfg"
  env LB200_PIPE_CHUNKS=${ch:-1} LB200_PIPE_GREEN=${gr:-1} timeout 200 python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu --pipe $s --pipe-sms ${sms:-56} 2>>gpurun_out/pipe_sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('$cfg', 'MLUPS %.0f  ms/step %.4f |' % (d['value'], d['ms_per_step']), d['config'].get('slab_pipeline'), '|', ' '.join('%s %.3f x%d' % (a, b['ms_per_launch'], b['launches']) for a, b in k.items() if b['launches']), '| e2e %.0f' % d['e2e']['value'], '| clocks', d['clocks'])
" >> $out 2>&1
done
cat $out
