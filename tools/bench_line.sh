#!/bin/bash
# one compact line per bench run: tools/bench_line.sh LABEL [ENV=VAL ...] -- [bench args]
label=$1; shift
envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" 2>>gpurun_out/bench_line.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('$label', 'MLUPS %.0f  ms/step %.4f |' % (d['value'], d['ms_per_step']), ' '.join('%s %.4f x%d' % (a, b['ms_per_launch'], b['launches']) for a, b in k.items() if b['launches']), '| e2e %.0f' % d['e2e']['value'], '| launches', d['gpu_launches'], '| clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
" | tee -a gpurun_out/bench_lines.log
