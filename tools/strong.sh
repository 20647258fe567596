#!/bin/bash
# scaling point: tools/strong.sh N [extra bench.py flags, e.g. --strong]   (fixed 512x256x256 over N GPUs with --strong)
N=$1; shift
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu "$@" 2>/dev/null | tail -1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu "$@" 2>/dev/null | tail -1
fi > /tmp/strong_line.json
python - <<'PY'
import json
d = json.loads(open('/tmp/strong_line.json').read().strip().splitlines()[-1])
k = d['kernels']
print('N=%d %s %s [%s] MLUPS %.0f  ms/step %.3f |' % (d['n_gpus'], d['scaling'], d['config']['lattice_per_gpu'], d['config']['x_plane_exchange'][:6],
      d['value'], d['ms_per_step']), ' '.join('%s %.3f x%d' % (a, b['ms_per_launch'], b['launches']) for a, b in k.items() if b['launches']))
PY
