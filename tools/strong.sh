#!/bin/bash
# strong scaling point: fixed 512x256x256 lattice over N GPUs
N=$1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --strong --steps 50 --warmup 5 --no-cpu 2>/dev/null | tail -1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --strong --steps 50 --warmup 5 --no-cpu 2>/dev/null | tail -1
fi | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('N=%d %s MLUPS %.0f  ms/step %.3f |' % (d['n_gpus'], d['config']['lattice_per_gpu'], d['value'], d['ms_per_step']), ' '.join('%s %.3f x%d' % (a, b['ms_per_launch'], b['launches']) for a, b in k.items() if b['launches']))
"
