#!/bin/bash
# multi-GPU sweep: how deep to coarsen slabs / how many coarse iterations (8 ranks)
for cfg in "1 16" "2 16" "4 16" "4 40" "8 40"; do
  set -- $cfg
  ESIMP_MG_MIN_LAYERS=$1 ESIMP_MG_COARSE_ITS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$1 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('min_layers $1 coarse_its $2: %.1f ms/step (%.2f it/s), its %s, phases %s' % (d['ms_per_step'], d['value'], [h['pcg_iters'] for h in d['history']], {k:round(v,1) for k,v in d['phases_ms_per_step'].items()}))"
done
