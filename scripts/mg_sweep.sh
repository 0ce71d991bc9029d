#!/bin/bash
# sweep of the V-cycle knobs on the C5 bench (device-resident value only)
for cfg in "2 10 16" "1 10 16" "1 6 16" "3 10 16" "2 16 16" "1 16 16"; do
  set -- $cfg
  ESIMP_MG_NU=$1 ESIMP_MG_RATIO=$2 ESIMP_MG_COARSE_ITS=$3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nu $1 ratio $2 coarse $3: %.1f ms/step, its %s' % (d['ms_per_step'], [h['pcg_iters'] for h in d['history']]))"
done
