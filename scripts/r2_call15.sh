#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r2_mg_sweep.txt
run() { echo "== $*" | tee -a gpurun_out/r2_mg_sweep.txt; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('it/s %.3f  ms/step %.1f  CG its %.1f  solve ms %.1f' % (d['value'], d['ms_per_step'], d['config']['mean_pcg_iters_per_step'], d['phases_ms_per_step']['ms_solve']))" | tee -a gpurun_out/r2_mg_sweep.txt; }
run ESIMP_MG_NU=2 ESIMP_MG_RATIO=10
run ESIMP_MG_NU=1 ESIMP_MG_RATIO=10
run ESIMP_MG_NU=1 ESIMP_MG_RATIO=6
run ESIMP_MG_NU=3 ESIMP_MG_RATIO=10
run ESIMP_MG_NU=2 ESIMP_MG_RATIO=16
run ESIMP_MG_NU=2 ESIMP_MG_RATIO=6
run ESIMP_MG_NU=2 ESIMP_MG_RATIO=10 ESIMP_MG_COARSE_ITS=8
run ESIMP_MG_NU=2 ESIMP_MG_RATIO=10 ESIMP_SWEEP_SMOOTH=1
