#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 2 6; do timeout 200 python scripts/bench_operator_modes.py 512 256 256 $v 10 2>&1 | tail -1 | tee -a gpurun_out/r2_modes_isolated.jsonl; done
timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 2 0,2 > gpurun_out/plain_for_ncu.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_sweep -c 6 -f -o gpurun_out/prof_r2_sweep_v1 python scripts/bench_operator_modes.py 512 256 256 6 2 0,2 > gpurun_out/ncu_r2_sweep_v1.log 2>&1
tail -3 gpurun_out/ncu_r2_sweep_v1.log
