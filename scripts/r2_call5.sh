#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -3
echo "promo NONE"; timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 2>&1 | tail -1 | tee -a gpurun_out/r2_modes_isolated.jsonl
echo "promo 256"; ESIMP_TMA_L2PROMO=3 timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 2>&1 | tail -1
echo "promo 128"; ESIMP_TMA_L2PROMO=2 timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 2>&1 | tail -1
