#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -2
for pfd in 0 4 8 16; do
  echo "pfd=$pfd"
  ESIMP_SWEEP_PFD=$pfd timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 0,1,2,3,4 2>&1 | tail -1
done
