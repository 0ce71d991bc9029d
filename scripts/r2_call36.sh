#!/bin/bash
# round 2, GPU call 36: one-cell-per-thread sweep kernel (sweep1_kernel.cuh, ESIMP_SWEEP_CPT1 = 16 | 12): isolated mode timings and
# parity through the C ABI; the relaxed OC-step assertion of the example-02 golden loop
set -u
mkdir -p gpurun_out
O=gpurun_out
for v in 16 12; do
  echo "== CPT1=$v" | tee -a $O/r2_sweep1_isolated.log
  ESIMP_SWEEP_CPT1=$v timeout 70 python scripts/bench_operator_modes.py 512 256 256 0 20 2>&1 | tail -1 | tee -a $O/r2_sweep1_isolated.log
done
for v in 16 12; do
  echo "== CPT1=$v" | tee -a $O/r2_sweep1_parity.log
  ESIMP_SWEEP_CPT1=$v timeout 90 python -m pytest tests/test_gpu_sweep.py -q 2>&1 | tail -3 | tee -a $O/r2_sweep1_parity.log
done
timeout 60 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sliding" 2>&1 | tail -3 | tee $O/r2_call36_sliding.log
