#!/bin/bash
set -u
mkdir -p gpurun_out
for pf in 0 1; do for ns in 0 100 400; do
  echo "prefetch=$pf backoff_ns=$ns"
  ESIMP_SWEEP_PREFETCH=$pf ESIMP_SWEEP_BACKOFF_NS=$ns timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 0,2,3 2>&1 | tail -1
done; done
