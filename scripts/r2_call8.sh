#!/bin/bash
set -u
for ring in 4 5 6 7; do
  echo "ring=$ring"
  ESIMP_SWEEP_RING=$ring timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 0,1 2>&1 | tail -1
done
echo "ring=7 chunk 33"; ESIMP_FUSED_CHUNK=33 ESIMP_SWEEP_RING=7 timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 0,1 2>&1 | tail -1
echo "ring=5 chunk 33"; ESIMP_FUSED_CHUNK=33 ESIMP_SWEEP_RING=5 timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 0,1 2>&1 | tail -1
