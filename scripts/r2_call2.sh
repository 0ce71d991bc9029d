#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_tma_probe.jsonl
for v in 0 1 2 3 4 5 6 10 11 7; do
  timeout 60 scripts/microbench/tma_rows $v >> gpurun_out/r2_tma_probe.jsonl 2>&1; echo "variant $v rc=$?" >> gpurun_out/r2_tma_probe.jsonl
done
cat gpurun_out/r2_tma_probe.jsonl
