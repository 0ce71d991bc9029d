#!/bin/bash
set -u
mkdir -p gpurun_out
echo skip tests
for e in 0 1; do
ESIMP_UM_STREAM_KE=$e timeout 600 python bench.py --workload c3 --steps 5 --warmup 2 > gpurun_out/r2_bench_c3_stream$e.json 2> gpurun_out/r2_bench_c3.err; echo "c3 stream_ke=$e rc=$?"; tail -c 300 gpurun_out/r2_bench_c3.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_c3_stream$e.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
PY
done
