#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2_gpu_suite.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_c5_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"])
for k, v in d["roofline_others"].items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
print(d.get("cpu_baseline"))
print(json.dumps(d.get("cpu_baseline_measured"))[:1500])
PY
