#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweep.py tests/test_gpu_post_load.py -m gpu -x -q -k "gmg or sweep or golden or host_api" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"])
for k, v in d["roofline_others"].items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
