#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 20 2>&1 | tail -1 | tee -a gpurun_out/r2_modes_isolated.jsonl
timeout 200 python scripts/bench_operator_modes.py 512 256 256 2 20 2>&1 | tail -1 | tee -a gpurun_out/r2_modes_isolated.jsonl
timeout 400 python bench.py --precond gmg --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg_sweep2.json 2> gpurun_out/r2_bench_c5_gmg_sweep2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_gmg_sweep2.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["phases_ms_per_step"], d["e2e"])
PY
