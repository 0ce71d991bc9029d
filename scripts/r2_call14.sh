#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_mgpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r2_bench_c5_n2.json 2> gpurun_out/r2_bench_c5_n2.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/r2_bench_c5_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_n2.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["share_of_step"], d["config"]["mean_pcg_iters_per_step"])
for k, v in d["roofline_others"].items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
