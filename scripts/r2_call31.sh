#!/bin/bash
# round 2, GPU call 31 (8 GPUs): the final code under the driver's protocol at N = 8 and N = 4 (per-category clocks outside the timed region)
set -u
mkdir -p gpurun_out
O=gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_final_bench_c5_n$N.json 2> $O/r2_final_bench_c5_n$N.err; echo "bench n=$N rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_final_bench_c5_n$N.json") if l.startswith("{")][-1])
print("N=$N", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"], d["detail_pass"]["ms_per_step"])
b = d.get("multi_gpu_breakdown") or {}
print({k: (round(v["ms_per_step"], 2), round(v["count_per_step"], 1), round(v["avg_ms"], 4)) for k, v in b.items()})
PY
done
