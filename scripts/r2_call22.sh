#!/bin/bash
# round 2, GPU call 22 (2 GPUs): lagged vs blocking CG poll on slabs, per-category event clocks, multi-GPU parity under pytest
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r2_c22_tests_n$N.log
for lag in 1 0; do
  ESIMP_PCG_LAG=$lag timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$lag \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_c22_bench_n${N}_lag$lag.json 2> $O/r2_c22_bench_n${N}_lag$lag.err; echo "bench n=$N lag=$lag rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_c22_bench_n${N}_lag$lag.json") if l.startswith("{")][-1])
print("lag$lag", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"])
print(json.dumps(d.get("multi_gpu_breakdown"), indent=None))
PY
done
