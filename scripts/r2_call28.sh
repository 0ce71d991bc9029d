#!/bin/bash
# round 2, GPU call 28 (8 GPUs): slab parity under pytest (2 and 4 ranks), then the N = 8 bench line with its event-clock breakdown,
# default path and the opt-in peer-memory halo + lagged poll
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee $O/r2_c28_tests.log
i=0
for cfg in "ESIMP_P2P_HALO=0" "ESIMP_P2P_HALO=1 ESIMP_PCG_LAG=1"; do
  i=$((i+1))
  env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$i \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_c28_bench_n8_$i.json 2> $O/r2_c28_bench_n8_$i.err; echo "bench n=8 $cfg rc=$?"
  tail -c 300 $O/r2_c28_bench_n8_$i.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_c28_bench_n8_$i.json") if l.startswith("{")][-1])
print("$cfg", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
b = d.get("multi_gpu_breakdown") or {}
print({k: (round(v["ms_per_step"], 2), round(v["count_per_step"], 1), round(v["avg_ms"], 4)) for k, v in b.items()})
PY
done
