#!/bin/bash
# round 2, GPU call 27 (2 GPUs): peer-memory halo kernel with 96 / 32 CTAs against NCCL send / recv, bench breakdown
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
i=0
for cfg in "ESIMP_P2P_HALO=1 ESIMP_P2P_CTAS=96" "ESIMP_P2P_HALO=1 ESIMP_P2P_CTAS=32" "ESIMP_P2P_HALO=0"; do
  i=$((i+1))
  env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$i \
      bench.py --gpus $N --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c27_bench_n${N}_$i.json 2> $O/r2_c27_bench_n${N}_$i.err; echo "bench n=$N $cfg rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_c27_bench_n${N}_$i.json") if l.startswith("{")][-1])
print("$cfg", d["value"], d["ms_per_step"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
b = d.get("multi_gpu_breakdown") or {}
print({k: (round(v["ms_per_step"], 2), round(v["count_per_step"], 1), round(v["avg_ms"], 4)) for k, v in b.items() if k.startswith("comm") or k.startswith("mg_below") or k.startswith("host")})
PY
done
