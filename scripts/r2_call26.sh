#!/bin/bash
# round 2, GPU call 26 (2 GPUs): peer-memory halo exchange vs NCCL send / recv (parity under pytest first, then the bench breakdown)
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
nvidia-smi topo -m > $O/r2_c26_topo.txt 2>&1
ESIMP_P2P_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee $O/r2_c26_tests_n$N.log
for p2p in 1 0; do
  ESIMP_P2P_DEBUG=1 ESIMP_P2P_HALO=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$p2p \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_c26_bench_n${N}_p2p$p2p.json 2> $O/r2_c26_bench_n${N}_p2p$p2p.err; echo "bench n=$N p2p=$p2p rc=$?"
  grep -i "esimp rank\|error" $O/r2_c26_bench_n${N}_p2p$p2p.err | head -5
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_c26_bench_n${N}_p2p$p2p.json") if l.startswith("{")][-1])
print("p2p$p2p", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
b = d.get("multi_gpu_breakdown") or {}
print({k: (round(v["ms_per_step"], 2), round(v["count_per_step"], 1), round(v["avg_ms"], 4)) for k, v in b.items()})
PY
done
