#!/bin/bash
# round 2, GPU call 25: cluster tail with shared-memory slab staging (parity, in-step timing: levels <= 10000 / <= 2000 nodes / off)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gmg or golden" 2>&1 | tail -6 | tee $O/r2_c25_tests.log
run_bench() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c25_bench_$name.json 2> $O/r2_c25_bench_$name.err; echo "bench $name rc=$?"; tail -c 400 $O/r2_c25_bench_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_c25_bench_$name.json"))
print("$name", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
PY
}
run_bench tail10000 ESIMP_MG_TAIL=1
run_bench tail2000 ESIMP_MG_TAIL=1 ESIMP_MG_TAIL_NODES=2000
run_bench tail0 ESIMP_MG_TAIL=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_coarse_tail -c 6 python scripts/solver_robustness.py c4 1 2>&1 | grep -i "k_coarse_tail\|duration" | head -14
