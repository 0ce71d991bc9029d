#!/bin/bash
# round 2, GPU call 37: the committed end state once more (launcher generalised for the one-cell kernel, new golden loops, host mirror at C5):
# whole -m gpu suite, smoke, the driver-protocol bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2_last2_gpu_suite.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r2_last2_smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > $O/r2_last2_bench_c5_n1.json 2> $O/r2_last2_bench_c5_n1.err; echo "bench rc=$?"; tail -c 300 $O/r2_last2_bench_c5_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_last2_bench_c5_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
PY
