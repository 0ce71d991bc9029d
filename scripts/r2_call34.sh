#!/bin/bash
# round 2, GPU call 34: the committed end state -- whole -m gpu suite, smoke, the driver-protocol bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2_end_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r2_end_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_end_bench_c5_n1.json 2> $O/r2_end_bench_c5_n1.err; echo "bench rc=$?"; tail -c 300 $O/r2_end_bench_c5_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_end_bench_c5_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"], d["roofline"]["traffic"])
for k, v in d.get("roofline_others", {}).items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
timeout 300 python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline > $O/r2_end_bench_c3.json 2> $O/r2_end_bench_c3.err; echo "c3 rc=$?"; head -c 400 $O/r2_end_bench_c3.json; echo
