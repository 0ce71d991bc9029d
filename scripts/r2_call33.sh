#!/bin/bash
# round 2, GPU call 33: D^-1 as one double per node in the sweep kernel's smoother modes (parity, isolated modes, bench line)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -m gpu -x -q -k "sweep or gmg or golden" 2>&1 | tail -4 | tee $O/r2_c33_tests.log
timeout 200 python scripts/bench_operator_modes.py 512 256 256 0 20 2>&1 | tail -1 | tee $O/r2_c33_operator_modes_isolated.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_c33_bench_c5_n1.json 2> $O/r2_c33_bench_c5_n1.err; echo "bench rc=$?"; tail -c 300 $O/r2_c33_bench_c5_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_c33_bench_c5_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
for k, v in d.get("roofline_others", {}).items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
