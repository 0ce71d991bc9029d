#!/bin/bash
# round 2, GPU call 30: final state -- whole -m gpu suite, smoke, the driver-protocol bench line (+ reference arm), launch list and
# ncu --set full of the operator modes as they run now (all on the sweep kernel, 64-byte TMA promotion, operand preload)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r2_final_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2_final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_final_bench_c5_n1.json 2> $O/r2_final_bench_c5_n1.err; echo "bench rc=$?"; tail -c 300 $O/r2_final_bench_c5_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench_c5_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
for k, v in d.get("roofline_others", {}).items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
print(d.get("phases_ms_per_step")); print(d.get("cpu_baseline")); print(d.get("detail_pass"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_final_bench_reference_arm.json 2> $O/r2_final_bench_reference_arm.err; echo "reference arm rc=$?"; head -c 600 $O/r2_final_bench_reference_arm.json; echo
timeout 200 python scripts/bench_operator_modes.py 512 256 256 0 20 2>&1 | tail -1 | tee $O/r2_final_operator_modes_isolated.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file $O/r2_final_launches_bench_c5.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2_final_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 11 -c 10 -o $O/r2_final_ncu_operator_modes -f \
    python scripts/ncu_operator_modes.py 512 256 256 0 > $O/r2_final_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -12
