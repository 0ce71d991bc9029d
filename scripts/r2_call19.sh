#!/bin/bash
# round 2, GPU call 19 (first of the re-entered session): re-establish the evidence that gpurun_out/ lost --
# whole -m gpu suite, the driver-protocol bench line, launch list, ncu --set full of the operator modes, FP64 roof.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt
timeout 120 scripts/microbench/fp64_peak > $O/r2_fp64_peak.json 2> $O/r2_fp64_peak.err; echo "fp64_peak rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/r2_gpu_suite.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_c5_n1.json 2> $O/r2_bench_c5_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
for k, v in d.get("roofline_others", {}).items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
print(d.get("phases_ms_per_step"))
PY
for v in 2 6; do timeout 200 python scripts/bench_operator_modes.py 512 256 256 $v 10 2>&1 | tail -1 | tee -a $O/r2_operator_modes_isolated.log; done
# launch list of the same command (short: 1 timed step after 3 warm-ups; -c bounds the replay)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file $O/r2_launches_bench_c5_gmg.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# ncu --set full of the sweep kernel (plain, residual) and the round-1 kernel in the smoother modes, isolated driver
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 11 -c 10 -o $O/r2_ncu_operator_modes -f \
    python scripts/ncu_operator_modes.py 512 256 256 0 > $O/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O
