#!/bin/bash
# round 2, GPU call 3: first run of the sweep kernel (parity, isolated timing, bench) and the whole -m gpu suite on the padded layout
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2_sweep_tests.log
if ! grep -q "passed" gpurun_out/r2_sweep_tests.log || grep -q "failed\|error" gpurun_out/r2_sweep_tests.log; then
    echo "sweep tests not green: stopping"; 
    ESIMP_SWEEP=0 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_gpu_suite_nosweep.log
    exit 0
fi
for v in 2 6; do timeout 120 python scripts/ncu_matvec.py 512 256 256 $v 10 2>&1 | tail -1 | tee -a gpurun_out/r2_matvec_isolated.log; done
for v in 2 6; do timeout 120 python scripts/ncu_matvec.py 240 80 80 $v 20 2>&1 | tail -1 | tee -a gpurun_out/r2_matvec_isolated.log; done
timeout 400 python bench.py --precond gmg --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg_sweep.json 2> gpurun_out/r2_bench_c5_gmg_sweep.err; echo "bench sweep rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_c5_gmg_sweep.json",):
    try:
        d = json.load(open("gpurun_out/" + f))
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["phases_ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_gpu_suite.log
