#!/bin/bash
# round 2, GPU call 1: FP64 roof + TMA row-box probe, then the never-run FP32 V-cycle tests and gmg vs gmg32 bench lines
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 120 scripts/microbench/fp64_peak > gpurun_out/r2_fp64_peak.json 2> gpurun_out/r2_fp64_peak.err; echo "fp64_peak rc=$?"
timeout 120 scripts/microbench/tma_rows > gpurun_out/r2_tma_rows.json 2> gpurun_out/r2_tma_rows.err; echo "tma_rows rc=$?"
cat gpurun_out/r2_tma_rows.json
ESIMP_TEST_GMG_FP32=1 timeout 600 python -m pytest tests/test_gpu_gmg_fp32.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_gmg_fp32_tests.log
timeout 400 python bench.py --precond gmg --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg.json 2> gpurun_out/r2_bench_c5_gmg.err; echo "gmg rc=$?"
if grep -q "passed" gpurun_out/r2_gmg_fp32_tests.log && ! grep -q "failed" gpurun_out/r2_gmg_fp32_tests.log; then
    timeout 400 python bench.py --precond gmg32 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg32.json 2> gpurun_out/r2_bench_c5_gmg32.err; echo "gmg32 rc=$?"
    ESIMP_F32_CTAS=3 timeout 400 python bench.py --precond gmg32 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg32_3ctas.json 2> gpurun_out/r2_bench_c5_gmg32_3ctas.err
else
    ESIMP_MG_DEBUG=1 ESIMP_TEST_GMG_FP32=1 timeout 120 python -m pytest tests/test_gpu_gmg_fp32.py -m gpu -x -q 2>&1 | tail -80 > gpurun_out/r2_gmg_fp32_debug.log
fi
head -c 700 gpurun_out/r2_bench_c5_gmg.json; echo
head -c 700 gpurun_out/r2_bench_c5_gmg32.json 2>/dev/null; echo
cat gpurun_out/r2_fp64_peak.json
