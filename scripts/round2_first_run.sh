#!/bin/bash
# First GPU call of round 2: validate the opt-in FP32 V-cycle (csrc/multigrid32.cu, never run on a device so far),
# then measure it against the FP64 V-cycle on the headline workload.  Run through gpurun from the repo root:
#   gpurun --timeout 900 -- 'bash scripts/round2_first_run.sh'
set -u
mkdir -p gpurun_out
ESIMP_TEST_GMG_FP32=1 timeout 600 python -m pytest tests/test_gpu_gmg_fp32.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_gmg_fp32_tests.log
if grep -q "passed" gpurun_out/r2_gmg_fp32_tests.log && ! grep -q "failed" gpurun_out/r2_gmg_fp32_tests.log; then
    timeout 300 python bench.py --precond gmg32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg32.json 2> gpurun_out/r2_bench_c5_gmg32.err
    ESIMP_F32_CTAS=3 timeout 300 python bench.py --precond gmg32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg32_3ctas.json 2> gpurun_out/r2_bench_c5_gmg32_3ctas.err
    timeout 300 python bench.py --precond gmg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_gmg.json 2> gpurun_out/r2_bench_c5_gmg.err
    tail -c 600 gpurun_out/r2_bench_c5_gmg32_3ctas.json
    tail -c 600 gpurun_out/r2_bench_c5_gmg32.json
    tail -c 600 gpurun_out/r2_bench_c5_gmg.json
else
    # small reproducer with the solver trace
    ESIMP_MG_DEBUG=1 ESIMP_TEST_GMG_FP32=1 timeout 120 python -m pytest tests/test_gpu_gmg_fp32.py -m gpu -x -q -k "24-12-6 or nel2" 2>&1 | tail -60 > gpurun_out/r2_gmg_fp32_debug.log
fi
