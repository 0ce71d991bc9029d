#!/bin/bash
# round 2, GPU call 21: A/B of the FUSE_DOT implementations (variant libraries), sweep smoother vs round-1 smoother inside the step with
# the 64-byte TMA promotion
# The variant libraries were built beforehand (not part of `make`): for v in 0 1 2 (a macro of that experiment, removed afterwards: 0 node-by-node dot, 1 mode-space dot, 2 one accumulator per component): nvcc -O3 -std=c++17 -lineinfo -gencode
# arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DSW_DOT_IMPL=$v -c csrc/hex8_structured.cu -o build/hex8_v$v.o, linked with the other objects
# of the library into easysimp.jl_b200/variants/*.so and selected with ESIMP_LIB.
set -u
mkdir -p gpurun_out
O=gpurun_out
V=easysimp.jl_b200/variants
for v in 0 1 2; do
  echo "== dot impl $v" | tee -a $O/r2_c21_dot.log
  ESIMP_LIB=$V/libesimp_dot$v.so ESIMP_TMA_PROMO=1 timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 20 0,1,3 2>&1 | tail -1 | tee -a $O/r2_c21_dot.log
done
ESIMP_LIB=$V/libesimp_dot1.so timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -m gpu -x -q -k "sweep or gmg or matvec" 2>&1 | tail -4 | tee $O/r2_c21_tests.log
run_bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c21_bench_$name.json 2> $O/r2_c21_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_c21_bench_$name.json"))
print("$name", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"])
for k, v in d.get("roofline_others", {}).items():
    if k.startswith("operator"): print("   ", k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
}
run_bench dot0_promo1 ESIMP_LIB=$V/libesimp_dot0.so ESIMP_TMA_PROMO=1
run_bench dot1_promo1 ESIMP_LIB=$V/libesimp_dot1.so ESIMP_TMA_PROMO=1
run_bench dot1_promo1_sweepsmooth ESIMP_LIB=$V/libesimp_dot1.so ESIMP_TMA_PROMO=1 ESIMP_SWEEP_SMOOTH=1
run_bench dot2_promo1 ESIMP_LIB=$V/libesimp_dot2.so ESIMP_TMA_PROMO=1
