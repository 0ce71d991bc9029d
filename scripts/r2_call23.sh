#!/bin/bash
# round 2, GPU call 23: store-phase operand preloading variants of the sweep kernel (isolated + in step), marching restriction kernel
# The variant libraries were built beforehand (not part of `make`): for v in 0 1 3 7: nvcc -O3 -std=c++17 -lineinfo -gencode
# arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DSW_PRELOAD=$v -c csrc/hex8_structured.cu -o build/hex8_v$v.o, linked with the other objects
# of the library into easysimp.jl_b200/variants/*.so and selected with ESIMP_LIB.
set -u
mkdir -p gpurun_out
O=gpurun_out
V=easysimp.jl_b200/variants
export ESIMP_TMA_PROMO=1 ESIMP_SWEEP_SMOOTH=1
for v in 0 1 3 7; do
  echo "== preload $v" | tee -a $O/r2_c23_pre.log
  ESIMP_LIB=$V/libesimp_pre$v.so timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 20 2>&1 | tail -1 | tee -a $O/r2_c23_pre.log
done
ESIMP_LIB=$V/libesimp_pre3.so timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -m gpu -x -q -k "sweep or gmg or matvec or golden" 2>&1 | tail -4 | tee $O/r2_c23_tests.log
run_bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c23_bench_$name.json 2> $O/r2_c23_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_c23_bench_$name.json"))
print("$name", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"])
for k, v in d.get("roofline_others", {}).items():
    if k.startswith("operator") or "restrict" in k or "prolong" in k: print("   ", k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
}
run_bench pre0 ESIMP_LIB=$V/libesimp_pre0.so
run_bench pre1 ESIMP_LIB=$V/libesimp_pre1.so
run_bench pre3 ESIMP_LIB=$V/libesimp_pre3.so
run_bench pre7 ESIMP_LIB=$V/libesimp_pre7.so
