#!/bin/bash
# round 2, GPU call 20: sweep-kernel micro-optimisations (3 dot accumulators, predicated arrives, producer suspend hint), the TMA
# L2-promotion sweep, the lagged CG poll; parity first
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -m gpu -x -q -k "sweep or gmg or golden or matvec" 2>&1 | tail -4 | tee $O/r2_c20_tests.log
for p in 3 2 1 0; do
  echo "== ESIMP_TMA_PROMO=$p" | tee -a $O/r2_c20_promo.log
  ESIMP_TMA_PROMO=$p timeout 200 python scripts/bench_operator_modes.py 512 256 256 6 10 2>&1 | tail -1 | tee -a $O/r2_c20_promo.log
done
for p in 3 0; do
  ESIMP_TMA_PROMO=$p timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_c20_bench_promo$p.json 2> $O/r2_c20_bench_promo$p.err; echo "bench promo=$p rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_c20_bench_promo$p.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["config"]["mean_pcg_iters_per_step"], d["gpu_launches"])
for k, v in d.get("roofline_others", {}).items(): print(k, v["launches"], round(v["avg_launch_ms"], 4), round(v["frac"], 3), round(v["share_of_step"], 3))
PY
done
ESIMP_PCG_LAG=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c20_bench_nolag.json 2> $O/r2_c20_bench_nolag.err; echo "bench nolag rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_c20_bench_nolag.json')); print('nolag', d['value'], d['ms_per_step'], d['roofline']['frac'])"
