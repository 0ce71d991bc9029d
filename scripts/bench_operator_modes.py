"""Isolated timing of every store mode of the structured operator kernel on one grid (CUDA events, `reps` launches each):
    python scripts/bench_operator_modes.py NX NY NZ VARIANT [REPS] [MODES]
VARIANT 2 = round-1 owner-computes kernel, 6 = z-sweeping TMA kernel.  Also the command profiled with ncu (profiles/)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1:4])
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
modes = [int(m) for m in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0, 1, 2, 3, 4]
peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
rng = np.random.default_rng(0)
ctx = es.Context(0)
ctx.set_params(matvec_variant=variant)
ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
ctx.set_dirichlet(np.arange(3 * (ny + 1) * (nz + 1), dtype=np.int64) * (nx + 1) * 0)
ctx.set_load(rng.standard_normal(ctx.n_dofs))
ctx.set_density(1e-3 + (1 - 1e-3) * rng.random(ctx.n_cells))
names = ["plain", "plain+dot", "smoother", "residual", "smoother+dot"]
out = {}
ctx.bench_operator_mode(0, 30)  # clocks up before anything is timed
for m in modes:
    ctx.bench_operator_mode(m, 5)
    ms, nbytes = ctx.bench_operator_mode(m, reps)
    out[names[m]] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3)}
print(json.dumps({"grid": [nx, ny, nz], "variant": variant, "reps": reps, "modes": out}))
ctx.close()
