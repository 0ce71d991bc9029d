"""The command profiled with `ncu --set full -k regex:k_matvec -s 11 -c 10`: 11 clock-raising launches of the plain mode, then two
launches of every store mode of the structured operator (plain, plain+dot, smoother, residual, smoother+dot) on one grid:
    python scripts/ncu_operator_modes.py NX NY NZ [VARIANT]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1:4])
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(0)
ctx = es.Context(0)
ctx.set_params(matvec_variant=variant)
ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
ctx.set_dirichlet(np.arange(3 * (ny + 1) * (nz + 1), dtype=np.int64) * (nx + 1) * 0)
ctx.set_load(rng.standard_normal(ctx.n_dofs))
ctx.set_density(1e-3 + (1 - 1e-3) * rng.random(ctx.n_cells))
ctx.bench_operator_mode(0, 10)
for m in range(5):
    ms, nbytes = ctx.bench_operator_mode(m, 1)
    print(m, round(ms, 4))
ctx.close()
