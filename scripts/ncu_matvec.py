"""One operator variant on one grid, few launches: the command profiled with ncu (see profiles/)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1:4])
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 2
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rng = np.random.default_rng(0)
ctx = es.Context(0)
ctx.set_params(matvec_variant=variant)
ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
ctx.set_dirichlet(np.arange(3 * (ny + 1) * (nz + 1), dtype=np.int64) * (nx + 1) * 0)
ctx.set_load(rng.standard_normal(ctx.n_dofs))
ctx.set_density(1e-3 + (1 - 1e-3) * rng.random(ctx.n_cells))
ms, nbytes = ctx.bench_matvec(reps)
print(f"variant {variant} {nx}x{ny}x{nz}: {ms:.4f} ms/apply, {nbytes / ms / 1e6:.1f} GB/s algorithmic")
ctx.close()
