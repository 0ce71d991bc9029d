"""Quick device-side timing of the operator variants (not the bench contract; see bench.py)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402

PEAK = 6541.8
try:
    PEAK = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def probe(nel, variants, reps=20):
    nx, ny, nz = nel
    ncell = nx * ny * nz
    ndof = 3 * (nx + 1) * (ny + 1) * (nz + 1)
    rng = np.random.default_rng(0)
    rho = 1e-3 + (1 - 1e-3) * rng.random(ncell)
    f = rng.standard_normal(ndof)
    for name, v in variants.items():
        ctx = es.Context(0)
        ctx.set_params(matvec_variant=v)
        ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
        ctx.set_dirichlet(np.arange(3 * (ny + 1) * (nz + 1), dtype=np.int64) * 0)  # dof 0 only
        ctx.set_load(f)
        ctx.set_density(rho)
        ms, nbytes = ctx.bench_matvec(reps)
        print(f"{nel} {name:18s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.1f} GB/s  frac {nbytes / ms / 1e6 / PEAK:.3f}", flush=True)
        ctx.close()


if __name__ == "__main__":
    V = {"fused16x16": 2, "fused_atomic8x32": 3, "fused8x32": 4, "fused8x16": 5}
    probe((240, 80, 80), dict(V, dense_atomic=1))
    t = time.time()
    probe((512, 256, 256), V, reps=10)
    print("c5 probe wall", time.time() - t)
