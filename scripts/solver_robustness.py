"""CG-iteration curve of the multigrid-preconditioned solve over a long SIMP run (the designs turn 0/1, the modulus contrast
reaches 1e9): python scripts/solver_robustness.py c4|c5 N_ITERATIONS -> one JSON line with the per-iteration CG counts."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402
from bench import WORKLOADS, cantilever  # noqa: E402

wl, n_it = sys.argv[1], int(sys.argv[2])
nel = WORKLOADS[wl]
pres, f = cantilever(nel)
ctx = es.Context(0)
ctx.set_params(volume_fraction=0.5, filter_radius=1.5, pcg_rtol=1e-8, pcg_maxit=5000, pcg_check_every=50, warm_start=1, precond=1)
ctx.set_grid_structured(*nel, 1.0, 1.0, 1.0, None)
ctx.set_dirichlet(pres)
ctx.set_load(f)
ctx.filter_build_structured()
ctx.init_design(None)
t0 = time.perf_counter()
st = [ctx.iterate() for _ in range(n_it)]
dt = time.perf_counter() - t0
print(json.dumps({"workload": wl, "grid": list(nel), "simp_iterations": n_it, "seconds": dt, "mean_it_per_s": n_it / dt,
                  "pcg_iters": [s.pcg_iters for s in st], "change": [round(s.change, 4) for s in st], "grayness": [round(s.grayness, 4) for s in st],
                  "energy": [s.energy for s in st], "ms_solve": [round(s.ms_solve, 1) for s in st], "status": [s.status for s in st]}))
ctx.close()
