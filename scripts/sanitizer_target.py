"""Small workload that drives every store mode of both structured operator kernels (round-1 owner-computes kernel and the
z-sweeping TMA kernel, double precision) plus the transfer / vector kernels of the multigrid CG, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitizer_target.py
    compute-sanitizer --tool racecheck python scripts/sanitizer_target.py
Prints the solver iteration counts and the agreement of the two kernels; exits non-zero on disagreement."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402

nel = (16, 32, 64)  # 2 y tiles of the sweep kernel, 2 z segments incl. the extra-column path; 3+ multigrid levels
nx, ny, nz = nel
node = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)  # noqa: E731
jj, kk = np.meshgrid(np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
pres = np.sort((3 * node(0, jj.ravel(), kk.ravel())[:, None] + np.arange(3)[None, :]).ravel()).astype(np.int64)
n_nodes = (nx + 1) * (ny + 1) * (nz + 1)
f = np.zeros(3 * n_nodes)
f[3 * node(nx, jj.ravel(), kk.ravel()) + 1] = -1.0 / jj.size
rng = np.random.default_rng(5)
rho = 1e-3 + (1 - 1e-3) * rng.random(nx * ny * nz)
u_test = rng.standard_normal(3 * n_nodes)
sol, its, mv = {}, {}, {}
for variant in (2, 6):
    ctx = es.Context(0)
    ctx.set_params(pcg_rtol=1e-8, pcg_check_every=1, precond=1, matvec_variant=variant, filter_radius=1.5)
    ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.filter_build_structured()
    ctx.set_density(rho)
    mv[variant] = ctx.matvec(u_test)                      # plain mode
    sol[variant], its[variant], relres, st = ctx.solve()  # dot, smoother, smoother + dot, residual modes on every level
    ctx.init_design(None)
    s = ctx.iterate()                                     # sensitivities, filter, OC, metrics
    ctx.close()
    print(f"variant {variant}: {its[variant]} CG iterations, relres {relres:.2e}, one SIMP iteration energy {s.energy:.6e}", flush=True)
e1 = np.abs(mv[6] - mv[2]).max() / np.abs(mv[2]).max()
e2 = np.abs(sol[6] - sol[2]).max() / np.abs(sol[2]).max()
print(f"sweep vs round-1 kernel: matvec {e1:.1e}, solution {e2:.1e}")
sys.exit(0 if e1 < 1e-12 and e2 < 1e-6 and abs(its[6] - its[2]) <= 1 else 1)
