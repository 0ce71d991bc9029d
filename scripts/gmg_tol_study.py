"""C1 golden loop with the multigrid-preconditioned solve at several PCG tolerances: which rtol keeps the final
density within 1e-3 of the oracle (direct solve) after 50 iterations?"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es
from oracle import simp_oracle as o
gz = np.load(ROOT / "tests/golden/c1_cantilever_60x20x4.npz")
g, dofs, f, pres = o.cantilever_problem((60, 20, 4))
for precond, rtol in [(0, 1e-8), (1, 1e-8), (1, 1e-10), (1, 1e-12), (1, 1e-13)]:
    ctx = es.Context(0)
    ctx.set_params(filter_radius=1.5, pcg_rtol=rtol, precond=precond, pcg_check_every=10)
    ctx.set_grid_structured(60, 20, 4, 1.0, 1.0, 1.0, dofs)
    ctx.set_dirichlet(pres); ctx.set_load(f); ctx.filter_build_structured(); ctx.init_design(None)
    st = [ctx.iterate() for _ in range(50)]
    e = np.array([s.energy for s in st])
    d = np.abs(ctx.get_physical_density() - gz["final_density"])
    print(f"precond {precond} rtol {rtol:g}: compliance rel err {np.max(np.abs(e-gz['energy'])/gz['energy']):.2e}  density Linf {d.max():.2e} (cells > 1e-3: {(d>1e-3).sum()})  pcg iters mean {np.mean([s.pcg_iters for s in st]):.0f} max {max(s.pcg_iters for s in st)}", flush=True)
    ctx.close()
