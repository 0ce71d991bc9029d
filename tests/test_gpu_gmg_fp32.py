"""-m gpu tests of the opt-in FP32 V-cycle preconditioner (csrc/multigrid32.cu, precond = ESIMP_PRECOND_GMG_FP32).

First GPU run (round 2, profiles/r2_gmg_fp32_first_gpu_run.log): every case converged to the FP64 answer, but the
anisotropic 24x12x6 grid needed 285 CG iterations against 184 with the FP64 V-cycle (bar: 1.5x + 2), so the path stays
experimental -- the library refuses it unless ESIMP_EXPERIMENTAL_GMG_FP32=1 -- and these tests run only on request:
    ESIMP_TEST_GMG_FP32=1 python -m pytest tests/test_gpu_gmg_fp32.py -m gpu -q
They hold the path to the same bars as the FP64 V-cycle: solve vs sparse direct 1e-6, loop compliance 1e-6 relative,
densities 1e-3, and an iteration count within 1.5x of the FP64 preconditioner.
"""
import os

import numpy as np
import pytest

from oracle import simp_oracle as o
from test_gpu_parity import _run_loop, make_ctx, relmax

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("ESIMP_TEST_GMG_FP32") != "1",
                                 reason="experimental path: first GPU run (round 2) missed the iteration-count bar on an anisotropic grid")]
if os.environ.get("ESIMP_TEST_GMG_FP32") == "1":
    os.environ["ESIMP_EXPERIMENTAL_GMG_FP32"] = "1"  # the library refuses the experimental preconditioner without it


@pytest.mark.parametrize("nel,right", [((32, 16, 16), (32.0, 16.0, 16.0)), ((40, 20, 20), (2.0, 1.0, 1.0)), ((24, 12, 6), (24.0, 6.0, 1.5))])
@pytest.mark.parametrize("rtol", [1e-8, 1e-10])
def test_gmg_fp32_solve_matches_direct(gpu_ctx_factory, nel, right, rtol):
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(4)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    pres = np.sort(dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel())
    f = o.apply_point_load(np.zeros(3 * g.n_nodes), dofs, sorted(o.select_nodes_by_plane(g, [right[0], 0, 0], [1, 0, 0], 1e-6)), [0.0, -1.0, 0.5])
    iters, sol = {}, {}
    for precond in (1, 2):
        ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=rtol, pcg_check_every=10, precond=precond)
        ctx.set_load(f)
        ctx.set_density(rho)
        sol[precond], iters[precond], relres, st = ctx.solve()
        assert st == 0 and relres <= rtol
        ctx.close()
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    Kbc, fbc, _ = o.apply_dirichlet(o.assemble_stiffness(g, dofs, E, o.unit_element_matrices(g, 0.3, 1.0)), f.copy(), pres)
    u_ref = o.solve_direct(Kbc, fbc)
    assert relmax(sol[2], u_ref) < (1e-6 if rtol <= 1e-10 else 1e-4) and np.all(sol[2][pres] == 0.0)
    assert iters[2] <= 1.5 * iters[1] + 2, iters


@pytest.mark.parametrize("name,problem,n_it", [("c1_cantilever_60x20x4", o.cantilever_problem, 50), ("c2_4legs_40x20x20", o.four_legs_problem, None)])
def test_loop_golden_gmg_fp32(gpu_ctx_factory, golden_dir, name, problem, n_it):
    gz = np.load(golden_dir / f"{name}.npz")
    nel = tuple(int(v) for v in gz["nel"])
    g, dofs, f, pres = problem(nel)
    E0, Emin, nu, p, vf, ratio, move, damp = gz["params"]
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, E0=E0, Emin=Emin, nu=nu, p=p, volume_fraction=vf, move_limit=move, damping=damp,
                   filter_radius=ratio * o.estimate_element_size(g), filter_type=int(gz["filter_type"][0]), pcg_rtol=1e-10,
                   pcg_check_every=10, precond=2)
    ctx.set_load(f)
    ctx.filter_build_structured()
    stats = _run_loop(ctx, int(gz["max_iterations"][0]))
    energy = np.array([s.energy for s in stats])
    assert np.max(np.abs(energy - gz["energy"]) / gz["energy"]) < 1e-6
    e_fin, _, _ = ctx.final_analysis()
    assert abs(e_fin - gz["final_energy"][0]) / gz["final_energy"][0] < 1e-6
    assert np.abs(ctx.get_physical_density() - gz["final_density"]).max() < 1e-3
