"""-m gpu parity tests: every stage of the CUDA hot path, called through the C ABI (ctypes, host
buffers), against the CPU oracle on the same seeded inputs, against the committed golden fixtures,
and -- at BASELINE sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): element sensitivities and filters 1e-10 relative;
per-iteration compliance 1e-6 relative (PCG rtol stated per test); final density L_inf <= 1e-3.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import simp_oracle as o

pytestmark = pytest.mark.gpu

VARIANTS = {"dense_atomic": 1, "fused16x16": 2, "fused_atomic8x32": 3, "fused8x32": 4, "fused8x16": 5, "sweep": 6}


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def random_problem(nel, right, seed, frac_fixed=0.05):
    g = o.generate_grid_hex(nel, (0.0, 0.0, 0.0), right)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(seed)
    nd = 3 * g.n_nodes
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    u = rng.standard_normal(nd)
    clamp = dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel()
    extra = rng.choice(nd, size=int(frac_fixed * nd), replace=False)
    pres = np.unique(np.concatenate([clamp, extra]))
    return g, dofs, rho, u, pres


def make_ctx(factory, g, dofs, pres=None, **params):
    nx, ny, nz, left, right = g.structured
    ctx = factory()
    if params:
        ctx.set_params(**params)
    ctx.set_grid_structured(nx, ny, nz, (right[0] - left[0]) / nx, (right[1] - left[1]) / ny, (right[2] - left[2]) / nz, dofs)
    if pres is not None:
        ctx.set_dirichlet(pres)
    return ctx


GRIDS = [
    ((60, 20, 4), (60.0, 20.0, 4.0)),      # BASELINE config 1
    ((40, 20, 20), (2.0, 1.0, 1.0)),       # BASELINE config 2 (h = 0.05)
    ((13, 7, 5), (13.0, 3.5, 1.25)),       # ragged sizes, anisotropic cuboid
    ((9, 33, 35), (9.0, 33.0, 35.0)),      # crosses (y,z) tile boundaries of every variant
    ((1, 1, 1), (1.0, 1.0, 1.0)),          # minimum size
]


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("nel,right", GRIDS)
def test_matvec_matches_oracle(gpu_ctx_factory, nel, right, variant):
    g, dofs, rho, u, pres = random_problem(nel, right, seed=sum(nel))
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, matvec_variant=VARIANTS[variant])
    ctx.set_density(rho)
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    d = o.operator_diagonal(g, dofs, E, k0)
    y_ref = o.operator_apply(g, dofs, E, k0, pres, u, float(np.mean(d)))
    y = ctx.matvec(u)
    assert relmax(y, y_ref) < 2e-13
    # Jacobi diagonal of the BC-modified operator
    keep = np.ones_like(d)
    keep[pres] = 0
    d_ref = d * keep + float(np.mean(d)) * (1 - keep)
    assert relmax(ctx.diagonal(), d_ref) < 1e-13


def test_matvec_native_numbering(gpu_ctx_factory):
    """dof_of_node = NULL: caller numbering is 3*node + component (synthetic benchmark path)."""
    g, dofs, rho, u, pres = random_problem((10, 6, 7), (10.0, 6.0, 7.0), seed=5)
    nat = 3 * np.arange(g.n_nodes)[:, None] + np.arange(3)[None, :]
    ctx = make_ctx(gpu_ctx_factory, g, None)
    pres_nat = np.unique(nat.ravel()[np.isin(dofs.ravel(), pres)])
    ctx.set_dirichlet(pres_nat)
    ctx.set_density(rho)
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    d = o.operator_diagonal(g, nat, E, k0)
    y_ref = o.operator_apply(g, nat, E, k0, pres_nat, u, float(np.mean(d)))
    assert relmax(ctx.matvec(u), y_ref) < 2e-13


@pytest.mark.parametrize("nel,right,rtol", [((60, 20, 4), (60.0, 20.0, 4.0), 1e-10), ((20, 10, 10), (2.0, 1.0, 1.0), 1e-10)])
def test_solve_matches_direct(gpu_ctx_factory, nel, right, rtol):
    """PCG (rtol stated) against the oracle's sparse direct solve of the assembled, BC-modified K."""
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(3)
    rho = 0.05 + 0.95 * rng.random(g.n_cells)
    fixed = o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)
    pres = np.sort(dofs[np.array(sorted(fixed))].ravel())
    f = rng.standard_normal(3 * g.n_nodes)
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=rtol, pcg_check_every=10)
    ctx.set_load(f)
    ctx.set_density(rho)
    u, iters, relres, st = ctx.solve()
    assert st == 0 and relres <= rtol
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    Kbc, fbc, _ = o.apply_dirichlet(o.assemble_stiffness(g, dofs, E, k0), f.copy(), pres)
    u_ref = o.solve_direct(Kbc, fbc)
    assert relmax(u, u_ref) < 1e-6
    assert np.all(u[pres] == 0.0)
    # energy = 0.5 u'Ku = 0.5 f'u at the solution (Optimization.jl:320)
    dc, energy = ctx.sensitivities()
    assert abs(energy - 0.5 * fbc @ u_ref) / abs(0.5 * fbc @ u_ref) < 1e-7
    # a solved field is dominated by rigid-body motion of each cell (|u|^2 K0 >> u_e'K0u_e), so the dense
    # u_e'K0u_e of the oracle itself cancels ~8 digits here; the 1e-10 bar is tested on generic u below
    dc_ref = o.sensitivities(g, dofs, rho, u, 1.0, 1e-9, 0.3, 3.0, k0)
    assert relmax(dc, dc_ref) < 1e-8


@pytest.mark.parametrize("nel,right", GRIDS[:4])
def test_sensitivities_match_oracle(gpu_ctx_factory, nel, right):
    """u_e'K0u_e and dc for an arbitrary displacement field: 1e-10 relative (north_star)."""
    g, dofs, rho, u, _ = random_problem(nel, right, seed=11)
    ctx = make_ctx(gpu_ctx_factory, g, dofs, np.zeros(0, dtype=np.int64), E0=2.0, Emin=1e-6, p=3.0)
    ctx.set_density(rho)
    ctx.set_displacement(u)
    dc, energy = ctx.sensitivities()
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    dc_ref = o.sensitivities(g, dofs, rho, u, 2.0, 1e-6, 0.3, 3.0, k0)
    assert relmax(dc, dc_ref) < 1e-10
    e_ref = 0.5 * float(o.simp_modulus(rho, 2.0, 1e-6, 3.0) @ o.element_energies(g, dofs, u, k0))
    assert abs(energy - e_ref) / e_ref < 1e-12


def test_sensitivities_noninteger_p(gpu_ctx_factory):
    g, dofs, rho, u, _ = random_problem((8, 6, 5), (8.0, 6.0, 5.0), seed=2)
    ctx = make_ctx(gpu_ctx_factory, g, dofs, np.zeros(0, dtype=np.int64), p=2.5)
    ctx.set_density(rho)
    ctx.set_displacement(u)
    dc, _ = ctx.sensitivities()
    assert relmax(dc, o.sensitivities(g, dofs, rho, u, 1.0, 1e-9, 0.3, 2.5)) < 1e-10


@pytest.mark.parametrize("ratio", [1.0, 1.5, 2.0, 2.5, 3.2])
@pytest.mark.parametrize("nel,right", [((60, 20, 4), (60.0, 20.0, 4.0)), ((20, 10, 10), (2.0, 1.0, 1.0)), ((7, 5, 9), (7.0, 2.5, 18.0))])
def test_filters_match_oracle(gpu_ctx_factory, nel, right, ratio):
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    V = o.element_volumes(g)
    fc = o.create_filter_cache(g, ratio, V)
    rng = np.random.default_rng(7)
    rho = 1e-4 + rng.random(g.n_cells)  # some cells below the 1e-3 guard of the sensitivity filter
    rho[:5] = 1e-4
    dc = -rng.random(g.n_cells) * 10.0 ** rng.uniform(-6, 2, g.n_cells)
    ctx = make_ctx(gpu_ctx_factory, g, None, None, filter_radius=fc.radius)
    ctx.filter_build_structured()
    assert relmax(ctx.filter_sens(rho, dc), o.filter_sensitivity(fc, rho, dc)) < 1e-10
    assert relmax(ctx.filter_dens(rho), o.filter_density(fc, rho)) < 1e-10
    assert relmax(ctx.filter_dens_T(dc), o.filter_density_transpose(fc, dc)) < 1e-10


@pytest.mark.parametrize("ftype", ["sensitivity", "density"])
def test_oc_matches_oracle(gpu_ctx_factory, ftype):
    g = o.generate_grid_hex((30, 10, 6))
    V = o.element_volumes(g)
    fc = o.create_filter_cache(g, 1.5, V)
    rng = np.random.default_rng(9)
    rho = np.clip(0.5 + 0.3 * rng.standard_normal(g.n_cells), 1e-3, 1.0)
    dc = -rng.random(g.n_cells) * 50.0
    use_df = ftype == "density"
    dv = o.filter_density_transpose(fc, V / V.sum()) if use_df else V / V.sum()
    ctx = make_ctx(gpu_ctx_factory, g, None, None, filter_radius=fc.radius, filter_type=int(use_df), volume_fraction=0.45)
    ctx.filter_build_structured()
    new, lam, steps, st = ctx.oc(rho, dc, dv)
    ref, lam_ref, steps_ref, conv = o.optimality_criteria_update(rho, dc, dv, 0.45, V.sum(), V, 0.2, 0.5, fc, use_df)
    # the stop test |dV| < 1e-6 is absolute (OptimalityCriteria.jl:128): a different summation order
    # may stop one bisection step earlier or later, which moves lambda by one interval halving
    assert abs(steps - steps_ref) <= 1 and (st == 0) == conv
    assert abs(lam - lam_ref) <= 1e-6 * lam_ref
    assert np.abs(new - ref).max() < 1e-6
    assert abs(float(o.filter_density(fc, new) @ V if use_df else new @ V) - 0.45 * V.sum()) < 1e-5


def _run_loop(ctx, n_iter):
    ctx.init_design(None)
    return [ctx.iterate() for _ in range(n_iter)]


def _golden_case(gpu_ctx_factory, golden_dir, name, problem, pcg_rtol, precond=0, check_max_u=True):
    gz = np.load(golden_dir / f"{name}.npz")
    nel = tuple(int(v) for v in gz["nel"])
    g, dofs, f, pres = problem(nel)
    E0, Emin, nu, p, vf, ratio, move, damp = gz["params"]
    R = ratio * o.estimate_element_size(g)
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, E0=E0, Emin=Emin, nu=nu, p=p, volume_fraction=vf, move_limit=move,
                   damping=damp, filter_radius=R, filter_type=int(gz["filter_type"][0]), pcg_rtol=pcg_rtol, pcg_check_every=10, precond=precond)
    ctx.set_load(f)
    ctx.filter_build_structured()
    n_it = int(gz["max_iterations"][0])
    stats = _run_loop(ctx, n_it)
    energy = np.array([s.energy for s in stats])
    # per-iteration compliance within 1e-6 relative (north_star), PCG rtol = pcg_rtol
    assert np.max(np.abs(energy - gz["energy"]) / gz["energy"]) < 1e-6
    assert np.max(np.abs(np.array([s.volume for s in stats]) - gz["volume"]) / gz["volume"]) < 1e-9
    assert np.max(np.abs(np.array([s.change for s in stats]) - gz["change"])) < 1e-4
    # the other logged scalars of the loop (Optimization.jl:374-379, OptimizationLogger.jl:50-60): grayness = fraction of cells
    # with 0.1 < rho~ < 0.9 (a cell within 1e-3 of a threshold may fall either side: a handful of cells), max |u| like the
    # compliance, and the Lagrange multiplier the bisection ends on (same midpoints; the |dV| < 1e-6 break may come a step apart)
    n_cells = g.n_cells
    assert np.max(np.abs(np.array([s.grayness for s in stats]) - gz["grayness"])) <= 5.0 / n_cells + 1e-12
    if check_max_u:  # not defined where K(rho) with its constraints is singular (rigid modes left free): any u + null vector solves it
        assert np.max(np.abs(np.array([s.max_u for s in stats]) - gz["max_u"]) / gz["max_u"]) < 1e-6
    assert np.max(np.abs(np.array([s.lam for s in stats]) - gz["lam"]) / gz["lam"]) < 1e-4
    # ... normally: where |dV| of an earlier midpoint sits at the absolute 1e-6 threshold itself (5e-10 of the volume), rounding decides
    # whether the bisection stops there or runs on to the interval-width floor (seen on a B200: 51 against 57 steps in one of 30
    # iterations of example 02), and lambda still agrees as asserted above -- so the step counts must agree to +-2 in 9 of 10 iterations
    d_steps = np.abs(np.array([s.oc_steps for s in stats]) - gz["oc_steps"])
    assert np.median(d_steps) == 0 and np.count_nonzero(d_steps > 2) <= max(1, n_it // 10)
    e_fin, v_fin, _ = ctx.final_analysis()
    assert abs(e_fin - gz["final_energy"][0]) / gz["final_energy"][0] < 1e-6
    # final density field L_inf <= 1e-3 after the fixed iteration count (north_star)
    dens = ctx.get_physical_density()
    assert np.abs(dens - gz["final_density"]).max() < 1e-3
    return stats, dens


def test_loop_c1_cantilever_golden(gpu_ctx_factory, golden_dir):
    """BASELINE config 1: 60x20x4, vf 0.5, p 3, r 1.5, sensitivity filter, 50 iterations."""
    _golden_case(gpu_ctx_factory, golden_dir, "c1_cantilever_60x20x4", o.cantilever_problem, 1e-10)


def test_loop_c1b_runtests_golden(gpu_ctx_factory, golden_dir):
    """test/runtests.jl:73-84 variant: E0=200, Emin=1e-6, vf 0.4, r 2.5, 20 iterations."""
    ctx_density = _golden_case(gpu_ctx_factory, golden_dir, "c1b_runtests_60x20x4", o.cantilever_problem, 1e-10)[1]
    # ... and the design the CUDA path grows is the truss of the reference's own data/cantilever_beam_opt.vtu
    # (qualitative anchor, see tests/test_oracle.py::test_reference_density_artifact_topology)
    from test_oracle import reference_artifact_agreement

    same, ref_cov, own_cov = reference_artifact_agreement(ctx_density, golden_dir)
    assert same > 0.90 and ref_cov > 0.95 and own_cov > 0.95


def test_loop_c2_small_density_filter_golden(gpu_ctx_factory, golden_dir):
    _golden_case(gpu_ctx_factory, golden_dir, "c2s_4legs_20x10x10", o.four_legs_problem, 1e-10)


def test_loop_c2_4legs_golden(gpu_ctx_factory, golden_dir):
    """BASELINE config 2: 40x20x20 over 2x1x1, density filter r=2.0."""
    _golden_case(gpu_ctx_factory, golden_dir, "c2_4legs_40x20x20", o.four_legs_problem, 1e-10)


@pytest.mark.parametrize("name,problem", [("e02_sliding_support_60x20x4", "sliding_support_problem"), ("e06_mbb_20x10x10", "mbb_problem"),
                                          ("e08_michell_half_20x10x10", "michell_half_problem")])
@pytest.mark.parametrize("precond", [0, 1])
def test_loop_example_scripts_with_sliding_constraints_golden(gpu_ctx_factory, golden_dir, name, problem, precond):
    """test/Examples/02_sliding_support.jl, 06_3D_2x1x1_MBB.jl, 08_3D_2x1x1_Michell-half_tol_study.jl (the latter two at half
    resolution): apply_sliding_boundary! handlers, i.e. nodes with one prescribed component, alone and mixed with clamped
    patches -- the Dirichlet masks of the operator, the diagonal, the multigrid levels and the fix-up see per-component sets.
    Jacobi and multigrid preconditioner."""
    # example 02 constrains U1 on the face x = 0 and U2 at ONE node: the z translation and a rotation about the x axis stay free, K is
    # singular and the load is consistent with it (no resultant along z, no moment about x).  Compliance, sensitivities and densities
    # do not see the rigid part of u; max|u| does (tests/test_oracle.py::test_sliding_support_example_is_singular_but_consistent)
    stats, _ = _golden_case(gpu_ctx_factory, golden_dir, name, getattr(o, problem), 1e-10, precond=precond, check_max_u=not name.startswith("e02"))
    assert max(s.pcg_iters for s in stats) < (400 if precond else 20000)


def test_host_api_simp_optimize_matches_oracle():
    """The reference-shaped host API (generate_grid .. simp_optimize) against a live oracle run."""
    import easysimp_b200 as es

    grid = es.generate_grid(es.Hexahedron, (24, 8, 4), (0.0, 0.0, 0.0), (24.0, 8.0, 4.0))
    dh, cellvalues, K, f = es.setup_problem(grid, device=0)
    fixed = es.select_nodes_by_plane(grid, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    force = es.select_nodes_by_circle(grid, [24.0, 0.0, 2.0], [1.0, 0.0, 0.0], 1.0)
    mm = es.create_simp_material_model(1.0, 0.3, 1e-9, 3.0)
    es.assemble_stiffness_matrix_simp_(K, f, dh, cellvalues, mm, np.full(grid.getncells(), 0.4))
    ch = es.apply_fixed_boundary_(K, f, dh, fixed)
    params = es.OptimizationParameters(max_iterations=8, tolerance=0.0, pcg_rtol=1e-10)
    res = es.simp_optimize(grid, dh, cellvalues, [es.PointLoad(dh, sorted(force), [0.0, -1.0, 0.0])], [ch], params, K=K)
    g, dofs, fo, pres = o.cantilever_problem((24, 8, 4))
    assert np.array_equal(dofs, dh.dof_of_node) and np.array_equal(np.sort(ch.prescribed_dofs), pres)
    ref = o.simp_optimize(g, dofs, fo, pres, o.OptimizationParameters(max_iterations=8, tolerance=0.0))
    e_ref = np.array([h.energy for h in ref["history"]])
    assert np.max(np.abs(np.array(res.energy_history) - e_ref) / e_ref) < 1e-6
    assert np.abs(res.densities - ref["densities"]).max() < 1e-3
    assert relmax(res.displacements, ref["displacements"]) < 1e-5
    assert res.iterations == 8 and not res.converged
    with pytest.raises(ValueError):
        es.apply_force_(np.zeros(dh.ndofs), dh, [], [0.0, -1.0, 0.0])  # error() in FEA.jl:358-360
    K.ctx.close()


# ----------------------------------------------------------------------------- multigrid-preconditioned CG
@pytest.mark.parametrize("variant", [0, 6])  # 6: the sweep kernel (all store modes) on every level with an even nz
@pytest.mark.parametrize("nel,right", [((32, 16, 16), (32.0, 16.0, 16.0)), ((40, 20, 20), (2.0, 1.0, 1.0)), ((24, 12, 6), (24.0, 6.0, 1.5))])
def test_gmg_solve_matches_direct(gpu_ctx_factory, nel, right, variant):
    """CG with the V-cycle preconditioner reaches the same solution as the sparse direct solve, in far fewer
    iterations than Jacobi-PCG (random high-contrast modulus field, rtol 1e-10)."""
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(4)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    pres = np.sort(dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel())
    f = o.apply_point_load(np.zeros(3 * g.n_nodes), dofs, sorted(o.select_nodes_by_plane(g, [right[0], 0, 0], [1, 0, 0], 1e-6)), [0.0, -1.0, 0.5])
    iters = {}
    for precond in (0, 1):
        ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=1e-10, pcg_check_every=10, precond=precond, matvec_variant=variant)
        ctx.set_load(f)
        ctx.set_density(rho)
        u, iters[precond], relres, st = ctx.solve()
        assert st == 0 and relres <= 1e-10
        if precond == 1:
            E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
            Kbc, fbc, _ = o.apply_dirichlet(o.assemble_stiffness(g, dofs, E, o.unit_element_matrices(g, 0.3, 1.0)), f.copy(), pres)
            assert relmax(u, o.solve_direct(Kbc, fbc)) < 1e-6 and np.all(u[pres] == 0.0)
        ctx.close()
    assert iters[1] * 5 < iters[0], iters


def test_gmg_lagged_poll_returns_the_same_iterate(gpu_ctx_factory, monkeypatch):
    """The lagged convergence poll of the multigrid-preconditioned CG (the host enqueues iteration k before it looks at <r,r> of
    iteration k-1, csrc/multigrid.cu) stops at the same iteration with bit-identical displacements as the blocking poll."""
    nel, right = (32, 16, 16), (32.0, 16.0, 16.0)
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(5)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    pres = np.sort(dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel())
    f = o.apply_point_load(np.zeros(3 * g.n_nodes), dofs, sorted(o.select_nodes_by_plane(g, [right[0], 0, 0], [1, 0, 0], 1e-6)), [0.0, -1.0, 0.5])
    res = {}
    for lag in ("0", "1"):
        monkeypatch.setenv("ESIMP_PCG_LAG", lag)
        for rtol in (1e-4, 1e-10):
            ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=rtol, precond=1)
            ctx.set_load(f)
            ctx.set_density(rho)
            u, it, relres, st = ctx.solve()
            assert st == 0 and relres <= rtol
            res[lag, rtol] = (u, it, relres)
            ctx.close()
    for rtol in (1e-4, 1e-10):
        assert res["0", rtol][1] == res["1", rtol][1] and res["0", rtol][2] == res["1", rtol][2]
        assert np.array_equal(res["0", rtol][0], res["1", rtol][0])
    # maxit reached inside the lagged loop: blocking fall-back on the last iteration, warning status
    monkeypatch.setenv("ESIMP_PCG_LAG", "1")
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=1e-14, pcg_maxit=5, precond=1)
    ctx.set_load(f)
    ctx.set_density(rho)
    _, it, _, st = ctx.solve()
    assert it == 5 and st != 0
    ctx.close()


@pytest.mark.parametrize("nel,right", [((64, 32, 32), (64.0, 32.0, 32.0)), ((48, 24, 12), (4.0, 2.0, 1.0))])
def test_gmg_cluster_tail_equals_level_by_level_cycle(gpu_ctx_factory, monkeypatch, nel, right):
    """The bottom levels of the V-cycle run as one thread-block-cluster kernel (csrc/coarse_tail.cuh).  It is the same preconditioner
    as the level-by-level path up to summation order: same CG iteration count (+-1) and the same displacement."""
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(6)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells) ** 3
    pres = np.sort(dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel())
    f = o.apply_point_load(np.zeros(3 * g.n_nodes), dofs, sorted(o.select_nodes_by_plane(g, [right[0], 0, 0], [1, 0, 0], 1e-6)), [0.0, -1.0, 0.5])
    res = {}
    for tail in ("0", "1"):
        monkeypatch.setenv("ESIMP_MG_TAIL", tail)
        ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=1e-10, precond=1)
        ctx.set_load(f)
        ctx.set_density(rho)
        u, it, relres, st = ctx.solve()
        assert st == 0 and relres <= 1e-10
        res[tail] = (u, it)
        ctx.close()
    assert abs(res["0"][1] - res["1"][1]) <= 1, (res["0"][1], res["1"][1])
    assert relmax(res["0"][0], res["1"][0]) < 1e-7


def test_loop_c1_cantilever_golden_gmg(gpu_ctx_factory, golden_dir):
    """BASELINE config 1 with the multigrid-preconditioned solve: same golden bars (compliance 1e-6, density 1e-3)."""
    gz = np.load(golden_dir / "c1_cantilever_60x20x4.npz")
    g, dofs, f, pres = o.cantilever_problem((60, 20, 4))
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, filter_radius=1.5 * o.estimate_element_size(g), pcg_rtol=1e-10, precond=1)
    ctx.set_load(f)
    ctx.filter_build_structured()
    stats = _run_loop(ctx, 50)
    energy = np.array([s.energy for s in stats])
    assert np.max(np.abs(energy - gz["energy"]) / gz["energy"]) < 1e-6
    e_fin, _, _ = ctx.final_analysis()  # refreshes the physical densities from the last design (Optimization.jl:498-503)
    assert abs(e_fin - gz["final_energy"][0]) / gz["final_energy"][0] < 1e-6
    assert np.abs(ctx.get_physical_density() - gz["final_density"]).max() < 1e-3
    assert max(s.pcg_iters for s in stats) < 200


# ----------------------------------------------------------------------------- properties at BASELINE sizes
@pytest.mark.parametrize("nel", [(240, 80, 80)])
def test_operator_properties_c4(gpu_ctx_factory, nel):
    """Size-independent properties at config 4 (240x80x80, 4.7 M dofs): symmetry <Ax,y> = <x,Ay>,
    linearity, rigid translations in the null space of the unconstrained operator, and agreement of
    the kernel variants with each other."""
    nx, ny, nz = nel
    n_nodes = (nx + 1) * (ny + 1) * (nz + 1)
    n_cells = nx * ny * nz
    idx = np.arange(3 * n_nodes, dtype=np.uint64)
    x = 2 * o.splitmix64(idx, 1) - 1
    y = 2 * o.splitmix64(idx, 2) - 1
    rho = 1e-3 + (1 - 1e-3) * o.splitmix64(np.arange(n_cells, dtype=np.uint64), 3)
    results = {}
    for name in ("fused16x16", "fused_atomic8x32", "fused8x32", "fused8x16", "sweep"):
        ctx = gpu_ctx_factory()
        ctx.set_params(matvec_variant=VARIANTS[name])
        ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
        ctx.set_dirichlet(np.zeros(0, dtype=np.int64))
        ctx.set_density(rho)
        Ax, Ay = ctx.matvec(x), ctx.matvec(y)
        assert abs(Ax @ y - x @ Ay) / abs(Ax @ y) < 1e-11
        assert relmax(ctx.matvec(2.0 * x - 3.0 * y), 2.0 * Ax - 3.0 * Ay) < 1e-12
        t = np.zeros(3 * n_nodes)
        t[1::3] = 1.0  # rigid y translation
        assert np.abs(ctx.matvec(t)).max() < 1e-11
        results[name] = Ax
        ctx.close()
    assert relmax(results["fused_atomic8x32"], results["fused16x16"]) < 1e-12
    assert relmax(results["fused8x32"], results["fused16x16"]) < 1e-12
    assert relmax(results["fused8x16"], results["fused16x16"]) < 1e-12
    assert relmax(results["sweep"], results["fused16x16"]) < 1e-12


def test_first_simp_iteration_c4_against_cpu_port(gpu_ctx_factory):
    """One SIMP iteration at BASELINE config 4 (240x80x80, 1.536 M cells, 4.74 M dofs), stage by stage, against the C / OpenMP
    restatement oracle/cpu_simp.c (itself held against the numpy oracle in tests/test_cpu_port.py): the CUDA solve
    (multigrid CG, rtol 1e-10) satisfies the CPU port's system to 1e-9, compliance 1e-6, element sensitivities and the
    sensitivity filter 1e-10, the OC update ends on the same multiplier and the same design."""
    from bench import cantilever
    from oracle import cpu_port as cp

    nel = (240, 80, 80)
    nx, ny, nz = nel
    pres, f = cantilever(nel)
    n_cells = nx * ny * nz
    rho = np.full(n_cells, 0.5)
    ctx = gpu_ctx_factory()
    ctx.set_params(volume_fraction=0.5, filter_radius=1.5, pcg_rtol=1e-10, pcg_maxit=20000, pcg_check_every=5, precond=1)
    ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.filter_build_structured()
    ctx.set_density(rho)
    u, iters, relres, st = ctx.solve()
    assert st == 0 and relres <= 1e-10
    dc, energy = ctx.sensitivities()
    dcf = ctx.filter_sens(rho, dc)
    dv = np.full(n_cells, 1.0 / n_cells)
    rho_new, lam, steps, _ = ctx.oc(rho, dcf, dv)
    ctx.close()
    # ---- the CPU port on the same inputs
    g1 = o.generate_grid_hex((1, 1, 1))
    K0 = np.ascontiguousarray(o.hex8_ke(g1.coords[g1.cells[0]], *o.lame(1.0, 0.3)))
    E = cp.modulus(rho, 1.0, 1e-9, 3.0)
    fixed = np.zeros(f.size, dtype=np.uint8)
    fixed[pres] = 1
    Ku = cp.matvec(nel, K0, E, fixed, 1.0, u)
    free = fixed == 0
    assert np.linalg.norm((f - Ku)[free]) / np.linalg.norm(f[free]) < 1e-9      # u solves the CPU port's system
    assert abs(0.5 * float(u @ Ku) - energy) / energy < 1e-6                    # compliance (Optimization.jl:320)
    dc_cpu, en_cpu = cp.sensitivities(nel, K0, rho, E, u, 1.0, 1e-9, 3.0)
    assert relmax(dc, dc_cpu) < 1e-10 and abs(en_cpu - energy) / energy < 1e-9
    assert relmax(dcf, cp.filter_sens(nel, (1.0, 1.0, 1.0), 1.5, rho, dc_cpu)) < 1e-10
    lo, hi, new_cpu, lam_cpu = 1e-9, 1e9, None, 0.0
    for _ in range(200):  # OptimalityCriteria.jl:98-137 on the CPU port's candidate / volume kernel
        lam_cpu = 0.5 * (lo + hi)
        new_cpu, v = cp.oc_step(rho, dcf, dv, lam_cpu, 0.2, 0.5, 1.0)
        err = v - 0.5 * n_cells
        if abs(err) < 1e-6:
            break
        lo, hi = (lam_cpu, hi) if err > 0 else (lo, lam_cpu)
    # at this size the volume sums carry ~1e-4 of absolute rounding noise, far above the |dV| < 1e-6 break: both bisections run until
    # the bracket collapses, possibly a few steps apart
    assert abs(lam - lam_cpu) / lam_cpu < 1e-4 and np.abs(rho_new - new_cpu).max() < 1e-4


def test_solution_properties_c4(gpu_ctx_factory):
    """240x80x80 cantilever: energy 0.5 u'Ku equals 0.5 f'u at the PCG solution, u = 0 on the clamp."""
    nx, ny, nz = 240, 80, 80
    n_nodes = (nx + 1) * (ny + 1) * (nz + 1)
    node = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)  # noqa: E731
    jj, kk = np.meshgrid(np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    clamp = node(0, jj.ravel(), kk.ravel())
    pres = np.sort((3 * clamp[:, None] + np.arange(3)[None, :]).ravel())
    f = np.zeros(3 * n_nodes)
    load_nodes = [node(nx, 0, nz // 2), node(nx, 0, nz // 2 - 1), node(nx, 0, nz // 2 + 1), node(nx, 1, nz // 2)]
    f[3 * np.array(load_nodes) + 1] = -0.25
    ctx = gpu_ctx_factory()
    ctx.set_params(pcg_rtol=1e-8, pcg_maxit=20000, pcg_check_every=50)
    ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.set_density(np.full(nx * ny * nz, 0.5))
    u, iters, relres, st = ctx.solve()
    assert st == 0 and relres <= 1e-8
    _, energy = ctx.sensitivities(want_dc=False)
    assert abs(energy - 0.5 * f @ u) / energy < 1e-6
    assert np.all(u[pres] == 0.0)
    ctx.close()


def test_operator_properties_c5_full_size(gpu_ctx_factory):
    """BASELINE config 5 (512x256x256, 101.6 M dofs) on one GPU: symmetry and the rigid-translation null space of the
    unconstrained operator."""
    nx, ny, nz = 512, 256, 256
    n_nodes = (nx + 1) * (ny + 1) * (nz + 1)
    idx = np.arange(3 * n_nodes, dtype=np.uint64)
    x = 2 * o.splitmix64(idx, 11) - 1
    y = 2 * o.splitmix64(idx, 12) - 1
    del idx
    rho = 1e-3 + (1 - 1e-3) * o.splitmix64(np.arange(nx * ny * nz, dtype=np.uint64), 13)
    ctx = gpu_ctx_factory()
    ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
    ctx.set_dirichlet(np.zeros(0, dtype=np.int64))
    ctx.set_density(rho)
    Ax = ctx.matvec(x)
    lhs = float(Ax @ y)
    del Ax
    Ay = ctx.matvec(y)
    assert abs(lhs - float(x @ Ay)) / abs(lhs) < 1e-10
    del Ay, x
    y[:] = 0.0
    y[2::3] = 1.0  # rigid z translation
    assert np.abs(ctx.matvec(y)).max() < 1e-10
    ctx.close()


def test_host_api_readme_script_c5_full_size():
    """the reference's README script (README.md:110-153) through the reference-shaped host calls at BASELINE
    config 5 -- generate_grid(512, 256, 256), setup_problem (Ferrite's first-encounter dof numbering in closed form, operator
    handle instead of a sparse K), node pickers, assemble / apply_fixed_boundary! on the handle, simp_optimize, create_results_data.
    The compliance of the first iterations must equal what the same problem gives when it is driven straight through the C ABI in
    the library's native numbering (bench.py's history, profiles/r2_end_bench_c5_n1.json: PCG rtol 1e-8 on both sides), the
    volume constraint must hold and the caller-side setup must stay a matter of seconds (scripts/host_mirror_c5.py prints it)."""
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "scripts"))
    import host_mirror_c5

    import json
    import os

    res = host_mirror_c5.run((512, 256, 256), 2)
    print(res)
    if os.environ.get("ESIMP_HOST_MIRROR_OUT"):  # evidence file of a builder run (profiles/r2_host_mirror_c5.json)
        Path(os.environ["ESIMP_HOST_MIRROR_OUT"]).write_text(json.dumps(res) + "\n")
    assert res["n_dofs"] == 101649411 and res["n_fixed_nodes"] == 257 * 257 and res["n_force_nodes"] == 4
    native = [6.411570549436242, 2.4987159902172644]
    for it, e_ref in zip(res["iterations"], native):
        assert abs(it["energy"] - e_ref) / e_ref < 1e-6
        assert abs(it["volume"] / (512.0 * 256.0 * 256.0) - 0.5) < 1e-6
    assert abs(res["final_volume_fraction"] - 0.5) < 1e-6 and res["max_von_mises"] > 0.0 and res["max_displacement"] > 0.0
    setup = sum(v for k, v in res["host_seconds"].items() if k not in ("simp_optimize", "create_results_data"))
    assert setup < 120.0


def test_error_paths_and_edge_cases(gpu_ctx_factory):
    """Error behaviour at the boundary (negative status + message, no crash) and degenerate inputs."""
    import easysimp_b200 as es

    g, dofs, f, pres = o.cantilever_problem((8, 4, 2))
    ctx = gpu_ctx_factory()
    with pytest.raises(es.EsimpError, match="no mesh"):
        # straight through the C ABI: the Python wrapper would trip over n_dofs == 0 first
        ctx._chk(ctx.lib.esimp_set_load(ctx._h, np.zeros(3).ctypes.data_as(es.lib._D)), "esimp_set_load")
    ctx.set_grid_structured(8, 4, 2, 1.0, 1.0, 1.0, dofs)
    with pytest.raises(es.EsimpError, match="set_density"):
        ctx.matvec(np.zeros(3 * g.n_nodes))
    with pytest.raises(es.EsimpError, match="out of range"):
        ctx.set_dirichlet(np.array([3 * g.n_nodes + 5]))
    with pytest.raises(es.EsimpError, match="invalid material"):
        ctx.set_params(nu=0.5)
    ctx.set_params(nu=0.3)
    with pytest.raises(es.EsimpError, match="filter"):
        ctx.init_design(None)  # filter not built yet
    # empty Dirichlet set and zero load: u = 0, zero iterations
    ctx.set_dirichlet(np.zeros(0, dtype=np.int64))
    ctx.set_load(np.zeros(3 * g.n_nodes))
    ctx.set_density(np.full(g.n_cells, 0.5))
    u, iters, relres, st = ctx.solve()
    assert iters == 0 and np.all(u == 0.0)
    # OC with vanishing sensitivities: every candidate sits on the lower move limit, the volume target is never met,
    # the bisection runs out of its 200 steps like the reference (@warn at OptimalityCriteria.jl:140-142)
    ctx.filter_build_structured()
    V = o.element_volumes(g)
    rho = np.full(g.n_cells, 0.5)
    new, lam, steps, st = ctx.oc(rho, np.zeros(g.n_cells), V / V.sum())
    ref, lam_ref, steps_ref, conv = o.optimality_criteria_update(rho, np.zeros(g.n_cells), V / V.sum(), 0.5, V.sum(), V)
    assert steps == steps_ref == 200 and st == es.lib.ESIMP_WARN_OC_MAXIT and not conv
    assert np.array_equal(new, ref) and lam == lam_ref
