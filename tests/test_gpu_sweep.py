"""-m gpu parity tests of the z-sweeping TMA / mbarrier operator kernel (csrc/sweep_kernel.cuh) through the C ABI:
the operator against the oracle's assembled matrix on grids that cross y tiles, z segments and x chunks, and the
multigrid-preconditioned solve (every fused store mode of the kernel) against the round-1 kernel, which is itself held
against the sparse direct solve in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import simp_oracle as o
from test_gpu_parity import make_ctx, random_problem, relmax

pytestmark = pytest.mark.gpu
SWEEP, FUSED = 6, 2


@pytest.mark.parametrize("nel,right", [((12, 33, 64), (12.0, 33.0, 64.0)),   # 3 y tiles, 2 z segments, nz % 32 == 0 (extra column)
                                         ((10, 40, 80), (5.0, 10.0, 10.0)),    # partial last segment, anisotropic cuboid
                                         ((70, 17, 34), (70.0, 17.0, 34.0)),   # several x chunks at one CTA per SM
                                         ((6, 14, 2), (6.0, 14.0, 2.0))])      # one tiny segment
def test_sweep_matvec_matches_oracle(gpu_ctx_factory, nel, right):
    g, dofs, rho, u, pres = random_problem(nel, right, seed=sum(nel))
    ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, matvec_variant=SWEEP)
    ctx.set_density(rho)
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    d = o.operator_diagonal(g, dofs, E, k0)
    y_ref = o.operator_apply(g, dofs, E, k0, pres, u, float(np.mean(d)))
    assert relmax(ctx.matvec(u), y_ref) < 2e-13
    ctx.close()


@pytest.mark.parametrize("chunk", [0, 5])
def test_sweep_solve_agrees_with_round1_kernel(gpu_ctx_factory, chunk, monkeypatch):
    """64 x 32 x 64 cantilever, random high-contrast modulus: CG + V-cycle with the sweep kernel (plain + dot, smoother,
    smoother + dot and residual modes, levels 64/32/16/8/4/2 deep in z) reaches the solution of the round-1 kernel in
    the same number of iterations."""
    nel = (64, 32, 64)
    g = o.generate_grid_hex(nel, (0, 0, 0), tuple(float(v) for v in nel))
    dofs = None  # native numbering
    rng = np.random.default_rng(11)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    nx, ny, nz = nel
    node = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)  # noqa: E731
    jj, kk = np.meshgrid(np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    pres = np.sort((3 * node(0, jj.ravel(), kk.ravel())[:, None] + np.arange(3)[None, :]).ravel()).astype(np.int64)
    f = np.zeros(3 * g.n_nodes)
    f[3 * node(nx, jj.ravel(), kk.ravel()) + 1] = -1.0 / jj.size
    if chunk:
        monkeypatch.setenv("ESIMP_FUSED_CHUNK", str(chunk))
    sol, iters = {}, {}
    for variant in (FUSED, SWEEP):
        ctx = make_ctx(gpu_ctx_factory, g, dofs, pres, pcg_rtol=1e-10, pcg_check_every=1, precond=1, matvec_variant=variant)
        ctx.set_load(f)
        ctx.set_density(rho)
        sol[variant], iters[variant], relres, st = ctx.solve()
        assert st == 0 and relres <= 1e-10
        ctx.close()
    assert abs(iters[SWEEP] - iters[FUSED]) <= 1, iters
    assert relmax(sol[SWEEP], sol[FUSED]) < 1e-8
