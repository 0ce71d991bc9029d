"""The C/OpenMP CPU restatement (oracle/cpu_simp.c, the cpu_baseline of bench.py) against the numpy oracle."""
import numpy as np

from oracle import cpu_port as cp
from oracle import simp_oracle as o


def native(g):
    return 3 * np.arange(g.n_nodes)[:, None] + np.arange(3)[None, :]


def test_cpu_port_matches_numpy_oracle():
    nel = (7, 5, 4)
    g = o.generate_grid_hex(nel, (0, 0, 0), (7.0, 2.5, 1.0))
    nat = native(g)
    rng = np.random.default_rng(0)
    rho = 1e-3 + rng.random(g.n_cells)
    rho = np.minimum(rho, 1.0)
    u = rng.standard_normal(3 * g.n_nodes)
    K0 = np.ascontiguousarray(o.hex8_ke(g.coords[g.cells[0]], *o.lame(1.0, 0.3)))
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    E = cp.modulus(rho, 1.0, 1e-9, 3.0)
    assert np.allclose(E, o.simp_modulus(rho, 1.0, 1e-9, 3.0), rtol=1e-15)
    pres = np.sort(nat[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel())
    fixed = np.zeros(3 * g.n_nodes, dtype=np.uint8)
    fixed[pres] = 1
    d, m = cp.diagonal(nel, K0, E)
    d_ref = o.operator_diagonal(g, nat, E, k0)
    assert np.allclose(d, d_ref, rtol=1e-13) and abs(m - d_ref.mean()) < 1e-13 * m
    y = cp.matvec(nel, K0, E, fixed, m, u)
    assert np.abs(y - o.operator_apply(g, nat, E, k0, pres, u, m)).max() < 1e-12 * np.abs(y).max()
    dc, en = cp.sensitivities(nel, K0, rho, E, u, 1.0, 1e-9, 3.0)
    assert np.abs(dc - o.sensitivities(g, nat, rho, u, 1.0, 1e-9, 0.3, 3.0, k0)).max() < 1e-11 * np.abs(dc).max()
    fc = o.create_filter_cache(g, 1.5, o.element_volumes(g))
    h = (1.0, 0.5, 0.25)
    assert np.abs(cp.filter_sens(nel, h, fc.radius, rho, dc) - o.filter_sensitivity(fc, rho, dc)).max() < 1e-10 * np.abs(dc).max()
    # solve
    f = np.zeros(3 * g.n_nodes)
    f[nat[g.n_nodes - 1, 1]] = -1.0
    x = np.zeros_like(f)
    it, rr = cp.pcg(nel, K0, E, fixed, f, x, 1e-12, 20000)
    K = o.assemble_stiffness(g, nat, E, k0)
    Kbc, fbc, _ = o.apply_dirichlet(K, f.copy(), pres)
    assert rr <= 1e-12 and np.abs(x - o.solve_direct(Kbc, fbc)).max() < 1e-8 * np.abs(x).max()
    # OC step
    dv = np.full(g.n_cells, 1.0 / g.n_cells)
    new, v = cp.oc_step(rho, -np.abs(dc), dv, 3.0, 0.2, 0.5, 0.125)
    Be = np.abs(dc) / (3.0 * dv)
    ref = np.maximum(1e-3, np.maximum(rho - 0.2, np.minimum(1.0, np.minimum(rho + 0.2, rho * Be ** 0.5))))
    assert np.abs(new - ref).max() < 1e-14 and abs(v - ref.sum() * 0.125) < 1e-10
    assert cp.num_threads() >= 1
