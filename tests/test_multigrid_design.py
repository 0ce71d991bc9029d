"""CPU tests of the statements csrc/multigrid.cu is built on (numpy prototype scripts/gmg_prototype.py):
  * for trilinear hexes on nested cuboid grids the rediscretised coarse operator with E' = 1/4 sum of the 8 children
    (mean modulus x the factor 2 of K0(2h) = 2 K0(h)) IS the Galerkin operator P'AP when E is constant per coarse cell;
  * the V-cycle with Chebyshev-Jacobi smoothing is a symmetric operator and cuts the CG iteration count several-fold."""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
spec = importlib.util.spec_from_file_location("gmg_prototype", ROOT / "scripts" / "gmg_prototype.py")
gp = importlib.util.module_from_spec(spec)
sys.argv = [sys.argv[0]]
spec.loader.exec_module(gp)
from oracle import simp_oracle as o  # noqa: E402


def test_rediscretised_coarse_operator_is_galerkin_for_piecewise_constant_modulus():
    nel = (8, 4, 4)
    rng = np.random.default_rng(0)
    Ec = 0.1 + rng.random((2, 2, 4))                                   # one modulus per COARSE cell [k][j][i]
    Ef = np.repeat(np.repeat(np.repeat(Ec, 2, axis=0), 2, axis=1), 2, axis=2).ravel()
    nofix = np.zeros(3 * 9 * 5 * 5)
    lv = gp.build_levels(nel, Ef, nofix, 2)
    A0, A1, P = lv[0]["A"], lv[1]["A"], lv[0]["P"]
    G = (P.T @ A0 @ P).toarray()
    assert np.abs(G - A1.toarray()).max() < 1e-12 * np.abs(G).max()
    # K0(2h) = 2 K0(h) for cuboids
    g1, g2 = o.generate_grid_hex((1, 1, 1), (0, 0, 0), (1.0, 0.5, 0.25)), o.generate_grid_hex((1, 1, 1), (0, 0, 0), (2.0, 1.0, 0.5))
    k1 = o.hex8_ke(g1.coords[g1.cells[0]], *o.lame(1.0, 0.3))
    k2 = o.hex8_ke(g2.coords[g2.cells[0]], *o.lame(1.0, 0.3))
    assert np.abs(k2 - 2 * k1).max() < 1e-14


def test_vcycle_is_symmetric_and_beats_jacobi():
    nel = (16, 8, 8)
    g, dofs, f, pres = o.cantilever_problem(nel)
    nat = gp.native(g)
    rng = np.random.default_rng(1)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    fixed = np.zeros((g.n_nodes, 3))
    fixed[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))] = 1
    lv = gp.build_levels(nel, o.simp_modulus(rho, 1.0, 1e-9, 3.0), fixed.ravel(), 3)
    for L in lv:
        L["lmax"] = 1.05 * gp.lam_max(L["A"], L["dinv"])
    keep = lv[0]["keep"]
    x, y = rng.standard_normal(keep.size) * keep, rng.standard_normal(keep.size) * keep
    Mx, My = gp.vcycle(lv, 0, x), gp.vcycle(lv, 0, y)
    assert abs(Mx @ y - x @ My) < 1e-10 * abs(Mx @ y)               # M^-1 symmetric
    assert Mx @ x > 0
    fn = np.zeros(3 * g.n_nodes)
    fn[nat.ravel()] = f[dofs.ravel()]
    fn *= keep
    _, it_j = gp.pcg(lv[0]["A"], fn, lambda r: lv[0]["dinv"] * r)
    xg, it_g = gp.pcg(lv[0]["A"], fn, lambda r: gp.vcycle(lv, 0, r))
    assert it_g * 5 < it_j
    assert np.linalg.norm(lv[0]["A"] @ xg - fn) <= 1e-8 * np.linalg.norm(fn)
