"""CPU tests of the body-force / stress-recovery rows: the oracle restatement against analytic
known-answers, and the library's own per-cell stress arithmetic (csrc/stress_math.cuh, compiled for the host into
libesimp_selftest.so) against the oracle.  The kernels that call that arithmetic are covered by the -m gpu tests."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import simp_oracle as o

ROOT = Path(__file__).resolve().parents[1]


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def jittered(nel, right, amp, seed):
    g = o.generate_grid_hex(nel, (0, 0, 0), right)
    rng = np.random.default_rng(seed)
    return o.Grid(g.coords + amp * (rng.random(g.coords.shape) - 0.5), g.cells.copy(), None)


def test_body_force_total_and_threshold():
    """sum of the nodal forces = b * (mass of the cells at or above 1e-6), FEA.jl:495-497"""
    g = o.generate_grid_hex((5, 4, 3), (0, 0, 0), (2.5, 1.0, 0.9))
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(0)
    dens = rng.random(g.n_cells) * 7850.0
    dens[[0, 7, 20]] = [1e-6, 0.999e-6, 0.0]
    b = np.array([0.5, -9.81, 1.0])
    f = o.body_force(g, dofs, b, dens)
    V = o.element_volumes(g)
    keep = dens >= 1e-6
    for c in range(3):
        assert abs(f[dofs[:, c]].sum() - b[c] * float((dens * V)[keep].sum())) < 1e-9 * abs(b[c]) * float((dens * V).sum())
    # uniform density on a cuboid grid: an interior node carries 8 * rho * V/8 * b
    f1 = o.body_force(g, dofs, b, np.ones(g.n_cells))
    nx, ny = 5, 4
    interior = 2 + (nx + 1) * (2 + (ny + 1) * 1)
    assert np.allclose(f1[dofs[interior]], b * V[0], rtol=1e-13)
    corner = 0
    assert np.allclose(f1[dofs[corner]], b * V[0] / 8, rtol=1e-13)


def test_body_force_general_hex_and_tet_partition_of_unity():
    g = jittered((3, 3, 2), (3.0, 3.0, 2.0), 0.2, 1)
    dofs = o.distribute_dofs(g)
    b = np.array([1.0, 2.0, -3.0])
    f = o.body_force(g, dofs, b, np.full(g.n_cells, 2.0))
    # 2x2x2 Gauss integrates det J of a trilinear hex exactly only up to its own order: compare with the same rule
    vol2 = np.array([sum(abs(np.linalg.det(g.coords[c].T @ o._hex_dN(xi))) for xi in o._HEX_QP) for c in g.cells])
    assert np.allclose([f[dofs[:, c]].sum() for c in range(3)], 2.0 * b * vol2.sum(), rtol=1e-12)
    hexes = o.generate_grid_hex((2, 2, 2), (0, 0, 0), (1.0, 1.0, 1.0))
    split = [(0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6)]
    tets = o.Grid(hexes.coords.copy(), np.concatenate([hexes.cells[:, list(s)] for s in split], axis=0), None)
    dt = o.distribute_dofs(tets)
    ft = o.body_force(tets, dt, b, np.ones(tets.n_cells))
    assert np.allclose([ft[dt[:, c]].sum() for c in range(3)], b * 1.0, rtol=1e-12)


def test_stress_known_answers():
    """uniaxial strain field u_x = a x on a cuboid grid: eps = diag(a,0,0), sigma = (lam+2mu, lam, lam) a everywhere"""
    g = o.generate_grid_hex((3, 2, 2), (0, 0, 0), (3.0, 1.0, 0.5))
    dofs = o.distribute_dofs(g)
    a = 1e-3
    u = np.zeros(3 * g.n_nodes)
    u[dofs[:, 0]] = a * g.coords[:, 0]
    rho = np.full(g.n_cells, 0.7)
    st = o.stresses_simp(g, dofs, u, 2.0, 0.3, 1e-9, 3.0, rho)
    lam, mu = o.lame(float(o.simp_modulus(0.7, 2.0, 1e-9, 3.0)), 0.3)
    want = np.array([(lam + 2 * mu) * a, 0, 0, lam * a, 0, lam * a])
    assert np.abs(st["sigma"] - want).max() < 1e-15
    vm = 2 * mu * a  # |s_xx - s_yy| for this state
    assert np.allclose(st["vm_q1"], vm, rtol=1e-12) and np.allclose(st["vm_avg"], vm, rtol=1e-12)
    V = o.element_volumes(g)
    assert np.allclose(st["element_energy"], 0.5 * (lam + 2 * mu) * a * a * V, rtol=1e-12)
    assert st["max_cell"] == int(np.argmax(st["vm_avg"])) and st["max_vm"] == st["vm_avg"].max()
    # pure shear u_x = g y: sigma_xy = mu g, von Mises = sqrt(3) mu g
    u = np.zeros(3 * g.n_nodes)
    u[dofs[:, 0]] = a * g.coords[:, 1]
    st = o.stresses_simp(g, dofs, u, 2.0, 0.3, 1e-9, 3.0, rho)
    assert np.allclose(st["sigma"][:, :, 1], mu * a, rtol=1e-12) and np.allclose(st["vm_q1"], np.sqrt(3) * mu * a, rtol=1e-12)
    # no displacement: the strict `>` of FEA.jl:586 never fires
    st0 = o.stresses_simp(g, dofs, np.zeros(3 * g.n_nodes), 2.0, 0.3, 1e-9, 3.0, rho)
    assert st0["max_cell"] == -1 and st0["max_vm"] == 0.0


@pytest.mark.parametrize("kind", ["cuboid", "general_hex", "tet"])
def test_element_energy_equals_half_uKu(kind):
    """1/2 sum_q sigma:eps dOmega (PostProcessing.jl:172-198) = 1/2 E_e u_e' K0 u_e with the assembly's own matrices"""
    if kind == "cuboid":
        g = o.generate_grid_hex((4, 3, 2), (0, 0, 0), (2.0, 1.5, 0.8))
    elif kind == "general_hex":
        g = jittered((3, 3, 3), (3.0, 3.0, 3.0), 0.2, 2)
    else:
        h = o.generate_grid_hex((2, 2, 2), (0, 0, 0), (1.0, 2.0, 1.5))
        split = [(0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6)]
        g = o.Grid(h.coords.copy(), np.concatenate([h.cells[:, list(s)] for s in split], axis=0), None)
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(5)
    u = rng.standard_normal(int(dofs.max()) + 1)
    rho = 1e-3 + rng.random(g.n_cells) * 0.999
    st = o.stresses_simp(g, dofs, u, 1.0, 0.3, 1e-9, 3.0, rho)
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    ee = 0.5 * o.simp_modulus(rho, 1.0, 1e-9, 3.0) * o.element_energies(g, dofs, u, k0)
    assert relmax(st["element_energy"], ee) < 1e-12


def test_oracle_loop_with_body_force_differs_and_is_consistent():
    g, dofs, f, pres = o.cantilever_problem((12, 4, 2))
    prm = o.OptimizationParameters(volume_fraction=0.4, max_iterations=3, tolerance=0.0)
    a = o.simp_optimize(g, dofs, f, pres, prm, acceleration_data=(np.array([0.0, -9.81, 0.0]), 2e-3))
    b = o.simp_optimize(g, dofs, f, pres, prm)
    z = o.simp_optimize(g, dofs, f, pres, prm, acceleration_data=(np.array([0.0, -9.81, 0.0]), 0.0))
    assert abs(a["history"][0].energy / b["history"][0].energy - 1) > 1e-3
    assert [h.energy for h in z["history"]] == [h.energy for h in b["history"]]  # density 0 < 1e-6: every cell skipped


# ----------------------------------------------------------------------------- the library's stress arithmetic on the CPU
@pytest.fixture(scope="module")
def selftest():
    path = ROOT / "easysimp.jl_b200" / "libesimp_selftest.so"
    assert path.exists(), "run __graft_entry__.build() first"
    lib = C.CDLL(str(path))
    D = C.POINTER(C.c_double)
    lib.selftest_cell_stress.argtypes = [C.c_int, D, D, C.c_double, C.c_double, D, D]
    lib.selftest_cuboid_stress.argtypes = [C.c_double, C.c_double, C.c_double, D, C.c_double, C.c_double, D, D]
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("kind", ["general_hex", "tet"])
def test_library_cell_stress_matches_oracle(selftest, kind):
    if kind == "general_hex":
        g = jittered((3, 2, 2), (3.0, 1.0, 0.7), 0.1, 3)
    else:
        h = o.generate_grid_hex((2, 2, 1), (0, 0, 0), (1.0, 2.0, 0.5))
        split = [(0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6)]
        g = o.Grid(h.coords.copy(), np.concatenate([h.cells[:, list(s)] for s in split], axis=0), None)
    dofs = o.distribute_dofs(g)
    cd = o.celldofs(g, dofs)
    rng = np.random.default_rng(8)
    u = rng.standard_normal(int(dofs.max()) + 1)
    rho = 0.05 + 0.95 * rng.random(g.n_cells)
    ref = o.stresses_simp(g, dofs, u, 1.0, 0.3, 1e-9, 3.0, rho)
    nq = ref["sigma"].shape[1]
    for e in range(g.n_cells):
        xyz = np.ascontiguousarray(g.coords[g.cells[e]])
        ue = np.ascontiguousarray(u[cd[e]])
        sig = np.zeros((nq, 6))
        out3 = np.zeros(3)
        E = float(o.simp_modulus(rho[e], 1.0, 1e-9, 3.0))
        assert selftest.selftest_cell_stress(g.npe, _p(xyz), _p(ue), E, 0.3, _p(sig), _p(out3)) == 0
        assert relmax(sig, ref["sigma"][e]) < 1e-12
        assert np.allclose(out3, [ref["vm_q1"][e], ref["vm_avg"][e], ref["element_energy"][e]], rtol=1e-11)


@pytest.mark.parametrize("h", [(1.0, 1.0, 1.0), (0.05, 0.05, 0.05), (1.0, 0.5, 0.25)])
def test_library_cuboid_stress_matches_oracle(selftest, h):
    g = o.generate_grid_hex((2, 2, 2), (0, 0, 0), tuple(2 * v for v in h))
    dofs = o.distribute_dofs(g)
    cd = o.celldofs(g, dofs)
    rng = np.random.default_rng(9)
    u = rng.standard_normal(3 * g.n_nodes)
    rho = 0.05 + 0.95 * rng.random(g.n_cells)
    ref = o.stresses_simp(g, dofs, u, 1.0, 0.3, 1e-9, 3.0, rho)
    for e in range(g.n_cells):
        ue = np.ascontiguousarray(u[cd[e]])
        sig = np.zeros((8, 6))
        out3 = np.zeros(3)
        E = float(o.simp_modulus(rho[e], 1.0, 1e-9, 3.0))
        assert selftest.selftest_cuboid_stress(h[0], h[1], h[2], _p(ue), E, 0.3, _p(sig), _p(out3)) == 0
        assert relmax(sig, ref["sigma"][e]) < 1e-12
        assert np.allclose(out3, [ref["vm_q1"][e], ref["vm_avg"][e], ref["element_energy"][e]], rtol=1e-11)
