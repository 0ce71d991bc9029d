"""
CPU ORACLE for the EasySIMP.jl per-iteration SIMP loop.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* of the reference's algorithm for the hot path
(SURVEY.md section 8a).  Only `tests/`, `__graft_entry__.smoke()` and the CPU legs of
`bench.py` may import it.  The product (libeasysimp_b200.so + easysimp_b200 host package)
never imports, links or executes anything in `oracle/`.

PARITY UNPINNED: the reference (Julia) cannot be executed in this environment (no julia
binary) and its own test-suite contains no numeric assertions (test/runtests.jl has zero
`@test`).  The oracle is pinned instead by analytic known-answers (tests/test_oracle.py):
K0[0,0] = 0.23504273504..., trace, rank 18, rigid-body null space, patch test, finite
difference check of dc, filter row sums, <Hx,y> = <x,H'y>, direct solve vs CG; stress recovery by
uniaxial / pure-shear known answers and 1/2 sum sigma:eps dOmega = 1/2 E u'K0u (tests/test_post_load_cpu.py).
The one numeric output the reference ships -- the `density` field of data/cantilever_beam_opt.vtu,
written by an earlier version of the package with unrecorded parameters -- is matched
QUALITATIVELY: the oracle's README-configuration run grows the same truss (92 % of the cells on the
same side of 0.5, 97 % of the solid cells within one cell; test_reference_density_artifact_topology).

Third-party arithmetic that is NOT under /root/reference and is restated from its
published behaviour (Ferrite.jl compat "1.1", Project.toml:20; no Manifest.toml):
  * generate_grid(Hexahedron, ...) node / cell ordering        -> generate_grid_hex
  * DofHandler close! dof numbering (first encounter)          -> distribute_dofs
  * Lagrange{RefHexahedron,1}^3, QuadratureRule(2)             -> hex8_ke / tet4_ke
  * apply!(K, f, ch)  (zero rows+cols, diag = mean|diag|)      -> apply_dirichlet
  * cholesky(Symmetric(K,:L)) \\ f (CHOLMOD)                   -> solve_direct (SuperLU)
  * NearestNeighbors.inrange (d <= r, includes self)           -> cKDTree.query_ball_point
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.spatial import cKDTree

# --------------------------------------------------------------------------------------
# Grid + dof numbering  (Ferrite conventions)
# --------------------------------------------------------------------------------------


@dataclass
class Grid:
    """Mirror of the pieces of Ferrite.Grid the hot path reads (0-based indices)."""

    coords: np.ndarray  # (n_nodes, 3) float64
    cells: np.ndarray  # (n_cells, npe) int64, npe = 8 (hex, VTK order) or 4 (tet)
    structured: tuple | None = None  # (nx, ny, nz, left(3), right(3)) when made by generate_grid_hex

    @property
    def n_cells(self):
        return self.cells.shape[0]

    @property
    def n_nodes(self):
        return self.coords.shape[0]

    @property
    def npe(self):
        return self.cells.shape[1]


def generate_grid_hex(nel, left=(0.0, 0.0, 0.0), right=None) -> Grid:
    """Ferrite `generate_grid(Hexahedron, nel, left, right)`: nodes x-fastest then y then z,
    cells x-fastest, VTK hexahedron vertex order (README.md:115, every test/Examples file)."""
    nx, ny, nz = (int(v) for v in nel)
    if right is None:
        right = (float(nx), float(ny), float(nz))
    xs = np.linspace(left[0], right[0], nx + 1)
    ys = np.linspace(left[1], right[1], ny + 1)
    zs = np.linspace(left[2], right[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # x fastest in C order
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    nid = np.arange((nx + 1) * (ny + 1) * (nz + 1), dtype=np.int64).reshape(nz + 1, ny + 1, nx + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    k, j, i = k.ravel(), j.ravel(), i.ravel()  # cell id = i + nx*(j + ny*k)
    cells = np.stack(
        [
            nid[k, j, i], nid[k, j, i + 1], nid[k, j + 1, i + 1], nid[k, j + 1, i],
            nid[k + 1, j, i], nid[k + 1, j, i + 1], nid[k + 1, j + 1, i + 1], nid[k + 1, j + 1, i],
        ],
        axis=1,
    )
    return Grid(coords, cells, (nx, ny, nz, tuple(left), tuple(right)))


def distribute_dofs(grid: Grid) -> np.ndarray:
    """Ferrite `DofHandler` + `close!` for a 3-component Lagrange-1 field
    (FiniteElementAnalysis.jl:147-149): walk cells in order, local vertices in order, give
    every not-yet-seen vertex the next 3 consecutive dofs.  Returns dof_of_node (n_nodes, 3).
    Nodes that belong to no cell get -1 (the reference's get_node_dofs skips them too,
    FiniteElementAnalysis.jl:257-279)."""
    flat = grid.cells.ravel()
    _, first = np.unique(flat, return_index=True)
    order = flat[np.sort(first)]  # nodes in order of first encounter
    dof_of_node = -np.ones((grid.n_nodes, 3), dtype=np.int64)
    dof_of_node[order] = 3 * np.arange(order.size, dtype=np.int64)[:, None] + np.arange(3)
    return dof_of_node


def celldofs(grid: Grid, dof_of_node: np.ndarray) -> np.ndarray:
    """celldofs(cell) = [dofs(v1)..., dofs(v_npe)...]  -> (n_cells, 3*npe)."""
    return dof_of_node[grid.cells].reshape(grid.n_cells, -1)


# --------------------------------------------------------------------------------------
# Material + element stiffness  (FiniteElementAnalysis.jl:52-56, 100-112, 174-193)
# --------------------------------------------------------------------------------------


def lame(E, nu):
    """compute_lame_parameters, FiniteElementAnalysis.jl:52-56."""
    return E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


def simp_modulus(rho, E0, Emin, p):
    """E(rho) = Emin + (E0 - Emin) rho^p, FiniteElementAnalysis.jl:107."""
    return Emin + (E0 - Emin) * np.power(rho, p)


_HEX_REF = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]],
    dtype=np.float64,
)


def _hex_dN(xi):
    """d/dxi of N_a = 1/8 (1+xi_a xi)(1+eta_a eta)(1+zeta_a zeta): (8, 3)."""
    s = _HEX_REF
    f = 1.0 + s * xi[None, :]
    return 0.125 * np.stack([s[:, 0] * f[:, 1] * f[:, 2], f[:, 0] * s[:, 1] * f[:, 2], f[:, 0] * f[:, 1] * s[:, 2]], axis=1)


def _ke_from_gradients(g, lam, mu, dvol, ke):
    """ke[(a,c),(b,d)] += dvol * (lam g_a[c] g_b[d] + mu (delta_cd g_a.g_b + g_a[d] g_b[c]))
    which is (sym grad N_i) : (lam tr(eps_j) I + 2 mu eps_j) for N_i = N_a e_c
    (FiniteElementAnalysis.jl:181-190, constitutive_relation :63-65)."""
    n = g.shape[0]
    gg = g @ g.T
    blk = lam * np.einsum("ac,bd->acbd", g, g) + mu * np.einsum("ad,bc->acbd", g, g)
    blk += mu * np.einsum("ab,cd->acbd", gg, np.eye(3))
    ke += dvol * blk.reshape(3 * n, 3 * n)


def hex8_ke(xyz, lam, mu):
    """8-node hexahedron stiffness, 2x2x2 Gauss (QuadratureRule{RefHexahedron}(2),
    FiniteElementAnalysis.jl:138), local dof = 3*a + c."""
    ke = np.zeros((24, 24))
    gp = 1.0 / math.sqrt(3.0)
    for sz in (-gp, gp):
        for sy in (-gp, gp):
            for sx in (-gp, gp):
                dN = _hex_dN(np.array([sx, sy, sz]))
                J = xyz.T @ dN  # J[c, d] = d x_c / d xi_d
                g = dN @ np.linalg.inv(J)
                _ke_from_gradients(g, lam, mu, abs(np.linalg.det(J)), ke)
    return ke


def tet4_ke(xyz, lam, mu):
    """4-node tetrahedron (constant gradients; any rule whose weights sum to 1/6 gives
    V * B'DB, FiniteElementAnalysis.jl:141-142)."""
    dN = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    J = xyz.T @ dN
    g = dN @ np.linalg.inv(J)
    ke = np.zeros((12, 12))
    _ke_from_gradients(g, lam, mu, abs(np.linalg.det(J)) / 6.0, ke)
    return ke


def element_ke(xyz, lam, mu):
    return hex8_ke(xyz, lam, mu) if xyz.shape[0] == 8 else tet4_ke(xyz, lam, mu)


def unit_element_matrices(grid: Grid, nu: float, E: float = 1.0) -> np.ndarray:
    """K_unit per cell (initialize_element_cache, FiniteElementAnalysis.jl:608-630) for
    Young's modulus E.  Structured grids reuse one matrix (all cells congruent)."""
    lam, mu = lame(E, nu)
    n = 3 * grid.npe
    if grid.structured is not None:
        k0 = element_ke(grid.coords[grid.cells[0]], lam, mu)
        return np.broadcast_to(k0, (grid.n_cells, n, n))
    out = np.empty((grid.n_cells, n, n))
    for e in range(grid.n_cells):
        out[e] = element_ke(grid.coords[grid.cells[e]], lam, mu)
    return out


def element_volumes(grid: Grid) -> np.ndarray:
    """calculate_element_volumes (FiniteElementAnalysis.jl:754-771): sum_q detJ w_q with an
    order-3 rule, exact for trilinear hexes and linear tets."""
    X = grid.coords[grid.cells]  # (n_c, npe, 3)
    if grid.npe == 4:
        J = np.einsum("eac,ad->ecd", X, np.array([[-1.0, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]]))
        return np.abs(np.linalg.det(J)) / 6.0
    vol = np.zeros(grid.n_cells)
    pts = [-math.sqrt(0.6), 0.0, math.sqrt(0.6)]
    wts = [5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0]
    for iz, z in enumerate(pts):
        for iy, y in enumerate(pts):
            for ix, x in enumerate(pts):
                dN = _hex_dN(np.array([x, y, z]))
                J = np.einsum("eac,ad->ecd", X, dN)
                vol += np.linalg.det(J) * (wts[ix] * wts[iy] * wts[iz])
    return vol


# --------------------------------------------------------------------------------------
# Assembly, boundary conditions, loads, solve
# --------------------------------------------------------------------------------------


def assemble_stiffness(grid, dof_of_node, Ee_over_Eunit, k_unit) -> sp.csc_matrix:
    """assemble_with_cache! (FiniteElementAnalysis.jl:678-712): K = sum_e scale_e * K_unit_e
    with scale_e = lambda(rho_e)/lambda_unit = E(rho_e)/E_unit, scattered by celldofs."""
    cd = celldofs(grid, dof_of_node)
    n = cd.shape[1]
    rows = np.repeat(cd, n, axis=1).ravel()
    cols = np.tile(cd, (1, n)).ravel()
    vals = (Ee_over_Eunit[:, None, None] * k_unit).reshape(-1)
    nd = int(dof_of_node.max()) + 1
    return sp.coo_matrix((vals, (rows, cols)), shape=(nd, nd)).tocsc()


def apply_dirichlet(K: sp.csc_matrix, f: np.ndarray, prescribed: np.ndarray):
    """Ferrite apply!(K, f, ch) for homogeneous Dirichlet: m = mean|diag K|,
    zero rows + columns of prescribed dofs, K[d,d] = m, f[d] = 0."""
    m = float(np.mean(np.abs(K.diagonal())))
    keep = np.ones(K.shape[0])
    keep[prescribed] = 0.0
    P = sp.diags(keep)
    K = (P @ K @ P + sp.diags((1.0 - keep) * m)).tocsc()
    f = f * keep
    return K, f, m


def apply_point_load(f, dof_of_node, nodes, force_vector):
    """apply_force! (FiniteElementAnalysis.jl:357-376): F / |nodes| on every listed node."""
    nodes = np.asarray(sorted(nodes), dtype=np.int64)
    per = np.asarray(force_vector, dtype=np.float64) / len(nodes)
    for c in range(3):
        np.add.at(f, dof_of_node[nodes, c], per[c])
    return f


def solve_direct(K, f):
    """u = K \\ f.  The reference uses CHOLMOD (Optimization.jl:317); SuperLU here."""
    return spla.splu(K.tocsc(), permc_spec="MMD_AT_PLUS_A").solve(f)


def solve_pcg(K, f, x0=None, rtol=1e-12, maxit=200000):
    """Jacobi-preconditioned CG on the assembled K (cross-check of the direct solve and the
    model of what the GPU PCG does)."""
    d = K.diagonal()
    x = np.zeros_like(f) if x0 is None else x0.copy()
    r = f - K @ x
    z = r / d
    p = z.copy()
    rz = r @ z
    nf = np.linalg.norm(f)
    it = 0
    while it < maxit and np.linalg.norm(r) > rtol * nf:
        q = K @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        z = r / d
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it


# --------------------------------------------------------------------------------------
# Sensitivities, filters, OC  (src/Optimization/*.jl)
# --------------------------------------------------------------------------------------


def element_energies(grid, dof_of_node, u, k0) -> np.ndarray:
    """u_e' K0 u_e per cell."""
    ue = u[celldofs(grid, dof_of_node)]
    return np.einsum("ei,eij,ej->e", ue, k0, ue)


def sensitivities(grid, dof_of_node, rho_phys, u, E0, Emin, nu, p, k0=None) -> np.ndarray:
    """calculate_sensitivities! (SensitivityAnalysis.jl:31-80):
    dc_e = -p rho^(p-1) (E0 - Emin) u_e' K0 u_e, K0 for unit modulus (:44-45)."""
    if k0 is None:
        k0 = unit_element_matrices(grid, nu, 1.0)
    return -p * np.power(rho_phys, p - 1) * (E0 - Emin) * element_energies(grid, dof_of_node, u, k0)


def cell_centers(grid):
    """calculate_cell_centers (FilterCommon.jl:189-199): mean of cell node coordinates."""
    return grid.coords[grid.cells].mean(axis=1)


def estimate_element_size(grid):
    """estimate_element_size (FilterCommon.jl:109-119): mean over the first <=10 cells of
    (e1 e2 e3)^(1/3) for hexes (:177-182) or the mean edge length for tets (:166-170)."""
    n = min(10, grid.n_cells)
    tot = 0.0
    for e in range(n):
        X = grid.coords[grid.cells[e]]
        if grid.npe == 8:
            e1 = np.linalg.norm(X[1] - X[0])
            e2 = np.linalg.norm(X[3] - X[0])
            e3 = np.linalg.norm(X[4] - X[0])
            tot += (e1 * e2 * e3) ** (1 / 3)
        else:
            pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
            tot += sum(np.linalg.norm(X[j] - X[i]) for i, j in pairs) / 6.0
    return tot / n


@dataclass
class FilterCache:
    """FilterCache (FilterCommon.jl:40-45) as CSR neighbour lists + the derived weights."""

    rowptr: np.ndarray
    col: np.ndarray
    centers: np.ndarray
    radius: float
    volumes: np.ndarray
    weight: np.ndarray = field(default=None)  # max(0, R - dist) per CSR entry

    def row_of(self):
        return np.repeat(np.arange(self.rowptr.size - 1), np.diff(self.rowptr))


def create_filter_cache(grid, filter_radius_ratio, volumes) -> FilterCache:
    """create_filter_cache (FilterCommon.jl:61-98): R = ratio * char size, KD-tree inrange
    (d <= R, includes the cell itself)."""
    c = cell_centers(grid)
    R = filter_radius_ratio * estimate_element_size(grid)
    lists = cKDTree(c).query_ball_point(c, R)
    counts = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
    rowptr = np.concatenate([[0], np.cumsum(counts)])
    col = np.fromiter((j for l in lists for j in sorted(l)), dtype=np.int64, count=int(rowptr[-1]))
    fc = FilterCache(rowptr, col, c, float(R), np.asarray(volumes, dtype=np.float64))
    row = fc.row_of()
    fc.weight = np.maximum(0.0, R - np.linalg.norm(c[col] - c[row], axis=1))
    return fc


def _rowsum(fc, vals):
    return np.bincount(fc.row_of(), weights=vals, minlength=fc.rowptr.size - 1)


def filter_sensitivity(fc, rho, dc):
    """apply_sensitivity_filter_cached! (SensitivityFilter.jl:33-67)."""
    w = fc.weight
    num = _rowsum(fc, w * rho[fc.col] * dc[fc.col] / fc.volumes[fc.col])
    den = _rowsum(fc, w)
    rho_safe = np.maximum(1e-3, rho)
    safe = den > 1e-12
    out = dc.copy()
    out[safe] = num[safe] / (rho_safe[safe] / fc.volumes[safe] * den[safe])
    return out


def filter_density(fc, rho):
    """apply_density_filter_cached! (DensityFilter.jl:30-60)."""
    wv = fc.weight * fc.volumes[fc.col]
    num = _rowsum(fc, wv * rho[fc.col])
    den = _rowsum(fc, wv)
    safe = den > 1e-12
    out = rho.copy()
    out[safe] = num[safe] / den[safe]
    return out


def filter_density_transpose(fc, s):
    """apply_density_filter_chain_rule_cached! (DensityFilter.jl:77-117): scatter form,
    out[e] += w_ie V_e / D_i * s_i with D_i = sum_j w_ij V_j (skipped when D_i <= 1e-12)."""
    wv = fc.weight * fc.volumes[fc.col]
    den = _rowsum(fc, wv)
    row = fc.row_of()
    ok = den[row] > 1e-12
    contrib = np.zeros_like(wv)
    contrib[ok] = wv[ok] / den[row][ok] * s[row][ok]
    return np.bincount(fc.col, weights=contrib, minlength=s.size)


def optimality_criteria_update(rho, dc, dv, volfrac, total_volume, volumes, move=0.2, damping=0.5,
                               fc: FilterCache | None = None, use_density_filter=False):
    """optimality_criteria_update (OptimalityCriteria.jl:69-146), loop restated verbatim:
    lambda in [1e-9, 1e9], arithmetic midpoint, |dV| < 1e-6 absolute, <= 200 steps, x_min 1e-3.
    Returns (new design densities of the LAST evaluated lambda, lambda, steps, converged)."""
    target = volfrac * total_volume
    x_min = 1e-3
    lo, hi = 1e-9, 1e9
    lam = float("nan")
    new = rho.copy()
    steps = 0
    converged = False
    for it in range(1, 201):
        steps = it
        lam = 0.5 * (lo + hi)
        Be = np.abs(dc) / (lam * dv)
        cand = rho * np.power(Be, damping)
        new = np.maximum(x_min, np.maximum(rho - move, np.minimum(1.0, np.minimum(rho + move, cand))))
        phys = filter_density(fc, new) if use_density_filter else new
        err = float(np.dot(phys, volumes)) - target
        if abs(err) < 1e-6:
            converged = True
            break
        if err > 0:
            lo = lam
        else:
            hi = lam
    return new, lam, steps, converged


# --------------------------------------------------------------------------------------
# The SIMP loop (Optimization.jl:178-565)
# --------------------------------------------------------------------------------------


@dataclass
class OptimizationParameters:
    """Optimization.jl:53-123 (defaults :87-102)."""

    E0: float = 1.0
    Emin: float = 1e-9
    nu: float = 0.3
    p: float = 3.0
    volume_fraction: float = 0.5
    max_iterations: int = 200
    tolerance: float = 0.01
    filter_radius: float = 1.5
    filter_type: str = "sensitivity"
    move_limit: float = 0.2
    damping: float = 0.5


@dataclass
class IterationRecord:
    energy: float
    volume: float
    change: float
    lam: float
    grayness: float
    max_u: float
    oc_steps: int


def simp_optimize(grid, dof_of_node, f_load, prescribed, params: OptimizationParameters,
                  solver="direct", pcg_rtol=1e-12, keep_fields=False, acceleration_data=None):
    """Restatement of the loop at Optimization.jl:271-489 plus the final analysis :494-535.
    `f_load` is the assembled load vector (apply_load_condition! output, density independent),
    `prescribed` the union of all ConstraintHandler prescribed dofs; `acceleration_data` =
    (acceleration_vector, base_density) adds the variable-density body force to f before the
    loads and constraints on every assembly (:300-311, :517-528).  Returns a dict."""
    n_c = grid.n_cells
    V = element_volumes(grid)
    Vtot = float(V.sum())
    fc = create_filter_cache(grid, params.filter_radius, V)
    use_df = params.filter_type == "density"
    dv_phys = V / Vtot
    dv = filter_density_transpose(fc, dv_phys) if use_df else dv_phys.copy()
    k_unit = unit_element_matrices(grid, params.nu, params.E0)  # cache[:unit_matrices]
    k0 = unit_element_matrices(grid, params.nu, 1.0)
    rho = np.full(n_c, params.volume_fraction)
    hist = []
    fields = []
    u = np.zeros(int(dof_of_node.max()) + 1)
    converged = False

    def solve_state(rho_phys, u0):
        E = simp_modulus(rho_phys, params.E0, params.Emin, params.p)
        K = assemble_stiffness(grid, dof_of_node, E / params.E0, k_unit)
        f = f_load.copy()
        if acceleration_data is not None:
            f += body_force(grid, dof_of_node, acceleration_data[0], rho_phys * acceleration_data[1])
        Kbc, fbc, _ = apply_dirichlet(K, f, prescribed)
        if solver == "direct":
            uu = solve_direct(Kbc, fbc)
        else:
            uu, _ = solve_pcg(Kbc, fbc, x0=u0, rtol=pcg_rtol)
        return uu, 0.5 * float(uu @ (Kbc @ uu))

    it = 0
    for it in range(1, params.max_iterations + 1):
        old = rho.copy()
        phys = filter_density(fc, rho) if use_df else rho
        u, energy = solve_state(phys, u)
        volume = float(phys @ V)
        dc = sensitivities(grid, dof_of_node, phys, u, params.E0, params.Emin, params.nu, params.p, k0)
        dcf = filter_density_transpose(fc, dc) if use_df else filter_sensitivity(fc, rho, dc)
        rho, lam, steps, _ = optimality_criteria_update(
            rho, dcf, dv, params.volume_fraction, Vtot, V, params.move_limit, params.damping, fc, use_df)
        change = float(np.max(np.abs(rho - old)))
        gray = float(np.count_nonzero((phys > 0.1) & (phys < 0.9))) / n_c
        hist.append(IterationRecord(energy, volume, change, lam, gray, float(np.max(np.abs(u))), steps))
        if keep_fields:
            fields.append(dict(rho_phys=phys.copy(), u=u.copy(), dc=dc.copy(), dcf=dcf.copy(), rho_new=rho.copy()))
        if change < params.tolerance:
            converged = True
            break
    phys = filter_density(fc, rho) if use_df else rho
    u, energy = solve_state(phys, u)
    return dict(densities=phys, design=rho, displacements=u, energy=energy, volume=float(phys @ V),
                iterations=it, converged=converged, history=hist, fields=fields)


# --------------------------------------------------------------------------------------
# Body force, stress recovery and the export fields (the density-dependent load inside the
# loop, Optimization.jl:300-311; the post-solve fields of Optimization.jl:407-477, :537-539)
# --------------------------------------------------------------------------------------

_GP = 1.0 / math.sqrt(3.0)
# QuadratureRule{RefHexahedron}(2): tensor product, first coordinate fastest; point 1 = (-g,-g,-g)
_HEX_QP = np.array([[sx, sy, sz] for sz in (-_GP, _GP) for sy in (-_GP, _GP) for sx in (-_GP, _GP)])
_TET_DN = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def _hex_N(xi):
    s = _HEX_REF
    f = 1.0 + s * xi[None, :]
    return 0.125 * f[:, 0] * f[:, 1] * f[:, 2]


def body_force(grid, dof_of_node, accel, density_data) -> np.ndarray:
    """apply_variable_density_volume_force! (FiniteElementAnalysis.jl:486-526) on a zero f:
    fe[(a,c)] = sum_q density * b_c * N_a(q) * dOmega_q, cells with density < 1e-6 skipped
    (:495-497), added into f in cell order (:519-521)."""
    cd = celldofs(grid, dof_of_node)
    f = np.zeros(int(dof_of_node.max()) + 1)
    accel = np.asarray(accel, dtype=np.float64)
    for e in range(grid.n_cells):
        dens = density_data[e]
        if dens < 1e-6:
            continue
        X = grid.coords[grid.cells[e]]
        fe = np.zeros(3 * grid.npe)
        if grid.npe == 8:
            for xi in _HEX_QP:
                dN = _hex_dN(xi)
                dvol = abs(np.linalg.det(X.T @ dN))  # weight 1
                N = _hex_N(xi)
                fe += (dens * accel[None, :] * N[:, None] * dvol).ravel()  # fe[3a+c] += density * b_c * N_a * dOmega (:505-512)
        else:
            # order-2 rule on the tetrahedron integrates the linear N_a exactly: int N_a = V / 4
            vol = abs(np.linalg.det(X.T @ _TET_DN)) / 6.0
            fe += np.tile(dens * accel * vol / 4.0, 4)
        np.add.at(f, cd[e], fe)
    return f


def stress_voigt(eps, lam, mu):
    """constitutive_relation (FiniteElementAnalysis.jl:63-65) on a symmetric 3x3 strain."""
    return lam * np.trace(eps) * np.eye(3) + 2.0 * mu * eps


def von_mises(sig):
    """sqrt(3/2 dev(s):dev(s)) (FiniteElementAnalysis.jl:584, PostProcessing.jl:216-217)."""
    dev = sig - np.trace(sig) / 3.0 * np.eye(3)
    return math.sqrt(1.5 * float(np.sum(dev * dev)))


# SymmetricTensor{2,3} storage order (lower triangle, column major): 11 21 31 22 32 33
_SYM_IJ = [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2)]


def stresses_simp(grid, dof_of_node, u, E0, nu, Emin, p, density_data):
    """calculate_stresses_simp (FiniteElementAnalysis.jl:566-597) + the two consumers of its
    output in PostProcessing.jl: calculate_von_mises_stresses (:204-222, uses the FIRST
    quadrature point of each cell) and calculate_element_energy (:172-198).
    Returns dict(sigma [n_c, n_q, 6], vm_avg, max_vm, max_cell (0-based, -1 if none),
    vm_q1, element_energy)."""
    cd = celldofs(grid, dof_of_node)
    nq = 8 if grid.npe == 8 else 1
    sigma = np.zeros((grid.n_cells, nq, 6))
    vm_avg = np.zeros(grid.n_cells)
    vm_q1 = np.zeros(grid.n_cells)
    energy = np.zeros(grid.n_cells)
    max_vm, max_cell = 0.0, -1
    for e in range(grid.n_cells):
        lam, mu = lame(float(simp_modulus(density_data[e], E0, Emin, p)), nu)  # material_model(density), :577
        X = grid.coords[grid.cells[e]]
        ue = u[cd[e]].reshape(grid.npe, 3)
        avg = np.zeros((3, 3))
        W = 0.0
        if grid.npe == 8:
            qps = [(_hex_dN(xi), 1.0) for xi in _HEX_QP]
        else:
            qps = [(_TET_DN, 1.0 / 6.0)]  # constant gradient: the 4 points of the rule coincide in value
        for q, (dN, w) in enumerate(qps):
            J = X.T @ dN
            g = dN @ np.linalg.inv(J)
            grad_u = ue.T @ g  # grad_u[c, d] = d u_c / d x_d   (function_gradient, :543)
            eps = 0.5 * (grad_u + grad_u.T)
            sig = stress_voigt(eps, lam, mu)
            sigma[e, q] = [sig[i, j] for i, j in _SYM_IJ]
            avg += sig
            W += float(np.sum(sig * eps)) * abs(np.linalg.det(J)) * w
            if q == 0:
                vm_q1[e] = von_mises(sig)
        avg /= len(qps)
        vm_avg[e] = von_mises(avg)
        energy[e] = 0.5 * W
        if vm_avg[e] > max_vm:  # strict: first cell wins ties (:586)
            max_vm, max_cell = vm_avg[e], e
    return dict(sigma=sigma, vm_avg=vm_avg, max_vm=max_vm, max_cell=max_cell, vm_q1=vm_q1, element_energy=energy)


# --------------------------------------------------------------------------------------
# Matrix-free operator (what the CUDA hot path computes), for direct comparison
# --------------------------------------------------------------------------------------


def operator_apply(grid, dof_of_node, Ee, k0, prescribed, u, diag_value):
    """y = P K(E) P u + m (I - P) u without assembling K: the matrix-free form of
    assembly (FiniteElementAnalysis.jl:678-712) followed by apply!."""
    nd = u.size
    keep = np.ones(nd)
    keep[prescribed] = 0.0
    cd = celldofs(grid, dof_of_node)
    ue = (u * keep)[cd]
    ye = Ee[:, None] * np.einsum("eij,ej->ei", k0, ue)
    y = np.zeros(nd)
    np.add.at(y, cd.ravel(), ye.ravel())
    return y * keep + diag_value * (1.0 - keep) * u


def operator_diagonal(grid, dof_of_node, Ee, k0):
    cd = celldofs(grid, dof_of_node)
    de = Ee[:, None] * np.einsum("eii->ei", k0)
    d = np.zeros(int(dof_of_node.max()) + 1)
    np.add.at(d, cd.ravel(), de.ravel())
    return d


# --------------------------------------------------------------------------------------
# Synthetic problem recipes and deterministic inputs
# --------------------------------------------------------------------------------------


def select_nodes_by_plane(grid, point, normal, tol=1e-4):
    """SelectNodesForBC.jl:18-46."""
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    d = np.abs((grid.coords - np.asarray(point, dtype=np.float64)) @ n)
    return set(np.nonzero(d < tol)[0].tolist())


def select_nodes_by_circle(grid, center, normal, radius, tol=1e-6):
    """SelectNodesForBC.jl:67-102."""
    on_plane = np.array(sorted(select_nodes_by_plane(grid, center, normal, tol)), dtype=np.int64)
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    v = grid.coords[on_plane] - np.asarray(center, dtype=np.float64)
    proj = v - np.outer(v @ n, n)
    return set(on_plane[np.linalg.norm(proj, axis=1) <= radius + tol].tolist())


def cantilever_problem(nel, left=(0.0, 0.0, 0.0), right=None):
    """README.md:110-153 cantilever: clamp x = x_min plane, total force [0,-1,0] on the nodes
    select_nodes_by_circle(grid, [Lx, 0, Lz/2], [1,0,0], 1.0) returns."""
    g = generate_grid_hex(nel, left, right)
    dofs = distribute_dofs(g)
    _, _, _, lft, rgt = g.structured
    fixed = select_nodes_by_plane(g, [lft[0], 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    force_nodes = select_nodes_by_circle(g, [rgt[0], lft[1], 0.5 * (lft[2] + rgt[2])], [1.0, 0.0, 0.0], 1.0)
    prescribed = np.sort(dofs[np.array(sorted(fixed))].ravel())
    f = apply_point_load(np.zeros(3 * g.n_nodes), dofs, force_nodes, [0.0, -1.0, 0.0])
    return g, dofs, f, prescribed


def four_legs_problem(nel=(40, 20, 20)):
    """test/Examples/05_3D_2x1x1_4Legs.jl:24,54-94: 2x1x1 box, four clamped 0.3x0.3 corner
    patches on x=0, total force [0,0,-1] on the disc r<=0.1 around (2,.5,.5)."""
    g = generate_grid_hex(nel, (0.0, 0.0, 0.0), (2.0, 1.0, 1.0))
    dofs = distribute_dofs(g)
    x, y, z = g.coords.T
    lo = lambda v: v <= 0.3 + 1e-6  # noqa: E731
    hi = lambda v: v >= 1.0 - 0.3 - 1e-6  # noqa: E731
    fixed = np.nonzero((np.abs(x) < 1e-6) & ((lo(y) | hi(y)) & (lo(z) | hi(z))))[0]
    force = np.nonzero((np.abs(x - 2.0) < 1e-6) & ((y - 0.5) ** 2 + (z - 0.5) ** 2 <= 0.1 ** 2 + 1e-6))[0]
    prescribed = np.sort(dofs[fixed].ravel())
    f = apply_point_load(np.zeros(3 * g.n_nodes), dofs, force.tolist(), [0.0, 0.0, -1.0])
    return g, dofs, f, prescribed


def _nearest_node(grid, target):
    """The closest-node fallback the example scripts carry (first minimum, 02_sliding_support.jl:74-86)."""
    return int(np.argmin(np.linalg.norm(grid.coords - np.asarray(target, dtype=np.float64), axis=1)))


def sliding_support_problem(nel=(60, 20, 4)):
    """test/Examples/02_sliding_support.jl:46-118: x = 0 face slides (U1 = 0 only), the nodes
    select_nodes_by_circle(grid, [Lx, 0, Lz/2], [0,1,0], 0.5) returns carry U2 = 0 only, total force [0,-1,0]
    on select_nodes_by_circle(grid, [0, Ly, Lz/2], [1,0,0], 1.0).  Partial (per-component) Dirichlet sets."""
    g = generate_grid_hex(nel, (0.0, 0.0, 0.0), tuple(float(n) for n in nel))
    dofs = distribute_dofs(g)
    Lx, Ly, Lz = (float(n) for n in nel)
    sliding = select_nodes_by_plane(g, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    support = select_nodes_by_circle(g, [Lx, 0.0, 0.5 * Lz], [0.0, 1.0, 0.0], 0.5) or {_nearest_node(g, [Lx, 0.0, 0.5 * Lz])}
    force = select_nodes_by_circle(g, [0.0, Ly, 0.5 * Lz], [1.0, 0.0, 0.0], 1.0) or {_nearest_node(g, [0.0, Ly, 0.5 * Lz])}
    prescribed = np.unique(np.concatenate([dofs[np.array(sorted(sliding)), 0], dofs[np.array(sorted(support)), 1]]))
    f = apply_point_load(np.zeros(3 * g.n_nodes), dofs, force, [0.0, -1.0, 0.0])
    return g, dofs, f, prescribed


def mbb_problem(nel=(40, 20, 20)):
    """test/Examples/06_3D_2x1x1_MBB.jl:29,53-113: 2x1x1 box, symmetry face x = 0 (U1 = 0), bottom edge strip
    y = 0, x >= 1.95 (U2 = 0), the node nearest (0, 1, 0.5) holds U3, total force [0,-1,0] on the half disc
    r <= 0.1 around (0, 1, 0.5) of the top face."""
    g = generate_grid_hex(nel, (0.0, 0.0, 0.0), (2.0, 1.0, 1.0))
    dofs = distribute_dofs(g)
    x, y, z = g.coords.T
    eps = np.finfo(np.float64).eps
    sym = np.nonzero(np.abs(x) < 1e-12)[0]  # select_nodes_by_plane(..., eps()): exact plane on generate_grid coordinates
    sup = np.nonzero((np.abs(y) < eps) & (x >= 2.0 - 0.05 - eps))[0]
    zfix = _nearest_node(g, [0.0, 1.0, 0.5])
    force = np.nonzero((np.abs(y - 1.0) < eps) & (np.sqrt(x ** 2 + (z - 0.5) ** 2) <= 0.1 + eps) & (x >= -eps))[0]
    prescribed = np.unique(np.concatenate([dofs[sym, 0], dofs[sup, 1], dofs[[zfix], 2]]))
    f = apply_point_load(np.zeros(3 * g.n_nodes), dofs, force.tolist(), [0.0, -1.0, 0.0])
    return g, dofs, f, prescribed


def michell_half_problem(nel=(40, 20, 20)):
    """test/Examples/08_3D_2x1x1_Michell-half_tol_study.jl:66,95-164: two clamped 0.2 x 0.2 patches of the bottom
    face at the corners (0,0,0) and (2,0,0), symmetry plane z = 1 (U3 = 0), total force [0,-1,0] on the disc r <= 0.1
    around (1, 0, 1) of the bottom face.  Mixes apply_fixed_boundary! and apply_sliding_boundary! handlers."""
    g = generate_grid_hex(nel, (0.0, 0.0, 0.0), (2.0, 1.0, 1.0))
    dofs = distribute_dofs(g)
    x, y, z = g.coords.T
    eps = np.finfo(np.float64).eps
    cs = 0.20
    bottom = np.abs(y) < eps
    left = np.nonzero(bottom & (x <= cs + eps) & (z <= cs + eps))[0]
    right = np.nonzero(bottom & (x >= 2.0 - cs - eps) & (z <= cs + eps))[0]
    symz = np.array(sorted(select_nodes_by_plane(g, [0.0, 0.0, 1.0], [0.0, 0.0, 1.0], 1e-6)))
    force = np.nonzero(bottom & (np.sqrt((x - 1.0) ** 2 + (z - 1.0) ** 2) <= 0.1 + eps))[0]
    prescribed = np.unique(np.concatenate([dofs[left].ravel(), dofs[right].ravel(), dofs[symz, 2]]))
    f = apply_point_load(np.zeros(3 * g.n_nodes), dofs, force.tolist(), [0.0, -1.0, 0.0])
    return g, dofs, f, prescribed


def splitmix64(idx, seed=20261018):
    """Counter-based hash -> U[0,1): bit-identical in C++, CUDA and numpy."""
    with np.errstate(over="ignore"):
        z = (np.asarray(idx, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
