"""Host-side mirror of the reference's interface for the SIMP hot path.

The reference is Julia (no toolchain in this image), so the host side above the C ABI is written
in Python with the SAME names, argument meaning and error behaviour as the reference's entry points
(Julia's trailing `!` becomes a trailing `_`).  Everything numeric on the per-iteration path is a
call into libeasysimp_b200.so (easysimp_b200.lib); what stays on the host is what stays on the host
in the reference drop-in too: grid bookkeeping, node selection, load vectors, the loop control of
`simp_optimize` (Optimization.jl:271-489).

    reference (Julia)                                      here
    generate_grid(Hexahedron, nel, left, right)  [Ferrite] generate_grid(Hexahedron, nel, left, right)
    setup_problem(grid)              FEA.jl:130            setup_problem(grid) -> dh, cellvalues, K, f
    create_simp_material_model       FEA.jl:100            create_simp_material_model
    assemble_stiffness_matrix_simp!  FEA.jl:229,637        assemble_stiffness_matrix_simp_  (K is an operator handle)
    apply_fixed_boundary!            FEA.jl:293            apply_fixed_boundary_
    apply_sliding_boundary!          FEA.jl:326            apply_sliding_boundary_
    apply_force! / PointLoad         FEA.jl:357 / LoadConditions.jl:19  apply_force_ / PointLoad
    select_nodes_by_plane/_circle    SelectNodesForBC.jl:18,67
    FilterCache, create_filter_cache FilterCommon.jl:40,61
    OptimizationParameters/Result    Optimization.jl:53,145
    simp_optimize                    Optimization.jl:178
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib

Hexahedron = "Hexahedron"
Tetrahedron = "Tetrahedron"


# ----------------------------------------------------------------------------- grid / dofs
@dataclass
class Grid:
    """The parts of Ferrite.Grid the hot path reads.  0-based ids."""

    coords: np.ndarray  # (n_nodes, 3)
    cells: np.ndarray  # (n_cells, 8 | 4)
    structured: tuple | None = None  # (nx, ny, nz, left, right) for generate_grid grids

    def getncells(self):
        return self.cells.shape[0]

    def getnnodes(self):
        return self.coords.shape[0]


def generate_grid(celltype, nel, left=(0.0, 0.0, 0.0), right=None) -> Grid:
    """Ferrite's generate_grid(Hexahedron, nel, left, right): nodes and cells x-fastest, VTK vertex
    order per cell."""
    if celltype != Hexahedron:
        raise ValueError("only Hexahedron grids are generated; import tets as Grid(coords, cells)")
    nx, ny, nz = map(int, nel)
    if right is None:
        right = (float(nx), float(ny), float(nz))
    ax = [np.linspace(left[d], right[d], n + 1) for d, n in enumerate((nx, ny, nz))]
    coords = np.empty(((nx + 1) * (ny + 1) * (nz + 1), 3))
    coords[:, 0] = np.tile(ax[0], (ny + 1) * (nz + 1))
    coords[:, 1] = np.tile(np.repeat(ax[1], nx + 1), nz + 1)
    coords[:, 2] = np.repeat(ax[2], (nx + 1) * (ny + 1))
    e = np.arange(nx * ny * nz, dtype=np.int64)
    i, j, k = e % nx, (e // nx) % ny, e // (nx * ny)
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    n0 = i + sy * j + sz * k
    cells = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + sz, n0 + sz + sx, n0 + sz + sx + sy, n0 + sz + sy], axis=1)
    return Grid(coords, cells, (nx, ny, nz, tuple(map(float, left)), tuple(map(float, right))))


@dataclass
class DofHandler:
    """Ferrite DofHandler for a 3-component Lagrange-1 field: dofs are handed out in order of first
    encounter walking cells then local vertices (close!, FEA.jl:147-149)."""

    grid: Grid
    dof_of_node: np.ndarray  # (n_nodes, 3), -1 for nodes in no cell

    @property
    def ndofs(self):
        return int(self.dof_of_node.max()) + 1

    def celldofs(self):
        return self.dof_of_node[self.grid.cells].reshape(self.grid.getncells(), -1)


def _distribute_dofs_structured(nx: int, ny: int, nz: int) -> np.ndarray:
    """Closed form of the first-encounter numbering on a generate_grid mesh (setup at 100 M dofs): node
    (I,J,K) is first met by cell (max(I-1,0), max(J-1,0), max(K-1,0)) -- the earliest cell in x-fastest order that
    contains it -- as that cell's local vertex VTK[(dx,dy)] + 4 dz.  One sort of n_nodes keys instead of a unique over
    8 n_cells connectivity entries."""
    I = np.arange(nx + 1, dtype=np.int64)[None, None, :]
    J = np.arange(ny + 1, dtype=np.int64)[None, :, None]
    K = np.arange(nz + 1, dtype=np.int64)[:, None, None]
    i0, j0, k0 = np.maximum(I - 1, 0), np.maximum(J - 1, 0), np.maximum(K - 1, 0)
    vtk = np.array([0, 1, 3, 2], dtype=np.int64)  # (dx,dy) = (0,0) (1,0) (0,1) (1,1) -> local vertex 0 1 3 2
    local = vtk[(I - i0) + 2 * (J - j0)] + 4 * (K - k0)
    key = ((i0 + nx * (j0 + ny * k0)) * 8 + local).reshape(-1)  # node ids are x-fastest like the flattening
    order = np.argsort(key, kind="stable")
    dofs = np.empty((key.size, 3), dtype=np.int64)
    dofs[order] = (3 * np.arange(key.size, dtype=np.int64))[:, None] + np.arange(3, dtype=np.int64)[None, :]
    return dofs


def _distribute_dofs(grid: Grid) -> np.ndarray:
    if grid.structured is not None:
        return _distribute_dofs_structured(*grid.structured[:3])
    flat = grid.cells.reshape(-1)
    seen = np.zeros(grid.getnnodes(), dtype=bool)
    uniq, first = np.unique(flat, return_index=True)
    order = uniq[np.argsort(first, kind="stable")]
    seen[order] = True
    dofs = np.full((grid.getnnodes(), 3), -1, dtype=np.int64)
    dofs[order] = (3 * np.arange(order.size, dtype=np.int64))[:, None] + np.arange(3, dtype=np.int64)[None, :]
    return dofs


class StiffnessOperator:
    """What `K = allocate_matrix(dh)` (FEA.jl:153) becomes: a handle on the matrix-free operator
    that lives in the library context.  `K @ u` applies P K(rho) P + m (I - P)."""

    def __init__(self, ctx: _lib.Context, dh: DofHandler):
        self.ctx = ctx
        self.dh = dh
        self.shape = (dh.ndofs, dh.ndofs)
        self.prescribed = np.zeros(0, dtype=np.int64)
        self.material = None
        self._density = None

    def __matmul__(self, u):
        return self.ctx.matvec(np.asarray(u, dtype=np.float64))

    def diagonal(self):
        return self.ctx.diagonal()


def setup_problem(grid: Grid, interpolation_order: int = 1, device: int = -1, ctx: _lib.Context | None = None):
    """setup_problem (FEA.jl:130-157): returns dh, cellvalues, K, f.  K is a StiffnessOperator
    (no sparse matrix is allocated); cellvalues is a placeholder (quadrature lives in the kernels)."""
    if interpolation_order != 1:
        raise ValueError("only first-order Lagrange elements are supported (as in the reference's tests)")
    dh = DofHandler(grid, _distribute_dofs(grid))
    ctx = ctx or _lib.Context(device)
    if grid.structured is not None:
        nx, ny, nz, left, right = grid.structured
        h = [(right[d] - left[d]) / n for d, n in enumerate((nx, ny, nz))]
        ctx.set_grid_structured(nx, ny, nz, h[0], h[1], h[2], dh.dof_of_node)
    else:
        npe = grid.cells.shape[1]
        ctx.set_mesh_unstructured(npe, grid.coords, grid.cells, dh.celldofs())
    K = StiffnessOperator(ctx, dh)
    # the quadrature lives in the kernels; the placeholder carries the context so that the functions the reference
    # calls with (dh, cellvalues) and no K -- body force, stress recovery -- find the device state
    cellvalues = {"celltype": Hexahedron if grid.cells.shape[1] == 8 else Tetrahedron, "quadrature_order": 2, "ctx": ctx}
    return dh, cellvalues, K, np.zeros(dh.ndofs)


@dataclass
class SimpMaterialModel:
    """create_simp_material_model (FEA.jl:100-112): E(rho) = Emin + (E0-Emin) rho^p."""

    E0: float
    nu: float
    Emin: float = 1e-6
    p: float = 3.0

    def __call__(self, density):
        E = self.Emin + (self.E0 - self.Emin) * density ** self.p
        return E * self.nu / ((1 + self.nu) * (1 - 2 * self.nu)), E / (2 * (1 + self.nu))


def create_simp_material_model(E0, nu, Emin=1e-6, p=3.0):
    return SimpMaterialModel(float(E0), float(nu), float(Emin), float(p))


def assemble_stiffness_matrix_simp_(K: StiffnessOperator, f, dh, cellvalues, material_model, density_data, *cache_args):
    """assemble_stiffness_matrix_simp! (6-arg FEA.jl:229 and 9-arg :637 forms).  Matrix-free: uploads
    the physical densities; E_e and the Jacobi diagonal are formed on the device.  `f` is zero-filled
    like start_assemble(K, f) does."""
    K.material = material_model
    K.ctx.set_params(E0=material_model.E0, Emin=material_model.Emin, nu=material_model.nu, p=material_model.p)
    K.ctx.set_dirichlet(K.prescribed)
    K._density = np.array(density_data, dtype=np.float64)
    K.ctx.set_density(K._density)
    if f is not None:
        f[:] = 0.0
    return K


@dataclass
class ConstraintHandler:
    dh: DofHandler
    prescribed_dofs: np.ndarray


def _constrain(K, f, dh, nodes, comps):
    nodes = np.asarray(sorted(nodes), dtype=np.int64)
    dofs = np.unique(dh.dof_of_node[nodes][:, comps].reshape(-1))
    K.prescribed = np.union1d(K.prescribed, dofs)
    K.ctx.set_dirichlet(K.prescribed)
    if f is not None:
        f[dofs] = 0.0  # apply!(K, f, ch): f[d] = 0 for homogeneous constraints
    if K._density is not None:  # the Jacobi diagonal depends on the constraint set
        K.ctx.set_density(K._density)
    return ConstraintHandler(dh, dofs)


def apply_fixed_boundary_(K, f, dh, nodes):
    """apply_fixed_boundary! (FEA.jl:293-309): all three components of `nodes` are prescribed to 0."""
    return _constrain(K, f, dh, nodes, [0, 1, 2])


def apply_sliding_boundary_(K, f, dh, nodes, fixed_dofs):
    """apply_sliding_boundary! (FEA.jl:326-340); fixed_dofs are 1-based directions like the reference."""
    return _constrain(K, f, dh, nodes, [d - 1 for d in fixed_dofs])


def apply_force_(f, dh, nodes, force_vector):
    """apply_force! (FEA.jl:357-376): force_vector / |nodes| on each node; error on an empty set."""
    nodes = list(nodes)
    if len(nodes) == 0:
        raise ValueError("No nodes provided for force application.")
    per = np.asarray(force_vector, dtype=np.float64) / len(nodes)
    for n in nodes:
        d = dh.dof_of_node[n]
        if d[0] >= 0:
            f[d] += per
    return f


def apply_variable_density_volume_force_(f, dh, cellvalues, body_force_vector, density_data):
    """apply_variable_density_volume_force! (FEA.jl:486-526): f += sum_e density_e * b * int N dOmega, cells with
    density < 1e-6 skipped.  Inside simp_optimize the library does this on the device at every assembly
    (`acceleration_data`); this is the stand-alone form."""
    fb = cellvalues["ctx"].body_force(body_force_vector, density_data)
    f += fb
    return f


class StressField:
    """What `stress_field::Dict{Int,Vector{SymmetricTensor}}` (FEA.jl:566-597) becomes: three cell arrays computed on
    the device plus on-demand Gauss-point tensors.  `field[cell]` -> (nq, 6) in SymmetricTensor storage order
    (11 21 31 22 32 33), valid while the context still holds the state the field was computed from."""

    def __init__(self, ctx, fields):
        self._ctx = ctx
        self.von_mises_q1 = fields["vm_q1"]  # calculate_von_mises_stresses, PostProcessing.jl:204-222
        self.von_mises_avg = fields["vm_avg"]  # FEA.jl:584
        self.element_energy = fields["element_energy"]  # calculate_element_energy, PostProcessing.jl:172-198
        self.max_von_mises = fields["max_vm"]
        self.max_cell = fields["max_cell"]

    def __len__(self):
        return 0 if self.von_mises_q1 is None else self.von_mises_q1.size

    def __getitem__(self, cell):
        return self._ctx.stress_tensors(int(cell), 1)[0]

    def tensors(self, cell0=0, n=None):
        return self._ctx.stress_tensors(cell0, n)


def calculate_stresses_simp(u, dh, cellvalues, material_model, density_data):
    """calculate_stresses_simp (FEA.jl:566-597) -> (stress_field, max_von_mises, max_vm_cell_id); cell ids are 0-based
    here (-1 = none).  Uploads u and the densities; simp_optimize calls the device-resident form instead."""
    ctx = cellvalues["ctx"]
    ctx.set_params(E0=material_model.E0, Emin=material_model.Emin, nu=material_model.nu, p=material_model.p)
    ctx.set_density(np.asarray(density_data, dtype=np.float64))
    ctx.set_displacement(np.asarray(u, dtype=np.float64))
    sf = StressField(ctx, ctx.stress_fields())
    return sf, sf.max_von_mises, sf.max_cell


def create_results_data(grid, dh, cellvalues, opt_result):
    """create_results_data + the cell/point arrays export_main_results writes (PostProcessing.jl:39-110).  The
    `displacement` array is indexed by node through the dof handler (the reference's extract_nodal_displacements,
    :147-165, assumes dof = 3(node-1)+c, which Ferrite's numbering does not guarantee)."""
    u = opt_result.displacements
    nodal = np.where(dh.dof_of_node >= 0, u[np.maximum(dh.dof_of_node, 0)], 0.0).T  # (3, n_nodes)
    st = opt_result.stresses
    return dict(density=opt_result.densities, von_mises_stress=st.von_mises_q1, element_energy=st.element_energy,
                displacement=nodal, displacement_magnitude=np.linalg.norm(nodal, axis=0), energy=opt_result.energy,
                volume_fraction=opt_result.volume / float(np.sum(calculate_element_volumes(grid))),
                iterations=opt_result.iterations, converged=opt_result.converged)


@dataclass
class PointLoad:
    """PointLoad (LoadConditions.jl:19-40)."""

    dh: DofHandler
    nodes: list
    force_vector: list


def apply_load_condition_(f, load: PointLoad):
    return apply_force_(f, load.dh, load.nodes, load.force_vector)


def select_nodes_by_plane(grid, point, normal, tolerance=1e-4):
    """SelectNodesForBC.jl:18-46."""
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    dist = np.abs((grid.coords - np.asarray(point, dtype=np.float64)) @ n)
    return set(np.nonzero(dist < tolerance)[0].tolist())


def select_nodes_by_circle(grid, center, normal, radius, tolerance=1e-6):
    """SelectNodesForBC.jl:67-102."""
    on_plane = np.array(sorted(select_nodes_by_plane(grid, center, normal, tolerance)), dtype=np.int64)
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    v = grid.coords[on_plane] - np.asarray(center, dtype=np.float64)
    proj = v - np.outer(v @ n, n)
    return set(on_plane[np.linalg.norm(proj, axis=1) <= radius + tolerance].tolist())


# ----------------------------------------------------------------------------- filter cache
def estimate_element_size(grid: Grid) -> float:
    """FilterCommon.jl:109-119 / :166-182: mean over the first <= 10 cells."""
    n = min(10, grid.getncells())
    X = grid.coords[grid.cells[:n]]
    if grid.cells.shape[1] == 8:
        e1 = np.linalg.norm(X[:, 1] - X[:, 0], axis=1)
        e2 = np.linalg.norm(X[:, 3] - X[:, 0], axis=1)
        e3 = np.linalg.norm(X[:, 4] - X[:, 0], axis=1)
        sizes = (e1 * e2 * e3) ** (1 / 3)
    else:
        pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        sizes = sum(np.linalg.norm(X[:, j] - X[:, i], axis=1) for i, j in pairs) / 6.0
    total = 0.0
    for v in sizes:  # sequential accumulation like the reference's loop (FilterCommon.jl:112-117)
        total += float(v)
    return total / n


def calculate_cell_centers(grid: Grid):
    """FilterCommon.jl:189-199."""
    return grid.coords[grid.cells].mean(axis=1)


@dataclass
class FilterCache:
    """FilterCache (FilterCommon.jl:40-45).  On structured grids the neighbour lists are implicit
    (a stencil built inside the library); otherwise CSR arrays mirror `neighbor_lists`."""

    filter_radius: float
    element_volumes: np.ndarray | None = None
    cell_centers: np.ndarray | None = None
    rowptr: np.ndarray | None = None
    col: np.ndarray | None = None
    structured: bool = False


def create_filter_cache(grid: Grid, filter_radius_ratio: float, element_volumes=None) -> FilterCache:
    """create_filter_cache (FilterCommon.jl:61-98): R = ratio * characteristic size; KD-tree inrange
    lists for imported meshes, analytic stencil for generate_grid grids."""
    R = float(filter_radius_ratio) * estimate_element_size(grid)
    if grid.structured is not None:
        return FilterCache(R, element_volumes, None, None, None, True)
    from scipy.spatial import cKDTree  # NearestNeighbors.jl's role (Project.toml:11)

    c = calculate_cell_centers(grid)
    lists = cKDTree(c).query_ball_point(c, R)
    counts = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
    rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    col = np.fromiter((j for l in lists for j in sorted(l)), dtype=np.int64, count=int(rowptr[-1]))
    return FilterCache(R, np.asarray(element_volumes, dtype=np.float64), c, rowptr, col, False)


def calculate_element_volumes(grid: Grid):
    """calculate_element_volumes (FEA.jl:754-771); exact closed forms of the order-3 rule."""
    if grid.structured is not None:  # no (n_cells, 8, 3) gather at 33 M cells
        nx, ny, nz, left, right = grid.structured
        v = np.prod([(right[d] - left[d]) / n for d, n in enumerate((nx, ny, nz))])
        return np.full(grid.getncells(), v)
    X = grid.coords[grid.cells]
    if grid.cells.shape[1] == 4:
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=2)
        return np.abs(np.linalg.det(J)) / 6.0
    ref = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    pts = [-np.sqrt(0.6), 0.0, np.sqrt(0.6)]
    wts = [5 / 9, 8 / 9, 5 / 9]
    vol = np.zeros(grid.getncells())
    for a3, z in enumerate(pts):
        for a2, y in enumerate(pts):
            for a1, x in enumerate(pts):
                f = 1 + ref * np.array([x, y, z])
                dN = 0.125 * np.stack([ref[:, 0] * f[:, 1] * f[:, 2], f[:, 0] * ref[:, 1] * f[:, 2], f[:, 0] * f[:, 1] * ref[:, 2]], axis=1)
                J = np.einsum("eac,ad->ecd", X, dN)
                vol += np.linalg.det(J) * wts[a1] * wts[a2] * wts[a3]
    return vol


def check_sensitivity_health(sensitivities) -> bool:
    """check_sensitivity_health (OptimalityCriteria.jl:19-40): the warn-only diagnostic the reference runs at the top
    of every OC update.  It never changes a result, so the device loop skips it; this is the same test for callers who
    want it on sensitivities they downloaded (`Context.sensitivities()`, `Context.filter_sens(...)`)."""
    import warnings

    s = np.asarray(sensitivities, dtype=np.float64)
    if np.count_nonzero(s < 0) < s.size * 0.5:
        warnings.warn("Less than 50% of sensitivities are negative. Check if energy sensitivities are computed correctly.")
        return False
    a = np.abs(s)
    med = float(np.median(a))
    eps = float(np.finfo(np.float64).eps)
    if med < eps:
        warnings.warn(f"Sensitivities are effectively zero (median: {med}).")
        return False
    ratio = float(a.max()) / max(med, eps)
    if ratio > 1e8:
        warnings.warn(f"Sensitivity range too large (max/median: {ratio}). Check problem scaling.")
        return False
    return True


# ----------------------------------------------------------------------------- parameters / result
@dataclass
class OptimizationParameters:
    """OptimizationParameters (Optimization.jl:53-123), same fields and defaults (:87-102), plus the
    PCG controls that replace the direct solve."""

    E0: float = 1.0
    Emin: float = 1e-9
    nu: float = 0.3
    p: float = 3.0
    volume_fraction: float = 0.5
    max_iterations: int = 200
    tolerance: float = 0.01
    filter_radius: float = 1.5
    filter_type: str = "sensitivity"  # :sensitivity | :density
    move_limit: float = 0.2
    damping: float = 0.5
    use_cache: bool = True
    export_interval: int = 0
    export_path: str = ""
    task_name: str = "SIMP_Optimization"
    tolerance_checkpoints: list = field(default_factory=list)
    # solver (no counterpart in the reference: CHOLMOD is exact)
    pcg_rtol: float = 1e-8
    pcg_maxit: int = 200000
    pcg_check_every: int = 25
    matvec_variant: int = 0
    # "jacobi" | "gmg" | "auto" (gmg on generate_grid grids, jacobi on imported meshes) | "gmg_fp32" (opt-in: the V-cycle
    # evaluated in single precision inside the FP64 CG, csrc/multigrid32.cu)
    preconditioner: str = "auto"


@dataclass
class OptimizationResult:
    """OptimizationResult (Optimization.jl:145-155); `stresses` is a StressField (device-computed von Mises /
    element-energy arrays + on-demand Gauss-point tensors) instead of a Dict of tensors."""

    densities: np.ndarray
    displacements: np.ndarray
    stresses: object
    energy: float
    volume: float
    iterations: int
    converged: bool
    energy_history: list
    volume_history: list
    iteration_stats: list = field(default_factory=list)
    checkpoints: dict = field(default_factory=dict)


def _configure(ctx, grid, params: OptimizationParameters, fcache: FilterCache):
    ftype = {"sensitivity": _lib.ESIMP_FILTER_SENSITIVITY, "density": _lib.ESIMP_FILTER_DENSITY}[str(params.filter_type).lstrip(":")]
    ctx.set_params(E0=params.E0, Emin=params.Emin, nu=params.nu, p=params.p, volume_fraction=params.volume_fraction,
                   move_limit=params.move_limit, damping=params.damping, filter_radius=fcache.filter_radius, filter_type=ftype,
                   pcg_rtol=params.pcg_rtol, pcg_maxit=params.pcg_maxit, pcg_check_every=params.pcg_check_every,
                   matvec_variant=params.matvec_variant, warm_start=1,
                   precond=2 if params.preconditioner == "gmg_fp32" else
                   int(params.preconditioner == "gmg" or (params.preconditioner == "auto" and fcache.structured)))
    if fcache.structured:
        ctx.filter_build_structured()
    else:
        ctx.filter_build_csr(fcache.rowptr, fcache.col, fcache.cell_centers, fcache.element_volumes, fcache.filter_radius)


def simp_optimize(grid, dh, cellvalues, loads, boundary_conditions, params: OptimizationParameters, acceleration_data=None,
                  K: StiffnessOperator | None = None, on_iteration=None) -> OptimizationResult:
    """simp_optimize (Optimization.jl:178-565).  The loop control is here; the loop BODY (:273-379)
    is one `esimp_iterate` call whose state never leaves the device."""
    if K is None:
        raise ValueError("pass the operator handle returned by setup_problem as K=")
    ctx = K.ctx
    volumes = calculate_element_volumes(grid)
    fcache = create_filter_cache(grid, params.filter_radius, volumes)  # :238
    # loads are density independent: assemble f once (the reference re-assembles it every iteration, :314)
    f = np.zeros(dh.ndofs)
    for load in loads:
        apply_load_condition_(f, load)
    prescribed = np.zeros(0, dtype=np.int64)
    for ch in boundary_conditions:
        prescribed = np.union1d(prescribed, ch.prescribed_dofs)
    _configure(ctx, grid, params, fcache)
    ctx.set_dirichlet(prescribed)
    ctx.set_load(f)
    if acceleration_data is not None:  # (acceleration_vector, base_density), :195-198, applied on every assembly :300-311
        ctx.set_body_force(acceleration_data[0], acceleration_data[1])
    else:
        ctx.set_body_force(None)
    ctx.init_design(None)  # fill(volume_fraction), :222
    energy_history, volume_history, stats = [], [], []
    triggered = [False] * len(params.tolerance_checkpoints)
    checkpoints = {}
    converged = False
    iteration = 0
    for iteration in range(1, params.max_iterations + 1):
        s = ctx.iterate()
        energy_history.append(s.energy)  # :324-325
        volume_history.append(s.volume)
        stats.append({k: getattr(s, k) for k, _ in s._fields_})
        if on_iteration is not None:
            on_iteration(iteration, s)
        for idx, cp in enumerate(params.tolerance_checkpoints):  # :407-445 (export stays with the caller)
            if not triggered[idx] and s.change < cp:
                triggered[idx] = True
                # :417-424: stresses of the checkpoint state, computed where u lives
                checkpoints[cp] = dict(iteration=iteration, densities=ctx.get_physical_density(), energy=s.energy,
                                       stresses=StressField(ctx, ctx.stress_fields()))
        if s.change < params.tolerance:  # :484-488
            converged = True
            break
    energy, volume, _ = ctx.final_analysis()  # :494-535
    stresses = StressField(ctx, ctx.stress_fields())  # :537-539
    return OptimizationResult(ctx.get_physical_density(), ctx.get_displacement(), stresses, energy, volume, iteration, converged,
                              energy_history, volume_history, stats, checkpoints)
