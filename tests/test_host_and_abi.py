"""CPU tests of the host logic and of the C-ABI boundary: the shared library loads and exports every
symbol include/easysimp_b200.h declares; without a GPU the product path fails loudly (no fallback);
the host mirror reproduces the oracle's grid / dof / node-selection / filter-cache conventions."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import easysimp_b200 as es
from easysimp_b200 import lib as eslib
from oracle import simp_oracle as o

ROOT = Path(__file__).resolve().parents[1]


def header_functions():
    text = (ROOT / "include" / "easysimp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(esimp_[a-z_A-Z0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = es.load_library()
    declared = header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(eslib.exported_symbols()) == declared, "ctypes prototypes and header are out of sync"
    assert lib.esimp_version() == 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(eslib.EsimpParams) == 8 * 8 + 4 * 2 + 8 + 4 * 4
    assert ctypes.sizeof(eslib.EsimpIterStats) == 7 * 8 + 4 * 4 + 5 * 8
    p = eslib.EsimpParams()
    es.load_library().esimp_default_params(ctypes.byref(p))
    # OptimizationParameters defaults, Optimization.jl:87-102
    assert (p.E0, p.Emin, p.nu, p.p, p.volume_fraction, p.move_limit, p.damping, p.filter_radius) == (1.0, 1e-9, 0.3, 3.0, 0.5, 0.2, 0.5, 1.5)
    assert p.filter_type == eslib.ESIMP_FILTER_SENSITIVITY


def test_no_cpu_fallback_without_device():
    """Without a usable sm_100 device esimp_create must fail with a message; nothing computes on CPU."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(es.EsimpError, match="no CUDA device|CPU fallback|sm_"):
        es.Context(0)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (it is test infrastructure)."""
    pkg = ROOT / "easysimp.jl_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.jl")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
        assert "simp_oracle" not in text and "oracle/" not in text, f


@pytest.mark.parametrize("nel,left,right", [((5, 3, 2), (0, 0, 0), (5, 3, 2)), ((4, 4, 3), (-1.0, 0.5, 0.0), (1.0, 1.5, 0.6))])
def test_host_grid_and_dofs_match_oracle(nel, left, right):
    g = es.generate_grid(es.Hexahedron, nel, left, right)
    go = o.generate_grid_hex(nel, left, right)
    assert np.allclose(g.coords, go.coords, atol=0, rtol=0) and np.array_equal(g.cells, go.cells)
    from easysimp_b200.host import _distribute_dofs

    assert np.array_equal(_distribute_dofs(g), o.distribute_dofs(go))
    assert abs(es.estimate_element_size(g) - o.estimate_element_size(go)) <= 1e-15
    assert np.allclose(es.calculate_element_volumes(g), o.element_volumes(go), rtol=1e-14)
    assert np.array_equal(es.calculate_cell_centers(g), o.cell_centers(go))


@pytest.mark.parametrize("nel", [(1, 1, 1), (2, 1, 1), (1, 3, 2), (5, 4, 3), (13, 7, 5), (2, 2, 9)])
def test_structured_dof_numbering_closed_form(nel):
    """The sort-free-of-connectivity numbering used at 100 M dofs equals Ferrite's first-encounter walk,
    stated generically by the oracle and by the host's own unstructured path."""
    from easysimp_b200.host import Grid, _distribute_dofs, _distribute_dofs_structured

    go = o.generate_grid_hex(nel)
    want = o.distribute_dofs(go)
    assert np.array_equal(_distribute_dofs_structured(*nel), want)
    assert np.array_equal(_distribute_dofs(Grid(go.coords, go.cells, None)), want)  # generic walk, same answer
    # every cell's dofs are 3 consecutive numbers per vertex and cell 0 holds 0..23 in local order
    assert np.array_equal(want[go.cells[0]].ravel(), np.arange(24))


def test_host_constraint_handlers_fixed_and_sliding():
    """apply_fixed_boundary! / apply_sliding_boundary! (FEA.jl:293-340): prescribed dofs per handler, 1-based
    directions, f zeroed there, and the operator handle receives the UNION of all handlers (Optimization.jl:581-583)."""
    from easysimp_b200.host import DofHandler, _distribute_dofs

    class FakeCtx:
        def __init__(self):
            self.calls = []

        def set_dirichlet(self, dofs):
            self.calls.append(np.array(dofs))

    class FakeK:
        def __init__(self):
            self.ctx, self.prescribed, self._density = FakeCtx(), np.zeros(0, dtype=np.int64), None

    g = es.generate_grid(es.Hexahedron, (4, 3, 2))
    dh = DofHandler(g, _distribute_dofs(g))
    K, f = FakeK(), np.ones(dh.ndofs)
    left = es.select_nodes_by_plane(g, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-6)
    bottom = es.select_nodes_by_plane(g, [0.0, 0.0, 0.0], [0.0, 1.0, 0.0], 1e-6)
    ch1 = es.apply_sliding_boundary_(K, f, dh, left, [1])
    assert np.array_equal(ch1.prescribed_dofs, np.sort(dh.dof_of_node[sorted(left), 0]))
    ch2 = es.apply_sliding_boundary_(K, f, dh, bottom, [2, 3])
    assert np.array_equal(ch2.prescribed_dofs, np.sort(dh.dof_of_node[sorted(bottom)][:, 1:].ravel()))
    ch3 = es.apply_fixed_boundary_(K, f, dh, {0})
    assert np.array_equal(ch3.prescribed_dofs, np.sort(dh.dof_of_node[0]))
    union = np.unique(np.concatenate([ch1.prescribed_dofs, ch2.prescribed_dofs, ch3.prescribed_dofs]))
    assert np.array_equal(K.prescribed, union) and np.array_equal(K.ctx.calls[-1], union) and len(K.ctx.calls) == 3
    assert not f[union].any() and f.sum() == dh.ndofs - union.size
    # reference-side oracle statement of the same sets
    go = o.generate_grid_hex((4, 3, 2))
    do = o.distribute_dofs(go)
    assert np.array_equal(ch1.prescribed_dofs, np.sort(do[sorted(o.select_nodes_by_plane(go, [0, 0, 0], [1, 0, 0], 1e-6)), 0]))


def test_check_sensitivity_health_conditions():
    """the three warn-only conditions of OptimalityCriteria.jl:19-40, in the reference's order"""
    import warnings

    rng = np.random.default_rng(0)
    good = -rng.random(1001) - 0.1
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert es.check_sensitivity_health(good) is True
    for bad, msg in ((np.abs(good), "negative"), (np.where(np.arange(1001) < 400, good, -1e-20), "effectively zero"),
                     (np.concatenate([good, [-1e9]]), "range too large")):
        with pytest.warns(UserWarning, match=msg):
            assert es.check_sensitivity_health(bad) is False
    # exactly half negative passes the first test (strict <), and is then judged by the median
    half = np.concatenate([-np.ones(500), np.ones(500)])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert es.check_sensitivity_health(half) is True


def test_host_node_selection_matches_reference_recipe():
    g = es.generate_grid(es.Hexahedron, (60, 20, 4), (0, 0, 0), (60.0, 20.0, 4.0))
    fixed = es.select_nodes_by_plane(g, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    force = es.select_nodes_by_circle(g, [60.0, 0.0, 2.0], [1.0, 0.0, 0.0], 1.0)
    assert len(fixed) == 21 * 5 and len(force) == 4  # README.md:122-124
    go = o.generate_grid_hex((60, 20, 4))
    assert fixed == o.select_nodes_by_plane(go, [0, 0, 0], [1, 0, 0], 1e-3)
    assert force == o.select_nodes_by_circle(go, [60.0, 0.0, 2.0], [1, 0, 0], 1.0)


def test_host_unstructured_filter_cache_matches_oracle():
    go = o.generate_grid_hex((6, 4, 3))
    rng = np.random.default_rng(0)
    coords = go.coords + 0.05 * rng.standard_normal(go.coords.shape)
    g = es.Grid(coords, go.cells.copy(), None)
    V = es.calculate_element_volumes(g)
    gu = o.Grid(coords, go.cells.copy(), None)
    assert np.allclose(V, o.element_volumes(gu), rtol=1e-13)
    fc = es.create_filter_cache(g, 1.5, V)
    fo = o.create_filter_cache(gu, 1.5, V)
    assert fc.filter_radius == fo.radius and np.array_equal(fc.rowptr, fo.rowptr) and np.array_equal(fc.col, fo.col)


def test_optimization_parameters_defaults_match_reference():
    p = es.OptimizationParameters()
    assert (p.E0, p.Emin, p.nu, p.p, p.volume_fraction, p.max_iterations, p.tolerance, p.filter_radius, p.filter_type,
            p.move_limit, p.damping, p.use_cache, p.export_interval, p.export_path, p.task_name, p.tolerance_checkpoints) == (
        1.0, 1e-9, 0.3, 3.0, 0.5, 200, 0.01, 1.5, "sensitivity", 0.2, 0.5, True, 0, "", "SIMP_Optimization", [])


def test_slab_partition_covers_grid():
    for nx, R, he in [(512, 8, 1), (60, 4, 2), (7, 3, 1), (5, 1, 1)]:
        slabs = [es.slab_partition(nx, r, R, he) for r in range(R)]
        assert slabs[0].ex0 == 0 and slabs[-1].ex1 == nx
        owned_planes = []
        for s in slabs:
            assert s.ex1 > s.ex0
            owned_planes += list(range(s.gp0 + s.own_p0, s.gp0 + s.own_p1))
            # layers the operator visits touch exactly the stored planes
            assert s.ge0 + s.mv_e0 == s.gp0 and s.ge0 + s.mv_e1 == s.gp0 + s.npx - 1
            assert s.nex == (s.ex1 - s.ex0) + s.halo_e * ((s.rank > 0) + (s.rank < R - 1))
        assert owned_planes == list(range(nx + 1))
        # the slab-thickness validation of esimp_set_grid_structured uses floor(nx / R), the thinnest slab of the partition, so that
        # every rank takes the same accept / reject decision (no rank is left waiting inside a collective)
        assert min(s.ex1 - s.ex0 for s in slabs) == nx // R


def test_multigrid_slab_hierarchy_invariants():
    """What the slab-aware V-cycle (csrc/multigrid.cu) relies on: on every level each node plane is owned by exactly
    one rank; coarse plane I and fine plane 2I are owned by the SAME rank (restriction / prolongation are local up to
    the one-plane halo); the fine planes 2I-1, 2I+1 a restriction reads and the coarse planes (i-1)/2, (i+1)/2 a
    prolongation reads are stored (owned or ghost) on that rank."""
    from easysimp_b200.partition import level_slab, multigrid_levels

    assert multigrid_levels((512, 256, 256), 1) == [(512 >> l, 256 >> l, 256 >> l) for l in range(9)]
    assert multigrid_levels((512, 256, 256), 8)[-1] == (8, 4, 4) and len(multigrid_levels((512, 256, 256), 8)) == 7
    assert multigrid_levels((60, 20, 4), 1) == [(60, 20, 4), (30, 10, 2), (15, 5, 1)]
    assert multigrid_levels((60, 20, 4), 4) == [(60, 20, 4)]  # 30 layers do not split over 4 ranks
    for nel, R in [((512, 256, 256), 8), ((512, 256, 256), 2), ((64, 16, 16), 2), ((40, 20, 20), 2), ((96, 8, 8), 3)]:
        levels = multigrid_levels(nel, R)
        for lf, lc in zip(levels[:-1], levels[1:]):
            owner_f, owner_c = {}, {}
            stored_f, stored_c = {}, {}
            for r in range(R):
                sf, sc = level_slab(lf[0], r, R), level_slab(lc[0], r, R)
                for p in range(sf.gp0 + sf.own_p0, sf.gp0 + sf.own_p1):
                    assert p not in owner_f
                    owner_f[p] = r
                for p in range(sc.gp0 + sc.own_p0, sc.gp0 + sc.own_p1):
                    assert p not in owner_c
                    owner_c[p] = r
                stored_f[r] = set(range(sf.gp0, sf.gp0 + sf.npx))
                stored_c[r] = set(range(sc.gp0, sc.gp0 + sc.npx))
            assert sorted(owner_f) == list(range(lf[0] + 1)) and sorted(owner_c) == list(range(lc[0] + 1))
            for I, r in owner_c.items():
                assert owner_f[2 * I] == r
                for p in (2 * I - 1, 2 * I + 1):
                    assert p < 0 or p > lf[0] or p in stored_f[r]
            for i, r in owner_f.items():
                for P in ({i // 2} if i % 2 == 0 else {(i - 1) // 2, (i + 1) // 2}):
                    assert P in stored_c[r]
