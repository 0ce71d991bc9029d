"""Reference-held parity vectors, when present.  tests/golden/make_julia_golden.jl runs the UNMODIFIED reference (Julia;
not available in the build image) and writes tests/golden/julia_<name>.json; with those files in place this module
holds the CPU oracle (-m "not gpu") and the CUDA path (-m gpu, through the C ABI) against the reference's own numbers:
element sensitivities and filter 1e-10 relative, per-iteration compliance 1e-6 relative (CG rtol 1e-10), final density
L_inf 1e-3 (BASELINE.json north_star).  Without the files every test skips, and the repository's parity status stays
"unpinned by the reference"."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import simp_oracle as o

GOLDEN = Path(__file__).resolve().parent / "golden"
FILES = sorted(GOLDEN.glob("julia_*.json"))
needs_vectors = pytest.mark.skipif(not FILES, reason="no tests/golden/julia_*.json: run tests/golden/make_julia_golden.jl where Julia is installed")


def load(path):
    return json.loads(path.read_text())


def problem_of(ref):
    nel = tuple(int(v) for v in ref["nel"])
    if ref["problem"] == "four_legs":
        return o.four_legs_problem(nel)
    return o.cantilever_problem(nel, (0.0, 0.0, 0.0), tuple(float(v) for v in ref["right"]))


def params_of(ref, g):
    return o.OptimizationParameters(E0=ref["E0"], Emin=ref["Emin"], nu=ref["nu"], p=ref["p"], volume_fraction=ref["volume_fraction"],
                                    max_iterations=int(ref["max_iterations"]), tolerance=ref["tolerance"],
                                    filter_radius=ref["filter_radius"] * o.estimate_element_size(g), filter_type=ref["filter_type"])


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / np.abs(b).max())


@needs_vectors
@pytest.mark.parametrize("path", FILES, ids=lambda p: p.stem)
def test_oracle_matches_the_reference(path):
    """the CPU restatement (direct solve) against the reference's own run: this is what pins the oracle"""
    ref = load(path)
    g, dofs, f, pres = problem_of(ref)
    prm = params_of(ref, g)
    r = o.simp_optimize(g, dofs, f, pres, prm, solver="direct", keep_fields=True)
    n = int(ref["iterations"])
    assert r["iterations"] == n
    assert rel([h.energy for h in r["history"]], ref["energy_history"]) < 1e-6
    assert rel([h.volume for h in r["history"]], ref["volume_history"]) < 1e-9
    assert abs(r["energy"] - ref["final_energy"]) / ref["final_energy"] < 1e-6
    assert np.abs(r["densities"] - np.asarray(ref["final_density"])).max() < 1e-3
    assert rel(r["fields"][0]["dc"], ref["first_dc"]) < 1e-10
    assert rel(r["fields"][0]["dcf"], ref["first_dc_filtered"]) < 1e-10


@needs_vectors
@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=lambda p: p.stem)
def test_cuda_path_matches_the_reference(path, gpu_ctx_factory):
    ref = load(path)
    g, dofs, f, pres = problem_of(ref)
    prm = params_of(ref, g)
    nx, ny, nz, left, right = g.structured
    ctx = gpu_ctx_factory()
    ctx.set_params(E0=prm.E0, Emin=prm.Emin, nu=prm.nu, p=prm.p, volume_fraction=prm.volume_fraction, move_limit=prm.move_limit,
                   damping=prm.damping, filter_radius=prm.filter_radius, filter_type=1 if prm.filter_type == "density" else 0,
                   pcg_rtol=1e-10, pcg_check_every=10, precond=1)
    ctx.set_grid_structured(nx, ny, nz, (right[0] - left[0]) / nx, (right[1] - left[1]) / ny, (right[2] - left[2]) / nz, dofs)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.filter_build_structured()
    # first iteration, stage by stage: sensitivities and filter 1e-10
    rho0 = np.full(g.n_cells, prm.volume_fraction)
    phys0 = ctx.filter_dens(rho0) if prm.filter_type == "density" else rho0
    ctx.set_density(phys0)
    ctx.solve(want_u=False)
    dc, energy = ctx.sensitivities()
    assert abs(energy - ref["first_energy"]) / ref["first_energy"] < 1e-6 and rel(dc, ref["first_dc"]) < 1e-8  # CG rtol 1e-10 in u
    dcf = ctx.filter_dens_T(dc) if prm.filter_type == "density" else ctx.filter_sens(rho0, dc)
    assert rel(dcf, ref["first_dc_filtered"]) < 1e-8
    # the loop
    ctx.init_design(None)
    stats = [ctx.iterate() for _ in range(int(ref["iterations"]))]
    assert rel([s.energy for s in stats], ref["energy_history"]) < 1e-6
    assert rel([s.volume for s in stats], ref["volume_history"]) < 1e-9
    e_fin, _, _ = ctx.final_analysis()
    assert abs(e_fin - ref["final_energy"]) / ref["final_energy"] < 1e-6
    assert np.abs(ctx.get_physical_density() - np.asarray(ref["final_density"])).max() < 1e-3
    ctx.close()
