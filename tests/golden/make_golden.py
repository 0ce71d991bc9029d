"""Generate the golden fixtures under tests/golden/ from the CPU oracle (oracle/simp_oracle.py).

The reference (Julia) cannot run in this environment and pins no numbers itself, so these vectors are
outputs of the oracle restatement -- which tests/test_oracle.py pins against analytic known-answers.
Run from the repo root:  python tests/golden/make_golden.py [case ...]
Each case stores inputs (grid size, parameters) and per-iteration outputs (energy, volume, change,
lambda, grayness, max|u|) plus the final density field, as a small .npz.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import simp_oracle as o  # noqa: E402

OUT = Path(__file__).resolve().parent


def run_case(name, problem, params, solver, iters_note=""):
    t0 = time.time()
    g, dofs, f, pres = problem
    r = o.simp_optimize(g, dofs, f, pres, params, solver=solver, pcg_rtol=1e-12)
    h = r["history"]
    np.savez_compressed(
        OUT / f"{name}.npz",
        nel=np.array(g.structured[:3]), left=np.array(g.structured[3]), right=np.array(g.structured[4]),
        params=np.array([params.E0, params.Emin, params.nu, params.p, params.volume_fraction, params.filter_radius,
                         params.move_limit, params.damping]),
        filter_type=np.array([0 if params.filter_type == "sensitivity" else 1]),
        max_iterations=np.array([params.max_iterations]),
        energy=np.array([x.energy for x in h]), volume=np.array([x.volume for x in h]),
        change=np.array([x.change for x in h]), lam=np.array([x.lam for x in h]),
        grayness=np.array([x.grayness for x in h]), max_u=np.array([x.max_u for x in h]),
        oc_steps=np.array([x.oc_steps for x in h]),
        final_density=r["densities"].astype(np.float64), final_energy=np.array([r["energy"]]),
        final_volume=np.array([r["volume"]]), solver=np.array([solver]),
    )
    print(f"{name}: {len(h)} iterations, final energy {r['energy']:.8e}, {time.time() - t0:.1f} s {iters_note}", flush=True)


CASES = {
    # BASELINE config 1: README cantilever as BASELINE.json states it (struct defaults, vf 0.5, r 1.5, 50 its)
    "c1_cantilever_60x20x4": lambda: run_case(
        "c1_cantilever_60x20x4", o.cantilever_problem((60, 20, 4)),
        o.OptimizationParameters(volume_fraction=0.5, max_iterations=50, tolerance=0.0, filter_radius=1.5), "direct"),
    # README / test/runtests.jl:73-84 variant: E0=200, Emin=1e-6, vf=0.4, r=2.5, 20 iterations
    "c1b_runtests_60x20x4": lambda: run_case(
        "c1b_runtests_60x20x4", o.cantilever_problem((60, 20, 4)),
        o.OptimizationParameters(E0=200.0, Emin=1e-6, volume_fraction=0.4, max_iterations=20, tolerance=0.0, filter_radius=2.5), "direct"),
    # BASELINE config 2: 4-legs 40x20x20 over 2x1x1, density filter r=2.0, 12 iterations (CPU CG at 1e-12)
    "c2_4legs_40x20x20": lambda: run_case(
        "c2_4legs_40x20x20", o.four_legs_problem((40, 20, 20)),
        o.OptimizationParameters(volume_fraction=0.4, max_iterations=12, tolerance=0.0, filter_radius=2.0, filter_type="density"),
        "pcg"),
    # small density-filter case solved directly
    "c2s_4legs_20x10x10": lambda: run_case(
        "c2s_4legs_20x10x10", o.four_legs_problem((20, 10, 10)),
        o.OptimizationParameters(volume_fraction=0.4, max_iterations=25, tolerance=0.0, filter_radius=2.0, filter_type="density"),
        "direct"),
    # test/Examples/02_sliding_support.jl: per-component constraints (x = 0 face U1 only, one node U2 only), E0 = 200, r 2.0
    "e02_sliding_support_60x20x4": lambda: run_case(
        "e02_sliding_support_60x20x4", o.sliding_support_problem((60, 20, 4)),
        o.OptimizationParameters(E0=200.0, Emin=1e-6, volume_fraction=0.4, max_iterations=30, tolerance=0.0, filter_radius=2.0), "direct"),
    # test/Examples/06_3D_2x1x1_MBB.jl at half resolution: symmetry face + support strip + one U3 node, half-disc load
    "e06_mbb_20x10x10": lambda: run_case(
        "e06_mbb_20x10x10", o.mbb_problem((20, 10, 10)),
        o.OptimizationParameters(E0=1.0, Emin=1e-6, volume_fraction=0.4, max_iterations=25, tolerance=0.0, filter_radius=2.0), "direct"),
    # test/Examples/08_3D_2x1x1_Michell-half_tol_study.jl at half resolution: clamped patches + symmetry plane z = 1
    "e08_michell_half_20x10x10": lambda: run_case(
        "e08_michell_half_20x10x10", o.michell_half_problem((20, 10, 10)),
        o.OptimizationParameters(E0=1.0, Emin=1e-9, volume_fraction=0.4, max_iterations=25, tolerance=0.0, filter_radius=2.0), "direct"),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or CASES):
        CASES[name]()
