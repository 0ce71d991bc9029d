"""The reference README script (README.md:110-153), line for line through the host mirror (easysimp_b200.host), at BASELINE
config 5: 512x256x256 unit cubes, 33.5 M cells, 101.6 M dofs, ONE GPU (caller-side setup at scale).

    python scripts/host_mirror_c5.py [--nel 512 256 256] [--iters 3] [--out gpurun_out/r2_host_mirror_c5.json]

Every caller-side stage is timed on the host clock: generate_grid, setup_problem (first-encounter dof numbering + the
library's grid setup), the two node pickers, assemble_stiffness_matrix_simp! on the operator handle, apply_fixed_boundary!,
simp_optimize (filter cache + volumes + loads inside it) and create_results_data.  The energies are compared with the ones
bench.py logs for the same problem driven straight through the C ABI in native numbering (profiles/r2_end_bench_c5_n1.json).
Needs a B200 (no CPU fallback)."""
import argparse
import json
import resource
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "easysimp.jl_b200"))
import easysimp_b200 as es  # noqa: E402


def run(nel, iters, rtol=1e-8):
    t = {}
    clock = time.perf_counter

    def stage(name, t0):
        t[name] = round(clock() - t0, 3)

    nx, ny, nz = nel
    t0 = clock()
    grid = es.generate_grid(es.Hexahedron, nel, (0.0, 0.0, 0.0), (float(nx), float(ny), float(nz)))
    stage("generate_grid", t0)
    t0 = clock()
    dh, cellvalues, K, f = es.setup_problem(grid, device=0)
    stage("setup_problem", t0)
    t0 = clock()
    fixed_nodes = es.select_nodes_by_plane(grid, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    force_nodes = es.select_nodes_by_circle(grid, [float(nx), 0.0, nz / 2.0], [1.0, 0.0, 0.0], 1.0)
    stage("select_nodes", t0)
    t0 = clock()
    material_model = es.create_simp_material_model(1.0, 0.3, 1e-9, 3.0)
    es.assemble_stiffness_matrix_simp_(K, f, dh, cellvalues, material_model, np.full(grid.getncells(), 0.5))
    stage("assemble_stiffness_matrix_simp", t0)
    t0 = clock()
    ch_fixed = es.apply_fixed_boundary_(K, f, dh, fixed_nodes)
    stage("apply_fixed_boundary", t0)
    per_iter = []
    last = [clock()]

    def on_iteration(it, s):
        now = clock()
        per_iter.append({"iteration": it, "wall_s": round(now - last[0], 3), "energy": s.energy, "volume": s.volume, "change": s.change,
                         "pcg_iters": s.pcg_iters, "device_ms": round(s.ms_filter + s.ms_solve + s.ms_sens + s.ms_oc, 2)})
        last[0] = now

    params = es.OptimizationParameters(E0=1.0, Emin=1e-9, volume_fraction=0.5, max_iterations=iters, tolerance=0.0, filter_radius=1.5,
                                       pcg_rtol=rtol)
    t0 = clock()
    last[0] = t0
    results = es.simp_optimize(grid, dh, cellvalues, [es.PointLoad(dh, sorted(force_nodes), [0.0, -1.0, 0.0])], [ch_fixed], params, K=K,
                               on_iteration=on_iteration)
    stage("simp_optimize", t0)
    t["simp_optimize_before_first_iteration"] = round(per_iter[0]["wall_s"] - per_iter[0]["device_ms"] / 1e3, 3)
    t0 = clock()
    data = es.create_results_data(grid, dh, cellvalues, results)
    stage("create_results_data", t0)
    out = {"nel": list(nel), "n_cells": grid.getncells(), "n_dofs": dh.ndofs, "n_fixed_nodes": len(fixed_nodes), "n_force_nodes": len(force_nodes),
           "host_seconds": t, "iterations": per_iter, "final_energy": results.energy, "final_volume_fraction": data["volume_fraction"],
           "max_von_mises": float(results.stresses.max_von_mises), "max_displacement": float(data["displacement_magnitude"].max()),
           "host_peak_rss_gb": round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0, 2)}
    K.ctx.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nel", type=int, nargs=3, default=[512, 256, 256])
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = run(tuple(a.nel), a.iters)
    ref = ROOT / "profiles" / "r2_end_bench_c5_n1.json"
    if tuple(a.nel) == (512, 256, 256) and ref.exists():
        hist = json.load(open(ref))["history"]
        res["energy_rel_diff_vs_c_abi_native_numbering"] = [abs(r["energy"] - h["energy"]) / h["energy"] for r, h in zip(res["iterations"], hist)]
    line = json.dumps(res)
    print(line)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(line + "\n")
