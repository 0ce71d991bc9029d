#!/usr/bin/env python
"""bench.py -- SIMP iterations/sec on the synthetic hex cantilever (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (C ABI, libeasysimp_b200.so)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path, restated (oracle/cpu_simp.c)
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1: one rank per GPU, x-slab partition, NCCL

A "step" is ONE SIMP iteration (Optimization.jl:271-379): modulus + Jacobi diagonal, FP64 CG solve to `--pcg-rtol`
(warm-started), element sensitivities + energy, sensitivity filter, OC bisection, metrics.  Workload at every N: the
512x256x256 unit-cube cantilever (33.5 M cells, 101.6 M dofs; strong scaling), clamped at x = 0, unit load at the tip
(SURVEY 8d recipe).

`value`   K iterations with all state resident in HBM, after W warm-up iterations from the uniform start design.
`e2e`     THE SAME K iterations (the design is reset, the same W warm-up iterations run again) driven through the C ABI with
          HOST buffers: design densities copied host->device before and device->host after every iteration (pinned memory).
`roofline`        the operator kernel's plain / plain+dot launches, CUDA-event timed inside the step.
`roofline_others` the same clock for the other fine-grid kernels of the step (smoother / residual modes of the operator,
                  filter, sensitivities, OC volumes, prolongation, restriction, the fused CG vector kernels).
`cpu_baseline`    oracle/cpu_simp.c (OpenMP port, Jacobi-PCG) on a bounded sample at the full size -- EXTRAPOLATED with the
                  Jacobi-PCG iteration count of the same SIMP iteration window (bench_calibration.json).
`cpu_baseline_measured`  the oracle's direct-solve path (sparse assembly + sparse LU, what the reference does with CHOLMOD)
                  run to completion on the README cantilever 60x20x4, per phase, beside the CUDA path on the same problem.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]

WORKLOADS = {"c5": (512, 256, 256), "c4": (240, 80, 80), "c1": (60, 20, 4), "c3": None}  # c3: imported mesh (stul15 fixture), replicas only
CALIB = ROOT / "bench_calibration.json"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=list(WORKLOADS))
    ap.add_argument("--pcg-rtol", type=float, default=1e-8)
    ap.add_argument("--filter-radius", type=float, default=1.5)
    ap.add_argument("--volfrac", type=float, default=0.5)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--precond", default="gmg", choices=["jacobi", "gmg"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-measured-iters", type=int, default=3, help="SIMP iterations of the direct-solve oracle on 60x20x4 (0 = skip)")
    return ap.parse_args()


def host_threads():
    """threads the CPU legs use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1 for nproc > 1;
    the CPU port is told the count explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cantilever(nel):
    """Dirichlet dofs (x = 0 plane, all components) and the load vector, native numbering
    (node = i + (nx+1)(j + (ny+1)k), dof = 3 node + c); load = [0,-1,0] split over the nodes
    select_nodes_by_circle(grid, [Lx, 0, Lz/2], [1,0,0], 1.0) picks (SelectNodesForBC.jl:67-102)."""
    nx, ny, nz = nel
    node = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)  # noqa: E731
    jj, kk = np.meshgrid(np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    clamp = node(0, jj.ravel(), kk.ravel())
    pres = np.sort((3 * clamp[:, None] + np.arange(3)[None, :]).ravel()).astype(np.int64)
    zc = nz / 2.0
    load_nodes = [node(nx, j, k) for j in range(min(ny, 1) + 1) for k in range(nz + 1)
                  if (j ** 2 + (k - zc) ** 2) ** 0.5 <= 1.0 + 1e-6]
    f = np.zeros(3 * (nx + 1) * (ny + 1) * (nz + 1))
    f[3 * np.array(load_nodes) + 1] = -1.0 / len(load_nodes)
    return pres, f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if not self.rows:
            time.sleep(0.25)  # very short timed regions: wait for at least one sample
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def measured_peak():
    try:
        return json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ CPU legs
def jacobi_pcg_iters(workload, warmup, steps):
    """Jacobi-PCG iterations per SIMP iteration of the window [warmup, warmup + steps), measured once with the CUDA arm's own
    Jacobi-PCG (same algorithm as the CPU port, same rtol, same design trajectory): bench_calibration.json"""
    try:
        c = json.load(open(CALIB))[workload]
        hist = c.get("pcg_iters_by_simp_iteration")
        if hist:
            win = hist[warmup:warmup + steps]
            if len(win) == steps:
                return float(np.mean(win)), f"bench_calibration.json, SIMP iterations {warmup + 1}..{warmup + steps} (the timed window)", True
            return float(np.mean(hist[-min(len(hist), steps):])), "bench_calibration.json, last calibrated iterations (window extends past the calibration)", False
        return float(c["mean_pcg_iters_per_step"]), "bench_calibration.json (mean of its own window)", False
    except Exception:
        return 2000.0, "nominal (no calibration file)", False


def cpu_sample(nel, args, pcg_iters_per_step, iters_note, n_pcg=2):
    """Bounded sample of the CPU port at full size: `n_pcg` Jacobi-PCG iterations + sensitivities + filter + the whole OC
    bisection, timed with all host threads; the PCG part is scaled to `pcg_iters_per_step` (an EXTRAPOLATION)."""
    from oracle import cpu_port as cp
    from oracle import simp_oracle as o

    nthreads = host_threads()
    cp.lib().cpu_set_threads(nthreads)
    nx, ny, nz = nel
    pres, f = cantilever(nel)
    fixed = np.zeros(f.size, dtype=np.uint8)
    fixed[pres] = 1
    g1 = o.generate_grid_hex((1, 1, 1))
    K0 = np.ascontiguousarray(o.hex8_ke(g1.coords[g1.cells[0]], *o.lame(1.0, 0.3)))
    rho = np.full(nx * ny * nz, args.volfrac)
    u = np.zeros_like(f)
    prm = dict(E0=1.0, Emin=1e-9, p=3.0, R=args.filter_radius, move=0.2, damping=0.5, volfrac=args.volfrac)
    work = np.empty(4 * f.size)
    t, _, _ = cp.simp_iteration_sample(nel, (1.0, 1.0, 1.0), K0, rho, f, fixed, u, prm, n_pcg, work)
    # cpu_pcg does one extra operator apply for the initial residual and builds the diagonal: count n_pcg+1 applies
    per_pcg = t["pcg"] / (t["pcg_iters"] + 1)
    other = t["modulus"] + t["sens"] + t["filter"] + t["oc"]
    t_step = per_pcg * pcg_iters_per_step + other
    return {"value": 1.0 / t_step, "unit": "iter/s", "cores": cp.num_threads(), "kind": "port", "extrapolated": True,
            "sample": f"{nx}x{ny}x{nz}: {n_pcg} Jacobi-PCG iterations measured ({per_pcg:.3f} s each, OpenMP x{cp.num_threads()}, dense 24x24 K_e u_e per cell), "
                      f"EXTRAPOLATED to {pcg_iters_per_step:.0f} PCG iterations/step ({iters_note}; rtol {args.pcg_rtol:g}), plus one full sensitivity pass "
                      f"({t['sens']:.2f} s), sensitivity filter ({t['filter']:.2f} s) and OC bisection ({t['oc']:.2f} s, {t['oc_steps']} steps), all measured; "
                      "the reference's own Ferrite assembly + CHOLMOD solve cannot exist at this size (K alone ~100 GB) and Julia is not installed",
            "seconds_per_step": t_step, "seconds_measured": t["pcg"] + other}


def cpu_direct_path(n_iter, threads):
    """The oracle's restatement of the reference's CPU path (sparse assembly, apply!, sparse direct solve, sensitivities,
    filter, OC: Optimization.jl:271-379) run to completion for `n_iter` SIMP iterations on the README cantilever 60x20x4
    (BASELINE config 1: E0 = 1, Emin = 1e-9, nu = 0.3, vf = 0.5, r = 1.5, sensitivity filter); seconds per phase."""
    from oracle import simp_oracle as o

    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    g, dofs, f, pres = o.cantilever_problem((60, 20, 4))
    prm = o.OptimizationParameters(volume_fraction=0.5, filter_radius=1.5 * o.estimate_element_size(g))
    V = o.element_volumes(g)
    Vtot = float(V.sum())
    fc = o.create_filter_cache(g, prm.filter_radius, V)
    k_unit = o.unit_element_matrices(g, prm.nu, prm.E0)
    k0 = o.unit_element_matrices(g, prm.nu, 1.0)
    dv = V / Vtot
    rho = np.full(g.n_cells, prm.volume_fraction)
    ph = {k: 0.0 for k in ("assembly", "apply_bc", "solve", "energy", "sensitivities", "filter", "oc")}
    energies = []

    def loop():
        nonlocal rho
        for _ in range(n_iter):
            t0 = time.perf_counter()
            E = o.simp_modulus(rho, prm.E0, prm.Emin, prm.p)
            K = o.assemble_stiffness(g, dofs, E / prm.E0, k_unit)
            t1 = time.perf_counter()
            Kbc, fbc, _ = o.apply_dirichlet(K, f.copy(), pres)
            t2 = time.perf_counter()
            u = o.solve_direct(Kbc, fbc)
            t3 = time.perf_counter()
            energies.append(0.5 * float(u @ (Kbc @ u)))
            t4 = time.perf_counter()
            dc = o.sensitivities(g, dofs, rho, u, prm.E0, prm.Emin, prm.nu, prm.p, k0)
            t5 = time.perf_counter()
            dcf = o.filter_sensitivity(fc, rho, dc)
            t6 = time.perf_counter()
            rho, _, _, _ = o.optimality_criteria_update(rho, dcf, dv, prm.volume_fraction, Vtot, V, prm.move_limit, prm.damping, fc, False)
            t7 = time.perf_counter()
            for k, d in zip(ph, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6)):
                ph[k] += d

    if threadpool_limits is not None:
        with threadpool_limits(limits=threads):
            loop()
    else:
        loop()
    total = sum(ph.values())
    return {"threads": threads, "iterations": n_iter, "seconds_per_iteration": total / n_iter,
            "phase_seconds_per_iteration": {k: v / n_iter for k, v in ph.items()}, "energies": energies}


def gpu_c1(n_iter):
    """the CUDA path on the same 60x20x4 problem, same number of SIMP iterations (Jacobi-PCG to rtol 1e-10, device-resident)"""
    import easysimp_b200 as es
    from oracle import simp_oracle as o

    g, dofs, f, pres = o.cantilever_problem((60, 20, 4))
    ctx = es.Context(0)
    ctx.set_params(volume_fraction=0.5, filter_radius=1.5 * o.estimate_element_size(g), pcg_rtol=1e-10, pcg_maxit=200000, warm_start=1)
    ctx.set_grid_structured(60, 20, 4, 1.0, 1.0, 1.0, dofs)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.filter_build_structured()
    ctx.init_design(None)
    ctx.iterate()  # lazy allocations, constant uploads
    ctx.init_design(None)
    t0 = time.perf_counter()
    st = [ctx.iterate() for _ in range(n_iter)]
    dt = time.perf_counter() - t0
    ctx.close()
    return {"seconds_per_iteration": dt / n_iter, "energies": [s.energy for s in st], "pcg_iters": [s.pcg_iters for s in st], "solver": "Jacobi-PCG fp64, rtol 1e-10"}


def cpu_baseline_measured(n_iter, with_gpu):
    out = {"config": "README cantilever 60x20x4 (BASELINE config 1), 4 800 cells, 19 215 dofs; the reference's path: sparse assembly + apply! + "
                     "sparse direct solve (SuperLU here, CHOLMOD in the reference) + sensitivities + filter + OC, run to completion",
           "kind": "port (oracle/simp_oracle.py)", "single_thread": cpu_direct_path(n_iter, 1)}
    nt = host_threads()
    if nt > 1:
        out["all_cores"] = cpu_direct_path(n_iter, nt)
    if with_gpu:
        out["cuda_path_same_problem"] = gpu_c1(n_iter)
        e_cpu, e_gpu = np.array(out["single_thread"]["energies"]), np.array(out["cuda_path_same_problem"]["energies"])
        out["compliance_max_rel_diff"] = float(np.max(np.abs(e_cpu - e_gpu) / e_cpu))
    return out


def run_reference(args):
    """the reference arm: the CPU port on the box's host cores (rank 0 only), a bounded sample per step"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nel = WORKLOADS[args.workload]
    iters, note, same_window = jacobi_pcg_iters(args.workload, args.warmup, args.steps)
    # 1 warm-up + 2 measured samples (the sample does not depend on the SIMP iteration: repeating it K times measures nothing new)
    n_warm, n_meas = min(args.warmup, 1), max(1, min(args.steps, 2))
    vals, base = [], None
    for i in range(n_warm + n_meas):
        base = cpu_sample(nel, args, iters, note, n_pcg=2)
        if i >= n_warm:
            vals.append(base["seconds_per_step"])
    t_step = float(np.mean(vals))
    base["value"] = 1.0 / t_step
    out = {"impl": "reference", "metric": "simp_iters_per_sec", "value": 1.0 / t_step, "unit": "iter/s", "n_gpus": args.gpus,
           "steps": n_meas, "warmup": n_warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
           "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "extrapolated": True, "same_config": False,
           "same_config_note": "CPU port = Jacobi-PCG (neither the reference's CHOLMOD nor this repo's multigrid CG); its PCG iteration count per step is "
                               "taken from a Jacobi-PCG run of the CUDA arm on the same SIMP iterations" + ("" if same_window else " (window not covered by the calibration)"),
           "config": workload_config(args, nel, iters, note, solver="Jacobi-PCG fp64 (CPU port), warm start"),
           "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")},
           "e2e": {"value": 1.0 / t_step, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, nel, pcg_iters, pcg_note, solver=None):
    nx, ny, nz = nel
    if solver is None:
        solver = ("CG fp64 + geometric-multigrid V(2,2) Chebyshev-Jacobi preconditioner (fp64)" if args.precond == "gmg" else "Jacobi-PCG fp64") + ", warm start"
    return {"workload": f"hex cantilever {nx}x{ny}x{nz} unit cubes, {nx * ny * nz} cells, {3 * (nx + 1) * (ny + 1) * (nz + 1)} dofs",
            "volume_fraction": args.volfrac, "p": 3.0, "E0": 1.0, "Emin": 1e-9, "nu": 0.3, "filter": "sensitivity",
            "filter_radius": args.filter_radius, "move_limit": 0.2, "damping": 0.5, "solver": solver,
            "pcg_rtol": args.pcg_rtol, "mean_pcg_iters_per_step": pcg_iters, "pcg_iters_source": pcg_note,
            "simp_iterations_timed": f"{args.warmup + 1}..{args.warmup + args.steps} from the uniform start design",
            "partition": f"x-slabs over {args.gpus} GPU(s)", "l2": "inputs_larger_than_l2 (813 MB per vector vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------ imported-mesh workload
def run_b200_c3(args):
    """BASELINE config 3 as a bench line: the reference's data/stul15.vtu (131 920 general hexahedra, 147 922 nodes; fixture
    tests/golden/mesh_stul15_hex.npz) with the boundary conditions of test/Examples/04_gripper_complex.jl minus the body force,
    element-by-element gather-scatter operator with graph colouring, CSR filter, Jacobi-PCG.  Replicas only (no slab partition)."""
    import easysimp_b200 as es

    z = np.load(ROOT / "tests" / "golden" / "mesh_stul15_hex.npz")
    grid = es.Grid(z["points"].astype(np.float64), z["cells"].astype(np.int64), None)
    ctx = es.Context(0)
    ctx.set_params(E0=2.4e3, Emin=1e-6, nu=0.35, p=3.0, volume_fraction=0.3, pcg_rtol=args.pcg_rtol, pcg_maxit=200000, pcg_check_every=50,
                   warm_start=1, precond=0)  # material before the mesh: the element data are built with nu
    dh, cellvalues, K, f = es.setup_problem(grid, ctx=ctx)
    fixed = es.select_nodes_by_circle(grid, [0.0, 75.0, 115.0], [0.0, -1.0, 0.0], 16.11, 1e-3)
    sym = es.select_nodes_by_plane(grid, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    leg = es.select_nodes_by_plane(grid, [0.0, 0.0, -90.0], [0.0, 0.0, 1.0], 1.0)
    cam = es.select_nodes_by_circle(grid, [0.0, 0.0, 5.0], [0.0, 0.0, 1.0], 21.5, 1e-3)
    don = dh.dof_of_node
    pres = np.unique(np.concatenate([don[np.array(sorted(fixed))].ravel(), don[np.array(sorted(sym)), 0]])).astype(np.int64)
    es.apply_force_(f, dh, sorted(leg), [0.0, 0.0, -13000.0])
    es.apply_force_(f, dh, sorted(cam), [0.0, 0.0, -500.0])
    vols = es.calculate_element_volumes(grid)
    fc = es.create_filter_cache(grid, 1.5, vols)
    ctx.set_params(filter_radius=fc.filter_radius)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.filter_build_csr(fc.rowptr, fc.col, fc.cell_centers, fc.element_volumes, fc.filter_radius)
    ctx.init_design(None)
    stats = [ctx.iterate() for _ in range(args.warmup)]
    sampler = ClockSampler(0)
    sampler.start()
    ctx.profile_enable(2)  # one imported-mesh operator apply = memset + one launch per colour: clocked as one category
    launches0 = ctx.launch_count
    import torch

    torch.cuda.synchronize()
    sampler.rows.clear()
    t0 = time.perf_counter()
    timed = [ctx.iterate() for _ in range(args.steps)]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    ms, cnt = ctx.profile_kernels()["um_matvec"]
    ctx.profile_enable(0)
    n_cells, n_dofs = grid.getncells(), dh.ndofs
    # e2e: the same iterations again with the design travelling through host buffers
    rho_host = torch.empty(n_cells, dtype=torch.float64).pin_memory().numpy()
    ctx.init_design(None)
    for _ in range(args.warmup):
        ctx.iterate()
    ctx.get_design(out=rho_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.set_design(rho_host)
        ctx.iterate()
        ctx.get_design(out=rho_host)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    peak, peak_src = measured_peak()
    alg = 16.0 * n_dofs + (8.0 + 24 * 4 + 24 * 8) * n_cells  # u in, y out, E, the 24 dof ids and the 24 node coordinates of every cell
    out = {"metric": "simp_iters_per_sec", "value": args.steps / wall, "unit": "iter/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f64",
           "data": "reference mesh data/stul15.vtu (fixture), example-04 boundary conditions",
           "config": {"workload": f"imported hex8 mesh stul15: {n_cells} cells, {n_dofs} dofs, gather-scatter operator with graph colouring",
                      "E0": 2.4e3, "Emin": 1e-6, "nu": 0.35, "p": 3.0, "volume_fraction": 0.3, "filter": "sensitivity (CSR)", "filter_radius": 1.5,
                      "solver": "Jacobi-PCG fp64, warm start", "pcg_rtol": args.pcg_rtol,
                      "mean_pcg_iters_per_step": float(np.mean([s.pcg_iters for s in timed])), "l2": "working set (13 MB) fits L2: L2-assisted"},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": args.steps / t_e2e, "unit": "iter/s", "h2d_bytes_per_step": 8 * n_cells, "d2h_bytes_per_step": 8 * n_cells + 56},
           "roofline": {"bound": "hbm", "kernel": "k_um_matvec_hex_geom (one apply = memset + one launch per colour)", "achieved": alg / (ms / cnt) / 1e6 if cnt else None,
                        "peak": peak, "unit": "GB/s", "frac": alg / (ms / cnt) / 1e6 / peak if cnt else None, "traffic": None,
                        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "algorithmic_bytes_per_launch": alg, "launches_timed": cnt,
                        "avg_launch_ms": ms / cnt if cnt else None, "share_of_step": ms / (1e3 * wall) if cnt else None,
                        "note": "latency / launch bound at this size (15-20 colour launches of a few thousand cells each); FP64 bound per cell"},
           "history": [{"energy": s.energy, "change": s.change, "pcg_iters": s.pcg_iters} for s in stats + timed]}
    print(json.dumps(out), flush=True)
    ctx.close()


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch

    import easysimp_b200 as es

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nel = WORKLOADS[args.workload]
    nx, ny, nz = nel
    pres, f = cantilever(nel)
    ctx = es.Context(local)
    ctx.set_params(volume_fraction=args.volfrac, filter_radius=args.filter_radius, pcg_rtol=args.pcg_rtol, pcg_maxit=200000,
                   pcg_check_every=50, matvec_variant=args.variant, warm_start=1, precond={"jacobi": 0, "gmg": 1}[args.precond])
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(es.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    ctx.set_grid_structured(nx, ny, nz, 1.0, 1.0, 1.0, None)
    ctx.set_dirichlet(pres)
    ctx.set_load(f)
    ctx.filter_build_structured()
    del f

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- timed region 1: K device-resident iterations after W warm-up iterations from the uniform design
    ctx.init_design(None)
    stats = [ctx.iterate() for _ in range(args.warmup)]
    sampler = ClockSampler(local)
    sampler.start()  # started before the barrier so that it is already polling when the timed region begins
    ctx.profile_enable(1)  # only the plain / dot operator launches are clocked inside the timed region (the roofline figure)
    launches0 = ctx.launch_count
    barrier()
    sampler.rows.clear()
    t0 = time.perf_counter()
    timed = [ctx.iterate() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    mv_ms, mv_n, mv_bytes = ctx.profile_matvec()
    ctx.profile_enable(0)
    dev_ms = sum(s.ms_total for s in timed)  # CUDA events on the library's stream, per iterate
    t_timed = max_over_ranks(max(wall, dev_ms * 1e-3))
    # ---- timed region 2 (e2e): the SAME iterations through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        n_cells = nx * ny * nz
        rho_host = torch.empty(n_cells, dtype=torch.float64).pin_memory().numpy()
        ctx.init_design(None)  # back to the uniform design, displacement guess zero: the same W + K iterations again
        for _ in range(args.warmup):
            ctx.iterate()
        ctx.get_design(out=rho_host)
        barrier()
        t0 = time.perf_counter()
        e2e_stats = []
        for _ in range(args.steps):
            ctx.set_design(rho_host)          # H2D: design densities, 8 B/cell
            e2e_stats.append(ctx.iterate())
            ctx.get_design(out=rho_host)      # D2H into the pinned buffer: updated design (Optimization.jl:360) + 7 scalars
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": args.steps / t_e2e, "unit": "iter/s", "h2d_bytes_per_step": 8 * n_cells, "d2h_bytes_per_step": 8 * n_cells + 8 * 7,
               "same_iterations_as_value": True,
               "max_rel_energy_diff_vs_value_run": float(max(abs(a.energy - b.energy) / abs(b.energy) for a, b in zip(e2e_stats, timed)))}
    # ---- detail pass (not timed for the headline): the same W + K iterations once more with every kernel category and, at N > 1,
    # every halo exchange / allreduce bracketed by CUDA events.  Thousands of event pairs per step cost a few per cent where the
    # step is latency bound (N = 8: 8 %), which is why they are kept out of timed region 1.
    ctx.init_design(None)
    for _ in range(args.warmup):
        ctx.iterate()
    ctx.profile_enable(2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.iterate()
    barrier()
    t_detail = max_over_ranks(time.perf_counter() - t0)
    op_ms_all, op_n_all = ctx.profile_operator_all()
    kernels = ctx.profile_kernels()
    ctx.profile_enable(0)
    if rank == 0:
        pcg_iters = float(np.mean([s.pcg_iters for s in timed]))
        peak, peak_src = measured_peak()
        achieved = mv_bytes / (mv_ms / max(mv_n, 1)) / 1e6 if mv_n else None  # GB/s per rank
        traffic = None
        try:
            traffic = json.load(open(ROOT / "profiles" / "traffic.json")).get(f"{args.workload}_n{world}")
        except Exception:
            pass
        # algorithmic bytes per launch of the other fine-grid kernels on this rank (SURVEY 8d conventions)
        n_cell_r = nx * ny * nz / world
        n_dof_r = (mv_bytes - 8.0 * n_cell_r) / 16.0 if mv_bytes else 3.0 * (nx + 1) * (ny + 1) * (nz + 1) / world
        # operator_smoother: V(2,2) without a stored direction = (z, r -> z') twice and once with z_prev as well: 24 / 24 / 32 B/dof,
        # plus D^-1 as one double per NODE (8/3 B/dof, sweep_kernel.cuh sw_dinv) and the modulus
        per_launch = {"operator_smoother": ((24 + 24 + 32) / 3.0 + 8.0 / 3.0) * n_dof_r + 8 * n_cell_r, "operator_residual": 24 * n_dof_r + 8 * n_cell_r,
                      "filter_stencil": 24 * n_cell_r, "sens_energy": 8 * n_dof_r + 16 * n_cell_r, "oc_volumes": 16 * n_cell_r,
                      "prolong_fine": (16 + 1 + 1 / 3) * n_dof_r, "restrict_fine": (8 + 1) * n_dof_r, "mg_update_first": 40 * n_dof_r,
                      "mg_direction": 40 * n_dof_r}
        others = {}
        for k, nb in per_launch.items():
            ms, cnt = kernels.get(k, (0.0, 0))
            if cnt:
                gbs = nb / (ms / cnt) / 1e6
                others[k] = {"launches": cnt, "avg_launch_ms": ms / cnt, "algorithmic_bytes_per_launch": nb, "achieved": gbs, "unit": "GB/s",
                             "frac": gbs / peak, "share_of_step": ms / (1e3 * t_detail)}
        out = {"metric": "simp_iters_per_sec", "value": args.steps / t_timed, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * t_timed / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(args, nel, pcg_iters, "measured in the timed region"),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": "hex8 operator apply, plain / plain + <p,Ap> store modes (k_matvec_sweep; k_matvec_fused on coarse levels)",
                            "achieved": achieved, "peak": peak,
                            "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src}); sustained (kernel timed inside the step)",
                            "algorithmic_bytes_per_launch": mv_bytes, "launches_timed": int(mv_n), "avg_launch_ms": mv_ms / max(mv_n, 1),
                            "share_of_step": op_ms_all / (1e3 * t_detail) if op_n_all else None,
                            "share_note": f"all {op_n_all} fine-grid launches of the operator kernel (plain/dot applies + the modes that fuse the multigrid "
                                          f"smoother step / residual into the store); achieved/frac use the {mv_n} plain/dot launches only"},
               "roofline_others": others,
               "detail_pass": {"ms_per_step": 1e3 * t_detail / args.steps,
                               "note": "roofline_others, share_of_step and multi_gpu_breakdown come from a second run of the same iterations with every "
                                       "kernel category clocked by CUDA events; value / ms_per_step / roofline.achieved come from timed region 1"},
               "phases_ms_per_step": {k: float(np.mean([getattr(s, k) for s in timed])) for k in ("ms_filter", "ms_solve", "ms_sens", "ms_oc")},
               "history": [{"energy": s.energy, "change": s.change, "pcg_iters": s.pcg_iters, "pcg_relres": s.pcg_relres, "oc_steps": s.oc_steps}
                           for s in stats + timed]}
        if world > 1:
            # rank 0's clocks of what the slab partition adds (CUDA events on the library's stream; a halo / allreduce time
            # includes waiting for the slower neighbour) and two host clocks of the PCG loop, all per SIMP step
            def per_step(k):
                ms, cnt = kernels.get(k, (0.0, 0))
                return {"ms_per_step": ms / args.steps, "count_per_step": cnt / args.steps, "avg_ms": ms / cnt if cnt else None}
            out["multi_gpu_breakdown"] = {k: per_step(k) for k in ("comm_halo", "comm_allreduce_scalars", "comm_level_sum", "mg_below_fine",
                                                                   "host_pcg_wall", "host_pcg_wait", "operator_plain_dot",
                                                                   "operator_smoother", "operator_residual", "mg_update_first",
                                                                   "mg_direction", "prolong_fine", "restrict_fine")}
        if world == 1 and not args.no_cpu_baseline:
            # the CPU port is Jacobi-PCG: scale its sample with ITS iteration count on the same SIMP iterations (measured once with
            # the same Jacobi-PCG on the GPU, bench_calibration.json), not with this arm's multigrid count
            if args.precond == "jacobi":
                it_cpu, note = pcg_iters, "this run's own Jacobi-PCG count"
            else:
                it_cpu, note, _ = jacobi_pcg_iters(args.workload, args.warmup, args.steps)
            cb = cpu_sample(nel, args, it_cpu, note)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")}
    ctx.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.cpu_measured_iters > 0:
        try:
            out["cpu_baseline_measured"] = cpu_baseline_measured(args.cpu_measured_iters, with_gpu=True)
        except Exception as e:  # the measured small-problem leg must never cost the headline line
            out["cpu_baseline_measured"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "c3":
        run_b200_c3(a)
    else:
        run_b200(a)
