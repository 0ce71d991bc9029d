#!/usr/bin/env python
"""bench.py -- SIMP iterations/sec on the synthetic hex cantilever (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (C ABI, libeasysimp_b200.so)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path, restated (oracle/cpu_simp.c)
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1: one rank per GPU, x-slab partition, NCCL

A "step" is ONE SIMP iteration (Optimization.jl:271-379): modulus + Jacobi diagonal, FP64 PCG solve to
`--pcg-rtol` (warm-started), element sensitivities + energy, sensitivity filter, OC bisection, metrics.
Workload at every N: the 512x256x256 unit-cube cantilever (33.5 M cells, 101.6 M dofs; strong scaling),
clamped at x = 0, unit load at the tip (SURVEY 8d recipe).  `value` times esimp_iterate with all state
resident in HBM; `e2e` times the same step driven through the C ABI with HOST buffers (design densities
copied host->device before and device->host after every iteration, from/to pinned memory).
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

WORKLOADS = {"c5": (512, 256, 256), "c4": (240, 80, 80), "c1": (60, 20, 4)}
CALIB = ROOT / "bench_calibration.json"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=list(WORKLOADS))
    ap.add_argument("--pcg-rtol", type=float, default=1e-8)
    ap.add_argument("--filter-radius", type=float, default=1.5)
    ap.add_argument("--volfrac", type=float, default=0.5)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--precond", default="gmg", choices=["jacobi", "gmg", "gmg32"],
                    help="gmg32 = the V-cycle evaluated in fp32 inside the fp64 CG (opt-in, csrc/multigrid32.cu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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


def cpu_sample(nel, args, pcg_iters_per_step, n_pcg=3):
    """Bounded sample of the CPU path at full size: `n_pcg` PCG iterations + sensitivities + filter + the
    whole OC bisection, timed with all host threads; the PCG part is scaled to `pcg_iters_per_step`."""
    from oracle import cpu_port as cp
    from oracle import simp_oracle as o

    nx, ny, nz = nel
    pres, f = cantilever(nel)
    fixed = np.zeros(f.size, dtype=np.uint8)
    fixed[pres] = 1
    K0 = np.ascontiguousarray(o.hex8_ke(o.generate_grid_hex((1, 1, 1)).coords[o.generate_grid_hex((1, 1, 1)).cells[0]], *o.lame(1.0, 0.3)))
    rho = np.full(nx * ny * nz, args.volfrac)
    u = np.zeros_like(f)
    prm = dict(E0=1.0, Emin=1e-9, p=3.0, R=args.filter_radius, move=0.2, damping=0.5, volfrac=args.volfrac)
    work = np.empty(4 * f.size)
    t, _, _ = cp.simp_iteration_sample(nel, (1.0, 1.0, 1.0), K0, rho, f, fixed, u, prm, n_pcg, work)
    # cpu_pcg does one extra operator apply for the initial residual and builds the diagonal: count n_pcg+1 applies
    per_pcg = t["pcg"] / (t["pcg_iters"] + 1)
    other = t["modulus"] + t["sens"] + t["filter"] + t["oc"]
    t_step = per_pcg * pcg_iters_per_step + other
    return {"value": 1.0 / t_step, "unit": "iter/s", "cores": cp.num_threads(), "kind": "port",
            "sample": f"{nx}x{ny}x{nz}: {n_pcg} Jacobi-PCG iterations ({per_pcg:.3f} s each, OpenMP, dense 24x24 K_e u_e per cell) scaled to "
                      f"{pcg_iters_per_step:.0f} PCG iterations/step (count of the same PCG at rtol {args.pcg_rtol:g}), plus one full sensitivity pass "
                      f"({t['sens']:.2f} s), sensitivity filter ({t['filter']:.2f} s) and OC bisection ({t['oc']:.2f} s, {t['oc_steps']} steps); "
                      "the reference's own Ferrite assembly + CHOLMOD solve cannot exist at this size (K alone ~100 GB) and Julia is not installed",
            "seconds_per_step": t_step}


def calibrated_pcg_iters(workload, rtol):
    try:
        c = json.load(open(CALIB))[workload]
        return float(c["mean_pcg_iters_per_step"]), c
    except Exception:
        return 2000.0, None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nel = WORKLOADS[args.workload]
    iters, calib = calibrated_pcg_iters(args.workload, args.pcg_rtol)
    vals = []
    base = None
    for _ in range(args.warmup + args.steps):
        base = cpu_sample(nel, args, iters, n_pcg=2)
        vals.append(base["seconds_per_step"])
    t_step = float(np.mean(vals[args.warmup:]))
    base["value"] = 1.0 / t_step
    out = {"impl": "reference", "metric": "simp_iters_per_sec", "value": 1.0 / t_step, "unit": "iter/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, nel, iters, "calibrated (bench_calibration.json)" if calib else "nominal"),
           "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": 1.0 / t_step, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, nel, pcg_iters, pcg_note):
    nx, ny, nz = nel
    return {"workload": f"hex cantilever {nx}x{ny}x{nz} unit cubes, {nx * ny * nz} cells, {3 * (nx + 1) * (ny + 1) * (nz + 1)} dofs",
            "volume_fraction": args.volfrac, "p": 3.0, "E0": 1.0, "Emin": 1e-9, "nu": 0.3, "filter": "sensitivity",
            "filter_radius": args.filter_radius, "move_limit": 0.2, "damping": 0.5,
            "solver": ("CG fp64 + geometric-multigrid V(2,2) Chebyshev-Jacobi preconditioner" if args.precond == "gmg" and args.impl == "b200"
                       else "CG fp64 + geometric-multigrid V(2,2) Chebyshev-Jacobi preconditioner evaluated in fp32"
                       if args.precond == "gmg32" and args.impl == "b200" else "Jacobi-PCG fp64") + ", warm start",
            "pcg_rtol": args.pcg_rtol, "mean_pcg_iters_per_step": pcg_iters, "pcg_iters_source": pcg_note,
            "partition": f"x-slabs over {args.gpus} GPU(s)", "l2": "inputs_larger_than_l2 (813 MB per vector vs 126 MB L2)"}


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
                   pcg_check_every=50, matvec_variant=args.variant, warm_start=1, precond={"jacobi": 0, "gmg": 1, "gmg32": 2}[args.precond])
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
    ctx.init_design(None)
    del f

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stats = []
    for _ in range(args.warmup):
        stats.append(ctx.iterate())
    # ---- timed region: K device-resident iterations
    sampler = ClockSampler(local)
    sampler.start()  # started before the barrier so that it is already polling when the timed region begins
    ctx.profile_enable(True)
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
    op_ms_all, op_n_all = ctx.profile_operator_all()
    ctx.profile_enable(False)
    dev_ms = sum(s.ms_total for s in timed)  # CUDA events on the library's stream, per iterate
    t_local = torch.tensor([max(wall, dev_ms * 1e-3)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    t_timed = float(t_local.item())
    # ---- e2e: the same step through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        n_cells = nx * ny * nz
        rho_host = torch.empty(n_cells, dtype=torch.float64).pin_memory().numpy()
        rho_host[:] = ctx.get_design()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.set_design(rho_host)          # H2D: design densities, 8 B/cell
            s = ctx.iterate()
            ctx.get_design(out=rho_host)      # D2H into the pinned buffer: updated design (Optimization.jl:360) + 7 scalars
            timed.append(s)
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": args.steps / float(t_e2e.item()), "unit": "iter/s", "h2d_bytes_per_step": 8 * n_cells,
               "d2h_bytes_per_step": 8 * n_cells + 8 * 7}
    if rank == 0:
        pcg_iters = float(np.mean([s.pcg_iters for s in timed[:args.steps]]))
        peak, peak_src = measured_peak()
        achieved = mv_bytes / (mv_ms / max(mv_n, 1)) / 1e6 if mv_n else None  # GB/s per rank
        traffic = None
        try:
            traffic = json.load(open(ROOT / "profiles" / "traffic.json")).get(f"{args.workload}_n{world}")
        except Exception:
            pass
        out = {"metric": "simp_iters_per_sec", "value": args.steps / t_timed, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * t_timed / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(args, nel, pcg_iters, "measured in the timed region"),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": "k_matvec_fused (hex8 operator apply)", "achieved": achieved, "peak": peak,
                            "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src}); sustained (kernel timed inside the step)",
                            "algorithmic_bytes_per_launch": mv_bytes, "launches_timed": int(mv_n), "avg_launch_ms": mv_ms / max(mv_n, 1),
                            "share_of_step": op_ms_all / (1e3 * t_timed) if op_n_all else None,
                            "share_note": f"all {op_n_all} fine-grid launches of k_matvec_fused (plain/dot applies + the modes that fuse the multigrid "
                                          f"smoother step / residual into the store); achieved/frac use the {mv_n} plain/dot launches only"},
               "phases_ms_per_step": {k: float(np.mean([getattr(s, k) for s in timed[:args.steps]])) for k in ("ms_filter", "ms_solve", "ms_sens", "ms_oc")},
               "history": [{"energy": s.energy, "change": s.change, "pcg_iters": s.pcg_iters, "pcg_relres": s.pcg_relres, "oc_steps": s.oc_steps}
                           for s in stats + timed]}
        if world == 1 and not args.no_cpu_baseline:
            # the CPU port is Jacobi-PCG: scale its sample with ITS iteration count (measured once with the same
            # Jacobi-PCG on the GPU, bench_calibration.json), not with this arm's multigrid count
            cb = cpu_sample(nel, args, pcg_iters if args.precond == "jacobi" else calibrated_pcg_iters(args.workload, args.pcg_rtol)[0])
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
