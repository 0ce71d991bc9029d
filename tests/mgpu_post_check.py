"""Multi-GPU parity check of the body-force and stress-field rows (csrc/post_load.cu) on the x-slab path.

Run on a box with >= 2 GPUs (no torch needed: the ranks are plain processes, the NCCL id travels over a pipe):
    python tests/mgpu_post_check.py [n_ranks]
Every rank drives its own context with GLOBAL host arrays; the library cuts out its slab and the getters fill only the
entries the rank owns, so the parent sums the per-rank arrays and holds them against the oracle: stage calls 1e-12 /
1e-10 relative, loops with `acceleration_data` compliance 1e-6 / densities 1e-3 (the 1-GPU bars).
"""
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]

B = np.array([0.3, -9.81, 2.5])
ACC, BASE = np.array([0.0, -9.81, 0.0]), 2.0e-3
STAGE_GRID = ((23, 9, 7), (23.0, 4.5, 7.0))
LOOP_NEL = (24, 8, 4)


def stage_inputs():
    from oracle import simp_oracle as o

    g = o.generate_grid_hex(STAGE_GRID[0], (0, 0, 0), STAGE_GRID[1])
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(0)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    rho[rng.choice(g.n_cells, g.n_cells // 7, replace=False)] = 1e-9
    u = rng.standard_normal(3 * g.n_nodes)
    pres = np.sort(dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel())
    return g, dofs, rho, u, pres


def make_ctx(es, rank, world, uid, g, dofs, pres, **params):
    nx, ny, nz, left, right = g.structured
    ctx = es.Context(rank)
    if params:
        ctx.set_params(**params)
    ctx.comm_init(rank, world, uid)
    ctx.set_grid_structured(nx, ny, nz, (right[0] - left[0]) / nx, (right[1] - left[1]) / ny, (right[2] - left[2]) / nz, dofs)
    ctx.set_dirichlet(pres)
    return ctx


def worker(rank, world, uids, conn):
    import easysimp_b200 as es
    from oracle import simp_oracle as o

    out = {}
    g, dofs, rho, u, pres = stage_inputs()
    ctx = make_ctx(es, rank, world, uids[0], g, dofs, pres)
    ctx.set_density(rho)
    ctx.set_displacement(u)
    out["body"] = ctx.body_force(B, rho * 2700.0)
    sf = ctx.stress_fields()
    out["stress"] = sf
    c0, n = g.n_cells // 3, g.n_cells // 2
    out["sigma"] = ctx.stress_tensors(c0, n)
    ctx.close()
    g, dofs, f, pres = o.cantilever_problem(LOOP_NEL)
    R = 1.5 * o.estimate_element_size(g)
    for i, precond in enumerate((0, 1)):
        ctx = make_ctx(es, rank, world, uids[1 + i], g, dofs, pres, volume_fraction=0.4, filter_radius=R, pcg_rtol=1e-10,
                       pcg_check_every=10, precond=precond)
        ctx.set_load(f)
        ctx.set_body_force(ACC, BASE)
        ctx.filter_build_structured()
        ctx.init_design(None)
        stats = [ctx.iterate() for _ in range(8)]
        e_fin, _, _ = ctx.final_analysis()
        out[f"loop{precond}"] = dict(energy=[s.energy for s in stats], e_fin=e_fin, dens=ctx.get_physical_density(),
                                     stress=ctx.stress_fields())
        ctx.close()
    conn.send(out)
    conn.close()


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    import easysimp_b200 as es
    from oracle import simp_oracle as o

    uids = [es.comm_unique_id() for _ in range(3)]
    mpctx = mp.get_context("spawn")
    pipes, procs = [], []
    for r in range(world):
        a, b = mpctx.Pipe(duplex=False)
        p = mpctx.Process(target=worker, args=(r, world, uids, b))
        p.start()
        pipes.append(a)
        procs.append(p)
    # the oracle runs here while the ranks work
    g, dofs, rho, u, pres = stage_inputs()
    body_ref = o.body_force(g, dofs, B, rho * 2700.0)
    st_ref = o.stresses_simp(g, dofs, u, 1.0, 0.3, 1e-9, 3.0, rho)
    gl, dl, fl, pl = o.cantilever_problem(LOOP_NEL)
    loop_ref = o.simp_optimize(gl, dl, fl, pl, o.OptimizationParameters(volume_fraction=0.4, max_iterations=8, tolerance=0.0),
                               acceleration_data=(ACC, BASE))
    outs = []
    for a in pipes:
        if not a.poll(240):  # a rank died or hangs in a collective: do not wait for ever
            for p in procs:
                p.terminate()
            print("[mgpu-post] FAILED: no answer from a rank", flush=True)
            sys.exit(1)
        outs.append(a.recv())
    for p in procs:
        p.join()
    ok = all(p.exitcode == 0 for p in procs)
    tot = lambda key, sub=None: sum((o_[key] if sub is None else o_[key][sub]) for o_ in outs)  # noqa: E731
    errs = dict(body=relmax(tot("body"), body_ref), vm_q1=relmax(tot("stress", "vm_q1"), st_ref["vm_q1"]),
                vm_avg=relmax(tot("stress", "vm_avg"), st_ref["vm_avg"]),
                energy=relmax(tot("stress", "element_energy"), st_ref["element_energy"]))
    c0, n = g.n_cells // 3, g.n_cells // 2
    errs["sigma"] = relmax(tot("sigma"), st_ref["sigma"][c0:c0 + n])
    same_max = all(o_["stress"]["max_cell"] == st_ref["max_cell"] and abs(o_["stress"]["max_vm"] - st_ref["max_vm"]) <= 1e-12 * st_ref["max_vm"]
                   for o_ in outs)
    print(f"[mgpu-post] {world} GPUs stage calls: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()) + f" max/argmax equal on all ranks: {same_max}")
    ok &= errs["body"] < 1e-12 and max(errs[k] for k in ("vm_q1", "vm_avg", "energy", "sigma")) < 1e-10 and same_max
    e_ref = np.array([h.energy for h in loop_ref["history"]])
    for precond in (0, 1):
        key = f"loop{precond}"
        e_err = max(float(np.max(np.abs(np.array(o_[key]["energy"]) - e_ref) / e_ref)) for o_ in outs)
        f_err = max(abs(o_[key]["e_fin"] - loop_ref["energy"]) / loop_ref["energy"] for o_ in outs)
        d_err = float(np.abs(tot(key, "dens") - loop_ref["densities"]).max())
        cells = {o_[key]["stress"]["max_cell"] for o_ in outs}
        print(f"[mgpu-post] loop with acceleration_data, {'multigrid' if precond else 'Jacobi'}-PCG: compliance rel err {e_err:.2e}, "
              f"final energy {f_err:.2e}, final density Linf {d_err:.2e}, max von Mises cell agreed: {len(cells) == 1}")
        ok &= e_err < 1e-6 and f_err < 1e-6 and d_err < 1e-3 and len(cells) == 1
    print("[mgpu-post] ALL OK" if ok else "[mgpu-post] FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
