"""Multi-GPU parity check of the x-slab path (NCCL halo exchange + allreduces inside libeasysimp_b200.so).

Run on a box with >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank drives its own context (one process per GPU) with GLOBAL host arrays; the library cuts out
its slab.  Checks against the committed golden fixtures (tests/golden/*.npz, oracle outputs):
per-iteration compliance 1e-6 relative, final density L_inf 1e-3 -- the same bars as the 1-GPU tests --
plus operator / filter stage calls against the oracle on rank 0.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es  # noqa: E402
from oracle import simp_oracle as o  # noqa: E402


def allsum(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def make_ctx(rank, world, local, uid, g, dofs, pres, **params):
    nx, ny, nz, left, right = g.structured
    ctx = es.Context(local)
    ctx.set_params(**params)
    ctx.comm_init(rank, world, uid)
    ctx.set_grid_structured(nx, ny, nz, (right[0] - left[0]) / nx, (right[1] - left[1]) / ny, (right[2] - left[2]) / nz, dofs)
    ctx.set_dirichlet(pres)
    return ctx


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def new_uid():
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(es.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().numpy().tobytes())

    ok = True
    # ---- stage calls: operator, diagonal, filters on random data (results are owned-entries-only per rank)
    g = o.generate_grid_hex((23, 9, 7), (0, 0, 0), (23.0, 4.5, 7.0))
    dofs = o.distribute_dofs(g)
    rng = np.random.default_rng(0)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    u = rng.standard_normal(3 * g.n_nodes)
    dcr = -rng.random(g.n_cells)
    pres = np.unique(np.concatenate([dofs[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))].ravel(),
                                     rng.choice(3 * g.n_nodes, 200, replace=False)]))
    fc = o.create_filter_cache(g, 2.5, o.element_volumes(g))
    for variant in (2, 4, 3, 1):
        ctx = make_ctx(rank, world, local, new_uid(), g, dofs, pres, matvec_variant=variant, filter_radius=fc.radius)
        ctx.set_density(rho)
        ctx.filter_build_structured()
        y = allsum(ctx.matvec(u))
        d = allsum(ctx.diagonal())
        fs, fd, ft = allsum(ctx.filter_sens(rho, dcr)), allsum(ctx.filter_dens(rho)), allsum(ctx.filter_dens_T(dcr))
        if rank == 0:
            k0 = o.unit_element_matrices(g, 0.3, 1.0)
            E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
            dd = o.operator_diagonal(g, dofs, E, k0)
            y_ref = o.operator_apply(g, dofs, E, k0, pres, u, float(np.mean(dd)))
            errs = [np.abs(y - y_ref).max() / np.abs(y_ref).max(),
                    np.abs(fs - o.filter_sensitivity(fc, rho, dcr)).max(), np.abs(fd - o.filter_density(fc, rho)).max(),
                    np.abs(ft - o.filter_density_transpose(fc, dcr)).max()]
            print(f"[mgpu] variant {variant}: matvec {errs[0]:.2e} filt_sens {errs[1]:.2e} filt_dens {errs[2]:.2e} filt_densT {errs[3]:.2e}", flush=True)
            ok &= errs[0] < 2e-13 and max(errs[1:]) < 1e-10
        ctx.close()

    # ---- full loops against the golden fixtures
    for name, problem, rtol, precond in (("c1_cantilever_60x20x4", o.cantilever_problem, 1e-10, 0),
                                          ("c2s_4legs_20x10x10", o.four_legs_problem, 1e-10, 0),
                                          ("c1_cantilever_60x20x4", o.cantilever_problem, 1e-10, 1),
                                          ("c2_4legs_40x20x20", o.four_legs_problem, 1e-10, 1)):
        gz = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
        nel = tuple(int(v) for v in gz["nel"])
        g, dofs, f, pres = problem(nel)
        E0, Emin, nu, p, vf, ratio, move, damp = gz["params"]
        R = ratio * o.estimate_element_size(g)
        ctx = make_ctx(rank, world, local, new_uid(), g, dofs, pres, E0=E0, Emin=Emin, nu=nu, p=p, volume_fraction=vf, move_limit=move,
                       damping=damp, filter_radius=R, filter_type=int(gz["filter_type"][0]), pcg_rtol=rtol, pcg_check_every=10,
                       precond=precond)
        ctx.set_load(f)
        ctx.filter_build_structured()
        ctx.init_design(None)
        stats = [ctx.iterate() for _ in range(int(gz["max_iterations"][0]))]
        e_fin, v_fin, _ = ctx.final_analysis()
        dens = allsum(ctx.get_physical_density())
        disp = allsum(ctx.get_displacement())
        if rank == 0:
            energy = np.array([s.energy for s in stats])
            e_err = np.max(np.abs(energy - gz["energy"]) / gz["energy"])
            v_err = np.max(np.abs(np.array([s.volume for s in stats]) - gz["volume"]) / gz["volume"])
            c_err = np.max(np.abs(np.array([s.change for s in stats]) - gz["change"]))
            d_err = np.abs(dens - gz["final_density"]).max()
            print(f"[mgpu] {name} on {world} GPUs, {'multigrid' if precond else 'Jacobi'}-PCG (mean {np.mean([s.pcg_iters for s in stats]):.0f} its): "
                  f"compliance rel err {e_err:.2e}, volume {v_err:.2e}, change {c_err:.2e}, "
                  f"final density Linf {d_err:.2e}, final energy rel {abs(e_fin - gz['final_energy'][0]) / gz['final_energy'][0]:.2e}, "
                  f"max|u| {np.abs(disp).max():.4e} vs {stats[-1].max_u:.4e}", flush=True)
            ok &= e_err < 1e-6 and v_err < 1e-9 and d_err < 1e-3 and c_err < 1e-4
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("[mgpu] ALL OK" if ok else "[mgpu] FAILED", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
