import os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
import easysimp_b200 as es
from oracle import simp_oracle as o
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for nel in [(16, 8, 8), (64, 16, 16)]:
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0: uid.copy_(torch.frombuffer(bytearray(es.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    g, dofs, f, pres = o.cantilever_problem(nel)
    ctx = es.Context(local)
    ctx.set_params(pcg_rtol=1e-10, precond=1, filter_radius=1.5)
    ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    ctx.set_grid_structured(*nel, 1.0, 1.0, 1.0, dofs)
    ctx.set_dirichlet(pres); ctx.set_load(f)
    ctx.set_density(np.full(g.n_cells, 0.5))
    try:
        u, it, rr, st = ctx.solve()
        t = torch.from_numpy(u).cuda(); dist.all_reduce(t); u = t.cpu().numpy()
        if rank == 0:
            k0 = o.unit_element_matrices(g, 0.3, 1.0)
            Kbc, fbc, _ = o.apply_dirichlet(o.assemble_stiffness(g, dofs, np.full(g.n_cells, 0.125 + 1e-9*(1-0.125)), k0), f.copy(), pres)
            ur = o.solve_direct(Kbc, fbc)
            print(nel, "iters", it, "relres", rr, "err vs direct", np.abs(u-ur).max()/np.abs(ur).max(), flush=True)
    except Exception as e:
        print(rank, nel, "FAILED", e, flush=True)
    ctx.close()
dist.destroy_process_group()
