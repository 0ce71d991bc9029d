"""world_size-2/3 gloo tests (CPU) of the N>1 path's host logic: the x-slab partition, the one-plane
node halo, the redundant left interface layer and the owned-range dot products -- exactly the scheme
csrc/api.cu + csrc/comm.cu implement with NCCL.  Each rank holds its slab of u in the device layout
[plane][comp][y][z], exchanges ghost planes with torch.distributed send/recv, applies the oracle's
element operator to the layers `slab_partition` says it visits, and the gathered owned planes must
equal the single-rank operator; the allreduced owned-range dot must equal the global dot."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nel, halo_e, out_dir):
    for p in (str(ROOT), str(ROOT / "easysimp.jl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from easysimp_b200.partition import slab_partition
    from oracle import simp_oracle as o

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = nel
    s = slab_partition(nx, rank, world, halo_e)
    plane = (ny + 1) * (nz + 1)
    k0 = o.hex8_ke(o.generate_grid_hex((1, 1, 1)).coords[o.generate_grid_hex((1, 1, 1)).cells[0]], *o.lame(1.0, 0.3))
    # global deterministic inputs, every rank cuts out its slab (what upload_dofs / upload_cells do)
    node = np.arange((nx + 1) * plane, dtype=np.uint64)
    u_glob = (2 * o.splitmix64(node[:, None] * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, :], 7) - 1).reshape(nx + 1, ny + 1, nz + 1, 3)
    E_glob = (1e-3 + o.splitmix64(np.arange(nx * ny * nz, dtype=np.uint64), 9)).reshape(nx, ny, nz)
    u_loc = np.zeros((s.npx, 3, ny + 1, nz + 1))
    own = slice(s.own_p0, s.own_p1)
    u_loc[own] = np.transpose(u_glob[s.gp0 + s.own_p0: s.gp0 + s.own_p1], (0, 3, 1, 2))  # ghosts left at 0
    E_loc = E_glob[s.ge0: s.ge0 + s.nex]

    # ---- halo exchange of one node plane each way (comm_halo_nodes)
    t = torch.from_numpy(u_loc)
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(t[s.own_p0].contiguous(), rank - 1))
        left_ghost = torch.empty_like(t[0])
        reqs.append(dist.irecv(left_ghost, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(t[s.own_p1 - 1].contiguous(), rank + 1))
        right_ghost = torch.empty_like(t[0])
        reqs.append(dist.irecv(right_ghost, rank + 1))
    for r in reqs:
        r.wait()
    if rank > 0:
        t[0] = left_ghost
    if rank < world - 1:
        t[s.npx - 1] = right_ghost
    assert np.array_equal(u_loc, np.transpose(u_glob[s.gp0: s.gp0 + s.npx], (0, 3, 1, 2))), "halo planes wrong"

    # ---- operator on the visited layers (owned + redundant left interface layer), store owned planes only
    y_loc = np.zeros_like(u_loc)
    e2p = s.ge0 - s.gp0
    sx = [0, 1, 1, 0, 0, 1, 1, 0]
    sy = [0, 0, 1, 1, 0, 0, 1, 1]
    sz = [0, 0, 0, 0, 1, 1, 1, 1]
    for le in range(s.mv_e0, s.mv_e1):
        lp = le + e2p
        ue = np.stack([u_loc[lp + sx[a], c, sy[a]:ny + sy[a], sz[a]:nz + sz[a]] for a in range(8) for c in range(3)], axis=-1)
        ye = E_loc[le][..., None] * (ue @ k0.T)
        for a in range(8):
            for c in range(3):
                y_loc[lp + sx[a], c, sy[a]:ny + sy[a], sz[a]:nz + sz[a]] += ye[..., 3 * a + c]
    # ---- owned-range dot product + allreduce (vec_dot)
    d = torch.tensor([float(np.vdot(u_loc[own], y_loc[own]))], dtype=torch.float64)
    dist.all_reduce(d)
    np.savez(Path(out_dir) / f"rank{rank}.npz", y=y_loc[own], p0=s.gp0 + s.own_p0, dot=d.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nel,halo_e", [(2, (6, 3, 4), 1), (3, (7, 2, 3), 2)])
def test_slab_operator_matches_global(tmp_path, world, nel, halo_e):
    from oracle import simp_oracle as o

    port = _free_port()
    mp.spawn(_worker, args=(world, port, nel, halo_e, str(tmp_path)), nprocs=world, join=True)
    nx, ny, nz = nel
    g = o.generate_grid_hex(nel)
    plane = (ny + 1) * (nz + 1)
    node = np.arange((nx + 1) * plane, dtype=np.uint64)
    u_glob = (2 * o.splitmix64(node[:, None] * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, :], 7) - 1).reshape(nx + 1, ny + 1, nz + 1, 3)
    E_glob = (1e-3 + o.splitmix64(np.arange(nx * ny * nz, dtype=np.uint64), 9)).reshape(nx, ny, nz)
    # oracle in caller numbering: node id = i + (nx+1)(j + (ny+1)k), cell id = i + nx (j + ny k)
    nat = 3 * np.arange(g.n_nodes)[:, None] + np.arange(3)[None, :]
    u = np.transpose(u_glob, (2, 1, 0, 3)).reshape(-1)
    E = np.transpose(E_glob, (2, 1, 0)).reshape(-1)
    k0 = o.unit_element_matrices(g, 0.3, 1.0)
    y = o.operator_apply(g, nat, E, k0, np.zeros(0, dtype=np.int64), u, 0.0)
    y_glob = np.transpose(y.reshape(nz + 1, ny + 1, nx + 1, 3), (2, 3, 1, 0))  # [x][c][y][z]
    got = np.zeros_like(y_glob)
    seen = np.zeros(nx + 1, dtype=int)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        p0 = int(z["p0"])
        got[p0:p0 + z["y"].shape[0]] = z["y"]
        seen[p0:p0 + z["y"].shape[0]] += 1
        assert abs(z["dot"][0] - u @ y) < 1e-10 * abs(u @ y)
    assert np.all(seen == 1), "every node plane is owned by exactly one rank"
    assert np.abs(got - y_glob).max() < 1e-12 * np.abs(y_glob).max()
