"""Round-2 study: how much of the growth of the CG iteration count on 0/1 designs is due to the coarse operators?
Compares, on a SIMP design grown by the oracle loop, GMG-PCG with the rediscretised coarse operator for several ways of
averaging the modulus of the 8 children (arithmetic = the CUDA path, geometric, harmonic, power means) against the
Galerkin operator P'AP (the best a V-cycle with this P can do).  numpy / scipy, CPU only.
    python scripts/gmg_coarse_study.py [nx ny nz] [simp_iters]"""
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))
import gmg_prototype as gp  # noqa: E402
from oracle import simp_oracle as o  # noqa: E402


def coarsen_E(E, how):
    nz, ny, nx = E.shape
    c = E.reshape(nz // 2, 2, ny // 2, 2, nx // 2, 2).transpose(0, 2, 4, 1, 3, 5).reshape(nz // 2, ny // 2, nx // 2, 8)
    if how == "arith":
        m = c.mean(axis=-1)
    elif how == "geom":
        m = np.exp(np.log(c).mean(axis=-1))
    elif how == "harm":
        m = 1.0 / (1.0 / c).mean(axis=-1)
    elif how.startswith("pow"):
        q = float(how[3:])
        m = (c ** q).mean(axis=-1) ** (1.0 / q)
    elif how == "ag":  # mean of arithmetic and geometric
        m = 0.5 * (c.mean(axis=-1) + np.exp(np.log(c).mean(axis=-1)))
    else:
        raise ValueError(how)
    return 2.0 * m  # K0(2h) = 2 K0(h)


def build(nel, Efine, fixed, nlev, how):
    levels = []
    nx, ny, nz = nel
    E = Efine.reshape(nz, ny, nx)
    fixmask = fixed.reshape(nz + 1, ny + 1, nx + 1, 3)
    k0 = None
    Aprev = None
    for l in range(nlev):
        g = o.generate_grid_hex((nx, ny, nz))
        nat = gp.native(g)
        if k0 is None:
            k0 = o.unit_element_matrices(g, 0.3, 1.0)[0]
        keep = 1.0 - fixmask.reshape(-1)
        if how == "galerkin" and l > 0:
            P = levels[-1]["P"]
            A = (P.T @ Aprev @ P).tocsr()
            Pm = sp.diags(keep)
            A = (Pm @ A @ Pm + sp.diags((1 - keep) * A.diagonal().mean())).tocsr()
        else:
            K = o.assemble_stiffness(g, nat, E.ravel(), np.broadcast_to(k0, (g.n_cells, 24, 24)))
            Pm = sp.diags(keep)
            A = (Pm @ K @ Pm + sp.diags((1 - keep) * K.diagonal().mean())).tocsr()
        levels.append(dict(A=A, dinv=1.0 / A.diagonal(), keep=keep, nel=(nx, ny, nz)))
        Aprev = A
        if l == nlev - 1:
            break
        Px, Py, Pz = gp.prolong_1d(nx // 2), gp.prolong_1d(ny // 2), gp.prolong_1d(nz // 2)
        levels[-1]["P"] = sp.kron(sp.kron(Pz, sp.kron(Py, Px)), sp.eye(3)).tocsr()
        if how != "galerkin":
            E = coarsen_E(E, how)
        else:
            E = coarsen_E(E, "arith")
        fixmask = fixmask[::2, ::2, ::2]
        nx, ny, nz = nx // 2, ny // 2, nz // 2
    for L in levels:
        L["lmax"] = 1.05 * gp.lam_max(L["A"], L["dinv"])
    return levels


def main():
    nel = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 32, 16)
    its = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    g, dofs, f, pres = o.cantilever_problem(nel)
    cache = Path(f"/tmp/rho_{nel[0]}x{nel[1]}x{nel[2]}_{its}.npy")
    if cache.exists():
        rho = np.load(cache)
    else:
        t = time.time()
        r = o.simp_optimize(g, dofs, f, pres, o.OptimizationParameters(max_iterations=its, tolerance=0.0, volume_fraction=0.5), solver="pcg", pcg_rtol=1e-7)
        rho = r["design"]
        np.save(cache, rho)
        print(f"oracle loop {time.time() - t:.0f} s", flush=True)
    print(f"{nel} after {its} SIMP iterations: gray fraction {np.mean((rho > 0.1) & (rho < 0.9)):.3f}, min {rho.min():.1e}", flush=True)
    nat = gp.native(g)
    fixed = np.zeros((g.n_nodes, 3))
    fixed[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))] = 1
    fn = np.zeros(3 * g.n_nodes)
    fn[nat.ravel()] = f[dofs.ravel()]
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    nlev = 4
    for how in ("arith", "ag", "pow0.5", "geom", "pow-0.5", "harm", "galerkin"):
        levels = build(nel, E, fixed.ravel(), nlev, how)
        A = levels[0]["A"]
        b = fn * levels[0]["keep"]
        for nu in (2, 3):
            t = time.time()
            x, it = gp.pcg(A, b, lambda r: gp.vcycle(levels, 0, r, nu), rtol=1e-8, maxit=400)
            print(f"  coarse modulus {how:9s} V({nu},{nu}): {it:4d} CG iterations ({time.time() - t:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
