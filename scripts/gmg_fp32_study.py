"""Round-2 design study (CPU, numpy/scipy): does the multigrid V-cycle still precondition the FP64 CG when it runs in
FP32?  The outer CG (operator, dot products, x and r updates) stays FP64, so the converged displacement is unchanged;
only the preconditioner z = V(r) is evaluated in single precision, with r scaled to unit max-norm on entry (the V-cycle
is linear) so that the SIMP modulus contrast of 1e9 cannot push intermediates out of the FP32 range.  On the device this
halves the bytes of every smoother / residual / transfer kernel and moves their arithmetic from the FP64 pipe
(64 FMA/clk/SM on B200) to the FP32 pipe.

    python scripts/gmg_fp32_study.py            # 48x16x16
    python scripts/gmg_fp32_study.py big        # 96x32x32 (minutes)
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))
sys.argv, argv = [sys.argv[0], "noop"], sys.argv  # keep gmg_prototype's __main__ blocks quiet on import
import gmg_prototype as gp  # noqa: E402
from oracle import simp_oracle as o  # noqa: E402

sys.argv = argv


def to_fp32(levels):
    out = []
    for L in levels:
        M = dict(A=L["A"].astype(np.float32), dinv=L["dinv"].astype(np.float32), keep=L["keep"].astype(np.float32),
                 lmax=np.float32(L["lmax"]), nel=L["nel"])
        if "P" in L:
            M["P"] = L["P"].astype(np.float32)
        out.append(M)
    return out


def vcycle32(levels32, r, nu, coarse_its):
    s = float(np.abs(r).max())
    if s == 0.0:
        return np.zeros_like(r)
    z = gp.vcycle(levels32, 0, (r / s).astype(np.float32), nu, coarse_its)
    assert z.dtype == np.float32
    return z.astype(np.float64) * s


def problem(nel, rho, nlev):
    g, dofs, f, pres = o.cantilever_problem(nel)
    nat = gp.native(g)
    fixed = np.zeros((g.n_nodes, 3))
    fixed[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))] = 1
    fn = np.zeros(3 * g.n_nodes)
    fn[nat.ravel()] = f[dofs.ravel()]
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    levels = gp.build_levels(nel, E, fixed.ravel(), nlev)
    for L in levels:
        L["lmax"] = 1.05 * gp.lam_max(L["A"], L["dinv"])
    return levels, fn * levels[0]["keep"]


def compare(nel, rho, nlev, label, nu=2, cits=16):
    levels, fn = problem(nel, rho, nlev)
    l32 = to_fp32(levels)
    A = levels[0]["A"]
    for rtol in (1e-8, 1e-10):
        t = time.time()
        x64, it64 = gp.pcg(A, fn, lambda r: gp.vcycle(levels, 0, r, nu, cits), rtol=rtol, maxit=500)
        t64 = time.time() - t
        t = time.time()
        x32, it32 = gp.pcg(A, fn, lambda r: vcycle32(l32, r, nu, cits), rtol=rtol, maxit=500)
        t32 = time.time() - t
        res32 = np.linalg.norm(fn - A @ x32) / np.linalg.norm(fn)
        print(f"{label} rtol {rtol:g}: FP64 V-cycle {it64} its ({t64:.1f}s) | FP32 V-cycle {it32} its ({t32:.1f}s), true relres {res32:.1e}, "
              f"|x32-x64|/|x64| {np.linalg.norm(x32 - x64) / np.linalg.norm(x64):.1e}, energy rel diff "
              f"{abs(fn @ x32 - fn @ x64) / abs(fn @ x64):.1e}", flush=True)


def simp_design(nel, its, vf=0.4):
    g, dofs, f, pres = o.cantilever_problem(nel)
    r = o.simp_optimize(g, dofs, f, pres, o.OptimizationParameters(max_iterations=its, tolerance=0.0, volume_fraction=vf),
                        solver="pcg", pcg_rtol=1e-7)
    return r["design"]


if __name__ == "__main__":
    big = len(sys.argv) > 1 and sys.argv[1] == "big"
    nel, nlev = ((96, 32, 32), 5) if big else ((48, 16, 16), 4)
    n = nel[0] * nel[1] * nel[2]
    compare(nel, np.full(n, 0.5), nlev, f"{nel} uniform 0.5")
    rng = np.random.default_rng(1)
    compare(nel, 1e-3 + (1 - 1e-3) * rng.random(n), nlev, f"{nel} random rho")
    if big:
        from scipy.ndimage import gaussian_filter

        fld = gaussian_filter(np.random.default_rng(3).standard_normal((32, 32, 96)), 3.0)
        rho = np.where(fld > np.quantile(fld, 0.5), 1.0, 1e-3).ravel()
        compare(nel, rho, nlev, f"{nel} binary blob design (contrast 1e9)")
    else:
        rho = simp_design(nel, 25)
        print("gray fraction", float(np.mean((rho > 0.1) & (rho < 0.9))), "min", float(rho.min()))
        compare(nel, rho, nlev, f"{nel} SIMP design after 25 its")
