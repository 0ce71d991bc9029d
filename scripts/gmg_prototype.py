"""numpy prototype of the GMG-PCG planned for the CUDA path: rediscretised coarse operators (E' = 1/4 sum of the 8
children: mean E times the factor 2 of K0(2h) = 2 K0(h)), trilinear prolongation P, restriction P', Chebyshev-Jacobi
smoothing, injected Dirichlet masks.  Compares iteration counts with Jacobi-PCG on SIMP-like density fields."""
import sys, time
from pathlib import Path
import numpy as np, scipy.sparse as sp
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import simp_oracle as o

def native(g): return 3*np.arange(g.n_nodes)[:,None]+np.arange(3)[None,:]

def prolong_1d(nc):  # (2nc+1) x (nc+1)
    nf = 2*nc
    P = sp.lil_matrix((nf+1, nc+1))
    for I in range(nc+1): P[2*I, I] = 1.0
    for I in range(nc): P[2*I+1, I] = 0.5; P[2*I+1, I+1] = 0.5
    return P.tocsr()

def build_levels(nel, Efine, fixed_nodes_fine, nlev):
    levels = []
    nx, ny, nz = nel
    E = Efine.reshape(nz, ny, nx)  # cell id = i + nx (j + ny k) -> [k][j][i]
    fixmask = fixed_nodes_fine.reshape(nz+1, ny+1, nx+1, 3)
    k0 = None
    for l in range(nlev):
        g = o.generate_grid_hex((nx, ny, nz))
        nat = native(g)
        if k0 is None: k0 = o.unit_element_matrices(g, 0.3, 1.0)[0]
        K = o.assemble_stiffness(g, nat, E.ravel(), np.broadcast_to(k0, (g.n_cells, 24, 24)))
        keep = 1.0 - fixmask.reshape(-1)
        Pm = sp.diags(keep)
        d = K.diagonal(); m = d.mean()
        A = (Pm @ K @ Pm + sp.diags((1-keep)*m)).tocsr()
        levels.append(dict(A=A, dinv=1.0/A.diagonal(), keep=keep, nel=(nx, ny, nz)))
        if l == nlev-1: break
        # coarsen
        Px, Py, Pz = prolong_1d(nx//2), prolong_1d(ny//2), prolong_1d(nz//2)
        Pn = sp.kron(Pz, sp.kron(Py, Px))                    # node id = i + (nx+1)(j + (ny+1)k)
        P = sp.kron(Pn, sp.eye(3)).tocsr()
        levels[-1]['P'] = P
        E = 0.25 * E.reshape(nz//2, 2, ny//2, 2, nx//2, 2).sum(axis=(1, 3, 5))
        fixmask = fixmask[::2, ::2, ::2]
        nx, ny, nz = nx//2, ny//2, nz//2
    return levels

def lam_max(A, dinv, iters=30):
    rng = np.random.default_rng(0); x = rng.standard_normal(A.shape[0])
    for _ in range(iters):
        x = dinv*(A@x); lam = np.linalg.norm(x); x /= lam
    return lam

def cheby(A, dinv, r, z, k, lmax, zero_guess):
    lmin = lmax/ 6.0   # smooth the upper part of the spectrum
    theta, delta = 0.5*(lmax+lmin), 0.5*(lmax-lmin)
    sigma = theta/delta; rho = 1.0/sigma
    res = r if zero_guess else r - A@z
    d = dinv*res/theta
    z = z + d
    for _ in range(k-1):
        rho_new = 1.0/(2*sigma - rho)
        res = r - A@z
        d = rho_new*rho*d + (2*rho_new/delta)*dinv*res
        z = z + d; rho = rho_new
    return z

def vcycle(levels, l, r, nu=2, coarse_its=40):
    L = levels[l]
    if l == len(levels)-1:
        return cheby(L['A'], L['dinv'], r, np.zeros_like(r), coarse_its, L['lmax'], True)
    z = cheby(L['A'], L['dinv'], r, np.zeros_like(r), nu, L['lmax'], True)
    res = (r - L['A']@z)
    rc = (L['P'].T @ res) * levels[l+1]['keep']
    zc = vcycle(levels, l+1, rc, nu, coarse_its)
    z = z + (L['P'] @ zc) * L['keep']
    return cheby(L['A'], L['dinv'], r, z, nu, L['lmax'], False)

def pcg(A, f, M, rtol=1e-8, maxit=20000):
    x = np.zeros_like(f); r = f.copy(); z = M(r); p = z.copy(); rz = r@z; nf = np.linalg.norm(f); it = 0
    while np.linalg.norm(r) > rtol*nf and it < maxit:
        q = A@p; a = rz/(p@q); x += a*p; r -= a*q; z = M(r); rzn = r@z; p = z + (rzn/rz)*p; rz = rzn; it += 1
    return x, it

def run(nel, rho, nlev, label, nu=2):
    g, dofs, f, pres = o.cantilever_problem(nel)
    nat = native(g)
    # native numbering problem
    fixed = np.zeros((g.n_nodes, 3)); fixed[np.array(sorted(o.select_nodes_by_plane(g, [0,0,0],[1,0,0],1e-6)))] = 1
    fn = np.zeros(3*g.n_nodes); fn[nat.ravel()] = f[dofs.ravel()]
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    levels = build_levels(nel, E, fixed.ravel(), nlev)
    for L in levels: L['lmax'] = 1.05*lam_max(L['A'], L['dinv'])
    A = levels[0]['A']; fn = fn*levels[0]['keep']
    t = time.time(); x1, it1 = pcg(A, fn, lambda r: levels[0]['dinv']*r); t1 = time.time()-t
    t = time.time(); x2, it2 = pcg(A, fn, lambda r: vcycle(levels, 0, r, nu)); t2 = time.time()-t
    print(f"{label}: lmax {[round(L['lmax'],3) for L in levels]} jacobi-pcg {it1} its ({t1:.1f}s)  gmg-pcg(V({nu},{nu}) cheby) {it2} its ({t2:.1f}s)  |dx|/|x| {np.linalg.norm(x1-x2)/np.linalg.norm(x1):.1e}", flush=True)

if __name__ == "__main__" and len(sys.argv) == 1:
    nel = (48, 16, 16)
    n = nel[0]*nel[1]*nel[2]
    run(nel, np.full(n, 0.5), 4, "uniform 0.5")
    rng = np.random.default_rng(1)
    run(nel, 1e-3 + (1-1e-3)*rng.random(n), 4, "random rho")
    # SIMP-like 0/1 design: run the oracle loop for some iterations on this grid and use its density
    g, dofs, f, pres = o.cantilever_problem(nel)
    r = o.simp_optimize(g, dofs, f, pres, o.OptimizationParameters(max_iterations=25, tolerance=0.0), solver="pcg", pcg_rtol=1e-8)
    rho = r["design"]
    print("gray fraction", np.mean((rho > 0.1) & (rho < 0.9)))
    run(nel, rho, 4, "SIMP design after 25 its")
    run(nel, rho, 4, "SIMP design after 25 its, V(3,3)", nu=3)
    run(nel, rho, 3, "SIMP design after 25 its, 3 levels")

def study():
    nel = (96, 32, 32)
    n = nel[0]*nel[1]*nel[2]
    g, dofs, f, pres = o.cantilever_problem(nel)
    # SIMP-like design: cheap surrogate = thresholded smooth random field + solid skin, contrast 1e9
    rng = np.random.default_rng(3)
    fld = rng.standard_normal((32, 32, 96))
    from scipy.ndimage import gaussian_filter
    fld = gaussian_filter(fld, 3.0)
    rho = np.where(fld > np.quantile(fld, 0.5), 1.0, 1e-3).ravel()
    for nlev, cits in [(5, 40), (4, 40), (3, 40), (3, 200), (4, 10)]:
        global COARSE_ITS
        run2(nel, rho, nlev, f"96x32x32 binary design, {nlev} levels, coarse cheb {cits}", cits)

def run2(nel, rho, nlev, label, cits):
    g, dofs, f, pres = o.cantilever_problem(nel)
    nat = native(g)
    fixed = np.zeros((g.n_nodes, 3)); fixed[np.array(sorted(o.select_nodes_by_plane(g, [0,0,0],[1,0,0],1e-6)))] = 1
    fn = np.zeros(3*g.n_nodes); fn[nat.ravel()] = f[dofs.ravel()]
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0)
    levels = build_levels(nel, E, fixed.ravel(), nlev)
    for L in levels: L['lmax'] = 1.05*lam_max(L['A'], L['dinv'])
    A = levels[0]['A']; fn = fn*levels[0]['keep']
    t = time.time(); x2, it2 = pcg(A, fn, lambda r: vcycle(levels, 0, r, 2, cits), maxit=400); t2 = time.time()-t
    print(f"{label}: gmg-pcg {it2} its ({t2:.1f}s) coarsest {levels[-1]['nel']}", flush=True)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "study":
    study()

def study2():
    nel = (64, 32, 16)
    g, dofs, f, pres = o.cantilever_problem(nel)
    t = time.time()
    r = o.simp_optimize(g, dofs, f, pres, o.OptimizationParameters(max_iterations=60, tolerance=0.0, volume_fraction=0.3), solver="pcg", pcg_rtol=1e-7)
    rho = r["design"]
    print("oracle loop", time.time()-t, "s; gray fraction", np.mean((rho > 0.1) & (rho < 0.9)), "min", rho.min(), flush=True)
    np.save("/tmp/rho_64x32x16.npy", rho)
    for nlev in (5, 3):
        run(nel, rho, nlev, f"SIMP design 60 its vf 0.3, {nlev} levels")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "study2":
    study2()
