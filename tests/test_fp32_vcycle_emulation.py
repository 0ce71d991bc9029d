"""CPU run of the FP32 V-cycle's device code (csrc/multigrid32.cu, opt-in, not yet run on a GPU): every kernel it
launches -- the float instantiation of the fused operator (smoother step, residual) and the vector / transfer kernels of
csrc/mg32_kernels.cuh -- is the library's own source compiled for the host (csrc/selftest_kernel_host.cpp).  The
orchestration below restates mg32_smooth / mg32_vcycle / pcg_solve_mg32 step by step in Python, so what is checked is:
  * one V-cycle of the float kernels equals a float64 numpy V-cycle of the same algorithm (assembled sparse operators,
    Kronecker-product transfer operators) to single-precision round-off;
  * FP64 CG preconditioned with it converges in the same number of iterations as with the float64 V-cycle and to the
    sparse direct solution.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import simp_oracle as o
from test_kernel_emulation import SMOOTH, RESID, blocks_of, mode_space

ROOT = Path(__file__).resolve().parents[1]
D, F, U8, I64 = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_longlong)
NU, COARSE_ITS, RATIO = 2, 16, 10.0  # esimp_ctx defaults: mg_nu, mg_coarse_its, mg_ratio


@pytest.fixture(scope="module")
def emu():
    path = ROOT / "easysimp.jl_b200" / "libesimp_kernel_emu.so"
    assert path.exists(), "run __graft_entry__.build() first"
    lib = C.CDLL(str(path))
    lib.emu_fused.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, D, D, D, D, D, D, C.c_double, C.c_double, D, D]
    lib.emu_set_deferred_copies(1)  # the stricter cp.async model (see test_kernel_emulation.py)
    lib.emu_k32_restrict.argtypes = [C.c_int, C.c_int, C.c_int, F, U8, F]
    lib.emu_k32_prolong_add.argtypes = [C.c_int, C.c_int, C.c_int, F, U8, F]
    lib.emu_k32_cheb_first.argtypes = [C.c_longlong, F, F, C.c_float, F, F]
    lib.emu_k32_enter.argtypes = [C.c_longlong, D, C.c_double, F, C.c_float, F, F, F]
    lib.emu_k32_fixup.argtypes = [I64, C.c_longlong, C.c_float, F, F]
    lib.emu_k32_convert.argtypes = [C.c_longlong, D, F]
    lib.emu_k32_dot_rz.argtypes = [C.c_longlong, D, F, C.c_double]
    lib.emu_k32_dot_rz.restype = C.c_double
    lib.emu_k32_update.argtypes = [C.c_longlong, D, D, C.c_double, C.c_double]
    lib.emu_k32_update.restype = C.c_double
    lib.emu_k32_init.argtypes = [C.c_longlong, D, D, D, D]
    lib.emu_k32_direction.argtypes = [C.c_longlong, F, C.c_double, D, D, C.c_double, C.c_double, C.c_double, C.c_int]
    return lib


def pd(a):
    return a.ctypes.data_as(D)


def pf(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(F)


def prolong_1d(nc):
    P = sp.lil_matrix((2 * nc + 1, nc + 1))
    for i in range(nc + 1):
        P[2 * i, i] = 1.0
    for i in range(nc):
        P[2 * i + 1, i] = 0.5
        P[2 * i + 1, i + 1] = 0.5
    return P.tocsr()


class Level:
    """one grid level, single rank; vectors of the emulated path live in the library's device layout"""

    def __init__(self, nel, E, fixed_nodes, k0):
        nx, ny, nz = self.nel = nel
        self.g = o.generate_grid_hex(nel)
        self.nat = 3 * np.arange(self.g.n_nodes)[:, None] + np.arange(3)[None, :]
        self.E = E  # cell id order i + nx (j + ny k)
        K = o.assemble_stiffness(self.g, self.nat, E, np.broadcast_to(k0, (self.g.n_cells, 24, 24)))
        self.keep = 1.0 - fixed_nodes.reshape(-1)  # (n_nodes, 3) flags -> native dof order
        diag = K.diagonal()
        self.m = float(diag.mean())
        Pm = sp.diags(self.keep)
        self.A = (Pm @ K @ Pm + sp.diags((1 - self.keep) * self.m)).tocsr()
        self.dinv = np.where(self.keep > 0, 1.0 / diag, 0.0)  # 0 on prescribed dofs, like k_diag_invert
        x = np.random.default_rng(0).standard_normal(self.A.shape[0]) * self.keep
        for _ in range(30):
            x = self.dinv * (self.A @ x)
            lam = np.linalg.norm(x)
            x /= lam
        self.lmax = 1.1 * lam
        self.nd = 3 * self.g.n_nodes
        # device-layout mirrors (float)
        self.fE = self.cells_dev(E).astype(np.float32)
        self.fdinv = self.to_dev(self.dinv).astype(np.float32)
        mask_nodes = (fixed_nodes.reshape(self.g.n_nodes, 3) * np.array([1, 2, 4])).sum(axis=1).astype(np.uint8)
        self.mask = np.ascontiguousarray(mask_nodes.reshape(nz + 1, ny + 1, nx + 1).transpose(2, 1, 0)).ravel()
        self.fixed = np.ascontiguousarray(np.nonzero(self.to_dev(1.0 - self.keep) > 0)[0].astype(np.int64))
        self.fr, self.fz, self.fz2, self.fd, self.fq = (np.zeros(self.nd, dtype=np.float32) for _ in range(5))

    def to_dev(self, v):
        nx, ny, nz = self.nel
        return np.ascontiguousarray(np.asarray(v).reshape(nz + 1, ny + 1, nx + 1, 3).transpose(2, 3, 1, 0)).ravel()

    def from_dev(self, v):
        nx, ny, nz = self.nel
        return np.ascontiguousarray(np.asarray(v).reshape(nx + 1, 3, ny + 1, nz + 1).transpose(3, 2, 0, 1)).ravel()

    def cells_dev(self, a):
        nx, ny, nz = self.nel
        return np.ascontiguousarray(a.reshape(nz, ny, nx).transpose(2, 1, 0)).ravel()


def build_hierarchy(nel, rho, nlev):
    g = o.generate_grid_hex(nel)
    k0 = o.unit_element_matrices(g, 0.3, 1.0)[0]
    fixed = np.zeros((g.n_nodes, 3))
    fixed[np.array(sorted(o.select_nodes_by_plane(g, [0, 0, 0], [1, 0, 0], 1e-6)))] = 1
    nx, ny, nz = nel
    E = o.simp_modulus(rho, 1.0, 1e-9, 3.0).reshape(nz, ny, nx)
    fix = fixed.reshape(nz + 1, ny + 1, nx + 1, 3)
    levels = []
    for l in range(nlev):
        levels.append(Level((nx, ny, nz), E.ravel().copy(), fix.reshape(-1, 3).copy(), k0))
        if l == nlev - 1:
            break
        Pn = sp.kron(prolong_1d(nz // 2), sp.kron(prolong_1d(ny // 2), prolong_1d(nx // 2)))
        levels[-1].P = sp.kron(Pn, sp.eye(3)).tocsr()
        E = 0.25 * E.reshape(nz // 2, 2, ny // 2, 2, nx // 2, 2).sum(axis=(1, 3, 5))  # k_restrict_E: K0(2h) = 2 K0(h) folded in
        fix = fix[::2, ::2, ::2]
        nx, ny, nz = nx // 2, ny // 2, nz // 2
    return levels, np.ascontiguousarray(blocks_of(mode_space(k0))), k0


def cheb_coeffs(lmax):
    lmin = lmax / RATIO
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    return theta, delta, theta / delta


# ----------------------------------------------------------------------------- float64 numpy statement of the cycle
def ref_smooth(L, r, z, k, zero_guess):
    theta, delta, sigma = cheb_coeffs(L.lmax)
    rho = 1.0 / sigma
    d = np.zeros_like(r)
    done = 0
    if zero_guess:
        d = L.dinv * r / theta
        z = d.copy()
        done = 1
    for i in range(done, k):
        if i == 0:
            c1, c2 = 0.0, 1.0 / theta
        else:
            rho_new = 1.0 / (2.0 * sigma - rho)
            c1, c2, rho = rho_new * rho, 2.0 * rho_new / delta, rho_new
        d = c1 * d + c2 * L.dinv * (r - L.A @ z)
        z = z + d
    return z


def ref_vcycle(levels, l, r):
    L = levels[l]
    if l == len(levels) - 1:
        return ref_smooth(L, r, None, NU if l == 0 else COARSE_ITS, True)
    z = ref_smooth(L, r, None, NU, True)
    res = (r - L.A @ z) * L.keep
    rc = (L.P.T @ res) * levels[l + 1].keep
    z = z + (L.P @ ref_vcycle(levels, l + 1, rc)) * L.keep
    return ref_smooth(L, r, z, NU, False)


# ----------------------------------------------------------------------------- the same with the library's kernels (emulated)
class Emulated:
    def __init__(self, emu, levels, blocks):
        self.emu, self.levels, self.blocks = emu, levels, blocks

    def fused(self, L, mode, z, r, dinv=None, d=None, c1=0.0, c2=0.0):
        nx, ny, nz = L.nel
        y = np.zeros(L.nd)
        dd = None if d is None else d.astype(np.float64)
        st = self.emu.emu_fused(1, mode, nx, ny, nz, 0, pd(self.blocks.ravel()), pd(L.fE.astype(np.float64)), pd(z.astype(np.float64)),
                                pd(r.astype(np.float64)), None if dinv is None else pd(dinv.astype(np.float64)), None if dd is None else pd(dd),
                                c1, c2, pd(y), None)
        assert st == 0
        return y.astype(np.float32), (None if dd is None else dd.astype(np.float32))

    def smooth(self, L, k, zero_guess, first_done=False):  # mg32_smooth
        theta, delta, sigma = cheb_coeffs(L.lmax)
        rho = 1.0 / sigma
        done = 0
        if zero_guess:
            if not first_done:
                self.emu.emu_k32_cheb_first(L.nd, pf(L.fr), pf(L.fdinv), np.float32(1.0 / theta), pf(L.fd), pf(L.fz))
            done = 1
        for i in range(done, k):
            if i == 0:
                c1, c2 = 0.0, 1.0 / theta
            else:
                rho_new = 1.0 / (2.0 * sigma - rho)
                c1, c2, rho = rho_new * rho, 2.0 * rho_new / delta, rho_new
            L.fz2, L.fd = self.fused(L, SMOOTH, L.fz, L.fr, L.fdinv, L.fd, float(np.float32(c1)), float(np.float32(c2)))
            L.fz, L.fz2 = L.fz2, L.fz

    def vcycle(self, l, first_done=False):  # mg32_vcycle
        L = self.levels[l]
        if l == len(self.levels) - 1:
            return self.smooth(L, NU if l == 0 else COARSE_ITS, True, first_done)
        Cl = self.levels[l + 1]
        self.smooth(L, NU, True, first_done)
        L.fq, _ = self.fused(L, RESID, L.fz, L.fr)
        if L.fixed.size:
            self.emu.emu_k32_fixup(L.fixed.ctypes.data_as(I64), L.fixed.size, np.float32(L.m), pf(L.fz), pf(L.fq))
        nx, ny, nz = L.nel
        self.emu.emu_k32_restrict(nx, ny, nz, pf(L.fq), Cl.mask.ctypes.data_as(U8), pf(Cl.fr))
        self.vcycle(l + 1)
        self.emu.emu_k32_prolong_add(nx, ny, nz, pf(Cl.fz), L.mask.ctypes.data_as(U8), pf(L.fz))
        self.smooth(L, NU, False)

    def precondition(self, r_dev, rr):  # mg32_precondition: returns (z32 of level 0, scale)
        L0 = self.levels[0]
        s = float(np.sqrt(rr))
        theta0 = 0.5 * (L0.lmax + L0.lmax / RATIO)
        self.emu.emu_k32_enter(L0.nd, pd(r_dev), 1.0 / s, pf(L0.fdinv), np.float32(1.0 / theta0), pf(L0.fr), pf(L0.fd), pf(L0.fz))
        self.vcycle(0, True)
        return L0.fz, s


def problem(nel=(16, 8, 8), nlev=3, seed=7):
    g, dofs, f, pres = o.cantilever_problem(nel)
    rng = np.random.default_rng(seed)
    rho = 1e-3 + (1 - 1e-3) * rng.random(g.n_cells)
    levels, blocks, k0 = build_hierarchy(nel, rho, nlev)
    nat = levels[0].nat
    fn = np.zeros(3 * g.n_nodes)
    fn[nat.ravel()] = f[dofs.ravel()]
    return levels, blocks, fn * levels[0].keep


def test_one_fp32_vcycle_matches_the_float64_statement(emu):
    levels, blocks, f = problem()
    L0 = levels[0]
    rng = np.random.default_rng(1)
    r = rng.standard_normal(L0.nd) * L0.keep
    z_ref = ref_vcycle(levels, 0, r / np.linalg.norm(r)) * np.linalg.norm(r)
    E = Emulated(emu, levels, blocks)
    z32, s = E.precondition(L0.to_dev(r), float(r @ r))
    z = L0.from_dev(z32.astype(np.float64)) * s
    assert not z[L0.keep == 0].any()
    assert np.linalg.norm(z - z_ref) < 2e-5 * np.linalg.norm(z_ref), np.linalg.norm(z - z_ref) / np.linalg.norm(z_ref)
    # and it is a useful preconditioner: one cycle reduces the error of A z = r substantially
    assert np.linalg.norm(r - L0.A @ z) < 0.7 * np.linalg.norm(r)


def test_fp64_cg_with_the_emulated_fp32_vcycle(emu):
    """pcg_solve_mg32 restated with the emulated kernels: same iteration count as the float64 V-cycle, direct-solve answer"""
    levels, blocks, f = problem()
    L0 = levels[0]
    A = L0.A
    rtol = 1e-8

    def pcg_ref():
        x = np.zeros_like(f)
        r = f.copy()
        z = ref_vcycle(levels, 0, r)
        p = z.copy()
        rz = r @ z
        it = 0
        while r @ r > rtol**2 * (f @ f) and it < 200:
            q = A @ p
            a = rz / (p @ q)
            x += a * p
            r -= a * q
            it += 1
            if r @ r <= rtol**2 * (f @ f):
                break
            z = ref_vcycle(levels, 0, r)
            rzn = r @ z
            p = z + (rzn / rz) * p
            rz = rzn
        return x, it

    x_ref, it_ref = pcg_ref()
    E = Emulated(emu, levels, blocks)
    n = L0.nd
    fdev = L0.to_dev(f)
    Adev = lambda v: L0.to_dev(A @ L0.from_dev(v))  # the FP64 operator of the outer CG (checked in test_kernel_emulation.py)  # noqa: E731
    u, p, r = np.zeros(n), np.zeros(n), np.zeros(n)
    q = Adev(u)
    rr_ff = np.zeros(2)
    emu.emu_k32_init(n, pd(fdev), pd(q), pd(r), pd(rr_ff))
    rr, ff = rr_ff
    tol2 = rtol**2 * ff
    RZ = [0.0, 0.0]
    z32, s = E.precondition(r, rr)
    RZ[0] = emu.emu_k32_dot_rz(n, pd(r), pf(z32), s)
    emu.emu_k32_direction(n, pf(z32), s, pd(p), pd(u), RZ[0], RZ[0], 1.0, 1)
    it = 0
    while rr > tol2 and it < 200:
        cur, nxt = it & 1, (it & 1) ^ 1
        q = Adev(p)
        pq = float(p @ q)
        rr = emu.emu_k32_update(n, pd(q), pd(r), RZ[cur], pq)
        it += 1
        if rr <= tol2:
            emu.emu_k32_direction(n, pf(z32), s, pd(p), pd(u), RZ[cur], RZ[cur], pq, 2)
            break
        z32, s = E.precondition(r, rr)
        RZ[nxt] = emu.emu_k32_dot_rz(n, pd(r), pf(z32), s)
        emu.emu_k32_direction(n, pf(z32), s, pd(p), pd(u), RZ[nxt], RZ[cur], pq, 0)
    x = L0.from_dev(u)
    assert abs(it - it_ref) <= 1, (it, it_ref)
    assert np.linalg.norm(f - A @ x) <= 1.0001 * rtol * np.linalg.norm(f)
    x_direct = o.solve_direct(A.tocsc(), f)
    assert np.linalg.norm(x - x_direct) < 1e-6 * np.linalg.norm(x_direct)
    assert np.linalg.norm(x - x_ref) < 1e-6 * np.linalg.norm(x_ref)
