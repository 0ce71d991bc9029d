"""CPU execution of the fused hex8 operator kernel: csrc/fused_kernel.cuh -- the source nvcc compiles into the product
library -- is compiled for the host by csrc/selftest_kernel_host.cpp (block threads = host threads, __syncthreads = a
barrier, cp.async = immediate copies) and every store mode is held against the oracle's assembled operator.
Covers what the -m gpu tests cover for this kernel (tiling across tile boundaries, x chunks, ring indexing, butterflies,
the smoother / residual / dot store modes) plus the single-precision instantiation the FP32 V-cycle uses."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import simp_oracle as o

ROOT = Path(__file__).resolve().parents[1]
D = C.POINTER(C.c_double)
PLAIN, DOT, SMOOTH, RESID, SMOOTH_DOT = range(5)


@pytest.fixture(scope="module", params=["copies_land_at_issue", "copies_land_at_wait_group"])
def emu(request):
    """both extremes of cp.async completion: at once, or only when cp.async.wait_group retires the group (a kernel that
    is right under both never reads an asynchronously copied datum before its wait + barrier)"""
    path = ROOT / "easysimp.jl_b200" / "libesimp_kernel_emu.so"
    assert path.exists(), "run __graft_entry__.build() first"
    lib = C.CDLL(str(path))
    lib.emu_fused.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, D, D, D, D, D, D, C.c_double, C.c_double, D, D]
    lib.emu_pat.argtypes = [C.c_int] * 3
    lib.emu_set_deferred_copies(int(request.param == "copies_land_at_wait_group"))
    lib.emu_fused_tile.argtypes = [C.c_int] * 6 + [D, D, D, D]
    lib.emu_fused_slab.argtypes = [C.c_int] * 12 + [D, D, D, D]
    lib.emu_set_smoother_direction.argtypes = [C.c_int, C.c_int, D]
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(D)


HEXREF = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)


def mode_space(k0):
    """K~ = T' K0 T / 64 with T[3a+c][3m+c] = sx^bx sy^by sz^bz, m = bx + 2 by + 4 bz (hex8_upload_constants)"""
    H = np.array([[(HEXREF[a, 0] if m & 1 else 1.0) * (HEXREF[a, 1] if m & 2 else 1.0) * (HEXREF[a, 2] if m & 4 else 1.0)
                   for m in range(8)] for a in range(8)])
    T = np.kron(H, np.eye(3))
    return T.T @ k0 @ T / 64.0


def blocks_of(kt):
    B = np.zeros((8, 3, 3))
    for q in range(8):
        for c in range(3):
            for d in range(3):
                B[q, c, d] = kt[3 * (q ^ (1 << c)) + c, 3 * (q ^ (1 << d)) + d]
    return B


class Problem:
    def __init__(self, nel, right, seed=0):
        self.nel = nel
        self.g = o.generate_grid_hex(nel, (0, 0, 0), right)
        nx, ny, nz = nel
        self.nat = 3 * np.arange(self.g.n_nodes)[:, None] + np.arange(3)[None, :]
        rng = np.random.default_rng(seed)
        self.rho = 1e-3 + (1 - 1e-3) * rng.random(self.g.n_cells)
        self.E = o.simp_modulus(self.rho, 1.0, 1e-9, 3.0)
        self.k0 = o.unit_element_matrices(self.g, 0.3, 1.0)
        self.B = np.ascontiguousarray(blocks_of(mode_space(self.k0[0])))
        self.nd = 3 * self.g.n_nodes
        self.rng = rng
        self.none = np.zeros(0, dtype=np.int64)

    def to_dev(self, v):  # native node-major (3 node + c) -> [plane][component][y][z]
        nx, ny, nz = self.nel
        return np.ascontiguousarray(v.reshape(nz + 1, ny + 1, nx + 1, 3).transpose(2, 3, 1, 0)).ravel()

    def from_dev(self, v):
        nx, ny, nz = self.nel
        return np.ascontiguousarray(v.reshape(nx + 1, 3, ny + 1, nz + 1).transpose(3, 2, 0, 1)).ravel()

    def cells_dev(self, a):  # cell id = i + nx (j + ny k) -> [layer][y][z]
        nx, ny, nz = self.nel
        return np.ascontiguousarray(a.reshape(nz, ny, nx).transpose(2, 1, 0)).ravel()

    def A(self, u):
        return o.operator_apply(self.g, self.nat, self.E, self.k0, self.none, u, 1.0)

    def run(self, emu, mode, u, is_float=False, chunk=0, r=None, dinv=None, d=None, c1=0.0, c2=0.0):
        nx, ny, nz = self.nel
        y = np.full(self.nd, np.nan)
        dot = np.zeros(1)
        dv = [None if a is None else self.to_dev(a) for a in (r, dinv, d)]
        st = emu.emu_fused(int(is_float), mode, nx, ny, nz, chunk, ptr(self.B.ravel()), ptr(self.cells_dev(self.E)), ptr(self.to_dev(u)),
                           ptr(dv[0]), ptr(dv[1]), ptr(dv[2]), c1, c2, ptr(y), ptr(dot))
        assert st == 0
        return self.from_dev(y), (None if dv[2] is None else self.from_dev(dv[2])), float(dot[0])


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_parity_pattern_of_the_mode_space_matrix(emu):
    """the structural zeros hard-wired in pat() hold for any cuboid and Poisson ratio; 45 non-zeros remain"""
    assert sum(emu.emu_pat(q, c, d) for q in range(8) for c in range(3) for d in range(3)) == 45
    for right, nu in (((1.0, 1.0, 1.0), 0.3), ((1.0, 0.5, 0.25), 0.3), ((0.05, 0.07, 0.03), 0.45), ((2.0, 1.0, 3.0), 0.0)):
        g = o.generate_grid_hex((1, 1, 1), (0, 0, 0), right)
        kt = mode_space(o.unit_element_matrices(g, nu, 1.0)[0])
        kmax = np.abs(kt).max()
        for i in range(24):
            for j in range(24):
                m1, c, m2, d = i // 3, i % 3, j // 3, j % 3
                q1, q2 = m1 ^ (1 << c), m2 ^ (1 << d)
                if not (q1 == q2 and emu.emu_pat(q1, c, d)):
                    assert abs(kt[i, j]) < 1e-12 * kmax, (right, nu, i, j)


GRIDS = [((13, 7, 5), (13.0, 3.5, 1.25), 0), ((5, 18, 17), (5.0, 18.0, 17.0), 0), ((13, 7, 5), (13.0, 7.0, 5.0), 4),
         ((1, 1, 1), (1.0, 1.0, 1.0), 0), ((9, 16, 15), (0.9, 1.6, 1.5), 3)]


@pytest.mark.parametrize("nel,right,chunk", GRIDS)
def test_plain_and_dot_modes_match_the_assembled_operator(emu, nel, right, chunk):
    """y = K(E) u of the oracle's sparse assembly; tiles across (y,z) tile boundaries, x cut into chunks"""
    p = Problem(nel, right)
    u = p.rng.standard_normal(p.nd)
    want = p.A(u)
    y, _, _ = p.run(emu, PLAIN, u, chunk=chunk)
    assert not np.isnan(y).any() and relmax(y, want) < 2e-13
    y2, _, dot = p.run(emu, DOT, u, chunk=chunk)
    assert np.array_equal(y, y2) and abs(dot - u @ want) < 1e-11 * abs(u @ want)


@pytest.mark.parametrize("ty,tz", [(8, 16), (8, 32), (16, 16)])
def test_other_tile_shapes(emu, ty, tz):
    """the 8x16 and 8x32 owner-computes tiles (ESIMP_MATVEC_FUSED_8x16 / _8x32) on a grid that needs several tiles each way"""
    p = Problem((4, 17, 33), (4.0, 17.0, 33.0), seed=3)
    u = p.rng.standard_normal(p.nd)
    y = np.full(p.nd, np.nan)
    assert emu.emu_fused_tile(ty, tz, 4, 17, 33, 0, ptr(p.B.ravel()), ptr(p.cells_dev(p.E)), ptr(p.to_dev(u)), ptr(y)) == 0
    assert relmax(p.from_dev(y), p.A(u)) < 2e-13


@pytest.mark.parametrize("nel,right,chunk", GRIDS[:3])
def test_smoother_and_residual_modes(emu, nel, right, chunk):
    """z' = z + v, d' = v, v = c1 d + c2 Dinv (r - A z); residual r - A z; <r, z'> of the fused-dot variant"""
    p = Problem(nel, right, seed=1)
    z, r, d = (p.rng.standard_normal(p.nd) for _ in range(3))
    dinv = 1.0 / (0.5 + p.rng.random(p.nd))
    Az = p.A(z)
    c1, c2 = 0.37, 0.81
    v = c1 * d + c2 * dinv * (r - Az)
    y, dnew, _ = p.run(emu, SMOOTH, z, chunk=chunk, r=r, dinv=dinv, d=d, c1=c1, c2=c2)
    assert relmax(y, z + v) < 1e-12 and relmax(dnew, v) < 1e-12
    y2, dnew2, dot = p.run(emu, SMOOTH_DOT, z, chunk=chunk, r=r, dinv=dinv, d=d, c1=c1, c2=c2)
    assert np.array_equal(y2, y) and abs(dot - r @ (z + v)) < 1e-10 * np.abs(r * (z + v)).sum()
    # c1 = 0 restarts the recurrence: d is not read at all (NaN in, finite out)
    v0 = c2 * dinv * (r - Az)
    y, dnew, _ = p.run(emu, SMOOTH, z, chunk=chunk, r=r, dinv=dinv, d=np.full(p.nd, np.nan), c1=0.0, c2=c2)
    assert relmax(y, z + v0) < 1e-12 and relmax(dnew, v0) < 1e-12
    y, _, _ = p.run(emu, RESID, z, chunk=chunk, r=r)
    assert relmax(y, r - Az) < 1e-12


@pytest.mark.parametrize("nel,right,chunk", GRIDS[:2])
def test_single_precision_instantiation(emu, nel, right, chunk):
    """the float kernel of the FP32 V-cycle (csrc/multigrid32.cu): same modes, float32 round-off"""
    p = Problem(nel, right, seed=2)
    z, r, d = (p.rng.standard_normal(p.nd) for _ in range(3))
    dinv = 1.0 / (0.5 + p.rng.random(p.nd))
    Az = p.A(z)
    y, _, _ = p.run(emu, PLAIN, z, is_float=True, chunk=chunk)
    assert relmax(y, Az) < 2e-5
    c1, c2 = 0.37, 0.81
    v = c1 * d + c2 * dinv * (r - Az)
    y, dnew, _ = p.run(emu, SMOOTH, z, is_float=True, chunk=chunk, r=r, dinv=dinv, d=d, c1=c1, c2=c2)
    assert relmax(y, z + v) < 2e-5 and relmax(dnew, v) < 2e-5
    y, _, _ = p.run(emu, RESID, z, is_float=True, chunk=chunk, r=r)
    assert relmax(y, r - Az) < 2e-5


@pytest.mark.parametrize("nranks,halo_e,chunk", [(2, 1, 0), (3, 2, 0), (4, 1, 2)])
def test_x_slab_partition_reproduces_the_global_operator(emu, nranks, halo_e, chunk):
    """every rank runs the kernel on its slab (owned planes + one ghost plane per interior side, the redundant left
    interface layer, `halo_e` ghost cell layers: easysimp_b200/partition.py = esimp_set_grid_structured); the owned
    planes of all ranks together are the global y, and nothing outside them is written"""
    import easysimp_b200.partition as part

    p = Problem((13, 6, 5), (13.0, 6.0, 5.0), seed=4)
    nx, ny, nz = p.nel
    u = p.rng.standard_normal(p.nd)
    want = p.to_dev(p.A(u)).reshape(nx + 1, -1)
    udev, Edev = p.to_dev(u).reshape(nx + 1, -1), p.cells_dev(p.E).reshape(nx, -1)
    covered = np.zeros(nx + 1, dtype=int)
    for r in range(nranks):
        sl = part.slab_partition(nx, r, nranks, halo_e)
        u_loc = np.ascontiguousarray(udev[sl.gp0:sl.gp0 + sl.npx]).ravel()
        E_loc = np.ascontiguousarray(Edev[sl.ge0:sl.ge0 + sl.nex]).ravel()
        y_loc = np.full(u_loc.size, np.nan)
        assert emu.emu_fused_slab(nx, ny, nz, sl.gp0, sl.npx, sl.ge0, sl.nex, sl.own_p0, sl.own_p1, sl.mv_e0, sl.mv_e1, chunk,
                                  ptr(p.B.ravel()), ptr(E_loc), ptr(u_loc), ptr(y_loc)) == 0
        y_loc = y_loc.reshape(sl.npx, -1)
        own = slice(sl.own_p0, sl.own_p1)
        assert relmax(y_loc[own], want[sl.gp0 + sl.own_p0:sl.gp0 + sl.own_p1]) < 2e-13
        ghost = np.ones(sl.npx, dtype=bool)
        ghost[own] = False
        assert np.isnan(y_loc[ghost]).all()  # ghost planes are never stored: only u moves between ranks
        covered[sl.gp0 + sl.own_p0:sl.gp0 + sl.own_p1] += 1
    assert (covered == 1).all()


@pytest.mark.parametrize("nel,right,chunk", GRIDS[1:3])
def test_direction_free_smoother_steps(emu, nel, right, chunk):
    """the Chebyshev direction need not be a stored vector: right after the zero-guess first step it is the iterate itself
    (FUSE_D_IS_Z = 2), later it is z - z_prev (FUSE_D_FROM_ZPREV = 3); the d vector is neither read (NaN in) nor written"""
    p = Problem(nel, right, seed=5)
    r = p.rng.standard_normal(p.nd)
    dinv = 1.0 / (0.5 + p.rng.random(p.nd))
    c1, c2 = 0.37, 0.81
    z1 = 0.3 * dinv * r                               # first step from a zero guess: d1 = z1
    z2_want = z1 + c1 * z1 + c2 * dinv * (r - p.A(z1))
    emu.emu_set_smoother_direction(2, 0, None)
    try:
        y, dnew, _ = p.run(emu, SMOOTH, z1, chunk=chunk, r=r, dinv=dinv, d=None if False else np.full(p.nd, np.nan), c1=c1, c2=c2)
        assert relmax(y, z2_want) < 1e-12 and np.isnan(dnew).all()
        # next step: d2 = z2 - z1, z_prev handed in the device layout of the kernel under test
        z3_want = z2_want + c1 * (z2_want - z1) + c2 * dinv * (r - p.A(z2_want))
        zprev = p.to_dev(z1)
        emu.emu_set_smoother_direction(3, 0, ptr(zprev))
        z1, keep = z2_want, zprev  # noqa: F841  (keep the buffer alive during the call)
        y, dnew, dot = p.run(emu, SMOOTH_DOT, z1, chunk=chunk, r=r, dinv=dinv, d=None if False else np.full(p.nd, np.nan), c1=c1, c2=c2)
        assert relmax(y, z3_want) < 1e-12 and np.isnan(dnew).all() and abs(dot - r @ z3_want) < 1e-10 * np.abs(r * z3_want).sum()
    finally:
        emu.emu_set_smoother_direction(1, 1, None)
