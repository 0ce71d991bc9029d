"""CPU execution of the z-sweeping TMA / mbarrier operator kernel (csrc/sweep_kernel.cuh, the source nvcc compiles into the
product library): csrc/selftest_kernel_host.cpp runs every thread of a block as a host thread and emulates mbarriers
(pending arrivals, transaction bytes, phase parity), 3-D TMA boxes with out-of-bounds zero fill, warp shuffles and warp
barriers.  Every store mode is held against the oracle's assembled operator on grids that cross y tiles (15 owned node
rows), z segments (32 columns, incl. nz % 32 == 0 where lane 31 takes the last node column along) and x chunks, on the
padded node-row pitch the library uses, on x slabs, and under two TMA models: boxes land at issue, or a DMA thread
delivers them later (a read before the mbarrier wait would then see stale data)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from test_kernel_emulation import DOT, PLAIN, RESID, SMOOTH, SMOOTH_DOT, Problem, ptr, relmax

ROOT = Path(__file__).resolve().parents[1]
D = C.POINTER(C.c_double)


# kernel under test: the two-cells-per-thread sweep kernel (8 consumer warps, tiles of 15 owned node rows) and its one-cell-per-thread
# sibling csrc/sweep1_kernel.cuh with 16 / 12 consumer warps (tiles of 15 / 11 owned node rows)
@pytest.fixture(scope="module", params=[("tma_lands_at_issue", 0), ("tma_lands_later", 0), ("tma_lands_at_issue", 12), ("tma_lands_later", 16)],
                ids=lambda p: f"{p[0]}-{'two_cells' if p[1] == 0 else 'one_cell_%d_warps' % p[1]}")
def emu(request):
    path = ROOT / "easysimp.jl_b200" / "libesimp_kernel_emu.so"
    assert path.exists(), "run __graft_entry__.build() first"
    lib = C.CDLL(str(path))
    lib.emu_sweep_set_variant.argtypes = [C.c_int]
    lib.emu_sweep_set_variant(request.param[1])
    lib.emu_sweep.argtypes = [C.c_int] * 14 + [D] * 6 + [C.c_double, C.c_double, D, D]
    lib.emu_set_smoother_direction.argtypes = [C.c_int, C.c_int, D]
    lib.emu_sweep_set_ik.argtypes = [C.c_double] * 3
    lib.emu_sweep_set_ik(*IK)
    lib.emu_set_deferred_tma(int(request.param[0] == "tma_lands_later"))
    return lib


IK = (1.3, 0.7, 2.1)  # stands for 1 / K0[3a+c, 3a+c]


def node_dinv(p, seed=11):
    """D^-1 the way the sweep kernel's smoother modes take it: ONE double per node (s, think 1 / sum of E around the node) with the
    prescribed-component mask in its three lowest mantissa bits; Dinv[node, c] = 0 if bit c else IK[c] * s.  Returns the dof vector
    (caller numbering) for the expected values and the node array in the padded device layout for the kernel."""
    nx, ny, nz = p.nel
    rng = np.random.default_rng(seed)
    shape = (nx + 1, ny + 1, nz + 1)
    sval = 1.0 / (0.5 + rng.random(shape))
    mask = np.where(rng.random(shape) < 0.15, rng.integers(1, 8, shape), 0).astype(np.int64)
    bits = (sval.view(np.int64) & ~np.int64(7)) | mask
    clean = (bits & ~np.int64(7)).view(np.float64)
    dinv_dev = np.stack([np.where((mask >> c) & 1, 0.0, clean * IK[c]) for c in range(3)], axis=1)  # [plane][c][y][z]
    sinv = np.zeros((nx + 1, ny + 1, pitch(nz)))
    sinv[..., :nz + 1] = bits.view(np.float64)
    return p.from_dev(np.ascontiguousarray(dinv_dev).ravel()), sinv.ravel()


def pitch(nz):
    return (nz + 1 + 7) & ~7


def pad_rows(v_dev, nel, npx=None):
    """[plane][c][y][z] with pitch nz + 1 -> the same with the library's padded pitch (pads = 0)"""
    nx, ny, nz = nel
    npx = nx + 1 if npx is None else npx
    a = v_dev.reshape(npx, 3, ny + 1, nz + 1)
    out = np.zeros((npx, 3, ny + 1, pitch(nz)))
    out[..., :nz + 1] = a
    return out.ravel()


def unpad_rows(v_pad, nel, npx=None):
    nx, ny, nz = nel
    npx = nx + 1 if npx is None else npx
    a = v_pad.reshape(npx, 3, ny + 1, pitch(nz))
    return np.ascontiguousarray(a[..., :nz + 1]).ravel(), a[..., nz + 1:]


def run(emu, p, mode, u, chunk=0, r=None, sinv=None, d=None, c1=0.0, c2=0.0):
    nx, ny, nz = p.nel
    y = np.full(3 * (nx + 1) * (ny + 1) * pitch(nz), np.nan)
    dot = np.zeros(1)
    dv = [None if a is None else pad_rows(p.to_dev(a), p.nel) for a in (r, None, d)]
    dv[1] = sinv  # node scalar array, already in the device layout (node_dinv)
    st = emu.emu_sweep(mode, nx, ny, nz, pitch(nz), 0, nx + 1, 0, nx, 0, nx + 1, 0, nx, chunk, ptr(p.B.ravel()), ptr(p.cells_dev(p.E)),
                       ptr(pad_rows(p.to_dev(u), p.nel)), ptr(dv[0]), ptr(dv[1]), ptr(dv[2]), c1, c2, ptr(y), ptr(dot))
    assert st == 0
    yv, ypad = unpad_rows(y, p.nel)
    assert np.isnan(ypad).all()  # the pad entries of a node row are never written
    dn = None if dv[2] is None else p.from_dev(unpad_rows(dv[2], p.nel)[0])
    return p.from_dev(yv), dn, float(dot[0])


# y tiles own 15 node rows, z segments hold 32 columns: (5, 18, 40) needs 2 tiles, 2 segments; (3, 33, 64) 3 tiles and the
# extra-column path (nz % 32 == 0); (4, 6, 32) one segment with the extra column; (13, 7, 5) a single partial segment
GRIDS = [((13, 7, 5), (13.0, 3.5, 1.25), 0), ((5, 18, 40), (5.0, 18.0, 40.0), 0), ((5, 18, 40), (5.0, 9.0, 10.0), 2),
         ((3, 33, 64), (3.0, 33.0, 64.0), 0), ((4, 6, 32), (4.0, 6.0, 32.0), 3), ((1, 1, 1), (1.0, 1.0, 1.0), 0)]


@pytest.mark.parametrize("nel,right,chunk", GRIDS)
def test_plain_and_dot_modes_match_the_assembled_operator(emu, nel, right, chunk):
    p = Problem(nel, right)
    u = p.rng.standard_normal(p.nd)
    want = p.A(u)
    y, _, _ = run(emu, p, PLAIN, u, chunk=chunk)
    assert not np.isnan(y).any() and relmax(y, want) < 2e-13
    y2, _, dot = run(emu, p, DOT, u, chunk=chunk)
    assert np.array_equal(y, y2) and abs(dot - u @ want) < 1e-11 * abs(u @ want)


@pytest.mark.parametrize("nel,right,chunk", GRIDS[:4])
def test_smoother_and_residual_modes(emu, nel, right, chunk):
    p = Problem(nel, right, seed=1)
    z, r, d = (p.rng.standard_normal(p.nd) for _ in range(3))
    dinv, sinv = node_dinv(p)
    Az = p.A(z)
    c1, c2 = 0.37, 0.81
    v = c1 * d + c2 * dinv * (r - Az)
    y, dnew, _ = run(emu, p, SMOOTH, z, chunk=chunk, r=r, sinv=sinv, d=d, c1=c1, c2=c2)
    assert relmax(y, z + v) < 1e-12 and relmax(dnew, v) < 1e-12
    y2, dnew2, dot = run(emu, p, SMOOTH_DOT, z, chunk=chunk, r=r, sinv=sinv, d=d, c1=c1, c2=c2)
    assert np.array_equal(y2, y) and abs(dot - r @ (z + v)) < 1e-10 * np.abs(r * (z + v)).sum()
    v0 = c2 * dinv * (r - Az)  # c1 = 0 restarts the recurrence: d is not read at all (NaN in, finite out)
    y, dnew, _ = run(emu, p, SMOOTH, z, chunk=chunk, r=r, sinv=sinv, d=np.full(p.nd, np.nan), c1=0.0, c2=c2)
    assert relmax(y, z + v0) < 1e-12 and relmax(dnew, v0) < 1e-12
    y, _, _ = run(emu, p, RESID, z, chunk=chunk, r=r)
    assert relmax(y, r - Az) < 1e-12


@pytest.mark.parametrize("nranks,halo_e,chunk", [(2, 1, 0), (3, 2, 2)])
def test_x_slab_partition_reproduces_the_global_operator(emu, nranks, halo_e, chunk):
    import easysimp_b200.partition as part

    p = Problem((9, 17, 34), (9.0, 17.0, 34.0), seed=4)
    nx, ny, nz = p.nel
    u = p.rng.standard_normal(p.nd)
    want = p.to_dev(p.A(u)).reshape(nx + 1, -1)
    udev, Edev = p.to_dev(u).reshape(nx + 1, -1), p.cells_dev(p.E).reshape(nx, -1)
    covered = np.zeros(nx + 1, dtype=int)
    for r in range(nranks):
        sl = part.slab_partition(nx, r, nranks, halo_e)
        u_loc = pad_rows(np.ascontiguousarray(udev[sl.gp0:sl.gp0 + sl.npx]).ravel(), p.nel, sl.npx)
        E_loc = np.ascontiguousarray(Edev[sl.ge0:sl.ge0 + sl.nex]).ravel()
        y_loc = np.full(u_loc.size, np.nan)
        assert emu.emu_sweep(PLAIN, nx, ny, nz, pitch(nz), sl.gp0, sl.npx, sl.ge0, sl.nex, sl.own_p0, sl.own_p1, sl.mv_e0, sl.mv_e1, chunk,
                             ptr(p.B.ravel()), ptr(E_loc), ptr(u_loc), None, None, None, 0.0, 0.0, ptr(y_loc), None) == 0
        y_loc = unpad_rows(y_loc, p.nel, sl.npx)[0].reshape(sl.npx, -1)
        own = slice(sl.own_p0, sl.own_p1)
        assert relmax(y_loc[own], want[sl.gp0 + sl.own_p0:sl.gp0 + sl.own_p1]) < 2e-13
        ghost = np.ones(sl.npx, dtype=bool)
        ghost[own] = False
        assert np.isnan(y_loc[ghost]).all()
        covered[sl.gp0 + sl.own_p0:sl.gp0 + sl.own_p1] += 1
    assert (covered == 1).all()


@pytest.mark.parametrize("nel,right,chunk", GRIDS[1:3])
def test_direction_free_smoother_steps(emu, nel, right, chunk):
    """the Chebyshev direction need not be a stored vector: right after the zero-guess first step it is the iterate itself
    (FUSE_D_IS_Z = 2), later it is z - z_prev (FUSE_D_FROM_ZPREV = 3); the d vector is neither read (NaN in) nor written"""
    p = Problem(nel, right, seed=5)
    r = p.rng.standard_normal(p.nd)
    dinv, sinv = node_dinv(p)
    c1, c2 = 0.37, 0.81
    z1 = 0.3 * dinv * r                               # first step from a zero guess: d1 = z1
    z2_want = z1 + c1 * z1 + c2 * dinv * (r - p.A(z1))
    emu.emu_set_smoother_direction(2, 0, None)
    try:
        y, dnew, _ = run(emu, p, SMOOTH, z1, chunk=chunk, r=r, sinv=sinv, d=np.full(p.nd, np.nan), c1=c1, c2=c2)
        assert relmax(y, z2_want) < 1e-12 and np.isnan(dnew).all()
        # next step: d2 = z2 - z1, z_prev handed in the device layout of the kernel under test
        z3_want = z2_want + c1 * (z2_want - z1) + c2 * dinv * (r - p.A(z2_want))
        zprev = pad_rows(p.to_dev(z1), p.nel)
        emu.emu_set_smoother_direction(3, 0, ptr(zprev))
        z1, keep = z2_want, zprev  # noqa: F841  (keep the buffer alive during the call)
        y, dnew, dot = run(emu, p, SMOOTH_DOT, z1, chunk=chunk, r=r, sinv=sinv, d=np.full(p.nd, np.nan), c1=c1, c2=c2)
        assert relmax(y, z3_want) < 1e-12 and np.isnan(dnew).all() and abs(dot - r @ z3_want) < 1e-10 * np.abs(r * z3_want).sum()
    finally:
        emu.emu_set_smoother_direction(1, 1, None)
