"""ctypes binding of libeasysimp_b200.so (include/easysimp_b200.h).

This is the same binding a Julia host makes with `ccall` (see INTEGRATION.md and
easysimp.jl_b200/julia/EasySIMPB200.jl); nothing here computes.  There is NO CPU fallback: if
the shared library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("ESIMP_LIB", PKG_DIR / "libeasysimp_b200.so"))
HEADER_PATH = PKG_DIR.parent / "include" / "easysimp_b200.h"

ESIMP_FILTER_SENSITIVITY, ESIMP_FILTER_DENSITY = 0, 1
ESIMP_CELL_TET4, ESIMP_CELL_HEX8 = 4, 8
ESIMP_MATVEC_AUTO, ESIMP_MATVEC_DENSE_ATOMIC, ESIMP_MATVEC_FUSED, ESIMP_MATVEC_FUSED_ATOMIC, ESIMP_MATVEC_FUSED_8x32, ESIMP_MATVEC_FUSED_8x16, ESIMP_MATVEC_SWEEP = range(7)
ESIMP_WARN_PCG_MAXIT, ESIMP_WARN_OC_MAXIT = 1, 2


class EsimpParams(C.Structure):
    _fields_ = [
        ("E0", C.c_double), ("Emin", C.c_double), ("nu", C.c_double), ("p", C.c_double),
        ("volume_fraction", C.c_double), ("move_limit", C.c_double), ("damping", C.c_double),
        ("filter_radius", C.c_double), ("filter_type", C.c_int32), ("pcg_maxit", C.c_int32),
        ("pcg_rtol", C.c_double), ("pcg_check_every", C.c_int32), ("matvec_variant", C.c_int32),
        ("warm_start", C.c_int32), ("precond", C.c_int32),
    ]


class EsimpIterStats(C.Structure):
    _fields_ = [
        ("energy", C.c_double), ("volume", C.c_double), ("change", C.c_double), ("lam", C.c_double),
        ("grayness", C.c_double), ("max_u", C.c_double), ("pcg_relres", C.c_double),
        ("pcg_iters", C.c_int32), ("oc_steps", C.c_int32), ("status", C.c_int32), ("reserved", C.c_int32),
        ("ms_filter", C.c_double), ("ms_solve", C.c_double), ("ms_sens", C.c_double), ("ms_oc", C.c_double),
        ("ms_total", C.c_double),
    ]


class EsimpError(RuntimeError):
    pass


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_SIGNATURES = {
    "esimp_version": (C.c_int, []),
    "esimp_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "esimp_destroy": (C.c_int, [_P]),
    "esimp_last_error": (C.c_char_p, [_P]),
    "esimp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "esimp_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "esimp_set_grid_structured": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _I64]),
    "esimp_set_mesh_unstructured": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _D, _I64, _I64]),
    "esimp_default_params": (C.c_int, [C.POINTER(EsimpParams)]),
    "esimp_set_params": (C.c_int, [_P, C.POINTER(EsimpParams)]),
    "esimp_set_dirichlet": (C.c_int, [_P, _I64, C.c_int64]),
    "esimp_set_load": (C.c_int, [_P, _D]),
    "esimp_filter_build_structured": (C.c_int, [_P]),
    "esimp_filter_build_csr": (C.c_int, [_P, _I64, _I64, _D, _D, C.c_double]),
    "esimp_set_density": (C.c_int, [_P, _D]),
    "esimp_matvec": (C.c_int, [_P, _D, _D]),
    "esimp_diagonal": (C.c_int, [_P, _D]),
    "esimp_solve": (C.c_int, [_P, _D, C.POINTER(C.c_int), _D]),
    "esimp_sensitivities": (C.c_int, [_P, _D, _D]),
    "esimp_filter_sens": (C.c_int, [_P, _D, _D, _D]),
    "esimp_filter_dens": (C.c_int, [_P, _D, _D]),
    "esimp_filter_dens_T": (C.c_int, [_P, _D, _D]),
    "esimp_oc": (C.c_int, [_P, _D, _D, _D, _D, _D, C.POINTER(C.c_int)]),
    "esimp_init_design": (C.c_int, [_P, _D]),
    "esimp_iterate": (C.c_int, [_P, C.POINTER(EsimpIterStats)]),
    "esimp_final_analysis": (C.c_int, [_P, _D, _D, C.POINTER(C.c_int)]),
    "esimp_get_design": (C.c_int, [_P, _D]),
    "esimp_get_physical_density": (C.c_int, [_P, _D]),
    "esimp_get_displacement": (C.c_int, [_P, _D]),
    "esimp_set_displacement": (C.c_int, [_P, _D]),
    "esimp_n_dofs": (C.c_int64, [_P]),
    "esimp_n_cells": (C.c_int64, [_P]),
    "esimp_bench_matvec": (C.c_int, [_P, C.c_int, _D, _D]),
    "esimp_bench_operator_mode": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "esimp_profile_kernels": (C.c_int, [_P, C.c_int, _D, C.POINTER(C.c_int64)]),
    "esimp_launch_count": (C.c_int64, [_P]),
    "esimp_set_design": (C.c_int, [_P, _D]),
    "esimp_profile_enable": (C.c_int, [_P, C.c_int]),
    "esimp_profile_matvec": (C.c_int, [_P, _D, C.POINTER(C.c_int64), _D]),
    "esimp_profile_operator_all": (C.c_int, [_P, _D, C.POINTER(C.c_int64)]),
    "esimp_set_body_force": (C.c_int, [_P, _D, C.c_double]),
    "esimp_body_force": (C.c_int, [_P, _D, _D, _D]),
    "esimp_stress_fields": (C.c_int, [_P, _D, _D, _D, _D, _I64]),
    "esimp_stress_tensors": (C.c_int, [_P, C.c_int64, C.c_int64, _D]),
    "esimp_stress_nq": (C.c_int, [_P]),
    "esimp_multi_create": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_int), C.c_int]),
    "esimp_multi_destroy": (C.c_int, [_P]),
    "esimp_multi_last_error": (C.c_char_p, [_P]),
    "esimp_multi_n_gpus": (C.c_int, [_P]),
    "esimp_multi_rank_ctx": (_P, [_P, C.c_int]),
    "esimp_multi_n_dofs": (C.c_int64, [_P]),
    "esimp_multi_n_cells": (C.c_int64, [_P]),
    "esimp_multi_set_params": (C.c_int, [_P, C.POINTER(EsimpParams)]),
    "esimp_multi_set_grid_structured": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _I64]),
    "esimp_multi_set_dirichlet": (C.c_int, [_P, _I64, C.c_int64]),
    "esimp_multi_set_load": (C.c_int, [_P, _D]),
    "esimp_multi_set_body_force": (C.c_int, [_P, _D, C.c_double]),
    "esimp_multi_filter_build_structured": (C.c_int, [_P]),
    "esimp_multi_init_design": (C.c_int, [_P, _D]),
    "esimp_multi_set_design": (C.c_int, [_P, _D]),
    "esimp_multi_iterate": (C.c_int, [_P, C.POINTER(EsimpIterStats)]),
    "esimp_multi_final_analysis": (C.c_int, [_P, _D, _D, C.POINTER(C.c_int)]),
    "esimp_multi_get_design": (C.c_int, [_P, _D]),
    "esimp_multi_get_physical_density": (C.c_int, [_P, _D]),
    "esimp_multi_get_displacement": (C.c_int, [_P, _D]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the shared library and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise EsimpError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_D)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_I64)


class Context:
    """Thin RAII wrapper around esimp_ctx*; every method is one C-ABI call with host buffers."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        self._h = _P()
        st = self.lib.esimp_create(C.byref(self._h), device)
        if st < 0:
            raise EsimpError(f"esimp_create failed ({st}): {self.lib.esimp_last_error(None).decode()}")
        self.params = EsimpParams()
        self.lib.esimp_default_params(C.byref(self.params))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.esimp_destroy(self._h)
            self._h = None

    __del__ = close

    def _chk(self, st, what):
        if st < 0:
            raise EsimpError(f"{what} failed ({st}): {self.lib.esimp_last_error(self._h).decode()}")
        return st

    # -- setup
    def comm_init(self, rank, nranks, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._chk(self.lib.esimp_comm_init(self._h, rank, nranks, buf), "esimp_comm_init")

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise AttributeError(k)
            setattr(self.params, k, v)
        self._chk(self.lib.esimp_set_params(self._h, C.byref(self.params)), "esimp_set_params")

    def set_grid_structured(self, nx, ny, nz, hx, hy, hz, dof_of_node=None):
        ptr = None
        if dof_of_node is not None:
            keep, ptr = _i64(np.asarray(dof_of_node).reshape(-1))
        self._chk(self.lib.esimp_set_grid_structured(self._h, nx, ny, nz, hx, hy, hz, ptr), "esimp_set_grid_structured")

    def set_mesh_unstructured(self, celltype, xyz, conn, celldofs):
        k1, p1 = _f64(np.asarray(xyz).reshape(-1))
        k2, p2 = _i64(np.asarray(conn).reshape(-1))
        k3, p3 = _i64(np.asarray(celldofs).reshape(-1))
        n_cells = k2.size // celltype
        self._chk(self.lib.esimp_set_mesh_unstructured(self._h, celltype, n_cells, k1.size // 3, p1, p2, p3),
                  "esimp_set_mesh_unstructured")

    def set_dirichlet(self, dofs):
        k, p = _i64(np.asarray(dofs).reshape(-1))
        self._chk(self.lib.esimp_set_dirichlet(self._h, p, k.size), "esimp_set_dirichlet")

    def set_load(self, f):
        k, p = _f64(f)
        assert k.size == self.n_dofs
        self._chk(self.lib.esimp_set_load(self._h, p), "esimp_set_load")

    def set_body_force(self, accel=None, base_density=0.0):
        """acceleration_data of simp_optimize (Optimization.jl:300-311); accel=None switches it off."""
        if accel is None:
            self._chk(self.lib.esimp_set_body_force(self._h, None, 0.0), "esimp_set_body_force")
            return
        k, p = _f64(np.asarray(accel, dtype=np.float64).reshape(3))
        self._chk(self.lib.esimp_set_body_force(self._h, p, float(base_density)), "esimp_set_body_force")

    def body_force(self, accel, density_data):
        a, pa = _f64(np.asarray(accel, dtype=np.float64).reshape(3))
        d, pd = _f64(density_data)
        assert d.size == self.n_cells
        out = np.empty(self.n_dofs)
        self._chk(self.lib.esimp_body_force(self._h, pa, pd, out.ctypes.data_as(_D)), "esimp_body_force")
        return out

    def stress_fields(self, want_fields=True):
        """-> dict(vm_q1, vm_avg, element_energy, max_vm, max_cell); the arrays are None with want_fields=False."""
        n = self.n_cells
        arrs = [np.empty(n) for _ in range(3)] if want_fields else [None] * 3
        ptrs = [a.ctypes.data_as(_D) if a is not None else None for a in arrs]
        mx = C.c_double(0)
        cell = C.c_int64(-1)
        self._chk(self.lib.esimp_stress_fields(self._h, ptrs[0], ptrs[1], ptrs[2], C.byref(mx), C.byref(cell)), "esimp_stress_fields")
        return dict(vm_q1=arrs[0], vm_avg=arrs[1], element_energy=arrs[2], max_vm=mx.value, max_cell=cell.value)

    @property
    def stress_nq(self):
        return int(self.lib.esimp_stress_nq(self._h))

    def stress_tensors(self, cell0=0, n=None):
        """sigma[n, nq, 6] of the caller's cells [cell0, cell0+n) (SymmetricTensor storage order 11 21 31 22 32 33)."""
        n = self.n_cells - cell0 if n is None else n
        out = np.empty((n, self.stress_nq, 6))
        self._chk(self.lib.esimp_stress_tensors(self._h, cell0, n, out.ctypes.data_as(_D)), "esimp_stress_tensors")
        return out

    def filter_build_structured(self):
        self._chk(self.lib.esimp_filter_build_structured(self._h), "esimp_filter_build_structured")

    def filter_build_csr(self, rowptr, col, centers, volumes, R):
        k1, p1 = _i64(rowptr)
        k2, p2 = _i64(col)
        k3, p3 = _f64(np.asarray(centers).reshape(-1))
        k4, p4 = _f64(volumes)
        self._chk(self.lib.esimp_filter_build_csr(self._h, p1, p2, p3, p4, float(R)), "esimp_filter_build_csr")

    @property
    def n_dofs(self):
        return int(self.lib.esimp_n_dofs(self._h))

    @property
    def n_cells(self):
        return int(self.lib.esimp_n_cells(self._h))

    @property
    def launch_count(self):
        return int(self.lib.esimp_launch_count(self._h))

    # -- stage calls
    def set_density(self, rho):
        k, p = _f64(rho)
        assert k.size == self.n_cells
        self._chk(self.lib.esimp_set_density(self._h, p), "esimp_set_density")

    def matvec(self, u):
        k, p = _f64(u)
        assert k.size == self.n_dofs
        y = np.empty(self.n_dofs)
        self._chk(self.lib.esimp_matvec(self._h, p, y.ctypes.data_as(_D)), "esimp_matvec")
        return y

    def diagonal(self):
        d = np.empty(self.n_dofs)
        self._chk(self.lib.esimp_diagonal(self._h, d.ctypes.data_as(_D)), "esimp_diagonal")
        return d

    def solve(self, want_u=True):
        u = np.empty(self.n_dofs) if want_u else None
        it = C.c_int(0)
        rr = C.c_double(0)
        st = self._chk(self.lib.esimp_solve(self._h, u.ctypes.data_as(_D) if want_u else None, C.byref(it), C.byref(rr)),
                       "esimp_solve")
        return u, it.value, rr.value, st

    def sensitivities(self, want_dc=True):
        dc = np.empty(self.n_cells) if want_dc else None
        en = C.c_double(0)
        self._chk(self.lib.esimp_sensitivities(self._h, dc.ctypes.data_as(_D) if want_dc else None, C.byref(en)),
                  "esimp_sensitivities")
        return dc, en.value

    def filter_sens(self, rho, dc):
        k1, p1 = _f64(rho)
        k2, p2 = _f64(dc)
        out = np.empty(self.n_cells)
        self._chk(self.lib.esimp_filter_sens(self._h, p1, p2, out.ctypes.data_as(_D)), "esimp_filter_sens")
        return out

    def filter_dens(self, rho):
        k1, p1 = _f64(rho)
        out = np.empty(self.n_cells)
        self._chk(self.lib.esimp_filter_dens(self._h, p1, out.ctypes.data_as(_D)), "esimp_filter_dens")
        return out

    def filter_dens_T(self, s):
        k1, p1 = _f64(s)
        out = np.empty(self.n_cells)
        self._chk(self.lib.esimp_filter_dens_T(self._h, p1, out.ctypes.data_as(_D)), "esimp_filter_dens_T")
        return out

    def oc(self, rho, dc, dv):
        k1, p1 = _f64(rho)
        k2, p2 = _f64(dc)
        k3, p3 = _f64(dv)
        out = np.empty(self.n_cells)
        lam = C.c_double(0)
        steps = C.c_int(0)
        st = self._chk(self.lib.esimp_oc(self._h, p1, p2, p3, out.ctypes.data_as(_D), C.byref(lam), C.byref(steps)), "esimp_oc")
        return out, lam.value, steps.value, st

    # -- loop
    def init_design(self, rho=None):
        ptr = None
        if rho is not None:
            k, ptr = _f64(rho)
        self._chk(self.lib.esimp_init_design(self._h, ptr), "esimp_init_design")

    def iterate(self) -> EsimpIterStats:
        s = EsimpIterStats()
        self._chk(self.lib.esimp_iterate(self._h, C.byref(s)), "esimp_iterate")
        return s

    def final_analysis(self):
        e = C.c_double(0)
        v = C.c_double(0)
        it = C.c_int(0)
        self._chk(self.lib.esimp_final_analysis(self._h, C.byref(e), C.byref(v), C.byref(it)), "esimp_final_analysis")
        return e.value, v.value, it.value

    def get_design(self, out=None):
        """`out`: optional caller buffer (e.g. pinned host memory) the library copies into directly."""
        if out is None:
            out = np.empty(self.n_cells)
        assert out.dtype == np.float64 and out.size == self.n_cells and out.flags.c_contiguous
        self._chk(self.lib.esimp_get_design(self._h, out.ctypes.data_as(_D)), "esimp_get_design")
        return out

    def get_physical_density(self):
        out = np.empty(self.n_cells)
        self._chk(self.lib.esimp_get_physical_density(self._h, out.ctypes.data_as(_D)), "esimp_get_physical_density")
        return out

    def get_displacement(self):
        out = np.empty(self.n_dofs)
        self._chk(self.lib.esimp_get_displacement(self._h, out.ctypes.data_as(_D)), "esimp_get_displacement")
        return out

    def set_displacement(self, u):
        k, p = _f64(u)
        assert k.size == self.n_dofs
        self._chk(self.lib.esimp_set_displacement(self._h, p), "esimp_set_displacement")

    def set_design(self, rho):
        k, p = _f64(rho)
        assert k.size == self.n_cells
        self._chk(self.lib.esimp_set_design(self._h, p), "esimp_set_design")

    def profile_enable(self, on=2):
        """0 off, 1 = clock the plain / dot operator launches only, 2 (or True... = 1!) every kernel category; pass the level as an int."""
        on = 2 if on is True else int(on)
        self._chk(self.lib.esimp_profile_enable(self._h, int(on)), "esimp_profile_enable")

    def profile_matvec(self):
        ms = C.c_double(0)
        n = C.c_int64(0)
        by = C.c_double(0)
        self._chk(self.lib.esimp_profile_matvec(self._h, C.byref(ms), C.byref(n), C.byref(by)), "esimp_profile_matvec")
        return ms.value, n.value, by.value

    def profile_operator_all(self):
        ms = C.c_double(0)
        n = C.c_int64(0)
        self._chk(self.lib.esimp_profile_operator_all(self._h, C.byref(ms), C.byref(n)), "esimp_profile_operator_all")
        return ms.value, n.value

    PROF_NAMES = ["operator_plain_dot", "operator_smoother", "operator_residual", "filter_stencil", "sens_energy", "oc_volumes",
                  "prolong_fine", "restrict_fine", "mg_update_first", "mg_direction", "um_matvec", "comm_halo", "comm_allreduce_scalars",
                  "comm_level_sum", "mg_below_fine", "host_pcg_wall", "host_pcg_wait"]

    def profile_kernels(self):
        """{category: (summed kernel ms, launches)} of the in-region CUDA-event timing (esimp_profile_kernels)"""
        n = len(self.PROF_NAMES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._chk(self.lib.esimp_profile_kernels(self._h, n, ms, cnt), "esimp_profile_kernels")
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.PROF_NAMES)}

    def bench_operator_mode(self, mode, reps):
        ms = C.c_double(0)
        by = C.c_double(0)
        self._chk(self.lib.esimp_bench_operator_mode(self._h, mode, reps, C.byref(ms), C.byref(by)), "esimp_bench_operator_mode")
        return ms.value, by.value

    def bench_matvec(self, reps):
        ms = C.c_double(0)
        by = C.c_double(0)
        self._chk(self.lib.esimp_bench_matvec(self._h, reps, C.byref(ms), C.byref(by)), "esimp_bench_matvec")
        return ms.value, by.value


def comm_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    st = lib.esimp_comm_unique_id(buf)
    if st < 0:
        raise EsimpError(f"esimp_comm_unique_id failed: {lib.esimp_last_error(None).decode()}")
    return buf.raw


class MultiContext:
    """esimp_multi*: N GPUs driven from this one process (one host thread per GPU inside the library).  Same global host arrays as
    Context; only the loop-path calls exist here, `rank_ctx` exposes a rank's own esimp_ctx* for anything else."""

    def __init__(self, devices=None, n_gpus=None):
        self.lib = load_library()
        self._h = _P()
        if devices is None:
            n = int(n_gpus)
            arr = None
        else:
            n = len(devices)
            arr = (C.c_int * n)(*devices)
        st = self.lib.esimp_multi_create(C.byref(self._h), arr, n)
        if st < 0:
            raise EsimpError(f"esimp_multi_create failed ({st}): {self.lib.esimp_multi_last_error(None).decode()}")
        self.params = EsimpParams()
        self.lib.esimp_default_params(C.byref(self.params))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.esimp_multi_destroy(self._h)
            self._h = None

    __del__ = close

    def _chk(self, st, what):
        if st < 0:
            raise EsimpError(f"{what} failed ({st}): {self.lib.esimp_multi_last_error(self._h).decode()}")
        return st

    @property
    def n_gpus(self):
        return self.lib.esimp_multi_n_gpus(self._h)

    @property
    def n_dofs(self):
        return self.lib.esimp_multi_n_dofs(self._h)

    @property
    def n_cells(self):
        return self.lib.esimp_multi_n_cells(self._h)

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise AttributeError(k)
            setattr(self.params, k, v)
        self._chk(self.lib.esimp_multi_set_params(self._h, C.byref(self.params)), "esimp_multi_set_params")

    def set_grid_structured(self, nx, ny, nz, hx, hy, hz, dof_of_node=None):
        ptr = None
        if dof_of_node is not None:
            keep, ptr = _i64(np.asarray(dof_of_node).reshape(-1))
        self._chk(self.lib.esimp_multi_set_grid_structured(self._h, nx, ny, nz, hx, hy, hz, ptr), "esimp_multi_set_grid_structured")

    def set_dirichlet(self, dofs):
        keep, ptr = _i64(dofs)
        self._chk(self.lib.esimp_multi_set_dirichlet(self._h, ptr, len(keep)), "esimp_multi_set_dirichlet")

    def set_load(self, f):
        keep, ptr = _f64(f)
        self._chk(self.lib.esimp_multi_set_load(self._h, ptr), "esimp_multi_set_load")

    def set_body_force(self, accel, base_density=1.0):
        if accel is None:
            self._chk(self.lib.esimp_multi_set_body_force(self._h, None, 0.0), "esimp_multi_set_body_force")
        else:
            keep, ptr = _f64(accel)
            self._chk(self.lib.esimp_multi_set_body_force(self._h, ptr, float(base_density)), "esimp_multi_set_body_force")

    def filter_build_structured(self):
        self._chk(self.lib.esimp_multi_filter_build_structured(self._h), "esimp_multi_filter_build_structured")

    def init_design(self, rho=None):
        ptr = None
        if rho is not None:
            keep, ptr = _f64(rho)
        self._chk(self.lib.esimp_multi_init_design(self._h, ptr), "esimp_multi_init_design")

    def set_design(self, rho):
        keep, ptr = _f64(rho)
        self._chk(self.lib.esimp_multi_set_design(self._h, ptr), "esimp_multi_set_design")

    def iterate(self):
        st = EsimpIterStats()
        self._chk(self.lib.esimp_multi_iterate(self._h, C.byref(st)), "esimp_multi_iterate")
        return st

    def final_analysis(self):
        e, v, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self._chk(self.lib.esimp_multi_final_analysis(self._h, C.byref(e), C.byref(v), C.byref(it)), "esimp_multi_final_analysis")
        return e.value, v.value, it.value

    def get_design(self):
        rho = np.zeros(self.n_cells)
        self._chk(self.lib.esimp_multi_get_design(self._h, rho.ctypes.data_as(_D)), "esimp_multi_get_design")
        return rho

    def get_physical_density(self):
        rho = np.zeros(self.n_cells)
        self._chk(self.lib.esimp_multi_get_physical_density(self._h, rho.ctypes.data_as(_D)), "esimp_multi_get_physical_density")
        return rho

    def get_displacement(self):
        u = np.zeros(self.n_dofs)
        self._chk(self.lib.esimp_multi_get_displacement(self._h, u.ctypes.data_as(_D)), "esimp_multi_get_displacement")
        return u
