"""ctypes wrapper of oracle/cpu_simp.c (OpenMP CPU restatement of the structured hot path).
TEST / BASELINE INFRASTRUCTURE ONLY -- see the header of cpu_simp.c.  Arrays use the caller's native
numbering: node = i + (nx+1)(j + (ny+1)k), dof = 3 node + c, cell = i + nx (j + ny k)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_D = C.POINTER(C.c_double)
_U8 = C.POINTER(C.c_uint8)
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "libcpu_simp.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(HERE)], check=True)
        L = C.CDLL(str(so))
        L.cpu_num_threads.restype = C.c_int
        L.cpu_set_threads.argtypes = [C.c_int]
        L.cpu_hex8_matvec.argtypes = [C.c_int] * 3 + [_D, _D, _U8, C.c_double, _D, _D]
        L.cpu_hex8_diagonal.argtypes = [C.c_int] * 3 + [_D, _D, _D]
        L.cpu_hex8_diagonal.restype = C.c_double
        L.cpu_pcg.argtypes = [C.c_int] * 3 + [_D, _D, _U8, _D, _D, C.c_double, C.c_int, _D, _D]
        L.cpu_pcg.restype = C.c_int
        L.cpu_sensitivities.argtypes = [C.c_int] * 3 + [_D, _D, _D, _D, C.c_double, C.c_double, C.c_double, _D]
        L.cpu_sensitivities.restype = C.c_double
        L.cpu_filter_sens.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [_D, _D, _D]
        L.cpu_oc_step.argtypes = [C.c_int64, _D, _D, _D, C.c_double, C.c_double, C.c_double, C.c_double, _D]
        L.cpu_oc_step.restype = C.c_double
        L.cpu_modulus.argtypes = [C.c_int64, _D, C.c_double, C.c_double, C.c_double, _D]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_D)


def num_threads():
    return lib().cpu_num_threads()


def matvec(nel, K0, E, fixed, m, u):
    y = np.empty_like(u)
    lib().cpu_hex8_matvec(*nel, _d(K0), _d(E), fixed.ctypes.data_as(_U8), m, _d(u), _d(y))
    return y


def diagonal(nel, K0, E):
    nx, ny, nz = nel
    d = np.empty(3 * (nx + 1) * (ny + 1) * (nz + 1))
    mean = lib().cpu_hex8_diagonal(nx, ny, nz, _d(K0), _d(E), _d(d))
    return d, mean


def pcg(nel, K0, E, fixed, f, x, rtol, maxit, work=None):
    if work is None:
        work = np.empty(4 * f.size)
    rr = C.c_double(0)
    it = lib().cpu_pcg(*nel, _d(K0), _d(E), fixed.ctypes.data_as(_U8), _d(f), _d(x), rtol, maxit, _d(work), C.byref(rr))
    return it, rr.value


def sensitivities(nel, K0, rho, E, u, E0, Emin, p):
    dc = np.empty_like(rho)
    en = lib().cpu_sensitivities(*nel, _d(K0), _d(rho), _d(E), _d(u), E0, Emin, p, _d(dc))
    return dc, en


def filter_sens(nel, h, R, rho, dc):
    out = np.empty_like(rho)
    lib().cpu_filter_sens(*nel, h[0], h[1], h[2], R, _d(rho), _d(dc), _d(out))
    return out


def oc_step(rho, dc, dv, lam, move, damping, vol):
    new = np.empty_like(rho)
    v = lib().cpu_oc_step(rho.size, _d(rho), _d(dc), _d(dv), lam, move, damping, vol, _d(new))
    return new, v


def modulus(rho, E0, Emin, p):
    E = np.empty_like(rho)
    lib().cpu_modulus(rho.size, _d(rho), E0, Emin, p, _d(E))
    return E


def simp_iteration_sample(nel, h, K0, rho, f, fixed, u, prm, n_pcg, work=None):
    """One SIMP iteration of the CPU path with the solve truncated to `n_pcg` PCG iterations: returns
    per-phase seconds.  Used by bench.py to time a BOUNDED sample of the workload."""
    import time

    t = {}
    t0 = time.perf_counter()
    E = modulus(rho, prm["E0"], prm["Emin"], prm["p"])
    t["modulus"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    it, rr = pcg(nel, K0, E, fixed, f, u, 0.0, n_pcg, work)
    t["pcg"] = time.perf_counter() - t0
    t["pcg_iters"] = it
    t0 = time.perf_counter()
    dc, en = sensitivities(nel, K0, rho, E, u, prm["E0"], prm["Emin"], prm["p"])
    t["sens"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    dcf = filter_sens(nel, h, prm["R"], rho, dc)
    t["filter"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    vol = h[0] * h[1] * h[2]
    dv = np.full(rho.size, 1.0 / rho.size)
    lo, hi = 1e-9, 1e9
    steps = 0
    for steps in range(1, 201):
        lam = 0.5 * (lo + hi)
        new, v = oc_step(rho, dcf, dv, lam, prm["move"], prm["damping"], vol)
        err = v - prm["volfrac"] * vol * rho.size
        if abs(err) < 1e-6:
            break
        if err > 0:
            lo = lam
        else:
            hi = lam
    t["oc"] = time.perf_counter() - t0
    t["oc_steps"] = steps
    return t, new, en
