"""CPU oracle package (test infrastructure only -- see pt_oracle.c / pt_numpy.py headers)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _cpu_id() -> str:
    """Identifies the host CPU's instruction set: the library is built -march=native, and it travels between machines."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                import hashlib
                return hashlib.sha1(line.encode()).hexdigest()[:16]
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpt_oracle.so")
    src = os.path.join(_HERE, "pt_oracle.c")
    stamp = os.path.join(_HERE, ".build_cpu")
    cpu = _cpu_id()
    built_for = open(stamp).read().strip() if os.path.exists(stamp) else ""
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src) or built_for != cpu:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libpt_oracle.so", "ARCH=-march=native"])
        with open(stamp, "w") as fh:
            fh.write(cpu)
    return so


def openblas_path():
    """The OpenBLAS bundled with the scipy wheel (LP64 build, `scipy_`-prefixed symbols); None if absent."""
    import glob
    try:
        import scipy
    except ImportError:
        return None
    hits = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas-*.so")))
    return hits[0] if hits else None


def use_blas(which: str = "own") -> str:
    """Route the GEMMs of pt_gemm through the built-in register-blocked kernel ("own") or a single-threaded OpenBLAS
    cblas_dgemm ("openblas" -- the reference's own arrangement, ijk.jl:45-46).  Returns what is active."""
    L = lib()
    if which == "openblas":
        path = openblas_path()
        if path and L.pt_oracle_use_blas(path.encode()) == 0:
            return "openblas"
    L.pt_oracle_use_blas(None)
    return "own"


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        L.pt_oracle_naive.argtypes = [ctypes.c_int, ctypes.c_int] + [dp] * 7 + [dp]
        L.pt_oracle_naive.restype = ctypes.c_int
        L.pt_oracle_gemm.argtypes = [ctypes.c_int, ctypes.c_int] + [dp] * 7 + [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, dp]
        L.pt_oracle_gemm.restype = ctypes.c_int
        L.pt_oracle_num_threads.restype = ctypes.c_int
        L.pt_oracle_use_blas.argtypes = [ctypes.c_char_p]
        L.pt_oracle_use_blas.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f(a):
    a = np.asfortranarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _args(T1, T2, OVVV, OOOV, OVOV, fo, fv):
    keep, ptrs = [], []
    for a in (T1, T2, OVVV, OOOV, OVOV, fo, fv):
        k, p = _f(a)
        keep.append(k)
        ptrs.append(p)
    return keep, ptrs


def pt_naive(T1, T2, OVVV, OOOV, OVOV, fo, fv) -> float:
    o, v = T1.shape
    keep, ptrs = _args(T1, T2, OVVV, OOOV, OVOV, fo, fv)
    e = ctypes.c_double(0.0)
    rc = lib().pt_oracle_naive(o, v, *ptrs, ctypes.byref(e))
    if rc:
        raise MemoryError("pt_oracle_naive failed")
    return e.value


def pt_gemm(T1, T2, OVVV, OOOV, OVOV, fo, fv, t_begin: int = 0, t_end: int = -1, nthreads: int = 0) -> float:
    o, v = T1.shape
    keep, ptrs = _args(T1, T2, OVVV, OOOV, OVOV, fo, fv)
    e = ctypes.c_double(0.0)
    rc = lib().pt_oracle_gemm(o, v, *ptrs, t_begin, t_end, nthreads, ctypes.byref(e))
    if rc:
        raise MemoryError("pt_oracle_gemm failed")
    return e.value


def num_threads() -> int:
    return lib().pt_oracle_num_threads()
