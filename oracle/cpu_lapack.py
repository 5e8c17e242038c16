"""Loader of oracle/cpu_lapack.c: the host-CPU baseline of SURVEY.md 8(d) "Baseline 2" -- single-threaded OpenBLAS LAPACK
(the one bundled with scipy) under an OpenMP loop over the batch.

TEST / BENCH INFRASTRUCTURE ONLY: bench.py's cpu_baseline leg and tests/ load it; the product never does.
Batches are numpy arrays in the reference's device layout: C-contiguous (k, cols, rows), i.e. column-major matrices with the
mats axis slowest (tensor.cuh:1278-1284). The routines work in place and return the seconds the loop took."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libcpu_lapack.so"
_lib = None


def openblas_path() -> str:
    import scipy
    hits = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    if not hits:
        raise RuntimeError("scipy's bundled OpenBLAS (scipy.libs/libscipy_openblas*.so) not found")
    return os.path.realpath(hits[0])


def build(force: bool = False) -> Path:
    src = _HERE / "cpu_lapack.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        _LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", str(_LIB), str(src), "-ldl"], check=True)
    return _LIB


def lib(threads: int | None = None) -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        for nm in ("cpu_chol_batch", "cpu_gemm_batch", "cpu_gels_batch", "cpu_geqrf_batch", "cpu_gesvd_batch"):
            getattr(_lib, nm).restype = C.c_double
        rc = _lib.cpu_lapack_open(openblas_path().encode())
        if rc != 0:
            raise RuntimeError(f"cpu_lapack_open failed ({rc})")
    if threads:
        _lib.cpu_lapack_set_threads(int(threads))
    return _lib


def host_threads() -> int:
    """every hardware thread this process may run on (torchrun sets OMP_NUM_THREADS=1: that is not the host's core count)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return 1 if a.dtype == np.float64 else 0


def chol_batch(A, b=None, info=None, threads=None) -> float:
    k, n = A.shape[0], A.shape[1]
    return lib(threads).cpu_chol_batch(_f64(A), n, C.c_size_t(k), _p(A), _p(b), _p(info))


def gemm_batch(A, B, Cm, alpha=1.0, beta=0.0, threads=None) -> float:
    k, ka, m = A.shape
    n = B.shape[1]
    return lib(threads).cpu_gemm_batch(_f64(A), m, n, ka, C.c_size_t(k), _p(A), _p(B), _p(Cm), C.c_double(alpha), C.c_double(beta))


def gels_batch(A, b, threads=None) -> float:
    k, n, m = A.shape
    return lib(threads).cpu_gels_batch(_f64(A), m, n, C.c_size_t(k), _p(A), _p(b))


def geqrf_batch(A, tau, threads=None) -> float:
    k, n, m = A.shape
    return lib(threads).cpu_geqrf_batch(_f64(A), m, n, C.c_size_t(k), _p(A), _p(tau))


def gesvd_batch(A, S, U, Vt, threads=None) -> float:
    k, n, m = A.shape
    return lib(threads).cpu_gesvd_batch(_f64(A), m, n, C.c_size_t(k), _p(A), _p(S), _p(U), _p(Vt))
