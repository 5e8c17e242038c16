"""CPU oracle for the batched linear-algebra hot path of GPUtils (numpy / scipy-LAPACK + oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs may import this module; the product path never does.

Batches are numpy arrays of shape (k, m, n) (matrix index first). Each function cites the reference call
site it restates. Parity pinning: PINNED by tests/golden/reference_vectors.json (the reference's own golden
vectors) and, on the GPU box, by oracle/_ref (the reference itself).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "liboracle.so"
_lib = None


def build_c_oracle(force: bool = False) -> Path:
    """gcc -O2 -fopenmp oracle.c -> oracle/_build/liboracle.so (committed recipe: oracle/Makefile)."""
    src = _HERE / "oracle.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        _LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", str(_LIB), str(src), "-lm"], check=True)
    return _LIB


def clib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_c_oracle()))
        for suf in ("f64", "f32"):
            for nm in ("dot", "nrm2", "asum"):
                getattr(_lib, f"oracle_{nm}_{suf}").restype = C.c_double
    return _lib


def _suf(a):
    return {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[np.asarray(a).dtype]


def _ct(a):
    return C.c_double if np.asarray(a).dtype == np.float64 else C.c_float


def _cm(a):
    """(k, m, n) batch -> flat column-major buffer in the reference's layout (tensor.cuh:1278-1284)."""
    return np.array(np.asarray(a).transpose(0, 2, 1), order="C", copy=True)  # always a private copy: the C routines work in place


def _from_cm(buf, k, m, n):
    return buf.reshape(k, n, m).transpose(0, 2, 1).copy()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(clib().oracle_num_threads())


# ---- GEMM: C_i <- beta C_i + alpha A_i B_i (ref: tensor.cuh:1286-1338) ---------------------------------------
def gemm_batched(A, B, Cin=None, alpha=1.0, beta=0.0, use_c=True):
    A = np.asarray(A); B = np.asarray(B)
    k, m, kk = A.shape
    n = B.shape[2]
    if not use_c:
        out = alpha * np.einsum("bil,blj->bij", A, B)
        return out if beta == 0 else out + beta * Cin
    c = _cm(Cin) if Cin is not None else np.zeros((k, n, m), dtype=A.dtype)
    a, b = _cm(A), _cm(B)
    ct = _ct(A)
    getattr(clib(), f"oracle_gemm_batched_{_suf(A)}")(C.c_size_t(m), C.c_size_t(n), C.c_size_t(kk), ct(alpha), _ptr(a), _ptr(b),
                                                   ct(beta), _ptr(c), C.c_size_t(k))
    return _from_cm(c, k, m, n)


# ---- Cholesky (ref: tensor.cuh:2135-2197, 1742-1783) ----------------------------------------------------------
def potrf_batched(A):
    """Returns (A with the lower triangle replaced by L, info[k])."""
    A = np.asarray(A)
    k, n, _ = A.shape
    a = _cm(A)
    info = np.zeros(k, dtype=np.int32)
    getattr(clib(), f"oracle_potrf_batched_{_suf(A)}")(C.c_size_t(n), _ptr(a), _ptr(info), C.c_size_t(k))
    return _from_cm(a, k, n, n), info


def potrs_batched(L, b):
    L = np.asarray(L); b = np.asarray(b)
    k, n, _ = L.shape
    l = _cm(L)
    x = np.ascontiguousarray(b.reshape(k, n)).copy()
    getattr(clib(), f"oracle_potrs_batched_{_suf(L)}")(C.c_size_t(n), _ptr(l), _ptr(x), C.c_size_t(k))
    return x.reshape(k, n, 1)


# ---- Householder QR / least squares (ref: tensor.cuh:1866-1927, 1340-1394) ----------------------------------------
def geqrf_batched(A):
    A = np.asarray(A)
    k, m, n = A.shape
    a = _cm(A)
    tau = np.zeros((k, n), dtype=A.dtype)
    getattr(clib(), f"oracle_geqrf_batched_{_suf(A)}")(C.c_size_t(m), C.c_size_t(n), _ptr(a), _ptr(tau), C.c_size_t(k))
    return _from_cm(a, k, m, n), tau


def ormqr_batched(trans, QR, tau, Cm):
    QR = np.asarray(QR); Cm = np.asarray(Cm)
    k, m, n = QR.shape
    nc = Cm.shape[2]
    a, c = _cm(QR), _cm(Cm)
    t = np.ascontiguousarray(tau)
    getattr(clib(), f"oracle_ormqr_batched_{_suf(QR)}")(C.c_int(1 if trans else 0), C.c_size_t(m), C.c_size_t(nc), C.c_size_t(n),
                                                    _ptr(a), _ptr(t), _ptr(c), C.c_size_t(k))
    return _from_cm(c, k, m, nc)


def gels_batched(A, b):
    """Returns (QR factors in A's storage, b with x in rows 0..n-1 and the Q^T b tail below, info)."""
    A = np.asarray(A); b = np.asarray(b)
    k, m, n = A.shape
    a = _cm(A)
    x = np.ascontiguousarray(b.reshape(k, m)).copy()
    info = np.zeros(k, dtype=np.int32)
    getattr(clib(), f"oracle_gels_batched_{_suf(A)}")(C.c_size_t(m), C.c_size_t(n), _ptr(a), _ptr(x), _ptr(info), C.c_size_t(k))
    return _from_cm(a, k, m, n), x.reshape(k, m, 1), info


# ---- SVD (ref: tensor.cuh:1624-1676: ?gesvd, jobu 'A'/'N', jobvt 'A') --------------------------------------------
def gesvd_batched(A, want_u=True):
    """LAPACK ?gesvd per matrix through scipy (the published algorithm cuSOLVER's gesvd follows).
    Returns S (k, n), U (k, m, m) or None, Vt (k, n, n)."""
    from scipy.linalg import lapack
    A = np.asarray(A)
    k, m, n = A.shape
    fn = lapack.dgesvd if A.dtype == np.float64 else lapack.sgesvd
    S = np.zeros((k, n), dtype=A.dtype); Vt = np.zeros((k, n, n), dtype=A.dtype)
    U = np.zeros((k, m, m), dtype=A.dtype) if want_u else None
    for i in range(k):
        u, s, vt, info = fn(np.asfortranarray(A[i]), compute_uv=1, full_matrices=1)
        assert info == 0
        S[i] = s; Vt[i] = vt
        if want_u:
            U[i] = u
    return S, U, Vt


def rank_batched(S, eps):
    """#{s > eps} per matrix (ref: tensor.cuh:1486-1491, 1600-1609)."""
    return (np.asarray(S) > eps).sum(axis=1).astype(np.uint32)


def nullspace_batched(A, eps=1e-6):
    """N_i (n, n): last n - rank_i columns of U of A_i^T, left-packed, zero padded; and N_i N_i^T
    (ref: tensor.cuh:2046-2079)."""
    A = np.asarray(A)
    k, m, n = A.shape
    S, U, _ = gesvd_batched(A.transpose(0, 2, 1), want_u=True)
    r = rank_batched(S, eps)
    N = np.zeros((k, n, n), dtype=A.dtype)
    for i in range(k):
        nul = n - int(r[i])
        if nul:
            N[i][:, :nul] = U[i][:, n - nul:]
    return N, N @ N.transpose(0, 2, 1), r


# ---- flat reductions (ref: tensor.cuh:968-1072) ---------------------------------------------------------------------
def dot(x, y):
    x = np.ascontiguousarray(x).ravel(); y = np.ascontiguousarray(y).ravel()
    return float(getattr(clib(), f"oracle_dot_{_suf(x)}")(C.c_size_t(x.size), _ptr(x), _ptr(y)))


def nrm2(x):
    x = np.ascontiguousarray(x).ravel()
    return float(getattr(clib(), f"oracle_nrm2_{_suf(x)}")(C.c_size_t(x.size), _ptr(x)))


def asum(x):
    x = np.ascontiguousarray(x).ravel()
    return float(getattr(clib(), f"oracle_asum_{_suf(x)}")(C.c_size_t(x.size), _ptr(x)))


# ---- the counter-based generator of gpub_fill_uniform_* / gpub_fill_spd_batched_*  -----------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def u01(seed: int, idx):
    with np.errstate(over="ignore"):
        h = _mix64(np.uint64(seed) ^ _mix64(np.asarray(idx, dtype=np.uint64)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def fill_uniform(n, lo, hi, seed, dtype=np.float64):
    return (lo + (hi - lo) * u01(seed, np.arange(n, dtype=np.uint64))).astype(dtype)


def fill_spd_batched(n, batch, shift, seed, dtype=np.float64):
    """A_i = G_i G_i^T + shift I, G ~ U[-1, 1]; returns (batch, n, n)."""
    g = 2.0 * u01(seed, np.arange(batch * n * n, dtype=np.uint64)) - 1.0
    G = g.reshape(batch, n, n).transpose(0, 2, 1)  # element (i, k) of matrix b at b*n*n + i + k*n
    A = G @ G.transpose(0, 2, 1) + shift * np.eye(n)
    return A.astype(dtype)
