"""ctypes binding of the C ABI declared in include/gputils_b200.h.

This is plumbing for tests and bench.py: it loads the in-tree libgputils_b200.so and passes raw device
pointers (torch is used only for device memory and streams). There is no fallback: if the library is
missing or a launcher returns non-zero, a GpubError is raised.

Tensor convention: a reference DTensor of shape (m rows, n cols, k mats) is column-major with the mats
axis slowest (ref: tensor.cuh:1278-1284). The equivalent torch tensor is a C-contiguous tensor of shape
(k, n, m); `from_numpy_batch` / `to_numpy_batch` convert from / to the usual numpy (k, m, n) batches.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_LIB_PATH = Path(os.environ.get("GPUB_LIB") or (Path(__file__).resolve().parent / "lib" / "libgputils_b200.so"))  # GPUB_LIB: tuning variants only


class GpubError(RuntimeError):
    pass


_sz, _int, _vp, _dbl, _flt, _u64 = C.c_size_t, C.c_int, C.c_void_p, C.c_double, C.c_float, C.c_uint64

# name -> (argtypes after (ctx, sidx)); every launcher returns int. {T} is c_double / c_float.
_TYPED = {
    "dot": [_sz, _vp, _vp, _vp],
    "nrm2": [_sz, _vp, _vp],
    "asum": [_sz, _vp, _vp],
    "amax_abs": [_sz, _vp, _vp, _vp],
    "amin_abs": [_sz, _vp, _vp, _vp],
    "scal": [_sz, "T", _vp],
    "axpy": [_sz, "T", _vp, _vp],
    "rot": [_sz, _vp, _sz, _vp, _sz, _vp, _vp, _int],
    "givens_rhypot": [_vp, _vp, _sz, _sz, _sz, _sz],
    "rot_batched": [_sz, _vp, _sz, _vp, _sz, _sz, _vp, _vp, _sz],
    "givens_annihilate_batched": [_vp, _sz, _sz, _sz, _sz, _sz, _sz, _sz],
    "gather_rows": [_vp, _sz, _sz, _sz, _sz, _vp],
    "transpose_batched": [_sz, _sz, _vp, _sz, _vp, _sz, _sz],
    "gemm_batched": [_sz, _sz, _sz, "T", _vp, _sz, _sz, _vp, _sz, _sz, "T", _vp, _sz, _sz, _sz],
    "potrf_batched": [_sz, _vp, _sz, _sz, _vp, _sz],
    "potrs_batched": [_sz, _vp, _sz, _sz, _vp, _sz, _sz],
    "potrs_allgather_batched": [_sz, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _int, _sz, _sz],
    "geqrf_batched": [_sz, _sz, _vp, _sz, _sz, _vp, _sz, _sz],
    "ormqr_batched": [_int, _sz, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _sz, _sz, _sz],
    "trsv_upper_batched": [_sz, _vp, _sz, _sz, _vp, _sz, _sz],
    "gels_batched": [_sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _sz],
    "gesvd_batched": [_int, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _sz],
    "count_gt_batched": [_vp, _sz, _sz, "T", _vp, _sz],
    "nullspace_pack_batched": [_sz, _vp, _sz, _vp, _vp, _sz, _sz],
    "aat_batched": [_sz, _vp, _sz, _vp, _sz, _sz],
    "nullspace_projector_batched": [_sz, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _sz],
    "nullspace_build_batched": [_sz, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _sz],
    "fill_uniform": [_sz, _vp, "T", "T", _u64],
    "fill_spd_batched": [_sz, _vp, _sz, "T", _u64, _sz],
    "chol_solve_from_host": [_sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz],
}

#: every symbol include/gputils_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTED = (
    ["gpub_version", "gpub_ctx_get", "gpub_ctx_ensure_streams", "gpub_ctx_num_streams", "gpub_ctx_stream",
     "gpub_ctx_bind_stream", "gpub_ctx_sync", "gpub_ctx_sync_all", "gpub_ctx_release", "gpub_ctx_release_all", "gpub_ctx_device", "gpub_ctx_sm_count",
     "gpub_multi_device_count", "gpub_multi_enable_peer_access", "gpub_multi_nccl_version", "gpub_multi_allgather", "gpub_multi_release",
     "gpub_fill_ptr_table", "gpub_gesvd_batched_worksize_f64", "gpub_gesvd_batched_worksize_f32",
     "gpub_mem_alloc", "gpub_mem_free", "gpub_mem_stats", "gpub_mem_trim", "gpub_upload", "gpub_download"]
    + [f"gpub_{n}_{s}" for n in _TYPED for s in ("f64", "f32")]
)

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Loads libgputils_b200.so (built by gputils_b200/build.py). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise GpubError(f"{_LIB_PATH} not found: run `python gputils_b200/build.py` (nvcc, sm_100a). "
                        "There is no CPU or library fallback.")
    lib = C.CDLL(str(_LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    lib.gpub_version.restype = C.c_char_p
    lib.gpub_ctx_get.argtypes = [_int, C.POINTER(_vp)]
    lib.gpub_ctx_ensure_streams.argtypes = [_vp, _int]
    lib.gpub_ctx_num_streams.argtypes = [_vp]
    lib.gpub_ctx_stream.argtypes = [_vp, _int, C.POINTER(_vp)]
    lib.gpub_ctx_bind_stream.argtypes = [_vp, _int, _vp]
    lib.gpub_ctx_sync.argtypes = [_vp, _int]
    lib.gpub_ctx_sync_all.argtypes = [_vp]
    lib.gpub_ctx_release.argtypes = [_vp]
    lib.gpub_ctx_device.argtypes = [_vp]
    lib.gpub_multi_device_count.argtypes = [C.POINTER(_int)]
    lib.gpub_multi_enable_peer_access.argtypes = [C.POINTER(_int), _int, C.POINTER(_int)]
    lib.gpub_multi_nccl_version.argtypes = [C.POINTER(_int)]
    lib.gpub_multi_allgather.argtypes = [C.POINTER(_vp), _int, _int, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), _int, C.POINTER(_int)]
    lib.gpub_ctx_sm_count.argtypes = [_vp]
    lib.gpub_fill_ptr_table.argtypes = [_vp, _int, _vp, _sz, _sz, _vp]
    lib.gpub_mem_alloc.argtypes = [_vp, _sz, C.POINTER(_vp)]
    lib.gpub_mem_free.argtypes = [_vp]
    lib.gpub_mem_stats.argtypes = [_vp, C.POINTER(_sz), C.POINTER(_sz)]
    lib.gpub_mem_trim.argtypes = [_vp, _sz]
    lib.gpub_upload.argtypes = [_vp, _int, _vp, _vp, _sz]
    lib.gpub_download.argtypes = [_vp, _int, _vp, _vp, _sz]
    for suf in ("f64", "f32"):
        fn = getattr(lib, f"gpub_gesvd_batched_worksize_{suf}")
        fn.argtypes = [_sz, _sz, _int, _sz]
        fn.restype = _sz
    for name, args in _TYPED.items():
        for suf, ct in (("f64", _dbl), ("f32", _flt)):
            fn = getattr(lib, f"gpub_{name}_{suf}")
            fn.argtypes = [_vp, _int] + [ct if a == "T" else a for a in args]
            fn.restype = _int
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        kind = "cudaError" if status > 0 else "GPUB_E"
        raise GpubError(f"libgputils_b200: {what} failed with {kind} {status}")


def _suffix(t) -> str:
    import torch
    if t.dtype == torch.float64:
        return "f64"
    if t.dtype == torch.float32:
        return "f32"
    raise GpubError(f"unsupported dtype {t.dtype}")


class Context:
    """Per-device stream context (the Session replacement). Stream 0 is bound to torch's current stream so
    that torch.cuda.Event timing and torch allocations order with the launches."""

    def __init__(self, device: int | None = None, bind_torch_stream: bool = True):
        import torch
        self.lib = load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = device
        h = _vp()
        check(self.lib.gpub_ctx_get(device, C.byref(h)), "gpub_ctx_get")
        self.h = h
        if bind_torch_stream:
            self.bind_torch_stream()

    def bind_torch_stream(self, sidx: int = 0):
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.gpub_ctx_bind_stream(self.h, sidx, _vp(s)), "gpub_ctx_bind_stream")

    def sync(self, sidx: int = 0):
        check(self.lib.gpub_ctx_sync(self.h, sidx), "gpub_ctx_sync")

    def call(self, name: str, t, *args, sidx: int = 0):
        """Calls gpub_<name>_<f32|f64> picked from tensor `t`'s dtype."""
        fn = getattr(self.lib, f"gpub_{name}_{_suffix(t)}")
        check(fn(self.h, sidx, *args), f"gpub_{name}_{_suffix(t)}")


def from_numpy_batch(a: np.ndarray, device="cuda"):
    """numpy (k, m, n) batch -> torch tensor in DTensor layout, shape (k, n, m) contiguous."""
    import torch
    a = np.asarray(a)
    if a.ndim == 2:
        a = a[None]
    return torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 1))).to(device)


def to_numpy_batch(t) -> np.ndarray:
    """torch tensor in DTensor layout (k, n, m) -> numpy (k, m, n)."""
    return t.detach().cpu().numpy().transpose(0, 2, 1).copy()


def _p(t):
    return _vp(t.data_ptr())


# ---- thin typed wrappers over the launchers; tensors are in DTensor layout (k, n, m) -------------------

def gemm_batched(ctx: Context, Cm, A, B, alpha=1.0, beta=0.0):
    k, ka, m = A.shape[0], A.shape[1], A.shape[2]
    n = B.shape[1]
    ctx.call("gemm_batched", A, m, n, ka, alpha, _p(A), m, m * ka, _p(B), ka, ka * n, beta, _p(Cm), m, m * n, k)


def potrf_batched(ctx: Context, A, info):
    k, n = A.shape[0], A.shape[1]
    ctx.call("potrf_batched", A, n, _p(A), n, n * n, _p(info), k)


def potrs_batched(ctx: Context, L, b):
    k, n = L.shape[0], L.shape[1]
    ctx.call("potrs_batched", L, n, _p(L), n, n * n, _p(b), n, k)


def gels_batched(ctx: Context, A, b, info=None):
    k, n, m = A.shape
    ctx.call("gels_batched", A, m, n, _p(A), m, m * n, _p(b), m, _p(info) if info is not None else None, k)


def geqrf_batched(ctx: Context, A, tau):
    k, n, m = A.shape
    ctx.call("geqrf_batched", A, m, n, _p(A), m, m * n, _p(tau), n, k)


def ormqr_batched(ctx: Context, trans: bool, A, tau, Cm):
    k, n, m = A.shape
    nc = Cm.shape[1]
    ctx.call("ormqr_batched", A, 1 if trans else 0, m, nc, n, _p(A), m, m * n, _p(tau), n, _p(Cm), m, m * nc, k)


def trsv_upper_batched(ctx: Context, R, n: int, ldr: int, stride_r: int, b, stride_b: int, batch: int):
    ctx.call("trsv_upper_batched", R, n, _p(R), ldr, stride_r, _p(b), stride_b, batch)


def gesvd_batched(ctx: Context, A, want_u: bool):
    """A (k, n, m) is destroyed. Returns S (k, n), U (k, m, m) or None, Vt (k, n, n), info (k,)."""
    import torch
    k, n, m = A.shape
    lib = ctx.lib
    ws = getattr(lib, f"gpub_gesvd_batched_worksize_{_suffix(A)}")(m, n, ord("A") if want_u else ord("N"), k)
    work = torch.empty(ws, dtype=torch.uint8, device=A.device)
    S = torch.empty((k, n), dtype=A.dtype, device=A.device)
    Vt = torch.empty((k, n, n), dtype=A.dtype, device=A.device)
    U = torch.empty((k, m, m), dtype=A.dtype, device=A.device) if want_u else None
    info = torch.zeros(k, dtype=torch.int32, device=A.device)
    ctx.call("gesvd_batched", A, ord("A") if want_u else ord("N"), m, n, _p(A), m, m * n, _p(S), n,
             _p(U) if want_u else None, m, m * m, _p(Vt), n, n * n, _p(work), ws, _p(info), k)
    return S, U, Vt, info


def transpose_batched(ctx: Context, A):
    import torch
    k, n, m = A.shape
    At = torch.empty((k, m, n), dtype=A.dtype, device=A.device)
    ctx.call("transpose_batched", A, m, n, _p(A), m * n, _p(At), m * n, k)
    return At


def reduce_scalar(ctx: Context, name: str, x, y=None):
    """name in {dot, nrm2, asum, amax_abs, amin_abs}; returns a python float (blocks, like cuBLAS)."""
    ct = _dbl if _suffix(x) == "f64" else _flt
    out = ct()
    n = x.numel()
    if name == "dot":
        ctx.call(name, x, n, _p(x), _p(y), C.byref(out))
    elif name in ("amax_abs", "amin_abs"):
        idx = C.c_longlong()
        ctx.call(name, x, n, _p(x), C.byref(out), C.byref(idx))
        return out.value, idx.value
    else:
        ctx.call(name, x, n, _p(x), C.byref(out))
    return out.value


def fill_uniform(ctx: Context, x, lo, hi, seed: int):
    ctx.call("fill_uniform", x, x.numel(), _p(x), lo, hi, seed)


def fill_spd_batched(ctx: Context, A, shift, seed: int):
    k, n = A.shape[0], A.shape[1]
    ctx.call("fill_spd_batched", A, n, _p(A), n * n, shift, seed, k)


def count_gt_batched(ctx: Context, S, eps: float):
    """rank_i = #{ j : S_i[j] > eps } (ref: tensor.cuh:1600-1609), one launch; returns int32 (k,)."""
    import torch
    k, length = S.shape[0], S.shape[1]
    count = torch.zeros(k, dtype=torch.int32, device=S.device)
    ctx.call("count_gt_batched", S, _p(S), length, length, eps, _p(count), k)
    return count


def nullspace_build(ctx: Context, a, eps: float = 1e-6):
    """The launch sequence of Nullspace<T>::Nullspace (include/gpub200/factorisers.cuh; ref: tensor.cuh:2046-2079) through
    the C ABI: tr -> gesvd(U) -> rank -> pack -> N N'. `a` is (k, n, m) in DTensor layout, i.e. m x n fat matrices (m <= n).
    Returns N (k, n, n), the projector N N' (k, n, n) and the ranks."""
    import torch
    k, n, m = a.shape
    assert m <= n
    at = transpose_batched(ctx, a)                       # n x m tall
    S, U, _, info = gesvd_batched(ctx, at, True)
    rank = count_gt_batched(ctx, S, eps)
    N = torch.empty((k, n, n), dtype=a.dtype, device=a.device)
    P = torch.empty((k, n, n), dtype=a.dtype, device=a.device)
    ctx.call("nullspace_build_batched", a, n, _p(U), n * n, _p(rank), _p(N), n * n, _p(P), n * n, k)
    return N, P, rank


def nullspace_project(ctx: Context, P, b):
    """b_i <- (N_i N_i') b_i in place: addAB with C aliasing B (ref: tensor.cuh:2081-2085)."""
    gemm_batched(ctx, b, P, b)


LOWER_ONLY = 1 << 32   # GPUB_LOWER_ONLY


def chol_solve_from_host(ctx: Context, A, b, info, A_host, b_host, x_host, info_host, chunks: int = 16, sidx: int = 0, lower_only: bool = False):
    """The product's host pipeline (gpub_chol_solve_from_host_*): upload / factorise + solve / download of successive chunks
    overlap on three streams. A, b, info are device tensors (k, n, n) / (k, 1, n) / (k,); the host tensors may be pinned."""
    k, n = A.shape[0], A.shape[1]
    hp = lambda t: _vp(t.data_ptr()) if t is not None else None
    ctx.call("chol_solve_from_host", A, n, _p(A), hp(b), hp(info), hp(A_host), hp(b_host), hp(x_host), hp(info_host), k,
             chunks | (LOWER_ONLY if lower_only else 0), sidx=sidx)
