"""GPU: same-box parity against the reference ITSELF. oracle/_ref/libgputils_ref.so is the untouched reference
header (GPUtils include/tensor.cuh) built against cuBLAS / cuSOLVER (oracle/Makefile target `ref`, built in the
build container). The same device buffers go through the reference's public API and through the new C ABI;
tolerances are the ones north_star states: relative Frobenius <= 1e-12 (fp64), <= 1e-5 (fp32); SVD compared on
singular values and subspaces, not signs."""
import ctypes as C

import numpy as np
import pytest

from conftest import REPO, TOL, rel_err

pytestmark = pytest.mark.gpu
REF_LIB = REPO / "oracle" / "_ref" / "libgputils_ref.so"


@pytest.fixture(scope="module")
def ref():
    if not REF_LIB.exists():
        pytest.fail(f"{REF_LIB} missing: run `make -C oracle ref` in the build container (needs /root/reference)")
    return C.CDLL(str(REF_LIB))


def _p(t):
    return C.c_void_p(t.data_ptr())


def _suf(dt):
    return "f64" if dt == np.float64 else "f32"


def _ct(dt):
    return C.c_double if dt == np.float64 else C.c_float


def dev(a):
    from gputils_b200 import capi
    return capi.from_numpy_batch(a)


def host(t):
    from gputils_b200 import capi
    return capi.to_numpy_batch(t)


SZ = C.c_size_t


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,k,batch", [(8, 8, 8, 4096), (4, 4, 4, 999), (32, 32, 32, 300), (64, 64, 64, 16), (128, 128, 128, 4), (3, 5, 2, 7)])
def test_addAB_matches_cublas(gpu_ctx, ref, oracle, dt, m, n, k, batch):
    import torch
    from gputils_b200 import capi
    A = oracle.fill_uniform(batch * m * k, -1.0, 1.0, 0x5EED0001, dt).reshape(batch, k, m)
    B = oracle.fill_uniform(batch * k * n, -1.0, 1.0, 0x5EED0101, dt).reshape(batch, n, k)
    dA = torch.from_numpy(A).cuda(); dB = torch.from_numpy(B).cuda()
    C_ref = torch.zeros((batch, n, m), dtype=dA.dtype, device="cuda"); C_new = torch.zeros_like(C_ref)
    getattr(ref, f"ref_addAB_{_suf(dt)}")(SZ(m), SZ(n), SZ(k), SZ(batch), _p(dA), _p(dB), _p(C_ref), _ct(dt)(1.0), _ct(dt)(0.0), 1, None)
    capi.gemm_batched(gpu_ctx, C_new, dA, dB)
    assert rel_err(C_new.cpu().numpy(), C_ref.cpu().numpy()) <= TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,batch", [(3, 2), (4, 2000), (8, 2000), (16, 1000), (32, 5000), (64, 32), (128, 8)])
def test_cholesky_batch_matches_cusolver(gpu_ctx, ref, dt, n, batch):
    import torch
    from gputils_b200 import capi
    tdt = torch.float64 if dt == np.float64 else torch.float32
    A = torch.empty((batch, n, n), dtype=tdt, device="cuda"); b = torch.empty((batch, 1, n), dtype=tdt, device="cuda")
    capi.fill_spd_batched(gpu_ctx, A, float(n), 0x5EED0002)
    capi.fill_uniform(gpu_ctx, b, -1.0, 1.0, 0x5EED0102)
    L_ref = torch.empty_like(A); x_ref = torch.empty_like(b)
    info_ref = torch.zeros(batch, dtype=torch.int32, device="cuda")
    getattr(ref, f"ref_chol_batch_{_suf(dt)}")(SZ(n), SZ(batch), _p(A), _p(L_ref), _p(b), _p(x_ref), _p(info_ref), 1, None, None)
    L_new = A.clone(); x_new = b.clone()
    info_new = torch.zeros(batch, dtype=torch.int32, device="cuda")
    capi.potrf_batched(gpu_ctx, L_new, info_new)
    capi.potrs_batched(gpu_ctx, L_new, x_new)
    assert torch.equal(info_ref, info_new)
    # only the lower triangle is specified (hard part 8); tensors are (k, col, row)
    low = torch.triu(torch.ones(n, n, device="cuda")).bool()      # [col, row] with row >= col
    assert rel_err(L_new[:, low].cpu().numpy(), L_ref[:, low].cpu().numpy()) <= TOL[np.dtype(dt)]
    assert rel_err(x_new.cpu().numpy(), x_ref.cpu().numpy()) <= 50 * TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,batch", [(2, 2, 3), (64, 16, 4096), (32, 8, 100), (100, 30, 8)])
def test_least_squares_batched_matches_cublas(gpu_ctx, ref, dt, m, n, batch):
    import torch
    from gputils_b200 import capi
    tdt = torch.float64 if dt == np.float64 else torch.float32
    A = torch.empty((batch, n, m), dtype=tdt, device="cuda"); b = torch.empty((batch, 1, m), dtype=tdt, device="cuda")
    capi.fill_uniform(gpu_ctx, A, -1.0, 1.0, 0x5EED0003)
    capi.fill_uniform(gpu_ctx, b, -1.0, 1.0, 0x5EED0103)
    b_ref = torch.empty_like(b)
    getattr(ref, f"ref_gels_{_suf(dt)}")(SZ(m), SZ(n), SZ(batch), _p(A), None, _p(b), _p(b_ref), 1, None)
    A_new = A.clone(); b_new = b.clone()
    capi.gels_batched(gpu_ctx, A_new, b_new)
    # the solution x = b[0:n] is what the API specifies
    assert rel_err(b_new[:, :, :n].cpu().numpy(), b_ref[:, :, :n].cpu().numpy()) <= 100 * TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,batch", [(4, 3, 2), (20, 3, 3), (128, 32, 2), (1024, 128, 2)])
def test_qr_factoriser_matches_cusolver(gpu_ctx, ref, dt, m, n, batch):
    import torch
    from gputils_b200 import capi
    tdt = torch.float64 if dt == np.float64 else torch.float32
    A = torch.empty((batch, n, m), dtype=tdt, device="cuda"); b = torch.empty((batch, 1, m), dtype=tdt, device="cuda")
    capi.fill_uniform(gpu_ctx, A, -1.0, 1.0, 0x5EED0004)
    capi.fill_uniform(gpu_ctx, b, -1.0, 1.0, 0x5EED0104)
    QR_ref = torch.empty_like(A); x_ref = torch.empty_like(b)
    getattr(ref, f"ref_qr_{_suf(dt)}")(SZ(m), SZ(n), SZ(batch), _p(A), _p(QR_ref), _p(b), _p(x_ref), 1, None, None)
    QR_new = A.clone(); x_new = b.clone()
    tau = torch.zeros((batch, n), dtype=tdt, device="cuda")
    capi.geqrf_batched(gpu_ctx, QR_new, tau)
    capi.ormqr_batched(gpu_ctx, True, QR_new, tau, x_new)
    capi.trsv_upper_batched(gpu_ctx, QR_new, n, m, m * n, x_new, m, batch)
    tol = 200 * TOL[np.dtype(dt)]
    assert rel_err(QR_new.cpu().numpy(), QR_ref.cpu().numpy()) <= tol        # same LAPACK storage: R and reflectors
    assert rel_err(x_new[:, :, :n].cpu().numpy(), x_ref[:, :, :n].cpu().numpy()) <= 10 * tol


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,batch", [(3, 2, 3), (8, 3, 2), (64, 16, 20), (200, 20, 2)])
def test_svd_matches_cusolver_values_and_subspaces(gpu_ctx, ref, dt, m, n, batch):
    import torch
    from gputils_b200 import capi
    tdt = torch.float64 if dt == np.float64 else torch.float32
    A = torch.empty((batch, n, m), dtype=tdt, device="cuda")
    capi.fill_uniform(gpu_ctx, A, -1.0, 1.0, 0x5EED0005)
    S_ref = torch.empty((batch, n), dtype=tdt, device="cuda"); Vt_ref = torch.empty((batch, n, n), dtype=tdt, device="cuda")
    U_ref = torch.empty((batch, m, m), dtype=tdt, device="cuda")
    getattr(ref, f"ref_svd_{_suf(dt)}")(SZ(m), SZ(n), SZ(batch), _p(A), _p(S_ref), _p(Vt_ref), _p(U_ref), None, None, _ct(dt)(1e-6), 1, None)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, A.clone(), True)
    tol = 100 * TOL[np.dtype(dt)]
    assert rel_err(S.cpu().numpy(), S_ref.cpu().numpy()) <= tol
    Un, Ur = host(U).astype(np.float64), host(U_ref).astype(np.float64)
    Vn, Vr = host(Vt).astype(np.float64), host(Vt_ref).astype(np.float64)
    for i in range(batch):
        assert np.abs(np.abs(Vn[i]) - np.abs(Vr[i])).max() <= 1e4 * tol
        assert np.abs(Un[i][:, :n] @ Un[i][:, :n].T - Ur[i][:, :n] @ Ur[i][:, :n].T).max() <= 1e3 * tol
        assert np.abs(Un[i][:, n:] @ Un[i][:, n:].T - Ur[i][:, n:] @ Ur[i][:, n:].T).max() <= 1e3 * tol


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_reductions_match_cublas(gpu_ctx, ref, dt):
    import torch
    from gputils_b200 import capi
    tdt = torch.float64 if dt == np.float64 else torch.float32
    n = 3_000_017
    x = torch.empty(n, dtype=tdt, device="cuda"); y = torch.empty(n, dtype=tdt, device="cuda")
    capi.fill_uniform(gpu_ctx, x, -1.0, 1.0, 31); capi.fill_uniform(gpu_ctx, y, -1.0, 1.0, 32)
    out = (C.c_double * 5)()
    getattr(ref, f"ref_reductions_{_suf(dt)}")(SZ(n), _p(x), _p(y), out)
    tol = 10 * TOL[np.dtype(dt)]
    assert abs(capi.reduce_scalar(gpu_ctx, "nrm2", x) / out[0] - 1) <= tol
    assert abs(capi.reduce_scalar(gpu_ctx, "asum", x) / out[1] - 1) <= tol
    assert abs(capi.reduce_scalar(gpu_ctx, "dot", x, y) - out[2]) <= tol * np.sqrt(n) * 10
    assert capi.reduce_scalar(gpu_ctx, "amax_abs", x)[0] == out[3]
    assert capi.reduce_scalar(gpu_ctx, "amin_abs", x)[0] == out[4]
