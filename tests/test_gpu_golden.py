"""GPU: the reference's own golden vectors (tests/golden/reference_vectors.json <- test/testTensor.cu) through the
C ABI, with the reference's tolerances (1e-10 fp64, 1e-4 fp32; testTensor.cu:5-6)."""
import numpy as np
import pytest

from conftest import mats, with_layout

pytestmark = pytest.mark.gpu
CASES = [(np.float64, 1e-10), (np.float32, 1e-4)]


def dev(a):
    from gputils_b200 import capi
    return capi.from_numpy_batch(a)


def host(t):
    from gputils_b200 import capi
    return capi.to_numpy_batch(t)


@pytest.mark.parametrize("dt,eps", CASES)
def test_cholesky_golden(gpu_ctx, golden, dt, eps):
    import torch
    from gputils_b200 import capi
    g = golden["cholesky"]
    A = np.tile(mats(with_layout(g, "A"), dt), (2, 1, 1))       # k = 2 like choleskyBatchFactorisation
    dA = dev(A)
    info = torch.zeros(2, dtype=torch.int32, device="cuda")
    capi.potrf_batched(gpu_ctx, dA, info)
    L = host(dA)
    for k in range(2):
        assert abs(L[k, 0, 0] - g["L00"]) < eps and abs(L[k, 2, 1] - g["L21"]) < eps and abs(L[k, 2, 2] - g["L22"]) < eps
    assert info.cpu().tolist() == [0, 0]
    b = np.tile(np.asarray(g["b"], dtype=dt).reshape(1, 3, 1), (2, 1, 1))
    db = dev(b)
    capi.potrs_batched(gpu_ctx, dA, db)
    assert np.abs(host(db).reshape(2, 3) - np.asarray(g["x"])).max() < eps
    # factor provided by the caller (choleskyBatchSolve, testTensor.cu:1364-1408)
    Lgiven = np.tile(np.asarray(g["L_rowmajor"], dtype=dt).reshape(1, 3, 3), (2, 1, 1))
    db = dev(b)
    capi.potrs_batched(gpu_ctx, dev(Lgiven), db)
    assert np.abs(host(db).reshape(2, 3) - np.asarray(g["x"])).max() < eps


@pytest.mark.parametrize("dt,eps", CASES)
def test_qr_least_squares_golden(gpu_ctx, golden, dt, eps):
    import torch
    from gputils_b200 import capi
    g = golden["qr_least_squares"]
    A = mats(with_layout(g, "A"), dt)
    b = np.asarray(g["b"], dtype=dt).reshape(1, 4, 1)
    dA = dev(A); db = dev(b)
    tau = torch.zeros((1, 3), dtype=dA.dtype, device="cuda")
    capi.geqrf_batched(gpu_ctx, dA, tau)
    capi.ormqr_batched(gpu_ctx, True, dA, tau, db)
    capi.trsv_upper_batched(gpu_ctx, dA, 3, 4, 12, db, 4, 1)
    x = host(db)[0, :3].astype(np.float64)
    res = np.linalg.norm(A[0].astype(np.float64) @ x - b[0])
    assert abs(res - g["residual_norm"]) < eps * 10
    # and through the fused batched path
    dA = dev(A); db = dev(b)
    capi.gels_batched(gpu_ctx, dA, db)
    x = host(db)[0, :3].astype(np.float64)
    assert abs(np.linalg.norm(A[0].astype(np.float64) @ x - b[0]) - g["residual_norm"]) < eps * 10


@pytest.mark.parametrize("dt,eps", CASES)
def test_svd_golden_signs(gpu_ctx, golden, dt, eps):
    from gputils_b200 import capi
    g = golden["svd_multiple"]
    A = mats(with_layout(g, "A"), dt)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A), True)
    eps = eps * (10 if dt == np.float32 else 1)          # the reference gives float SVD 10x slack (testTensor.cu:1169)
    assert np.abs(S.cpu().numpy().ravel() - np.asarray(g["S"])).max() < eps
    assert np.abs(U.cpu().numpy().ravel() - np.asarray(g["U"])).max() < eps       # device order == download order
    assert np.abs(Vt.cpu().numpy().ravel()[:4] - np.asarray(g["Vt_first4"])).max() < eps
    g2 = golden["svd_singular_values"]
    S, _, _, _ = capi.gesvd_batched(gpu_ctx, dev(mats(with_layout(g2, "B"), dt)), True)
    assert abs(float(S[0, 0]) - g2["S0"]) < eps and abs(float(S[0, 1]) - g2["S1"]) < eps


@pytest.mark.parametrize("dt,eps", CASES)
def test_reductions_golden(gpu_ctx, golden, dt, eps):
    import torch
    from gputils_b200 import capi
    A = torch.tensor(golden["data_234A"]["data"], dtype=torch.float64).to(torch.float64 if dt == np.float64 else torch.float32).cuda()
    B = torch.tensor(golden["data_234B"]["data"], dtype=A.dtype).cuda()
    AMB = torch.tensor(golden["data_234AMB"]["data"], dtype=A.dtype).cuda()
    r = golden["reductions"]
    assert capi.reduce_scalar(gpu_ctx, "dot", A, B) == r["dotF_A_B"]
    assert abs(capi.reduce_scalar(gpu_ctx, "nrm2", A) - r["normF_A"]) < eps
    assert abs(capi.reduce_scalar(gpu_ctx, "asum", A) - r["sumAbs_A"]) < 1e-10
    assert capi.reduce_scalar(gpu_ctx, "amax_abs", AMB)[0] == r["maxAbs_AMB"]
    assert capi.reduce_scalar(gpu_ctx, "amin_abs", AMB)[0] == r["minAbs_AMB"]
