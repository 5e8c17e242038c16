"""CPU: the oracle (oracle/oracle.c + oracle/oracle_np.py) against every golden vector the reference's own
tests hold for the path (tests/golden/reference_vectors.json <- test/testTensor.cu). This is what pins the
oracle (prompt section 3 / SURVEY.md 8c)."""
import numpy as np
import pytest

from conftest import mats, with_layout

HI = 1e-10  # PRECISION_HIGH of the reference tests (testTensor.cu:6)


def test_gemm_addAB_exact(golden, oracle):
    g = golden["addAB"]
    A, B, C = (mats(with_layout(g, k)) for k in "ABC")
    assert np.array_equal(oracle.gemm_batched(A, B), C)          # EXPECT_EQ in the reference (exact)
    assert np.array_equal(oracle.gemm_batched(A.astype(np.float32), B.astype(np.float32)), C.astype(np.float32))
    assert np.array_equal(oracle.gemm_batched(A, B, use_c=False), C)


def test_reductions(golden, oracle):
    A = np.asarray(golden["data_234A"]["data"], dtype=np.float64)
    B = np.asarray(golden["data_234B"]["data"], dtype=np.float64)
    AMB = np.asarray(golden["data_234AMB"]["data"], dtype=np.float64)
    r = golden["reductions"]
    assert oracle.dot(A, B) == r["dotF_A_B"]
    assert abs(oracle.nrm2(A) - r["normF_A"]) < HI
    assert abs(oracle.asum(A) - r["sumAbs_A"]) < HI
    assert np.abs(AMB).max() == r["maxAbs_AMB"] and np.abs(AMB).min() == r["minAbs_AMB"]
    assert np.array_equal(A - B, AMB)


def test_cholesky(golden, oracle):
    g = golden["cholesky"]
    A = mats(with_layout(g, "A"))
    L, info = oracle.potrf_batched(A)
    assert info[0] == 0
    assert abs(L[0, 0, 0] - g["L00"]) < HI and abs(L[0, 2, 1] - g["L21"]) < HI and abs(L[0, 2, 2] - g["L22"]) < HI
    assert np.abs(np.tril(L[0]) - np.asarray(g["L_rowmajor"]).reshape(3, 3)).max() < HI
    assert np.array_equal(np.triu(L[0], 1), np.triu(A[0], 1))     # strict upper triangle untouched
    x = oracle.potrs_batched(L, np.asarray(g["b"]).reshape(1, 3, 1))
    assert np.abs(x.ravel() - np.asarray(g["x"])).max() < HI
    # pre-factorised path of choleskyBatchSolve (testTensor.cu:1377-1397)
    Lgiven = np.asarray(g["L_rowmajor"]).reshape(1, 3, 3)
    x2 = oracle.potrs_batched(Lgiven, np.asarray(g["b"]).reshape(1, 3, 1))
    assert np.abs(x2.ravel() - np.asarray(g["x"])).max() < HI


def test_cholesky_not_positive_definite(oracle):
    A = np.array([[[4.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]]])   # second leading minor singular
    _, info = oracle.potrf_batched(A)
    assert info[0] == 2


def test_qr_least_squares(golden, oracle):
    g = golden["qr_least_squares"]
    A = mats(with_layout(g, "A"))
    b = np.asarray(g["b"]).reshape(1, 4, 1)
    _, xb, info = oracle.gels_batched(A, b)
    assert info[0] == 0
    res = np.linalg.norm(A[0] @ xb[0, :3] - b[0])
    assert abs(res - g["residual_norm"]) < HI
    # same through geqrf + Q^T b + trsv, the QRFactoriser route (tensor.cuh:1891-1927)
    qr, tau = oracle.geqrf_batched(A)
    qtb = oracle.ormqr_batched(True, qr, tau, b)
    x = np.linalg.solve(np.triu(qr[0, :3, :]), qtb[0, :3])
    assert abs(np.linalg.norm(A[0] @ x - b[0]) - g["residual_norm"]) < HI


def test_least_squares_1(golden, oracle):
    g = golden["least_squares_1"]
    A, b = mats(with_layout(g, "A")), mats(with_layout(g, "b"))
    _, xb, _ = oracle.gels_batched(A, b)
    assert np.linalg.norm(A @ xb[:, :2] - b) < HI


def test_svd_singular_values(golden, oracle):
    g = golden["svd_singular_values"]
    S, U, Vt = oracle.gesvd_batched(mats(with_layout(g, "B")))
    assert abs(S[0, 0] - g["S0"]) < HI and abs(S[0, 1] - g["S1"]) < HI


def test_svd_multiple_signs_and_nullspace_basis(golden, oracle):
    g = golden["svd_multiple"]
    A = mats(with_layout(g, "A"))
    S, U, Vt = oracle.gesvd_batched(A)
    assert np.abs(S.ravel() - np.asarray(g["S"])).max() < HI
    # U is (3,3,3) column-major in the reference's download order
    u_cm = np.ascontiguousarray(U.transpose(0, 2, 1)).ravel()
    assert np.abs(u_cm - np.asarray(g["U"])).max() < HI
    vt_cm = np.ascontiguousarray(Vt.transpose(0, 2, 1)).ravel()
    assert np.abs(vt_cm[:4] - np.asarray(g["Vt_first4"])).max() < HI


def test_svd_rank(golden, oracle):
    g = golden["svd_rank"]
    S, _, _ = oracle.gesvd_batched(mats(with_layout(g, "A")), want_u=False)
    assert list(oracle.rank_batched(S, HI)) == g["rank"]


def test_nullspace_properties(golden, oracle):
    A = mats(with_layout(golden["nullspace_tensor"], "A"))
    N, P, r = oracle.nullspace_batched(A)
    assert np.abs(A @ N).max() < HI
    NtN = N.transpose(0, 2, 1) @ N
    for i in range(A.shape[0]):
        assert abs(NtN[i, 0, 0] - 1) < HI
        assert np.abs(NtN[i] - np.diag(np.diag(NtN[i]))).max() < HI
    At = mats(with_layout(golden["nullspace_trivial"], "A"))
    Nt, _, _ = oracle.nullspace_batched(At)
    assert np.linalg.norm(Nt) == 0.0


def test_nullspace_projection(golden, oracle):
    g = golden["nullspace_project"]
    A = mats(with_layout(g, "A"))
    N, P, _ = oracle.nullspace_batched(A)
    x = np.asarray(g["x"], dtype=np.float64)
    proj = P[0] @ x
    assert np.linalg.norm(A[0] @ proj) < HI
    y = N[0] @ np.asarray(g["other"], dtype=np.float64)
    assert (y - proj) @ (proj - x) < HI


def test_transpose(golden, oracle):
    g = golden["transpose"]
    A, At = mats(with_layout(g, "A")), mats(with_layout(g, "At"))
    assert np.array_equal(A.transpose(0, 2, 1), At)


def test_generators_are_deterministic(oracle):
    a = oracle.fill_uniform(1000, -1.0, 1.0, 0x5EED0001)
    assert np.array_equal(a, oracle.fill_uniform(1000, -1.0, 1.0, 0x5EED0001))
    assert a.min() >= -1 and a.max() < 1 and abs(a.mean()) < 0.1
    S = oracle.fill_spd_batched(8, 3, 8.0, 7)
    assert np.all(np.linalg.eigvalsh(S) > 0)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_host_lapack_baseline_agrees_with_the_oracle(oracle, dt):
    """oracle/cpu_lapack.c (OpenBLAS LAPACK under an OpenMP loop: bench.py's cpu_baseline) computes what the oracle port computes:
    Cholesky factor + solve, GEMM, least squares, QR and singular values on small seeded batches."""
    import cpu_lapack as cl
    tol = 1e-12 if dt == np.float64 else 2e-5
    k, n = 64, 32
    A = oracle.fill_spd_batched(n, k, float(n), 5, dt)
    b = oracle.fill_uniform(k * n, -1.0, 1.0, 6, dt).reshape(k, n, 1)
    L_o, info_o = oracle.potrf_batched(A)
    x_o = oracle.potrs_batched(L_o, b)
    Ac = np.ascontiguousarray(A.transpose(0, 2, 1)); bc = np.ascontiguousarray(b.transpose(0, 2, 1)); info = np.ones(k, np.int32)
    assert cl.chol_batch(Ac, bc, info, threads=2) > 0 and not info.any()
    assert np.linalg.norm(np.tril(Ac.transpose(0, 2, 1)) - np.tril(L_o)) <= tol * np.linalg.norm(L_o)
    assert np.linalg.norm(bc.transpose(0, 2, 1) - x_o) <= 50 * tol * np.linalg.norm(x_o)
    B = oracle.fill_uniform(k * n * n, -1.0, 1.0, 7, dt).reshape(k, n, n)
    Cc = np.zeros((k, n, n), dt)
    assert cl.gemm_batch(np.ascontiguousarray(A.transpose(0, 2, 1)), np.ascontiguousarray(B.transpose(0, 2, 1)), Cc) > 0
    ref = oracle.gemm_batched(A, B)
    assert np.linalg.norm(Cc.transpose(0, 2, 1) - ref) <= tol * np.linalg.norm(ref)
    m, nn = 64, 16
    T = oracle.fill_uniform(k * m * nn, -1.0, 1.0, 8, dt).reshape(k, m, nn); r = oracle.fill_uniform(k * m, -1.0, 1.0, 9, dt).reshape(k, m, 1)
    _, xb_o, _ = oracle.gels_batched(T, r)
    Tc = np.ascontiguousarray(T.transpose(0, 2, 1)); rc = np.ascontiguousarray(r.transpose(0, 2, 1))
    assert cl.gels_batch(Tc, rc) > 0
    assert np.linalg.norm(rc[:, 0, :nn] - xb_o[:, :nn, 0]) <= 100 * tol * np.linalg.norm(xb_o[:, :nn])
    Tc = np.ascontiguousarray(T.transpose(0, 2, 1)); tau = np.zeros((k, nn), dt)
    assert cl.geqrf_batch(Tc, tau) > 0
    qr_o, tau_o = oracle.geqrf_batched(T)
    assert np.linalg.norm(Tc.transpose(0, 2, 1) - qr_o) <= 100 * tol * np.linalg.norm(qr_o) and np.linalg.norm(tau - tau_o) <= 100 * tol * np.linalg.norm(tau_o)
    Tc = np.ascontiguousarray(T.transpose(0, 2, 1)); S = np.zeros((k, nn), dt); Vt = np.zeros((k, nn, nn), dt)
    assert cl.gesvd_batch(Tc, S, None, Vt) > 0
    assert np.linalg.norm(S - np.linalg.svd(T.astype(np.float64), compute_uv=False)) <= 100 * tol * np.linalg.norm(S)
