"""GPU: the CUDA path, called through the C ABI (include/gputils_b200.h), against the CPU oracle on the same
seeded inputs. Tolerances (BASELINE.md / north_star): relative Frobenius error <= 1e-12 for fp64 and <= 1e-5
for fp32; exact equality where the arithmetic is exact (integer-valued data, copies, transposes, counts)."""
import numpy as np
import pytest

from conftest import TOL, rel_err

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]


def dev(a):
    from gputils_b200 import capi
    return capi.from_numpy_batch(a)


def host(t):
    from gputils_b200 import capi
    return capi.to_numpy_batch(t)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n,k,batch", [
    (8, 8, 8, 4096),          # BASELINE config 1
    (2, 2, 3, 3), (3, 2, 3, 5), (1, 1, 1, 7), (5, 7, 3, 33), (4, 4, 4, 1000), (16, 16, 16, 257), (32, 32, 32, 129),
    (7, 1, 7, 10),            # mat-vec, the example/main.cu shape class
    (33, 17, 9, 4), (64, 64, 64, 9), (128, 128, 128, 3), (100, 37, 51, 2), (64, 64, 64, 1), (200, 1, 300, 1),
])
def test_gemm_batched(gpu_ctx, oracle, dt, m, n, k, batch):
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(1000 + m + 7 * n + 13 * k)
    A = rng.uniform(-1, 1, (batch, m, k)).astype(dt); B = rng.uniform(-1, 1, (batch, k, n)).astype(dt)
    C0 = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
    dC = dev(C0)
    capi.gemm_batched(gpu_ctx, dC, dev(A), dev(B), 1.0, 0.0)
    assert rel_err(host(dC), oracle.gemm_batched(A, B)) <= TOL[np.dtype(dt)]
    dC = dev(C0)
    capi.gemm_batched(gpu_ctx, dC, dev(A), dev(B), -0.5, 2.0)
    assert rel_err(host(dC), oracle.gemm_batched(A, B, C0, -0.5, 2.0)) <= TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_exact_on_integer_data_and_empty_batch(gpu_ctx, golden, dt):
    import torch
    from conftest import mats, with_layout
    from gputils_b200 import capi
    g = golden["addAB"]
    A, B, C = (mats(with_layout(g, k), dt) for k in "ABC")
    dC = torch.zeros((3, 2, 2), dtype=dev(A).dtype, device="cuda")
    capi.gemm_batched(gpu_ctx, dC, dev(A), dev(B))
    assert np.array_equal(host(dC), C)
    empty = torch.zeros((0, 2, 2), dtype=dC.dtype, device="cuda")
    capi.gemm_batched(gpu_ctx, empty, torch.zeros((0, 3, 2), dtype=dC.dtype, device="cuda"), torch.zeros((0, 2, 3), dtype=dC.dtype, device="cuda"))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,batch", [(7, 1), (7, 64), (32, 100), (100, 3), (1024, 2)])
def test_gemm_output_aliasing_rhs(gpu_ctx, dt, n, batch):
    """Nullspace::project calls b.addAB(P, b): C aliases B (ref: tensor.cuh:2084)."""
    from gputils_b200 import capi
    rng = np.random.default_rng(n)
    P = rng.uniform(-1, 1, (batch, n, n)).astype(dt); b = rng.uniform(-1, 1, (batch, n, 1)).astype(dt)
    db = dev(b)
    capi.gemm_batched(gpu_ctx, db, dev(P), db)
    assert rel_err(host(db), P.astype(np.float64) @ b.astype(np.float64)) <= 10 * TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,batch", [(1, 5), (2, 9), (3, 2), (4, 1000), (5, 77), (8, 1000), (13, 50), (16, 513), (20, 33),
                                     (32, 1000), (32, 1), (33, 7), (64, 10), (100, 3), (128, 4), (200, 2),
                                     # more matrices than resident CTAs: every CTA of the persistent kernels walks several matrices
                                     (128, 2600), (96, 2500), (70, 2400), (48, 5000)])
def test_potrf_potrs(gpu_ctx, oracle, dt, n, batch):
    import torch
    from gputils_b200 import capi
    A = oracle.fill_spd_batched(n, batch, float(n), 0x5EED0002 + n, dt)
    b = oracle.fill_uniform(batch * n, -1.0, 1.0, 0x5EED0102 + n, dt).reshape(batch, n, 1)
    dA = dev(A); db = dev(b)
    info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    capi.potrf_batched(gpu_ctx, dA, info)
    L = host(dA)
    Lo, info_o = oracle.potrf_batched(A)
    assert np.array_equal(info.cpu().numpy(), info_o) and not info_o.any()
    assert np.array_equal(np.triu(L, 1), np.triu(A, 1)), "strict upper triangle must stay untouched"
    assert rel_err(np.tril(L), np.tril(Lo)) <= TOL[np.dtype(dt)]
    Lt = np.tril(L).astype(np.float64)
    assert rel_err(Lt @ Lt.transpose(0, 2, 1), A) <= TOL[np.dtype(dt)]
    capi.potrs_batched(gpu_ctx, dA, db)
    x = host(db)
    assert rel_err(x, oracle.potrs_batched(Lo, b)) <= 20 * TOL[np.dtype(dt)]
    assert rel_err(A.astype(np.float64) @ x, b) <= 20 * TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [3, 4, 8, 32, 40, 96, 128])
def test_potrf_reports_first_bad_pivot(gpu_ctx, oracle, dt, n):
    import torch
    from gputils_b200 import capi
    A = oracle.fill_spd_batched(n, 6, float(n), 11, dt)
    A[1, n // 2, n // 2] = -1.0          # leading minor n//2+1 not positive
    A[4] = 0.0                           # zero matrix: fails at pivot 1
    dA = dev(A)
    info = torch.zeros(6, dtype=torch.int32, device="cuda")
    capi.potrf_batched(gpu_ctx, dA, info)
    _, info_o = oracle.potrf_batched(A)
    assert info.cpu().tolist() == info_o.tolist() == [0, n // 2 + 1, 0, 0, 1, 0]
    good = [0, 2, 3, 5]
    Lo, _ = oracle.potrf_batched(A[good])
    assert rel_err(np.tril(host(dA)[good]), np.tril(Lo)) <= TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n,batch", [(2, 2, 3), (64, 16, 1000), (64, 16, 3), (32, 16, 65), (32, 8, 100), (20, 3, 5), (4, 3, 1),
                                       (100, 30, 4), (300, 40, 2), (16, 16, 20), (2000, 24, 2)])
def test_gels_batched(gpu_ctx, oracle, dt, m, n, batch):
    import torch
    from gputils_b200 import capi
    A = oracle.fill_uniform(batch * m * n, -1.0, 1.0, 0x5EED0003, dt).reshape(batch, n, m).transpose(0, 2, 1).copy()
    b = oracle.fill_uniform(batch * m, -1.0, 1.0, 0x5EED0103, dt).reshape(batch, m, 1)
    dA = dev(A); db = dev(b)
    info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    capi.gels_batched(gpu_ctx, dA, db, info)
    qr_o, xb_o, info_o = oracle.gels_batched(A, b)
    assert not info.cpu().numpy().any() and not info_o.any()
    tol = 50 * TOL[np.dtype(dt)]
    xb = host(db)
    assert rel_err(xb[:, :n], xb_o[:, :n]) <= tol           # the solution
    assert rel_err(xb, xb_o) <= tol                          # ... and the Q^T b tail the reference leaves behind
    assert rel_err(host(dA), qr_o) <= tol                    # A overwritten by the same QR factors
    x64 = xb[:, :n].astype(np.float64); A64 = A.astype(np.float64)
    grad = A64.transpose(0, 2, 1) @ (A64 @ x64 - b)          # normal equations
    assert np.linalg.norm(grad) <= tol * np.linalg.norm(A64) * np.linalg.norm(b)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n,batch", [(4, 3, 1), (20, 3, 1), (64, 16, 9), (128, 128, 2), (300, 20, 3), (1024, 128, 2),
                                       (512, 64, 150), (1000, 100, 3), (513, 37, 2), (257, 32, 5), (512, 48, 2), (300, 300, 1)])
def test_geqrf_ormqr_trsv(gpu_ctx, oracle, dt, m, n, batch):
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(m + n)
    A = rng.uniform(-100, 100, (batch, m, n)).astype(dt)
    b = rng.uniform(-1, 1, (batch, m, 1)).astype(dt)
    dA = dev(A)
    tau = torch.zeros((batch, n), dtype=dA.dtype, device="cuda")
    capi.geqrf_batched(gpu_ctx, dA, tau)
    qr_o, tau_o = oracle.geqrf_batched(A)
    tol = 100 * TOL[np.dtype(dt)]
    assert rel_err(host(dA), qr_o) <= tol and rel_err(tau.cpu().numpy(), tau_o) <= tol
    # Q from the reflectors: Q R = A  (QRFactoriser::getQR, tensor.cuh:1929-1995)
    eye = np.tile(np.eye(m, n, dtype=dt), (batch, 1, 1))
    dQ = dev(eye)
    capi.ormqr_batched(gpu_ctx, False, dA, tau, dQ)
    Q = host(dQ).astype(np.float64)
    R = np.triu(host(dA)[:, :n, :]).astype(np.float64)
    assert rel_err(Q @ R, A) <= tol
    assert np.abs(Q.transpose(0, 2, 1) @ Q - np.eye(n)).max() <= tol * 10
    # Q^T A = [R; 0] through the many-column path of ormqr (the blocked kernel for tall fp64 matrices)
    dQtA = dev(A)
    capi.ormqr_batched(gpu_ctx, True, dA, tau, dQtA)
    QtA = host(dQtA).astype(np.float64)
    Rfull = np.zeros_like(QtA); Rfull[:, :n, :] = R
    assert rel_err(QtA, Rfull) <= tol
    # least squares: Q^T b then R x = (Q^T b)[0:n]  (QRFactoriser::leastSquares, tensor.cuh:1891-1927)
    db = dev(b)
    capi.ormqr_batched(gpu_ctx, True, dA, tau, db)
    assert rel_err(host(db), oracle.ormqr_batched(True, qr_o, tau_o, b)) <= tol
    capi.trsv_upper_batched(gpu_ctx, dA, n, m, m * n, db, m, batch)
    x = host(db)[:, :n].astype(np.float64)
    xr = np.stack([np.linalg.lstsq(A[i].astype(np.float64), b[i].astype(np.float64), rcond=None)[0] for i in range(batch)])
    assert rel_err(x, xr) <= 1000 * TOL[np.dtype(dt)]


def _subspace_gap(U1, U2):
    return np.abs(U1 @ U1.T - U2 @ U2.T).max()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n,batch,want_u", [(3, 2, 3, True), (8, 3, 4, True), (4, 3, 3, False), (3, 3, 5, True), (64, 16, 40, True),
                                              (64, 32, 3, True), (200, 20, 3, True), (200, 20, 3, False), (500, 8, 2, True),
                                              (128, 64, 2, True), (256, 40, 3, False), (100, 33, 2, True), (300, 128, 1, True),
                                              (1024, 128, 2, False), (192, 96, 2, True), (256, 64, 5, False), (128, 128, 3, True)])
def test_gesvd_batched(gpu_ctx, oracle, dt, m, n, batch, want_u):
    from gputils_b200 import capi
    rng = np.random.default_rng(7 * m + n)
    A = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A), want_u)
    So, Uo, Vto = oracle.gesvd_batched(A.astype(np.float64), want_u)
    tol = 100 * TOL[np.dtype(dt)]
    assert not info.cpu().numpy().any()
    Sn = S.cpu().numpy().astype(np.float64)
    assert np.all(np.diff(Sn, axis=1) <= 0) and np.all(Sn >= 0)
    assert rel_err(Sn, So) <= tol
    Vn = host(Vt).astype(np.float64)
    assert np.abs(Vn @ Vn.transpose(0, 2, 1) - np.eye(n)).max() <= tol
    # singular vectors up to sign (singular values are distinct with probability one)
    assert np.abs(np.abs(Vn) - np.abs(Vto)).max() <= 1e4 * tol
    if want_u:
        Un = host(U).astype(np.float64)
        assert np.abs(Un.transpose(0, 2, 1) @ Un - np.eye(m)).max() <= tol
        assert rel_err(Un[:, :, :n] * Sn[:, None, :] @ Vn, A) <= tol
        for i in range(batch):   # orthogonal complement spans the same subspace as LAPACK's
            assert _subspace_gap(Un[i][:, n:], Uo[i][:, n:]) <= 1e3 * tol


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n", [(40, 12), (200, 48), (512, 128)])
def test_gesvd_rank_deficient_factors_stay_orthogonal(gpu_ctx, oracle, dt, m, n):
    """Rank-deficient input (duplicated columns, SURVEY.md 8d cfg4 variant): S has a numerically zero tail and
    both factors must still be complete orthogonal matrices -- Nullspace takes the trailing columns of U."""
    from gputils_b200 import capi
    rng = np.random.default_rng(m + n)
    A = rng.uniform(-1, 1, (2, m, n)).astype(dt)
    dup = n // 8 + 1
    A[:, :, n - dup:] = A[:, :, :dup]
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A), True)
    tol = 100 * TOL[np.dtype(dt)]
    Sn = S.cpu().numpy().astype(np.float64); Un = host(U).astype(np.float64); Vn = host(Vt).astype(np.float64)
    So, _, _ = oracle.gesvd_batched(A.astype(np.float64), False)
    assert np.abs(Sn - So).max() <= tol * So.max()
    assert np.all(Sn[:, n - dup:] <= 1e3 * tol * Sn[:, :1])
    assert np.abs(Un.transpose(0, 2, 1) @ Un - np.eye(m)).max() <= tol
    assert np.abs(Vn @ Vn.transpose(0, 2, 1) - np.eye(n)).max() <= tol
    assert rel_err(Un[:, :, :n] * Sn[:, None, :] @ Vn, A) <= tol
    # the trailing m - rank columns of U span the orthogonal complement of range(A)
    r = n - dup
    assert np.abs(Un[:, :, r:].transpose(0, 2, 1) @ A.astype(np.float64)).max() <= 1e3 * tol * So.max()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n", [(128, 64), (256, 128), (200, 20), (100, 33)])
def test_gesvd_graded_and_scaled_matrices(gpu_ctx, dt, m, n):
    """Singular values spread over many decades, and the same matrix scaled towards both ends of the exponent range: values relative to
    the largest, orthogonality and reconstruction must not notice the scale. Two guards are exercised: the Jacobi kernels bring the R
    factor to order one (squared norms and their products leave the range long before the entries do), and gesvd_batched scales a
    matrix whose largest entry is outside LAPACK's [sqrt(safmin) / eps, eps / sqrt(safmin)] before the QR step (k_prescale)."""
    from gputils_b200 import capi
    f64 = dt == np.float64
    rng = np.random.default_rng(m + n)
    Q, _ = np.linalg.qr(rng.normal(size=(m, n))); V, _ = np.linalg.qr(rng.normal(size=(n, n)))
    sig = np.logspace(0, -12 if f64 else -4, n)
    A0 = (Q * sig) @ V.T
    scales = (1.0, 1e100, 1e-100, 1e150, 1e-150) if f64 else (1.0, 1e10, 1e-10, 1e14, 1e-16)
    A = np.stack([s * A0 for s in scales] + [np.eye(m, n), np.zeros((m, n))]).astype(dt)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A.copy()), True)
    assert not info.cpu().numpy().any()
    Sn = S.cpu().numpy().astype(np.float64); Un = host(U).astype(np.float64); Vn = host(Vt).astype(np.float64)
    tol = 100 * TOL[np.dtype(dt)]
    for i, scale in enumerate(scales):
        es = np.abs(Sn[i] / scale - sig).max()
        assert es <= (1e-13 if f64 else 1e-5), (scale, es)
        assert rel_err((Un[i][:, :n] * Sn[i]) @ Vn[i], A[i].astype(np.float64)) <= tol, scale
    k = len(scales)
    assert np.abs(Sn[k] - 1.0).max() <= 10 * TOL[np.dtype(dt)] and np.abs(Sn[k + 1]).max() == 0.0
    for i in range(k + 2):
        eV, eU = np.abs(Vn[i] @ Vn[i].T - np.eye(n)).max(), np.abs(Un[i].T @ Un[i] - np.eye(m)).max()
        assert eV <= tol and eU <= tol, (i, eV, eU)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n", [(64, 16), (200, 24), (128, 64), (256, 128)])
def test_gesvd_and_qr_of_degenerate_matrices(gpu_ctx, dt, m, n):
    """All-ones (every column of the QR is rounding noise of the one before, down to subnormals: this used to give NaN in geqrf, an
    out-of-bounds read in the Jacobi tail and reflectors that were not orthogonal), rank one, rank one plus noise at 1e-14, duplicated
    columns, a permuted identity block and the zero matrix: finite results, singular values to eps of the largest, both factors
    orthogonal, reconstruction, info = 0; and the explicit Q of the all-ones QR is orthogonal."""
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(m * n)
    u = rng.normal(size=(m, 1)); v = rng.normal(size=(1, n))
    A = np.stack([np.ones((m, n)), u @ v, u @ v + 1e-14 * rng.normal(size=(m, n)), np.repeat(rng.normal(size=(m, n // 2)), 2, axis=1),
                  np.eye(m, n)[:, ::-1].copy(), np.zeros((m, n))]).astype(dt)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A.copy()), True)
    assert not info.cpu().numpy().any()
    Sn = S.cpu().numpy().astype(np.float64); Un = host(U).astype(np.float64); Vn = host(Vt).astype(np.float64)
    assert np.isfinite(Sn).all() and np.isfinite(Un).all() and np.isfinite(Vn).all()
    tol = 100 * TOL[np.dtype(dt)]
    A = A.astype(np.float64)
    for i in range(A.shape[0]):
        ref = np.linalg.svd(A[i], compute_uv=False)
        assert np.abs(Sn[i] - ref).max() <= 10 * TOL[np.dtype(dt)] * max(ref[0], 1.0), i
        assert np.abs(Un[i].T @ Un[i] - np.eye(m)).max() <= tol and np.abs(Vn[i] @ Vn[i].T - np.eye(n)).max() <= tol, i
        assert np.linalg.norm((Un[i][:, :n] * Sn[i]) @ Vn[i] - A[i]) <= tol * max(np.linalg.norm(A[i]), 1.0), i
    dA = dev(A[:1].astype(dt)); tau = torch.zeros((1, n), dtype=dA.dtype, device="cuda")
    capi.geqrf_batched(gpu_ctx, dA, tau)
    eye = dev(np.eye(m, dtype=dt)[None].copy())
    capi.ormqr_batched(gpu_ctx, False, dA, tau, eye)
    Q = host(eye)[0].astype(np.float64); R = np.triu(host(dA)[0][:n]).astype(np.float64)
    assert np.isfinite(Q).all() and np.abs(Q.T @ Q - np.eye(m)).max() <= tol and np.abs(Q[:, :n] @ R - A[0]).max() <= tol * m


@pytest.mark.parametrize("m,n", [(64, 16), (200, 24), (128, 64), (256, 128), (100, 40)])
def test_gesvd_nan_input_does_not_fault(gpu_ctx, m, n):
    """A NaN in the input poisons that matrix's result (as with LAPACK) but nothing else: no out-of-range index inside the kernels (the
    Jacobi tail ranks column norms into a permutation that indexes shared memory), the other matrices of the batch are untouched, and
    the context stays usable."""
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(m + 3 * n)
    A = rng.uniform(-1, 1, (3, m, n))
    A[1, m // 2, n // 3] = np.nan
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A.copy()), True)
    torch.cuda.synchronize()
    Sn = S.cpu().numpy()
    for i in (0, 2):
        assert np.abs(Sn[i] - np.linalg.svd(A[i], compute_uv=False)).max() <= 1e-12 * 10
    S2, _, _, info2 = capi.gesvd_batched(gpu_ctx, dev(A[:1].copy()), False)
    assert torch.equal(S2[0], S[0])


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n,want_u", [(64, 64, True), (128, 64, False), (192, 96, True)])
def test_gesvd_chunked_batch_equals_small_batches(gpu_ctx, dt, m, n, want_u):
    """A batch larger than the SM count is cut into sub-batches that run on the library's side streams (gesvd_batched): every
    matrix must get exactly what it gets in a batch small enough to run as one piece -- same kernels per matrix, so bit for bit --
    which pins the slicing of every operand and of the workspace, the fork and the join."""
    import torch
    from gputils_b200 import capi
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    batch = 2 * sms + 7
    rng = np.random.default_rng(5 * m + n)
    A = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A), want_u)
    nxt = (S.sum() + Vt.sum()).item()                    # queued behind the call on the same stream: must see every chunk
    assert np.isfinite(nxt) and not info.cpu().numpy().any()
    step = sms // 2
    for lo in range(0, batch, step):
        hi = min(batch, lo + step)
        S1, U1, Vt1, info1 = capi.gesvd_batched(gpu_ctx, dev(A[lo:hi]), want_u)
        assert torch.equal(S[lo:hi], S1) and torch.equal(Vt[lo:hi], Vt1) and not info1.cpu().numpy().any()
        if want_u:
            assert torch.equal(U[lo:hi], U1)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n", [(128, 64), (256, 128)])
def test_gesvd_clustered_singular_values(gpu_ctx, dt, m, n):
    """Multiple singular values (half of them 2, half 1, and the identity-like case of all equal): between columns of equal norm a
    tiny cosine still asks for a large rotation, the case the early-exit rule of k_jacobi_blk has to leave alone."""
    from gputils_b200 import capi
    rng = np.random.default_rng(m * n)
    Q, _ = np.linalg.qr(rng.normal(size=(m, n))); V, _ = np.linalg.qr(rng.normal(size=(n, n)))
    s1 = np.ones(n); s1[: n // 2] = 2.0
    A = np.stack([(Q * s1) @ V.T, Q @ V.T, (Q * np.linspace(3.0, 1.0, n)) @ V.T]).astype(dt)
    S, U, Vt, info = capi.gesvd_batched(gpu_ctx, dev(A), True)
    tol = 100 * TOL[np.dtype(dt)]
    assert not info.cpu().numpy().any()
    Sn = S.cpu().numpy().astype(np.float64); Un = host(U).astype(np.float64); Vn = host(Vt).astype(np.float64)
    So = np.stack([np.linalg.svd(A[i].astype(np.float64), compute_uv=False) for i in range(3)])
    assert np.abs(Sn - So).max() <= tol * So.max()
    assert np.abs(Vn @ Vn.transpose(0, 2, 1) - np.eye(n)).max() <= tol
    assert np.abs(Un.transpose(0, 2, 1) @ Un - np.eye(m)).max() <= tol
    assert rel_err(Un[:, :, :n] * Sn[:, None, :] @ Vn, A) <= tol


@pytest.mark.parametrize("dt", DTYPES)
def test_gesvd_rank_deficient_and_rank_count(gpu_ctx, oracle, golden, dt):
    import torch
    from conftest import mats, with_layout
    from gputils_b200 import capi
    A = mats(with_layout(golden["svd_rank"], "A"), dt)
    S, _, _, info = capi.gesvd_batched(gpu_ctx, dev(A), False)
    count = torch.zeros(3, dtype=torch.int32, device="cuda")
    eps = 1e-10 if dt == np.float64 else 1e-4
    gpu_ctx.call("count_gt_batched", S, capi._p(S), 3, 3, eps, capi._p(count), 3)
    assert count.cpu().tolist() == golden["svd_rank"]["rank"]
    gpu_ctx.call("count_gt_batched", S, capi._p(S), 3, 3, eps, capi._p(count), 3)   # accumulates, like the reference kernel
    assert count.cpu().tolist() == [2 * r for r in golden["svd_rank"]["rank"]]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [1, 5, 1000, 4097, 1_000_003])
def test_reductions_and_elementwise(gpu_ctx, oracle, dt, n):
    import torch
    from gputils_b200 import capi
    x = oracle.fill_uniform(n + 3, -1.0, 1.0, 21, dt); y = oracle.fill_uniform(n + 3, -1.0, 1.0, 22, dt)
    for off in (0, 1):                      # off = 1: a view that is not 16-byte aligned (slices do this)
        dx = torch.from_numpy(x).cuda()[off:off + n]; dy = torch.from_numpy(y).cuda()[off:off + n]
        xs, ys = x[off:off + n], y[off:off + n]
        tol = 10 * TOL[np.dtype(dt)]
        assert abs(capi.reduce_scalar(gpu_ctx, "nrm2", dx) - oracle.nrm2(xs)) <= tol * oracle.nrm2(xs)
        assert abs(capi.reduce_scalar(gpu_ctx, "asum", dx) - oracle.asum(xs)) <= tol * oracle.asum(xs)
        assert abs(capi.reduce_scalar(gpu_ctx, "dot", dx, dy) - oracle.dot(xs, ys)) <= tol * np.sqrt(n)
        v, i = capi.reduce_scalar(gpu_ctx, "amax_abs", dx)
        assert v == np.abs(xs).max() and i == int(np.abs(xs).argmax())
        v, i = capi.reduce_scalar(gpu_ctx, "amin_abs", dx)
        assert v == np.abs(xs).min() and i == int(np.abs(xs).argmin())
        dz = dy.clone()
        gpu_ctx.call("axpy", dz, n, -1.0, capi._p(dx), capi._p(dz))
        assert np.array_equal(dz.cpu().numpy(), ys - xs)
        gpu_ctx.call("scal", dz, n, 3.0, capi._p(dz))
        assert np.array_equal(dz.cpu().numpy(), dt(3.0) * (ys - xs))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n,batch", [(3, 2, 2), (1, 9, 4), (17, 33, 5), (64, 16, 100), (128, 1024, 3)])
def test_transpose_batched(gpu_ctx, dt, m, n, batch):
    from gputils_b200 import capi
    rng = np.random.default_rng(m * n)
    A = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
    At = capi.transpose_batched(gpu_ctx, dev(A))
    assert np.array_equal(host(At), A.transpose(0, 2, 1))


@pytest.mark.parametrize("dt", DTYPES)
def test_nullspace_pack_and_projector(gpu_ctx, dt):
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(5)
    n, batch = 6, 7
    U = rng.uniform(-1, 1, (batch, n, n)).astype(dt)
    rank = np.array([0, 1, 3, 6, 5, 2, 6], dtype=np.int32)
    dU = dev(U); dN = torch.empty_like(dU); dP = torch.empty_like(dU)
    dr = torch.from_numpy(rank).cuda()
    gpu_ctx.call("nullspace_pack_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(dN), n * n, batch)
    gpu_ctx.call("aat_batched", dU, n, capi._p(dN), n * n, capi._p(dP), n * n, batch)
    N = host(dN)
    for i in range(batch):
        nul = n - rank[i]
        assert np.array_equal(N[i][:, :nul], U[i][:, n - nul:]) and not N[i][:, nul:].any()
    assert rel_err(host(dP), N.astype(np.float64) @ N.astype(np.float64).transpose(0, 2, 1)) <= TOL[np.dtype(dt)]
    assert not host(dP)[3].any() and not host(dP)[6].any()          # full rank -> exactly zero projector


@pytest.mark.parametrize("n,batch", [(64, 5), (192, 3), (256, 2)])
def test_projector_on_the_tensor_pipe_is_symmetric_and_exact(gpu_ctx, n, batch):
    """N N^T for n a multiple of 64 (fp64) computes only the tiles on and below the diagonal and mirrors the rest."""
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(n)
    N = rng.uniform(-1, 1, (batch, n, n))
    N[:, :, n - n // 8:] = 0.0                        # the zero padding of a left-packed nullspace basis
    dN = dev(N); dP = torch.full_like(dN, float("nan"))
    gpu_ctx.call("aat_batched", dN, n, capi._p(dN), n * n, capi._p(dP), n * n, batch)
    P = host(dP)
    Nm = host(dN)
    assert np.array_equal(P, P.transpose(0, 2, 1)), "mirrored tiles must be bit-identical"
    assert rel_err(P, Nm @ Nm.transpose(0, 2, 1)) <= TOL[np.dtype(np.float64)]


def test_generators_match_the_numpy_mirror(gpu_ctx, oracle):
    import torch
    from gputils_b200 import capi
    x = torch.empty(10_001, dtype=torch.float64, device="cuda")
    capi.fill_uniform(gpu_ctx, x, -1.0, 1.0, 0x5EED0001)
    assert np.array_equal(x.cpu().numpy(), oracle.fill_uniform(10_001, -1.0, 1.0, 0x5EED0001))
    A = torch.empty((5, 8, 8), dtype=torch.float64, device="cuda")
    capi.fill_spd_batched(gpu_ctx, A, 8.0, 99)
    assert rel_err(host(A), oracle.fill_spd_batched(8, 5, 8.0, 99)) <= 1e-15


@pytest.mark.parametrize("n,batch,ld,pad", [(5, 11, 8, 8), (4, 9, 6, 4), (30, 7, 32, 0), (50, 5, 52, 12), (64, 3, 66, 2),
                                            (100, 4, 104, 16), (128, 3, 130, 6)])
def test_padded_strides_are_accepted(gpu_ctx, oracle, n, batch, ld, pad):
    """The C ABI takes explicit leading dimensions and batch strides (e.g. 128-byte aligned padded layouts): every size
    class of the Cholesky kernels (lane groups, warp per matrix, pipelined 32 x 32 blocks) factorises and solves in a padded
    buffer and leaves the padding untouched."""
    import torch
    from gputils_b200 import capi
    stride = ld * n + pad
    A = oracle.fill_spd_batched(n, batch, float(n), 3)
    b = oracle.fill_uniform(batch * n, -1.0, 1.0, 4).reshape(batch, n, 1)
    buf = torch.full((batch, stride), 777.0, dtype=torch.float64, device="cuda")
    for i in range(batch):
        buf[i, : ld * n].view(n, ld)[:, :n] = torch.from_numpy(A[i].T.copy()).cuda()
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    gpu_ctx.call("potrf_batched", buf, n, capi._p(buf), ld, stride, capi._p(info), batch)
    assert not info.cpu().numpy().any()
    Lo, _ = oracle.potrf_batched(A)
    out = buf.cpu().numpy()
    for i in range(batch):
        L = out[i, : ld * n].reshape(n, ld)[:, :n].T
        assert rel_err(np.tril(L), np.tril(Lo[i])) <= 1e-12
        assert np.array_equal(np.triu(L, 1), np.triu(A[i], 1))
        assert np.all(out[i, ld * n:] == 777.0) and np.all(out[i, : ld * n].reshape(n, ld)[:, n:] == 777.0)
    sb = n + 3
    rhs = torch.full((batch, sb), 555.0, dtype=torch.float64, device="cuda")
    rhs[:, :n] = torch.from_numpy(b[:, :, 0].copy()).cuda()
    gpu_ctx.call("potrs_batched", buf, n, capi._p(buf), ld, stride, capi._p(rhs), sb, batch)
    x = rhs.cpu().numpy()
    assert rel_err(x[:, :n, None], oracle.potrs_batched(Lo, b)) <= 2e-11
    assert np.all(x[:, n:] == 555.0)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [6, 64, 192])
def test_nullspace_projector_from_the_orthogonal_factor(gpu_ctx, dt, n):
    """gpub_nullspace_projector_batched: N N' computed as U2 U2' or as I - U1 U1', whichever side has fewer columns (rank read
    from the device, any value in 0..n, not a multiple of the panel depth) == gpub_aat_batched on the packed basis."""
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(n)
    ranks = np.array([0, 1, n // 2 - 1, n // 2, n // 2 + 1, n - 17 if n > 17 else 2, n - 1, n, 13 % n, 16 % n], dtype=np.int32)
    batch = len(ranks)
    U = np.linalg.qr(rng.uniform(-1, 1, (batch, n, n)))[0].astype(dt)
    dU = dev(U); dr = torch.from_numpy(ranks).cuda()
    dN = torch.empty_like(dU); dP = torch.full_like(dU, float("nan")); dP2 = torch.empty_like(dU)
    gpu_ctx.call("nullspace_pack_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(dN), n * n, batch)
    gpu_ctx.call("nullspace_projector_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(dN), n * n, capi._p(dP), n * n, batch)
    if dt == np.float64 and n % 64 == 0:
        # tensor-pipe path: both sides are read from U itself, the packed basis is not an input (gpub_nullspace_build packs beside it)
        dP3 = torch.empty_like(dU)
        gpu_ctx.call("nullspace_projector_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(torch.full_like(dU, float("nan"))), n * n,
                     capi._p(dP3), n * n, batch)
        assert torch.equal(dP3, dP)
    gpu_ctx.call("aat_batched", dU, n, capi._p(dN), n * n, capi._p(dP2), n * n, batch)
    P, P2, N = host(dP).astype(np.float64), host(dP2).astype(np.float64), host(dN).astype(np.float64)
    tol = 20 * TOL[np.dtype(dt)]
    assert np.abs(P - N @ N.transpose(0, 2, 1)).max() <= tol and np.abs(P - P2).max() <= tol
    assert np.array_equal(P, P.transpose(0, 2, 1))
    assert np.abs(P[0] - np.eye(n)).max() <= tol and np.abs(P[7]).max() <= tol     # rank 0: identity; full rank: zero


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [5, 64, 128])
def test_nullspace_build_equals_pack_then_projector(gpu_ctx, dt, n):
    """gpub_nullspace_build_batched (the packing beside the projector on a private stream, joined on the call's stream) gives the
    packed basis and the projector of the two separate calls, bit for bit, and work queued behind it on the call's stream sees both."""
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(3 * n)
    ranks = np.array([0, 1, n // 2, n // 2 + 1, n - 1, n, 3 % n, n - 2], dtype=np.int32)
    batch = len(ranks)
    U = np.linalg.qr(rng.uniform(-1, 1, (batch, n, n)))[0].astype(dt)
    dU = dev(U); dr = torch.from_numpy(ranks).cuda()
    dN = torch.empty_like(dU); dP = torch.empty_like(dU)
    gpu_ctx.call("nullspace_pack_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(dN), n * n, batch)
    gpu_ctx.call("nullspace_projector_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(dN), n * n, capi._p(dP), n * n, batch)
    for rep in range(3):
        dN2 = torch.full_like(dU, float("nan")); dP2 = torch.full_like(dU, float("nan"))
        gpu_ctx.call("nullspace_build_batched", dU, n, capi._p(dU), n * n, capi._p(dr), capi._p(dN2), n * n, capi._p(dP2), n * n, batch)
        s = dN2.sum() + dP2.sum()                  # queued on the same (current) stream right behind the call
        assert torch.equal(dN2, dN) and torch.equal(dP2, dP) and bool(torch.isfinite(s))


def test_nrm2_of_extreme_range_data_is_rescued_by_the_scaled_pass(gpu_ctx):
    """fp64 data whose squares overflow / underflow: the one-pass sum gives inf / 0, the scaled second pass the true norm."""
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(3)
    base = rng.uniform(0.5, 1.0, 100_003)
    for scale in (1e200, 1e-200, 1e160, 1e-170, 1.0):
        x = torch.from_numpy(base * scale).cuda()
        got = capi.reduce_scalar(gpu_ctx, "nrm2", x)
        want = float(np.linalg.norm(base)) * scale
        assert np.isfinite(got) and abs(got / want - 1) <= 1e-13, (scale, got, want)
    z = torch.zeros(1000, dtype=torch.float64, device="cuda")
    assert capi.reduce_scalar(gpu_ctx, "nrm2", z) == 0.0
    z[17] = float("inf")
    assert capi.reduce_scalar(gpu_ctx, "nrm2", z) == float("inf")


def test_pool_and_staged_copies_through_the_c_abi(gpu_ctx):
    """gpub_mem_alloc / free / stats and gpub_upload / gpub_download (pageable numpy memory, odd sizes, > 4 ring pieces)."""
    import ctypes as C
    lib, h = gpu_ctx.lib, gpu_ctx.h
    from gputils_b200 import capi
    rng = np.random.default_rng(11)
    for nbytes in (7, 300_001, (40 << 20) + 12345):
        src = rng.integers(0, 255, nbytes, dtype=np.uint8)
        dst = np.zeros_like(src)
        p = C.c_void_p()
        capi.check(lib.gpub_mem_alloc(h, nbytes, C.byref(p)), "gpub_mem_alloc")
        capi.check(lib.gpub_upload(h, 0, p, src.ctypes.data_as(C.c_void_p), nbytes), "gpub_upload")
        capi.check(lib.gpub_download(h, 0, dst.ctypes.data_as(C.c_void_p), p, nbytes), "gpub_download")
        assert np.array_equal(src, dst)
        reserved, used = C.c_size_t(), C.c_size_t()
        capi.check(lib.gpub_mem_stats(h, C.byref(reserved), C.byref(used)), "gpub_mem_stats")
        assert used.value >= nbytes and reserved.value >= used.value
        capi.check(lib.gpub_mem_free(p), "gpub_mem_free")
    assert lib.gpub_mem_free(C.c_void_p(12345)) != 0           # not a pool pointer: refused, not crashed


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,batch,pinned", [(32, 3001, True), (32, 3001, False), (16, 5, False), (64, 257, True)])
def test_host_pipeline_equals_potrf_potrs(gpu_ctx, oracle, dt, n, batch, pinned):
    """gpub_chol_solve_from_host_*: upload / factorise + solve / download of 7 chunks on three streams == the two plain launchers,
    bit for bit, from pinned and from pageable host memory; info reported for a matrix that is not positive definite."""
    import torch
    from gputils_b200 import capi
    tdt = torch.float64 if dt == np.float64 else torch.float32
    A = oracle.fill_spd_batched(n, batch, float(n), 21, dt); b = oracle.fill_uniform(batch * n, -1.0, 1.0, 22, dt).reshape(batch, n, 1)
    A[batch // 2, 0, 0] = -1.0
    dA = dev(A); db = dev(b); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    capi.potrf_batched(gpu_ctx, dA, info); capi.potrs_batched(gpu_ctx, dA, db)
    hA = torch.from_numpy(np.ascontiguousarray(A.transpose(0, 2, 1))); hb = torch.from_numpy(np.ascontiguousarray(b.transpose(0, 2, 1)))
    hx = torch.empty_like(hb); hi = torch.empty(batch, dtype=torch.int32)
    if pinned:
        hA, hb, hx, hi = hA.pin_memory(), hb.pin_memory(), hx.pin_memory(), hi.pin_memory()
    dA2 = torch.empty((batch, n, n), dtype=tdt, device="cuda"); db2 = torch.empty((batch, 1, n), dtype=tdt, device="cuda")
    info2 = torch.zeros(batch, dtype=torch.int32, device="cuda")
    capi.chol_solve_from_host(gpu_ctx, dA2, db2, info2, hA, hb, hx, hi, chunks=7)
    ok = (info.cpu() == 0)
    if pinned:
        # GPUB_LOWER_ONLY: the block wholly above the diagonal is not transferred (when its rows are >= 128 bytes) -- same factors and
        # solutions, and that block of the device tensor keeps what it held
        dA3 = torch.full((batch, n, n), 777.0, dtype=tdt, device="cuda"); db3 = torch.empty_like(db2); info3 = torch.zeros_like(info2)
        hx3 = torch.empty_like(hx); hi3 = torch.empty_like(hi)
        capi.chol_solve_from_host(gpu_ctx, dA3, db3, info3, hA, hb, hx3, hi3, chunks=5, lower_only=True)
        low3 = torch.triu(torch.ones(n, n)).bool()
        assert torch.equal(info3.cpu(), info2.cpu()) and torch.equal(hi3, hi)
        assert torch.equal(dA3.cpu()[ok][:, low3], dA2.cpu()[ok][:, low3]) and torch.equal(db3.cpu()[ok], db2.cpu()[ok]) and torch.equal(hx3[ok], hx[ok])
        skipped = (n // 2) * dA3.element_size() >= 128
        blk = dA3[:, n // 2:, : n // 2]                       # [mat][col >= n/2][row < n/2]
        assert bool((blk == 777.0).all()) == skipped
    assert torch.equal(info.cpu(), info2.cpu()) and torch.equal(info2.cpu(), hi) and int(hi[batch // 2]) == 1
    low = torch.triu(torch.ones(n, n)).bool()
    assert torch.equal(dA.cpu()[ok][:, low], dA2.cpu()[ok][:, low])
    assert torch.equal(db.cpu()[ok], db2.cpu()[ok]) and torch.equal(db2.cpu()[ok], hx[ok])


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,batch", [(32, 1000), (8, 77), (20, 33), (64, 40)])
def test_fused_solve_allgather_on_one_device(gpu_ctx, oracle, dt, n, batch):
    """gpub_potrs_allgather_batched_*: the solutions also land at (offset + i) * stride of every destination (here three buffers on
    the same device; two GPUs: tests/test_gpu_sharded.py), b itself is solved in place, nothing else is touched."""
    import ctypes as C
    import torch
    from gputils_b200 import capi
    A = oracle.fill_spd_batched(n, batch, float(n), 31, dt); b = oracle.fill_uniform(batch * n, -1.0, 1.0, 32, dt).reshape(batch, n, 1)
    dA = dev(A); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    capi.potrf_batched(gpu_ctx, dA, info)
    x_ref = dev(b); capi.potrs_batched(gpu_ctx, dA, x_ref)
    db = dev(b)
    offset, stride, total = 5, n + 2, batch + 9
    outs = [torch.full((total, stride), -7.0, dtype=db.dtype, device="cuda") for _ in range(3)]
    ptrs = (C.c_void_p * 3)(*[o.data_ptr() for o in outs])
    gpu_ctx.call("potrs_allgather_batched", db, n, capi._p(dA), n, n * n, capi._p(db), n, batch, ptrs, 3, offset, stride)
    # fp32 8 x 8 has a thread-per-matrix solve of its own: same solution up to rounding; every other shape runs the same kernel
    if n == 8 and dt == np.float32:
        assert rel_err(db.cpu().numpy(), x_ref.cpu().numpy()) <= 10 * TOL[np.dtype(dt)]
    else:
        assert torch.equal(db, x_ref)
    for o in outs:
        assert torch.equal(o[offset:offset + batch, :n], db[:, 0, :])
        assert bool((o[:offset] == -7.0).all()) and bool((o[offset + batch:] == -7.0).all()) and bool((o[:, n:] == -7.0).all())


@pytest.mark.parametrize("dt", DTYPES)
def test_batched_givens_through_the_c_abi(gpu_ctx, dt):
    import torch
    from gputils_b200 import capi
    rng = np.random.default_rng(9)
    m, n, batch = 7, 5, 129
    A = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
    dA = dev(A)
    gpu_ctx.call("givens_annihilate_batched", dA, capi._p(dA), m, n, m * n, 1, 4, 2, batch)
    got = host(dA).astype(np.float64)
    ref = A.astype(np.float64).copy()
    h = 1.0 / np.hypot(ref[:, 1, 2], ref[:, 4, 2]); c = ref[:, 1, 2] * h; s = ref[:, 4, 2] * h
    r1, r4 = ref[:, 1, :].copy(), ref[:, 4, :].copy()
    ref[:, 1, :] = c[:, None] * r1 + s[:, None] * r4
    ref[:, 4, :] = c[:, None] * r4 - s[:, None] * r1
    tol = 10 * TOL[np.dtype(dt)]
    assert np.abs(got - ref).max() <= tol and np.abs(got[:, 4, 2]).max() <= tol
    cs = torch.from_numpy(np.stack([np.cos(np.arange(batch) * 0.1), np.sin(np.arange(batch) * 0.1)]).astype(dt)).cuda()
    dB = dev(A)
    # columns 0 and 3 of every matrix: x = column 0 (stride 1), y = column 3
    gpu_ctx.call("rot_batched", dB, m, capi._p(dB), 1, C_void(dB, 3 * m), 1, m * n, capi._p(cs[0]), capi._p(cs[1]), batch)
    gb = host(dB).astype(np.float64); rb = A.astype(np.float64).copy()
    cc, ss = cs[0].cpu().numpy().astype(np.float64), cs[1].cpu().numpy().astype(np.float64)
    x, y = rb[:, :, 0].copy(), rb[:, :, 3].copy()
    rb[:, :, 0] = cc[:, None] * x + ss[:, None] * y
    rb[:, :, 3] = cc[:, None] * y - ss[:, None] * x
    assert np.abs(gb - rb).max() <= tol


def C_void(t, elem_offset):
    import ctypes as C
    return C.c_void_p(t.data_ptr() + elem_offset * t.element_size())
