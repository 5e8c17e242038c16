"""CPU: the LAPACK-faithful small-SVD core (gputils_b200/csrc/svd_small.cuh, __host__ __device__) compiled with
the host compiler and compared with LAPACK's ?gesvd through scipy and with the reference's sign-pinned golden
vectors (testTensor.cu:1126-1171). The shipped library runs the same code inside CUDA kernels; this harness
exists so the host logic is covered without a GPU."""
import ctypes as C
import subprocess

import numpy as np
import pytest
import scipy.linalg

from conftest import REPO, mats, with_layout


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("svdhost") / "svd_host.so"
    src = REPO / "tests" / "host_harness" / "svd_small_host.cpp"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    return C.CDLL(str(out))


def run(harness, A, want_u=True):
    m, n = A.shape
    dt = A.dtype
    a = np.asfortranarray(A).copy(order="F")
    S = np.zeros(n, dt); U = np.zeros((m, m), dt, order="F"); Vt = np.zeros((n, n), dt, order="F")
    scr = np.zeros(2 * n * n + 16 * n + 16, dt)
    fn = harness.harness_gesvd_small_f64 if dt == np.float64 else harness.harness_gesvd_small_f32
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    info = fn(m, n, p(a), p(S), p(U), p(Vt), 1 if want_u else 0, p(scr))
    return info, S, U, Vt


def test_golden_signs_and_nullspace_basis(harness, golden):
    g = golden["svd_multiple"]
    A = mats(with_layout(g, "A"))
    for i in range(3):
        info, S, U, Vt = run(harness, A[i])
        assert info == 0
        assert np.abs(S - np.asarray(g["S"][2 * i:2 * i + 2])).max() < 1e-10
        assert np.abs(U.ravel(order="F") - np.asarray(g["U"][9 * i:9 * i + 9])).max() < 1e-10
    _, _, _, Vt0 = run(harness, A[0])
    assert np.abs(Vt0.ravel(order="F") - np.asarray(g["Vt_first4"])).max() < 1e-10


@pytest.mark.parametrize("m,n", [(2, 2), (3, 2), (8, 3), (4, 3), (7, 3), (5, 5), (16, 8), (32, 16), (40, 32), (64, 20)])
def test_matches_lapack_gesvd_including_signs(harness, m, n):
    rng = np.random.default_rng(m * 100 + n)
    for t in range(10):
        A = rng.uniform(-1, 1, (m, n))
        info, S, U, Vt = run(harness, A)
        assert info == 0
        u, s, vt = scipy.linalg.svd(A, lapack_driver="gesvd")
        assert np.abs(S - s).max() < 1e-12 * max(1, s[0])
        assert np.abs(U[:, :n] * S @ Vt - A).max() < 1e-12
        assert np.abs(U.T @ U - np.eye(m)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(n)).max() < 1e-12
        # well separated singular values: vectors agree with LAPACK's, signs included
        if np.min(np.abs(np.diff(s))) > 1e-3 and s[-1] > 1e-3 and m >= int(1.6 * n):
            assert np.abs(vt - Vt).max() < 1e-9
            assert np.abs(u[:, :n] - U[:, :n]).max() < 1e-9


def test_rank_deficient_and_zero(harness):
    rng = np.random.default_rng(3)
    A = rng.uniform(-1, 1, (9, 4)); A[:, 3] = A[:, 0] - 2 * A[:, 1]
    info, S, U, Vt = run(harness, A)
    assert info == 0 and S[3] < 1e-14 * S[0]
    assert np.abs(U.T @ U - np.eye(9)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(4)).max() < 1e-12
    assert np.abs(U[:, :4] * S @ Vt - A).max() < 1e-12
    info, S, U, Vt = run(harness, np.zeros((4, 3)))
    assert info == 0 and np.all(S == 0) and np.array_equal(U, np.eye(4)) and np.array_equal(Vt, np.eye(3))


def test_float32(harness):
    rng = np.random.default_rng(4)
    A = rng.uniform(-1, 1, (12, 5)).astype(np.float32)
    info, S, U, Vt = run(harness, A)
    s = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    assert info == 0 and np.abs(S - s).max() < 1e-5
    assert np.abs(U[:, :5] * S @ Vt - A).max() < 1e-5
