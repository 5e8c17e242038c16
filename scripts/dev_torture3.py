"""Development aid: Cholesky status codes on non-SPD inputs against LAPACK, QR / least squares at the shapes where the kernel selection
changes, against numpy."""
import sys
import numpy as np, torch, scipy.linalg
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
dev = capi.from_numpy_batch; host = capi.to_numpy_batch
rng = np.random.default_rng(3)
bad = 0
def line(name, ok, **kv):
    global bad
    bad += 0 if ok else 1
    if not ok: print(name, " ".join(f"{k} {v:.1e}" if isinstance(v, float) else f"{k} {v}" for k, v in kv.items()), "  <-- BAD")
# ---- Cholesky info codes
for dt in (np.float64, np.float32):
    for n in (3, 4, 5, 8, 13, 16, 31, 32, 33, 48, 64, 65, 100, 128, 130):
        B = rng.normal(size=(n, n)); spd = B @ B.T + n * np.eye(n)
        mats = {"spd": spd, "zero": np.zeros((n, n)), "ones": np.ones((n, n)), "negdiag": spd - 3 * n * np.diag((np.arange(n) == n // 2).astype(float)) * 10,
                "lastneg": spd.copy(), "semidef": (lambda C: C @ C.T)(rng.normal(size=(n, max(1, n // 2))))}
        mats["lastneg"][n - 1, n - 1] = -1.0
        A = np.stack(list(mats.values())).astype(dt)
        dA = dev(A.copy()); info = torch.zeros(A.shape[0], dtype=torch.int32, device="cuda")
        capi.potrf_batched(ctx, dA, info)
        got = info.cpu().numpy()
        L = np.tril(host(dA)).astype(np.float64)
        for i, name in enumerate(mats):
            _, ref = scipy.linalg.lapack.dpotrf(A[i].astype(np.float64), lower=1) if dt == np.float64 else scipy.linalg.lapack.spotrf(A[i], lower=1)
            ok = int(got[i]) == int(ref) or (name == "semidef")      # a semidefinite matrix breaks down where rounding says
            if ref == 0:
                err = float(np.linalg.norm(L[i] @ L[i].T - A[i]) / np.linalg.norm(A[i]))
                ok = ok and err < (1e-12 if dt == np.float64 else 1e-5)
            line(f"potrf {np.dtype(dt).name} n={n} {name}", ok, info=int(got[i]), lapack=int(ref))
# ---- QR shapes
for dt in (np.float64, np.float32):
    tol = 1e-11 if dt == np.float64 else 2e-4
    for m in (17, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1024, 1025, 1500):
        for n in (3, 15, 16, 17, 31, 32, 33, 64, 100, 128, 130):
            if n > m: continue
            A = rng.normal(size=(2, m, n)).astype(dt)
            dA = dev(A.copy()); tau = torch.zeros((2, n), dtype=dA.dtype, device="cuda")
            capi.geqrf_batched(ctx, dA, tau)
            R = np.triu(host(dA)[:, :n, :]).astype(np.float64)
            Rref = np.stack([np.linalg.qr(A[i].astype(np.float64), mode="r") for i in range(2)])
            e1 = float(np.abs(np.abs(R) - np.abs(Rref)).max() / np.abs(Rref).max())
            b = rng.normal(size=(2, m, 1)).astype(dt); db = dev(b.copy())
            capi.ormqr_batched(ctx, True, dA, tau, db)
            qtb = host(db)[:, :, 0].astype(np.float64)
            e2 = float(abs(np.linalg.norm(qtb) - np.linalg.norm(b)) / np.linalg.norm(b))
            line(f"qr {np.dtype(dt).name} {m}x{n}", e1 < tol and e2 < tol, R=e1, normQtb=e2)
# ---- least squares shapes
for dt in (np.float64, np.float32):
    tol = 1e-9 if dt == np.float64 else 5e-3
    for (m, n) in ((4, 4), (8, 3), (33, 7), (64, 16), (64, 64), (100, 10), (200, 32), (300, 40), (1000, 20), (2000, 8)):
        A = rng.normal(size=(3, m, n)).astype(dt); b = rng.normal(size=(3, m, 1)).astype(dt)
        dA = dev(A.copy()); db = dev(b.copy())
        capi.gels_batched(ctx, dA, db)
        x = host(db)[:, :n, 0].astype(np.float64)
        xr = np.stack([np.linalg.lstsq(A[i].astype(np.float64), b[i].astype(np.float64), rcond=None)[0][:, 0] for i in range(3)])
        e = float(np.abs(x - xr).max() / np.abs(xr).max())
        line(f"gels {np.dtype(dt).name} {m}x{n}", e < tol, err=e)
print("BAD cases:", bad)
