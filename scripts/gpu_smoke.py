"""Quick on-GPU sanity run of every launcher against numpy (development aid; the real tests are tests/)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi

ctx = capi.Context()
rng = np.random.default_rng(0)
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
for dt, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
    for (m, n, k, batch) in [(8, 8, 8, 4096), (3, 2, 3, 5), (32, 32, 32, 100), (64, 64, 64, 10), (128, 128, 128, 3), (100, 37, 51, 7), (1024, 1, 1024, 4)]:
        A = rng.uniform(-1, 1, (batch, m, k)).astype(dt); B = rng.uniform(-1, 1, (batch, k, n)).astype(dt)
        dA, dB = capi.from_numpy_batch(A), capi.from_numpy_batch(B)
        dC = torch.zeros((batch, n, m), dtype=dA.dtype, device="cuda")
        capi.gemm_batched(ctx, dC, dA, dB)
        print("gemm", dt.__name__, (m, n, k, batch), rel(capi.to_numpy_batch(dC), A @ B))
    for n, batch in [(3, 2), (4, 1000), (8, 1000), (16, 500), (32, 1000), (20, 33), (64, 10), (128, 4), (200, 2)]:
        G = rng.uniform(-1, 1, (batch, n, n)); A = (G @ G.transpose(0, 2, 1) + n * np.eye(n)).astype(dt)
        b = rng.uniform(-1, 1, (batch, n, 1)).astype(dt)
        dA = capi.from_numpy_batch(A); db = capi.from_numpy_batch(b)
        info = torch.full((batch,), -7, dtype=torch.int32, device="cuda")
        capi.potrf_batched(ctx, dA, info)
        L = np.tril(capi.to_numpy_batch(dA))
        capi.potrs_batched(ctx, dA, db)
        x = capi.to_numpy_batch(db)
        print("chol", dt.__name__, (n, batch), "LLt", rel(L @ L.transpose(0, 2, 1), A), "solve", rel(A @ x, b), "info", int(info.abs().max()))
    for (m, n, batch) in [(2, 2, 3), (64, 16, 1000), (32, 8, 100), (20, 3, 5), (100, 30, 4), (300, 40, 2)]:
        A = rng.uniform(-1, 1, (batch, m, n)).astype(dt); b = rng.uniform(-1, 1, (batch, m, 1)).astype(dt)
        dA = capi.from_numpy_batch(A); db = capi.from_numpy_batch(b)
        capi.gels_batched(ctx, dA, db)
        x = capi.to_numpy_batch(db)[:, :n]
        xr = np.stack([np.linalg.lstsq(A[i].astype(np.float64), b[i].astype(np.float64), rcond=None)[0] for i in range(batch)])
        print("gels", dt.__name__, (m, n, batch), rel(x, xr))
    for (m, n, batch, wu) in [(3, 2, 3, True), (8, 3, 2, True), (64, 16, 50, True), (200, 20, 3, True), (200, 20, 3, False)]:
        A = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
        dA = capi.from_numpy_batch(A)
        S, U, Vt, info = capi.gesvd_batched(ctx, dA, wu)
        torch.cuda.synchronize()
        Sn = S.cpu().numpy(); Vtn = capi.to_numpy_batch(Vt)
        sref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
        msg = f"S {rel(Sn, sref):.2e}"
        if wu:
            Un = capi.to_numpy_batch(U)
            rec = Un[:, :, :n] * Sn[:, None, :] @ Vtn
            msg += f" rec {rel(rec, A):.2e} orth {np.abs(Un.transpose(0,2,1) @ Un - np.eye(m)).max():.2e}"
        print("svd", dt.__name__, (m, n, batch, wu), msg, "info", int(info.abs().max()))
    x = torch.from_numpy(rng.uniform(-1, 1, 1_000_003).astype(dt)).cuda(); y = torch.from_numpy(rng.uniform(-1, 1, 1_000_003).astype(dt)).cuda()
    xn, yn = x.cpu().numpy().astype(np.float64), y.cpu().numpy().astype(np.float64)
    print("nrm2", capi.reduce_scalar(ctx, "nrm2", x) / np.linalg.norm(xn) - 1, "asum", capi.reduce_scalar(ctx, "asum", x) / np.abs(xn).sum() - 1,
          "dot", capi.reduce_scalar(ctx, "dot", x, y) / (xn @ yn) - 1, "amax", capi.reduce_scalar(ctx, "amax_abs", x), np.abs(xn).max(), int(np.abs(xn).argmax()),
          "amin", capi.reduce_scalar(ctx, "amin_abs", x), np.abs(xn).min(), int(np.abs(xn).argmin()))
print("OK")
