"""Development aid: every size n = 1..130 through batched Cholesky factorise + solve and the square batched GEMM, two batch sizes."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
dev = capi.from_numpy_batch; host = capi.to_numpy_batch
rng = np.random.default_rng(7)
bad = 0
for dt, tol in ((np.float64, 1e-11), (np.float32, 3e-4)):
    for n in range(1, 131):
        for batch in (1, 37):
            B = rng.normal(size=(batch, n, n)); A = (B @ B.transpose(0, 2, 1) + n * np.eye(n)).astype(dt); b = rng.normal(size=(batch, n, 1)).astype(dt)
            dA = dev(A.copy()); db = dev(b.copy()); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
            capi.potrf_batched(ctx, dA, info); capi.potrs_batched(ctx, dA, db)
            x = host(db).astype(np.float64); L = np.tril(host(dA)).astype(np.float64)
            e1 = float(np.linalg.norm(L @ L.transpose(0, 2, 1) - A) / np.linalg.norm(A))
            e2 = float(np.linalg.norm(A.astype(np.float64) @ x - b) / np.linalg.norm(b))
            up_ok = np.array_equal(np.triu(host(dA), 1), np.triu(A, 1))           # the strict upper triangle is not referenced
            M1 = rng.uniform(-1, 1, (batch, n, n)).astype(dt); M2 = rng.uniform(-1, 1, (batch, n, n)).astype(dt)
            dC = dev(np.zeros((batch, n, n), dtype=dt)); capi.gemm_batched(ctx, dC, dev(M1), dev(M2))
            e3 = float(np.abs(host(dC).astype(np.float64) - M1.astype(np.float64) @ M2.astype(np.float64)).max())
            ok = e1 < tol and e2 < tol * 10 and up_ok and not info.cpu().numpy().any() and e3 < tol * 10 * n ** 0.5
            if not ok:
                bad += 1
                print(f"{np.dtype(dt).name} n={n} batch {batch}: LLt {e1:.1e} resid {e2:.1e} upper_untouched {up_ok} gemm {e3:.1e} info {int(info.abs().max())}  <-- BAD")
print("BAD cases:", bad)
