"""Development aid: the batched SVD at every n = 1..168 and several heights (path selection: thread-per-matrix, tall, Jacobi rt / blk)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
dev = capi.from_numpy_batch; host = capi.to_numpy_batch
rng = np.random.default_rng(8)
bad = 0
for dt, tol in ((np.float64, 1e-10), (np.float32, 2e-3)):
    nmax = 167 if dt == np.float64 else 200   # (fp64: the Jacobi tile of n = 168 no longer fits shared memory: GPUB_ENOTSUP)
    for n in list(range(1, 40)) + list(range(60, 70)) + list(range(92, 100)) + list(range(124, 132)) + [150, 160, nmax]:
        for m in sorted({n, n + 1, 2 * n, 300 if n <= 300 else n}):
            if m < n: continue
            for want_u in (True, False):
                A = rng.uniform(-1, 1, (2, m, n)).astype(dt)
                try:
                    S, U, Vt, info = capi.gesvd_batched(ctx, dev(A.copy()), want_u)
                except Exception as e:
                    bad += 1; print(f"{np.dtype(dt).name} {m}x{n} U={want_u}: raised {str(e)[:80]}  <-- BAD"); continue
                Sn = S.cpu().numpy().astype(np.float64); Vn = host(Vt).astype(np.float64)
                ref = np.stack([np.linalg.svd(A[i].astype(np.float64), compute_uv=False) for i in range(2)])
                e = [float(np.abs(Sn - ref).max() / ref.max()), float(max(np.abs(Vn[i] @ Vn[i].T - np.eye(n)).max() for i in range(2)))]
                if want_u:
                    Un = host(U).astype(np.float64)
                    e.append(float(max(np.abs(Un[i].T @ Un[i] - np.eye(m)).max() for i in range(2))))
                    e.append(float(max(np.abs((Un[i][:, :n] * Sn[i]) @ Vn[i] - A[i]).max() for i in range(2))))
                if not (max(e) < tol and not info.cpu().numpy().any()):
                    bad += 1; print(f"{np.dtype(dt).name} {m}x{n} U={want_u}: {['%.1e' % v for v in e]} info {int(info.abs().max())}  <-- BAD")
print("BAD cases:", bad)
