"""Development aid: batched GEMM over many shapes and alpha / beta, transposes, reductions on awkward lengths, against numpy."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
dev = capi.from_numpy_batch; host = capi.to_numpy_batch
rng = np.random.default_rng(4)
bad = 0
def line(name, ok, **kv):
    global bad
    bad += 0 if ok else 1
    if not ok: print(name, " ".join(f"{k} {v:.1e}" if isinstance(v, float) else f"{k} {v}" for k, v in kv.items()), "  <-- BAD")
shapes = [(1, 1, 1), (2, 3, 4), (3, 3, 3), (4, 4, 4), (5, 7, 3), (8, 8, 8), (8, 1, 8), (1, 8, 8), (16, 16, 16), (17, 16, 15), (32, 32, 32), (33, 31, 35),
          (64, 64, 64), (64, 1, 64), (65, 63, 64), (128, 128, 128), (128, 64, 16), (100, 100, 100), (128, 1, 128), (130, 129, 131), (256, 256, 64),
          (1024, 128, 128), (128, 128, 1024), (1024, 1, 1024), (200, 300, 10), (7, 500, 9)]
for dt, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
    for (m, n, k) in shapes:
        for batch in (1, 3, 300 if m * n * k <= 32 ** 3 else 5):
            for (alpha, beta) in ((1.0, 0.0), (-1.0, 1.0), (2.5, -0.5), (0.0, 1.0), (0.0, 0.0)):
                A = rng.uniform(-1, 1, (batch, m, k)).astype(dt); B = rng.uniform(-1, 1, (batch, k, n)).astype(dt); C0 = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
                dC = dev(C0.copy())
                if beta == 0.0 and alpha != 0.0: dC.fill_(float("nan"))         # beta = 0 must not read C
                capi.gemm_batched(ctx, dC, dev(A), dev(B), alpha, beta)
                got = host(dC).astype(np.float64)
                ref = alpha * (A.astype(np.float64) @ B.astype(np.float64)) + (beta * C0.astype(np.float64) if beta != 0.0 else 0.0)
                err = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
                line(f"gemm {np.dtype(dt).name} {m}x{n}x{k} batch {batch} alpha {alpha} beta {beta}", np.isfinite(got).all() and err < tol * max(1, k ** 0.5), err=err)
    # C aliasing B (Nullspace::project): square A, n = 1
    for m in (4, 64, 100, 1024):
        P = rng.uniform(-1, 1, (3, m, m)).astype(dt); b = rng.uniform(-1, 1, (3, m, 1)).astype(dt)
        db = dev(b.copy())
        capi.gemm_batched(ctx, db, dev(P), db)
        err = float(np.abs(host(db).astype(np.float64) - P.astype(np.float64) @ b.astype(np.float64)).max())
        line(f"gemm alias {np.dtype(dt).name} {m}", err < tol * 100 * m ** 0.5, err=err)
    for (m, n) in ((1, 1), (1, 9), (9, 1), (3, 4), (31, 33), (32, 32), (64, 65), (128, 1024), (1024, 128), (1000, 3), (3, 1000)):
        for batch in (1, 5, 70):
            A = rng.uniform(-1, 1, (batch, m, n)).astype(dt)
            At = host(capi.transpose_batched(ctx, dev(A)))
            line(f"transpose {np.dtype(dt).name} {m}x{n} batch {batch}", np.array_equal(At, A.transpose(0, 2, 1)))
    for n in (1, 2, 3, 31, 32, 33, 255, 256, 257, 1023, 1025, 4095, 65537, 1_000_001, 3_333_333):
        x = rng.uniform(-1, 1, n).astype(dt); y = rng.uniform(-1, 1, n).astype(dt)
        dx = torch.from_numpy(x).cuda(); dy = torch.from_numpy(y).cuda()
        x64, y64 = x.astype(np.float64), y.astype(np.float64)
        rt = (1e-13 if dt == np.float64 else 1e-5)
        e = [abs(capi.reduce_scalar(ctx, "nrm2", dx) - np.linalg.norm(x64)) / np.linalg.norm(x64),
             abs(capi.reduce_scalar(ctx, "asum", dx) - np.abs(x64).sum()) / np.abs(x64).sum(),
             abs(capi.reduce_scalar(ctx, "dot", dx, dy) - x64 @ y64) / max(1.0, n ** 0.5)]
        mx, imx = capi.reduce_scalar(ctx, "amax_abs", dx); mn, imn = capi.reduce_scalar(ctx, "amin_abs", dx)
        ok = max(e) < rt and mx == np.abs(x).max() and mn == np.abs(x).min()
        line(f"reduce {np.dtype(dt).name} n={n}", ok, nrm2=float(e[0]), asum=float(e[1]), dot=float(e[2]), amax=float(mx), ref=float(np.abs(x).max()))
print("BAD cases:", bad)
