"""Development aid: special inputs through the C ABI against numpy (rank-1 / zero / repeated columns, scaled least squares, scaled SPD
systems). Prints one line per case; '<-- BAD' marks anything outside the tolerances of the parity tests."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
dev = capi.from_numpy_batch; host = capi.to_numpy_batch
rng = np.random.default_rng(1)
bad = 0
def line(name, ok, **kv):
    global bad
    bad += 0 if ok else 1
    print(name, " ".join(f"{k} {v:.1e}" if isinstance(v, float) else f"{k} {v}" for k, v in kv.items()), "" if ok else "  <-- BAD")

# ---- SVD of special matrices
for (m, n) in ((64, 16), (200, 24), (128, 64), (256, 128)):
    u = rng.normal(size=(m, 1)); v = rng.normal(size=(1, n))
    cases = {"rank1": u @ v, "zero": np.zeros((m, n)), "ones": np.ones((m, n)), "dupcols": np.repeat(rng.normal(size=(m, n // 2)), 2, axis=1),
             "onehot": np.eye(m, n)[:, ::-1].copy(), "tinynoise": u @ v + 1e-14 * rng.normal(size=(m, n))}
    for name, A in cases.items():
        S, U, Vt, info = capi.gesvd_batched(ctx, dev(A[None].copy()), True)
        Sn = S.cpu().numpy()[0]; Un = host(U)[0]; Vn = host(Vt)[0]
        ref = np.linalg.svd(A, compute_uv=False)
        sc = max(ref[0], 1e-300)
        es = float(np.abs(Sn - ref).max() / sc)
        rec = float(np.linalg.norm((Un[:, :n] * Sn) @ Vn - A) / max(np.linalg.norm(A), 1e-300))
        ou = float(np.abs(Un.T @ Un - np.eye(m)).max()); ov = float(np.abs(Vn @ Vn.T - np.eye(n)).max())
        ok = np.isfinite(Sn).all() and es < 1e-12 and (rec < 1e-10 or np.linalg.norm(A) == 0) and ou < 1e-10 and ov < 1e-10 and int(info[0]) == 0
        line(f"svd {m}x{n} {name}", ok, sigma=es, recon=rec, orthU=ou, orthV=ov, info=int(info[0]))

# ---- least squares at scale (fp64 and fp32)
for dt, scales in ((np.float64, (1.0, 1e120, 1e-120)), (np.float32, (1.0, 1e12, 1e-12))):
    for (m, n) in ((64, 16), (200, 8)):
        A0 = rng.normal(size=(6, m, n)); b0 = rng.normal(size=(6, m, 1))
        for sc in scales:
            A = (A0 * sc).astype(dt); b = b0.astype(dt)
            dA = dev(A.copy()); db = dev(b.copy())
            capi.gels_batched(ctx, dA, db)
            x = host(db)[:, :n, 0].astype(np.float64)
            xr = np.stack([np.linalg.lstsq(A[i].astype(np.float64), b[i].astype(np.float64), rcond=None)[0][:, 0] for i in range(6)])
            err = float(np.abs(x - xr).max() / np.abs(xr).max())
            line(f"gels {np.dtype(dt).name} {m}x{n} scale {sc:g}", np.isfinite(x).all() and err < (1e-10 if dt == np.float64 else 2e-3), err=err)

# ---- SPD systems at scale
for dt, scales in ((np.float64, (1.0, 1e200, 1e-200)), (np.float32, (1.0, 1e25, 1e-25))):
    for n in (4, 8, 32, 64, 128):
        B = rng.normal(size=(5, n, n)); A0 = B @ B.transpose(0, 2, 1) + n * np.eye(n); b0 = rng.normal(size=(5, n, 1))
        for sc in scales:
            A = (A0 * sc).astype(dt); b = (b0 * sc).astype(dt)
            dA = dev(A.copy()); db = dev(b.copy()); info = torch.zeros(5, dtype=torch.int32, device="cuda")
            capi.potrf_batched(ctx, dA, info); capi.potrs_batched(ctx, dA, db)
            x = host(db)[:, :, 0].astype(np.float64)
            xr = np.linalg.solve(A0, b0)[:, :, 0]
            err = float(np.abs(x - xr).max() / np.abs(xr).max())
            line(f"chol {np.dtype(dt).name} n={n} scale {sc:g}", np.isfinite(x).all() and err < (1e-10 if dt == np.float64 else 5e-3) and not info.cpu().numpy().any(), err=err)
print("BAD cases:", bad)
