"""Development aid: more special inputs (fp32 degenerate SVDs, Nullspace of degenerate matrices, batch sizes around the SM count)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
dev = capi.from_numpy_batch; host = capi.to_numpy_batch
rng = np.random.default_rng(2)
bad = 0
def line(name, ok, **kv):
    global bad
    bad += 0 if ok else 1
    print(name, " ".join(f"{k} {v:.1e}" if isinstance(v, float) else f"{k} {v}" for k, v in kv.items()), "" if ok else "  <-- BAD")
for dt, tol in ((np.float32, 1e-3), (np.float64, 1e-10)):
    for (m, n) in ((64, 16), (200, 24), (128, 64), (256, 128), (128, 128), (96, 96)):
        u = rng.normal(size=(m, 1)); v = rng.normal(size=(1, n))
        cases = {"ones": np.ones((m, n)), "rank1": u @ v, "zero": np.zeros((m, n)), "dup": np.repeat(rng.normal(size=(m, n // 2)), 2, axis=1),
                 "random": rng.normal(size=(m, n))}
        for name, A in cases.items():
            A = A.astype(dt)
            S, U, Vt, info = capi.gesvd_batched(ctx, dev(A[None].copy()), True)
            Sn = S.cpu().numpy()[0].astype(np.float64); Un = host(U)[0].astype(np.float64); Vn = host(Vt)[0].astype(np.float64)
            ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
            es = float(np.abs(Sn - ref).max() / max(ref[0], 1e-30))
            rec = float(np.linalg.norm((Un[:, :n] * Sn) @ Vn - A) / max(np.linalg.norm(A), 1e-30))
            ou = float(np.abs(Un.T @ Un - np.eye(m)).max()); ov = float(np.abs(Vn @ Vn.T - np.eye(n)).max())
            ok = np.isfinite(Sn).all() and np.isfinite(Un).all() and es < tol and (rec < tol or not A.any()) and ou < tol and ov < tol and int(info[0]) == 0
            line(f"svd {np.dtype(dt).name} {m}x{n} {name}", ok, sigma=es, recon=rec, orthU=ou, orthV=ov, info=int(info[0]))
# Nullspace of degenerate fat matrices
for (m, n) in ((3, 7), (16, 64), (64, 128), (128, 256)):
    for name, a in {"zero": np.zeros((m, n)), "ones": np.ones((m, n)), "random": rng.normal(size=(m, n)), "rank2": rng.normal(size=(m, 2)) @ rng.normal(size=(2, n))}.items():
        N, P, rank = capi.nullspace_build(ctx, dev(a[None].copy()), 1e-8)
        Nn = host(N)[0]; Pn = host(P)[0]; r = int(rank[0]); rr = int(np.linalg.matrix_rank(a, tol=1e-8))
        e1 = float(np.abs(a @ Nn).max() / max(np.abs(a).max(), 1.0))
        e2 = float(np.abs(Pn - Nn @ Nn.T).max()); e3 = float(np.abs(Pn @ Pn - Pn).max())
        dim = int(round(np.trace(Pn)))
        ok = r == rr and e1 < 1e-10 and e2 < 1e-10 and e3 < 1e-10 and dim == n - rr and np.isfinite(Pn).all()
        line(f"nullspace {m}x{n} {name}", ok, rank=r, ref_rank=rr, aN=e1, P_NNt=e2, idem=e3, dim=dim)
# batch sizes around the SM count
sms = torch.cuda.get_device_properties(0).multi_processor_count
for batch in (1, 7, sms, sms + 1, sms + 8, 2 * sms - 1, 2 * sms + 1):
    A = rng.normal(size=(batch, 128, 64))
    S, U, Vt, info = capi.gesvd_batched(ctx, dev(A.copy()), True)
    Sn = S.cpu().numpy()
    ref = np.stack([np.linalg.svd(A[i], compute_uv=False) for i in range(batch)])
    es = float(np.abs(Sn - ref).max())
    Un = host(U); ou = float(max(np.abs(Un[i].T @ Un[i] - np.eye(128)).max() for i in (0, batch - 1, batch // 2)))
    line(f"svd batch {batch}", es < 1e-12 and ou < 1e-10 and not info.cpu().numpy().any(), sigma=es, orthU=ou)
print("BAD cases:", bad)
