import numpy as np, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from gputils_b200 import capi
ctx = capi.Context(0)
def dev(a): return capi.from_numpy_batch(a)
def host(t): return capi.to_numpy_batch(t)
rng = np.random.default_rng(0)
for (m, n) in ((8, 3), (64, 16), (64, 32), (200, 20), (500, 8), (100, 33), (256, 64)):
    Q, _ = np.linalg.qr(rng.normal(size=(m, n))); V, _ = np.linalg.qr(rng.normal(size=(n, n)))
    sig = np.logspace(0, -10, n)
    A0 = (Q * sig) @ V.T
    for scale in (1.0, 1e100, 1e-100, 1e150, 1e-150):
        A = (scale * A0)[None]
        S, U, Vt, info = capi.gesvd_batched(ctx, dev(A.copy()), True)
        Sn = S.cpu().numpy()[0] / scale; Un = host(U)[0]; Vn = host(Vt)[0]
        es = np.abs(Sn - sig).max()
        rec = np.linalg.norm((Un[:, :n] * Sn) @ Vn - A0) / np.linalg.norm(A0)
        ou = np.abs(Un.T @ Un - np.eye(m)).max(); ov = np.abs(Vn @ Vn.T - np.eye(n)).max()
        flag = "" if (es < 1e-12 and rec < 1e-10 and ou < 1e-10 and ov < 1e-10) else "   <-- BAD"
        print(f"{m}x{n} scale {scale:g}: sigma err {es:.1e} recon {rec:.1e} orthU {ou:.1e} orthV {ov:.1e} info {int(info[0])}{flag}")
# Cholesky / QR at scale
for scale in (1e100, 1e-100):
    n = 32
    B = rng.normal(size=(4, n, n)); Aspd = (B @ B.transpose(0, 2, 1) + n * np.eye(n)) * scale
    dA = dev(Aspd.copy()); info = torch.zeros(4, dtype=torch.int32, device="cuda")
    capi.potrf_batched(ctx, dA, info)
    L = np.tril(host(dA)); err = np.linalg.norm(L @ L.transpose(0, 2, 1) - Aspd) / np.linalg.norm(Aspd)
    print(f"potrf scale {scale:g}: recon {err:.1e} info {info.cpu().tolist()}")
    m, n = 256, 32
    M = rng.normal(size=(3, m, n)) * scale
    dM = dev(M.copy()); tau = torch.zeros((3, n), dtype=torch.float64, device="cuda")
    capi.geqrf_batched(ctx, dM, tau)
    Rr = np.triu(host(dM)[:, :n, :]); Rref = np.stack([np.linalg.qr(M[i], mode="r") for i in range(3)])
    print(f"geqrf scale {scale:g}: |R| err {np.abs(np.abs(Rr) - np.abs(Rref)).max() / np.abs(Rref).max():.1e}")
