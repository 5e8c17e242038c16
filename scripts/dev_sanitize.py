"""Small invocations of every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python scripts/dev_sanitize.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context()
rng = np.random.default_rng(0)
def spd(n, k, dt):
    G = rng.uniform(-1, 1, (k, n, n)); return (G @ G.transpose(0, 2, 1) + n * np.eye(n)).astype(dt)
for dt in (np.float64, np.float32):
    for n, k in [(4, 300), (8, 40), (16, 20), (32, 9), (64, 5), (100, 3), (128, 3)]:
        A = capi.from_numpy_batch(spd(n, k, dt)); b = capi.from_numpy_batch(rng.uniform(-1, 1, (k, n, 1)).astype(dt))
        info = torch.zeros(k, dtype=torch.int32, device="cuda")
        capi.potrf_batched(ctx, A, info)
        # GPUB_SANITIZE_SKIP_QUAD=1: synccheck stops at k_potrs_quad128 (role-dependent arrival at a named barrier)
        if not (n > 64 and os.environ.get("GPUB_SANITIZE_SKIP_QUAD") == "1"):
            capi.potrs_batched(ctx, A, b)
    for (m, n, kk, k) in [(8, 8, 8, 50), (32, 32, 32, 9), (64, 64, 64, 3), (128, 128, 128, 2), (256, 1, 256, 3), (100, 1, 70, 3)]:
        A = capi.from_numpy_batch(rng.uniform(-1, 1, (k, m, kk)).astype(dt)); B = capi.from_numpy_batch(rng.uniform(-1, 1, (k, kk, n)).astype(dt))
        Cm = torch.zeros((k, n, m), dtype=A.dtype, device="cuda"); capi.gemm_batched(ctx, Cm, A, B)
    A = capi.from_numpy_batch(rng.uniform(-1, 1, (69, 64, 16)).astype(dt)); b = capi.from_numpy_batch(rng.uniform(-1, 1, (69, 64, 1)).astype(dt))
    capi.gels_batched(ctx, A, b)
for (m, n, k) in [(1024, 128, 2), (300, 40, 2), (513, 38, 1)]:
    A = capi.from_numpy_batch(rng.uniform(-1, 1, (k, m, n))); tau = torch.zeros((k, n), dtype=torch.float64, device="cuda")
    capi.geqrf_batched(ctx, A, tau)
    Cq = capi.from_numpy_batch(rng.uniform(-1, 1, (k, m, 48)))
    capi.ormqr_batched(ctx, False, A, tau, Cq); capi.ormqr_batched(ctx, True, A, tau, Cq)
S, U, Vt, info = capi.gesvd_batched(ctx, capi.from_numpy_batch(rng.uniform(-1, 1, (2, 256, 64))), True)
S, U, Vt, info = capi.gesvd_batched(ctx, capi.from_numpy_batch(rng.uniform(-1, 1, (2, 200, 100)).astype(np.float32)), False)
a = capi.from_numpy_batch(rng.uniform(-1, 1, (3, 64, 192)))                 # fat 64 x 192: Nullspace through the block-reflector U assembly
N, P, rank = capi.nullspace_build(ctx, a)
bb = capi.from_numpy_batch(rng.uniform(-1, 1, (3, 192, 1))); capi.nullspace_project(ctx, P, bb)
G = capi.from_numpy_batch(rng.uniform(-1, 1, (70, 6, 5)))
ctx.call("givens_annihilate_batched", G, capi._p(G), 6, 5, 30, 0, 5, 0, 70)
torch.cuda.synchronize()
print("done")
