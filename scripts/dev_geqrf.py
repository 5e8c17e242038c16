"""Development aid: geqrf_batched against scipy's LAPACK geqrf on a few shapes (prints relative errors)."""
import sys
import numpy as np, torch
from scipy.linalg import lapack
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context()
rng = np.random.default_rng(0)
shapes = [(1024, 128, 3), (1000, 100, 2), (513, 38, 2), (513, 37, 2), (257, 32, 2), (512, 48, 2), (300, 300, 1), (1024, 16, 2), (640, 130, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in sys.argv[1].split(","))]
for (m, n, batch) in shapes:
    A = rng.uniform(-1, 1, (batch, m, n))
    dA = capi.from_numpy_batch(A)
    tau = torch.zeros((batch, min(m, n)), dtype=torch.float64, device="cuda")
    capi.geqrf_batched(ctx, dA, tau)
    torch.cuda.synchronize()
    QR = capi.to_numpy_batch(dA); T = tau.cpu().numpy()
    e1 = e2 = 0.0
    for i in range(batch):
        qr, t, _, info = lapack.dgeqrf(A[i])
        e1 = max(e1, np.linalg.norm(QR[i] - qr) / np.linalg.norm(qr)); e2 = max(e2, np.linalg.norm(T[i] - t) / np.linalg.norm(t))
    print((m, n, batch), "QR", e1, "tau", e2, flush=True)
