"""Development aid: the all-ones matrix (every column rounding noise of the one before) through geqrf, the explicit Q and gesvd."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
m, n = int(sys.argv[1]), int(sys.argv[2])
A = np.ones((1, m, n))
dA = capi.from_numpy_batch(A.copy()); tau = torch.zeros((1, n), dtype=torch.float64, device="cuda")
capi.geqrf_batched(ctx, dA, tau); torch.cuda.synchronize()
QR = capi.to_numpy_batch(dA)[0]; R = np.triu(QR[:n]); t = tau.cpu().numpy()[0]
print("geqrf: nan", int(np.isnan(QR).sum()), "tau range", t.min(), t.max(), "zero taus", int((t == 0).sum()))
# H_j orthogonality: tau_j (1 + |v_j|^2) == 2
V = np.tril(QR, -1)[:, :n] + np.eye(m, n)
dev_h = np.abs(t * (V * V).sum(axis=0) - 2.0) * (t != 0)
print("  max |tau (v'v) - 2| over reflectors", dev_h.max(), "at", int(dev_h.argmax()), "alpha there", QR[int(dev_h.argmax()), int(dev_h.argmax())])
I = capi.from_numpy_batch(np.eye(m)[None].copy())
capi.ormqr_batched(ctx, False, dA, tau, I); torch.cuda.synchronize()
Q = capi.to_numpy_batch(I)[0]
print("  explicit Q: |Q'Q - I|", np.abs(Q.T @ Q - np.eye(m)).max(), "|QR - A|", np.abs(Q[:, :n] @ R - A[0]).max())
S, U, Vt, info = capi.gesvd_batched(ctx, capi.from_numpy_batch(A.copy()), True); torch.cuda.synchronize()
Un = capi.to_numpy_batch(U)[0]
print("gesvd: S[:3]", S.cpu().numpy()[0][:3], "info", int(info[0]), "|U'U - I|", np.abs(Un.T @ Un - np.eye(m)).max(), " of the first n columns", np.abs(Un[:, :n].T @ Un[:, :n] - np.eye(n)).max())
