"""Development aid for compute-sanitizer: a batched SVD large enough to be cut into sub-batches on the side streams, and a Nullspace build."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(6)
sms = torch.cuda.get_device_properties(0).multi_processor_count
A = rng.uniform(-1, 1, (sms + 8, 64, 64))
S, U, Vt, info = capi.gesvd_batched(ctx, capi.from_numpy_batch(A.copy()), True)
torch.cuda.synchronize()
ref = np.linalg.svd(A[-1], compute_uv=False)
assert np.abs(S.cpu().numpy()[-1] - ref).max() < 1e-12 and not info.cpu().numpy().any()
a = rng.uniform(-1, 1, (sms + 3, 64, 128)).transpose(0, 2, 1).copy()      # fat 64 x 128 matrices in DTensor layout (k, cols, rows)
N, P, rank = capi.nullspace_build(ctx, capi.from_numpy_batch(rng.uniform(-1, 1, (sms + 3, 64, 128))), 1e-8)
torch.cuda.synchronize()
print("done", int(rank[0]))
