"""Development aid: per-phase cycle totals of k_potrf_blk (library built with -DGPUB_CHOL_PROFILE as a variant)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context()
lib = capi.load()
out = (C.c_ulonglong * 8)()
for n, dt, batch in [(128, torch.float64, 8192), (128, torch.float32, 16384), (96, torch.float64, 8192)]:
    A0 = torch.empty((batch, n, n), dtype=dt, device="cuda"); capi.fill_spd_batched(ctx, A0, float(n), 2)
    A = A0.clone(); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    for it in range(2):
        A.copy_(A0); lib.gpub_debug_chol_profile(out, 1)
        capi.potrf_batched(ctx, A, info)
        lib.gpub_debug_chol_profile(out, 0)
    names = ["load", "diag", "panel", "trailing", "store"]
    tot = sum(out[:5]) or 1
    print(n, dt, " ".join(f"{nm} {v / batch / 1e3:.1f}k ({100 * v / tot:.0f}%)" for nm, v in zip(names, out[:5])), f"| total {tot / batch / 1e3:.1f} kcycles per matrix")
