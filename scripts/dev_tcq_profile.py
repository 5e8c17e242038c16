"""Development aid: per-phase cycle totals of k_geqrf_tc (library built with -DGPUB_TCQ_PROFILE as a variant)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context()
A0 = torch.empty((256, 128, 1024), dtype=torch.float64, device="cuda"); capi.fill_uniform(ctx, A0, -1.0, 1.0, 8)
A = A0.clone(); tau = torch.zeros((256, 128), dtype=torch.float64, device="cuda")
lib = capi.load()
out = (C.c_ulonglong * 12)()
for it in range(3):
    A.copy_(A0); lib.gpub_debug_tcq_profile(out, 1)
    capi.geqrf_batched(ctx, A, tau)
    lib.gpub_debug_tcq_profile(out, 0)
names = ["load panel", "panel columns", "store panel + V", "G, T", "pass 1", "W'", "pass 2 + store", "A2 load wait", "panel -> smem + Gram", "panel recurrence (1 warp)", "panel replay", "-"]
tot = sum(out[:12])
for n_, v in zip(names, out[:12]):
    print(f"{n_:18s} {v/256/1e3:10.1f} kcycles per matrix  {100*v/tot:5.1f}%")
print("total per matrix", tot / 256 / 1e3, "kcycles =", tot / 256 / 1.965e9 * 1e6, "us")
