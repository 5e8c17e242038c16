"""Development aid: sweeps and cycles per matrix of k_jacobi_blk (library built as a variant with -DGPUB_JBLK_STATS):
GPUB_VARIANT=jstats GPUB_EXTRA_NVCC_FLAGS=-DGPUB_JBLK_STATS python gputils_b200/build.py
GPUB_LIB=gputils_b200/lib/variants/libgputils_b200_jstats.so python scripts/dev_jacobi_stats.py"""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context()
lib = capi.load()
out = (C.c_ulonglong * 4)()
for (m, n, batch) in ((1024, 128, 256), (256, 64, 256)):
    A0 = torch.empty((batch, n, m), dtype=torch.float64, device="cuda"); capi.fill_uniform(ctx, A0, -1.0, 1.0, 9)
    for it in range(2):
        A = A0.clone(); lib.gpub_debug_jacobi_stats(out, 1)
        capi.gesvd_batched(ctx, A, False)
        lib.gpub_debug_jacobi_stats(out, 0)
    mats = max(1, out[0])
    sw = out[1] / mats
    print(f"{m}x{n} batch {batch}: sweeps per matrix {sw:.2f}, sweep loop {out[2]/mats/1e3:.0f} kcycles per matrix "
          f"({out[2]/mats/sw/(2*(n//8)-1+0.75)/4:.0f} cycles per round), tail {out[3]/mats/1e3:.0f} kcycles")
