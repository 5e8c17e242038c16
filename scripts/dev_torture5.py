"""Development aid: gesvd_batched with padded leading dimensions and strides (the fall-back paths of the U assembly, the accumulated
rotations for a strided Vt) on small and chunked batches, against numpy."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gputils_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(5)
bad = 0
sms = torch.cuda.get_device_properties(0).multi_processor_count
for (m, n) in ((64, 16), (200, 24), (128, 64), (192, 96), (256, 128), (100, 40)):
    for batch in (3, sms + 9):
        for pads in ((0, 0, 0, 0, 0, 0, 0), (2, 6, 1, 2, 10, 2, 8), (0, 0, 3, 0, 0, 0, 0), (0, 0, 0, 0, 0, 2, 6), (0, 0, 0, 2, 4, 0, 0)):
            pl, psa, pss, plu, psu, plv, psv = pads
            lda, sA = m + pl, (m + pl) * n + psa
            sS = n + pss
            ldu, sU = m + plu, (m + plu) * m + psu
            ldvt, sVt = n + plv, (n + plv) * n + psv
            A = rng.uniform(-1, 1, (batch, m, n))
            bufA = torch.full((batch * sA,), 7.0, dtype=torch.float64, device="cuda")
            hostA = np.full(batch * sA, 7.0)
            for i in range(batch):
                for j in range(n): hostA[i * sA + j * lda: i * sA + j * lda + m] = A[i, :, j]
            bufA.copy_(torch.from_numpy(hostA))
            S = torch.full((batch * sS,), -5.0, dtype=torch.float64, device="cuda")
            U = torch.full((batch * sU,), -5.0, dtype=torch.float64, device="cuda")
            Vt = torch.full((batch * sVt,), -5.0, dtype=torch.float64, device="cuda")
            info = torch.zeros(batch, dtype=torch.int32, device="cuda")
            ws = ctx.lib.gpub_gesvd_batched_worksize_f64(m, n, ord("A"), batch)
            work = torch.empty(ws, dtype=torch.uint8, device="cuda")
            ctx.call("gesvd_batched", bufA, ord("A"), m, n, capi._p(bufA), lda, sA, capi._p(S), sS, capi._p(U), ldu, sU, capi._p(Vt), ldvt, sVt,
                     capi._p(work), ws, capi._p(info), batch)
            Sh, Uh, Vh = S.cpu().numpy(), U.cpu().numpy(), Vt.cpu().numpy()
            worst = 0.0; ok = not info.cpu().numpy().any()
            for i in (0, batch // 2, batch - 1):
                s_ = Sh[i * sS: i * sS + n]
                u_ = np.stack([Uh[i * sU + j * ldu: i * sU + j * ldu + m] for j in range(m)], axis=1)
                v_ = np.stack([Vh[i * sVt + j * ldvt: i * sVt + j * ldvt + n] for j in range(n)], axis=1)
                ref = np.linalg.svd(A[i], compute_uv=False)
                e = max(np.abs(s_ - ref).max(), np.abs(u_.T @ u_ - np.eye(m)).max(), np.abs(v_ @ v_.T - np.eye(n)).max(),
                        np.abs((u_[:, :n] * s_) @ v_ - A[i]).max())
                worst = max(worst, float(e))
            # padding untouched
            pad_ok = True
            if pss: pad_ok &= bool((Sh.reshape(batch, sS)[:, n:] == -5.0).all())
            if plv or psv:
                vv = Vh.reshape(batch, sVt)
                pad_ok &= bool((vv[:, ldvt * n:] == -5.0).all())
            ok = ok and worst < 1e-10 and pad_ok
            if not ok:
                bad += 1
                print(f"gesvd {m}x{n} batch {batch} pads {pads}: worst {worst:.1e} pad_ok {pad_ok} info {int(info.abs().max())}  <-- BAD")
print("BAD cases:", bad)
