"""Per-op throughput of the C-ABI launchers on the BASELINE shapes (development + profiles/ suite).

usage: python scripts/bench_ops.py [--ops potrf,potrs,gemm,gels,...] [--ref] [--reps 10] [--json out.json]
Timing: CUDA events on the launching stream, inputs restored from a pristine copy before every repetition
(outside the timed region; the restore also evicts L2), 3 warm-ups, median and best reported.
"""
import argparse, ctypes as C, json, statistics, sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from gputils_b200 import capi

PEAK = 6547.8
try:
    PEAK = float(json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timeit(fn, restore, reps):
    for _ in range(3):
        restore(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        restore()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="potrf,potrs")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--json", default="")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (quick runs)")
    args = ap.parse_args()
    ops = args.ops.split(",")
    ctx = capi.Context(0)
    ref = C.CDLL(str(REPO / "oracle" / "_ref" / "libgputils_ref.so")) if args.ref else None
    out = []
    SZ = C.c_size_t

    def report(name, shape, dt, batch, ms_med, ms_best, bytes_per, flops_per, ref_ms=None):
        r = {"op": name, "shape": shape, "dtype": dt, "batch": batch, "ms_median": ms_med, "ms_best": ms_best,
             "Mmat_per_s": batch / ms_med / 1e3, "GBps_algorithmic": bytes_per * batch / ms_med / 1e6,
             "frac_hbm_measured": bytes_per * batch / ms_med / 1e6 / PEAK, "GFLOPs": flops_per * batch / ms_med / 1e6}
        if ref_ms is not None:
            r["ref_ms"] = ref_ms; r["speedup_vs_ref"] = ref_ms / ms_med
        out.append(r)
        print(json.dumps(r), flush=True)

    def chol(n, dt, batch):
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        A0 = torch.empty((batch, n, n), dtype=tdt, device="cuda"); b0 = torch.empty((batch, 1, n), dtype=tdt, device="cuda")
        capi.fill_spd_batched(ctx, A0, float(n), 2); capi.fill_uniform(ctx, b0, -1.0, 1.0, 3)
        A = torch.empty_like(A0); b = torch.empty_like(b0); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        rf = rs = None
        if ref is not None:
            msf, mss = C.c_float(), C.c_float()
            getattr(ref, f"ref_chol_batch_{dt}")(SZ(n), SZ(batch), C.c_void_p(A0.data_ptr()), C.c_void_p(A.data_ptr()), C.c_void_p(b0.data_ptr()),
                                                 C.c_void_p(b.data_ptr()), C.c_void_p(info.data_ptr()), 3, C.byref(msf), C.byref(mss))
            rf, rs = msf.value, mss.value
        if "potrf" in ops:
            med, best = timeit(lambda: capi.potrf_batched(ctx, A, info), lambda: A.copy_(A0), args.reps)
            report("potrf", [n, n], dt, batch, med, best, 2 * n * n * s + 4, n ** 3 / 3 + n * n / 2 + n / 6, rf)
        A.copy_(A0); capi.potrf_batched(ctx, A, info); L = A.clone()
        if "potrs" in ops:
            med, best = timeit(lambda: capi.potrs_batched(ctx, L, b), lambda: b.copy_(b0), args.reps)
            report("potrs", [n, n], dt, batch, med, best, n * n * s + 2 * n * s, 2 * n * n, rs)
            x = b.transpose(1, 2)[:1024]; r = torch.linalg.norm(torch.bmm(A0[:1024], x) - b0.transpose(1, 2)[:1024]) / torch.linalg.norm(b0[:1024])
            assert float(r) < (1e-11 if dt == "f64" else 1e-3), float(r)

    def gemm(n, dt, batch):
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        A = torch.empty((batch, n, n), dtype=tdt, device="cuda"); B = torch.empty_like(A); Cm = torch.zeros_like(A)
        capi.fill_uniform(ctx, A, -1.0, 1.0, 4); capi.fill_uniform(ctx, B, -1.0, 1.0, 5)
        rm = None
        if ref is not None:
            ms = C.c_float()
            ct = C.c_double if dt == "f64" else C.c_float
            getattr(ref, f"ref_addAB_{dt}")(SZ(n), SZ(n), SZ(n), SZ(batch), C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(Cm.data_ptr()),
                                            ct(1.0), ct(0.0), 5, C.byref(ms))
            rm = ms.value
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        med, best = timeit(lambda: capi.gemm_batched(ctx, Cm, A, B), lambda: flush.zero_() if A.numel() * s * 3 < (400 << 20) else None, args.reps)
        report("gemm", [n, n, n], dt, batch, med, best, 3 * n * n * s, 2 * n ** 3, rm)

    def gels(m, n, dt, batch):
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        A0 = torch.empty((batch, n, m), dtype=tdt, device="cuda"); b0 = torch.empty((batch, 1, m), dtype=tdt, device="cuda")
        capi.fill_uniform(ctx, A0, -1.0, 1.0, 6); capi.fill_uniform(ctx, b0, -1.0, 1.0, 7)
        A = torch.empty_like(A0); b = torch.empty_like(b0)
        rm = None
        if ref is not None:
            ms = C.c_float()
            getattr(ref, f"ref_gels_{dt}")(SZ(m), SZ(n), SZ(batch), C.c_void_p(A0.data_ptr()), C.c_void_p(A.data_ptr()), C.c_void_p(b0.data_ptr()),
                                           C.c_void_p(b.data_ptr()), 3, C.byref(ms))
            rm = ms.value
        def restore():
            A.copy_(A0); b.copy_(b0)
        med, best = timeit(lambda: capi.gels_batched(ctx, A, b), restore, args.reps)
        report("gels", [m, n], dt, batch, med, best, 2 * m * n * s + 2 * m * s + 4, 2 * m * n * n - 2 * n ** 3 / 3 + 4 * m * n - n * n, rm)

    def qr(m, n, dt, batch):
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        A0 = torch.empty((batch, n, m), dtype=tdt, device="cuda")
        capi.fill_uniform(ctx, A0, -1.0, 1.0, 8)
        A = torch.empty_like(A0); tau = torch.zeros((batch, n), dtype=tdt, device="cuda")
        rm = None
        if ref is not None:
            ms = C.c_float()
            getattr(ref, f"ref_qr_{dt}")(SZ(m), SZ(n), SZ(batch), C.c_void_p(A0.data_ptr()), C.c_void_p(A.data_ptr()), None, None, 2, C.byref(ms), None)
            rm = ms.value
        med, best = timeit(lambda: capi.geqrf_batched(ctx, A, tau), lambda: A.copy_(A0), max(3, args.reps // 2))
        report("geqrf", [m, n], dt, batch, med, best, 2 * m * n * s + n * s, 2 * m * n * n - 2 * n ** 3 / 3, rm)

    def svd(m, n, dt, batch, want_u):
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        A0 = torch.empty((batch, n, m), dtype=tdt, device="cuda")
        capi.fill_uniform(ctx, A0, -1.0, 1.0, 9)
        A = torch.empty_like(A0)
        rm = None
        if ref is not None:
            S = torch.empty((batch, n), dtype=tdt, device="cuda"); Vt = torch.empty((batch, n, n), dtype=tdt, device="cuda")
            U = torch.empty((batch, m, m), dtype=tdt, device="cuda") if want_u else None
            ms = C.c_float()
            ct = C.c_double if dt == "f64" else C.c_float
            getattr(ref, f"ref_svd_{dt}")(SZ(m), SZ(n), SZ(batch), C.c_void_p(A0.data_ptr()), C.c_void_p(S.data_ptr()), C.c_void_p(Vt.data_ptr()),
                                          C.c_void_p(U.data_ptr()) if want_u else None, None, None, ct(1e-6), 1, C.byref(ms))
            rm = ms.value
            del S, Vt, U
        # the launcher is timed with its outputs and workspace already allocated (like the reference's factorise())
        lib = ctx.lib
        ws = getattr(lib, f"gpub_gesvd_batched_worksize_{dt}")(m, n, ord("A") if want_u else ord("N"), batch)
        work = torch.empty(ws, dtype=torch.uint8, device="cuda")
        S = torch.empty((batch, n), dtype=tdt, device="cuda"); Vt = torch.empty((batch, n, n), dtype=tdt, device="cuda")
        U = torch.empty((batch, m, m), dtype=tdt, device="cuda") if want_u else None
        info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        def run():
            ctx.call("gesvd_batched", A, ord("A") if want_u else ord("N"), m, n, capi._p(A), m, m * n, capi._p(S), n,
                     capi._p(U) if want_u else None, m, m * m, capi._p(Vt), n, n * n, capi._p(work), ws, capi._p(info), batch)
        med, best = timeit(run, lambda: A.copy_(A0), max(3, args.reps // 2))
        assert int(info.abs().max()) == 0
        flops = (4 * m * m * n + 22 * n ** 3) if want_u else (2 * m * n * n + 2 * n ** 3)       # SURVEY.md 8d estimates
        bytes_ = (m * n + n + n * n + (m * m if want_u else 0)) * s
        report("gesvd_U" if want_u else "gesvd", [m, n], dt, batch, med, best, bytes_, flops, rm)

    def nullspace(m, n, dt, batch):
        """Nullspace(a) for fat a (m x n, m <= n) mirrored through the C ABI exactly as include/gpub200/factorisers.cuh does:
        tr -> gesvd(U) -> count_gt -> nullspace_pack -> aat; then project = batched GEMM with C aliasing B."""
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        a0 = torch.empty((batch, n, m), dtype=tdt, device="cuda")            # (k, cols, rows) of the fat matrix
        capi.fill_uniform(ctx, a0, -1.0, 1.0, 10)
        b0 = torch.empty((batch, 1, n), dtype=tdt, device="cuda"); capi.fill_uniform(ctx, b0, -1.0, 1.0, 11)
        rb = rp = None
        if ref is not None:
            msb, msp = C.c_float(), C.c_float()
            proj = torch.empty_like(b0)
            getattr(ref, f"ref_nullspace_{dt}")(SZ(m), SZ(n), SZ(batch), C.c_void_p(a0.data_ptr()), None, C.c_void_p(b0.data_ptr()),
                                                C.c_void_p(proj.data_ptr()), 1, C.byref(msb), C.byref(msp))
            rb, rp = msb.value, msp.value
        lib = ctx.lib
        ws = getattr(lib, f"gpub_gesvd_batched_worksize_{dt}")(n, m, ord("A"), batch)
        work = torch.empty(ws, dtype=torch.uint8, device="cuda")
        at = torch.empty((batch, m, n), dtype=tdt, device="cuda")             # a^T: n x m tall
        S = torch.empty((batch, m), dtype=tdt, device="cuda"); Vt = torch.empty((batch, m, m), dtype=tdt, device="cuda")
        U = torch.empty((batch, n, n), dtype=tdt, device="cuda"); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
        N = torch.empty((batch, n, n), dtype=tdt, device="cuda"); P = torch.empty((batch, n, n), dtype=tdt, device="cuda")
        def build():
            ctx.call("transpose_batched", a0, m, n, capi._p(a0), m * n, capi._p(at), m * n, batch)
            ctx.call("gesvd_batched", at, ord("A"), n, m, capi._p(at), n, m * n, capi._p(S), m, capi._p(U), n, n * n, capi._p(Vt), m, m * m,
                     capi._p(work), ws, capi._p(info), batch)
            rank.zero_()
            ctx.call("count_gt_batched", S, capi._p(S), m, m, 1e-6, capi._p(rank), batch)
            ctx.call("nullspace_build_batched", U, n, capi._p(U), n * n, capi._p(rank), capi._p(N), n * n, capi._p(P), n * n, batch)
        med, best = timeit(build, lambda: None, max(2, args.reps // 3))
        assert int(info.abs().max()) == 0
        report("nullspace", [m, n], dt, batch, med, best, (m * n + 2 * n * n) * s, 4 * n * n * m + 22 * m ** 3 + 2 * n ** 3, rb)
        b = torch.empty_like(b0)
        med, best = timeit(lambda: capi.gemm_batched(ctx, b, P, b), lambda: b.copy_(b0), args.reps)
        report("project", [n, n], dt, batch, med, best, (n * n + 2 * n) * s, 2 * n * n, rp)
        # property: a * proj == 0
        x = b.transpose(1, 2)[:4]; r = torch.linalg.norm(torch.bmm(a0[:4].transpose(1, 2), x)) / torch.linalg.norm(b0[:4])
        assert float(r) < (1e-10 if dt == "f64" else 1e-3), float(r)

    def pitch(n, dt, batch):
        """dense (stride n*n) against 128-byte pitched storage for a shape that is not a multiple of 128 B: potrf + potrs through the strided C ABI"""
        tdt = torch.float64 if dt == "f64" else torch.float32
        s = 8 if dt == "f64" else 4
        A0 = torch.empty((batch, n, n), dtype=tdt, device="cuda"); capi.fill_spd_batched(ctx, A0, float(n), 2)
        b0 = torch.empty((batch, n), dtype=tdt, device="cuda"); capi.fill_uniform(ctx, b0, -1.0, 1.0, 3)
        info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        for name, stride in (("dense", n * n), ("pitched128", (n * n * s + 127) // 128 * 128 // s)):
            bs = n if name == "dense" else (n * s + 127) // 128 * 128 // s
            buf0 = torch.zeros((batch, stride), dtype=tdt, device="cuda"); buf0[:, : n * n] = A0.view(batch, n * n)
            rhs0 = torch.zeros((batch, bs), dtype=tdt, device="cuda"); rhs0[:, :n] = b0
            buf = torch.empty_like(buf0); rhs = torch.empty_like(rhs0)
            def run():
                ctx.call("potrf_batched", buf, n, capi._p(buf), n, stride, capi._p(info), batch)
                ctx.call("potrs_batched", buf, n, capi._p(buf), n, stride, capi._p(rhs), bs, batch)
            def restore():
                buf.copy_(buf0); rhs.copy_(rhs0)
            med, best = timeit(run, restore, args.reps)
            report("potrf+potrs_" + name, [n, n], dt, batch, med, best, 3 * n * n * s + 2 * n * s + 4, n ** 3 / 3 + 2 * n * n)

    sc = args.scale
    if "pitch" in ops:
        pitch(5, "f64", int(4_000_000 * sc))
        pitch(10, "f32", int(2_000_000 * sc))
    if "potrf" in ops or "potrs" in ops:
        chol(32, "f64", int(1_000_000 * sc))
    if "cholsweep" in ops:
        ops += ["potrf", "potrs"]
        for dt in ("f64", "f32"):
            s = 8 if dt == "f64" else 4
            for n in (4, 8, 16, 32, 64, 128):
                chol(n, dt, max(64, int((1 << 30) * sc / (n * n * s))))
    if "gemm" in ops:
        gemm(8, "f64", 4096)
        gemm(8, "f64", int(4_000_000 * sc))
    if "gemmsweep" in ops:
        for dt in ("f64", "f32"):
            s = 8 if dt == "f64" else 4
            for n in (4, 8, 16, 32, 64, 128):
                gemm(n, dt, max(64, int((1 << 30) * sc / (n * n * s))))
    if "gels" in ops:
        gels(64, 16, "f32", int((1 << 20) * sc))
    if "qr" in ops:
        qr(1024, 128, "f64", 256)
    if "svd" in ops:
        svd(1024, 128, "f64", max(2, int(256 * sc)), False)
    if "svdu" in ops:
        svd(1024, 128, "f64", max(2, int(256 * sc)), True)
    if "nullspace" in ops:
        nullspace(128, 1024, "f64", max(2, int(256 * sc)))
    if args.json:
        Path(args.json).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
