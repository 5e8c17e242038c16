#!/usr/bin/env python
"""bench.py -- benchmark of the batched linear-algebra hot path (BASELINE.json / SURVEY.md 8d), both arms.

Headline workload (`value`, N = 1): BASELINE config 2 -- CholeskyBatchFactoriser factorise + solve, 32 x 32 SPD fp64,
k = 1,000,000 matrices with one right-hand side each (8.19 GB of A + 256 MB of b, synthetic). A "step" is one factorise() +
one solve() over the whole batch through the C ABI (include/gputils_b200.h): two kernel launches. The operation is in place,
so A and b are restored from a pristine device copy between steps, outside the timed sub-regions (SURVEY.md 8d "Timing
method"); the 8.4 GB restore also evicts L2 (126 MB).

  value        matrices / s, whole job, inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e          the same metric through the product's host call (gpub_chol_solve_from_host_f64 = CholeskyBatchFactoriser::
               factoriseAndSolveFromHost): pinned host A, b -> device, factorise, solve, x and info -> pinned host, every step
               inside the timed region
  e2e_dropin   (N = 1) the reference-shaped call sequence upload -> factorise -> solve -> download from PAGEABLE host memory
               (DTensor::upload / download = gpub_upload / gpub_download here, cudaMemcpy in the reference)
  roofline     dominant kernel = potrf; algorithmic bytes = (2 n^2 s + 4) * k per launch over its mean device
               time; peak = MEASURED_PEAKS.json hbm_gbs
  configs      every other BASELINE config, each timed the same way (events, inputs restored outside the timed region):
               1 (GEMM 8x8 fp64: one call at k = 4096, CUDA-graph replay of that call, and the DRAM-bound k = 4e6 batch),
               3 (gels 64x16 fp32, k = 2^20), 4a/4b/4c (geqrf, Svd, Nullspace + project at 1024x128 fp64, k = 256) and
               5 (GEMM / potrf / potrs for n in 4..128, fp32 + fp64, >= 1 GiB per operand), with ms, matrices/s, GFLOP/s and
               the fraction of the bounding roofline
  strong       the FIXED k = 1e6 batch sharded over the N ranks (what north_star's ">= 7x at 8 GPUs" is about), beside the
               weak-scaling `value` (k matrices per GPU)
  allgather    (N > 1) NCCL all-gather of the x and L shards of the strong-scaling batch, GB/s into each device
  sharded_api  (N > 1) the product's one-process multi-GPU API (ShardedDTensor, gpub_multi_allgather over NCCL and over peer
               copies) run by rank 0 on all the devices of the box after the timed region
  cpu_baseline host LAPACK loop of SURVEY.md 8d "Baseline 2": single-threaded OpenBLAS potrf + potrs per matrix under an
               OpenMP loop over the batch (oracle/cpu_lapack.c), every host core, thread count stated
  --impl reference   the UNMODIFIED reference header built against cuBLAS / cuSOLVER (oracle/_ref/libgputils_ref.so), same
               inputs (generated with torch in both arms), same configs, same timing. GPUtils has no CPU implementation (it
               is a cuBLAS / cuSOLVER wrapper), so its arm runs on the same GPU; this arm never imports gputils_b200.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

N_MAT = 32
# DRAM bytes per 32 x 32 fp64 matrix of the potrf kernel from one `ncu --set full` capture (dram__bytes_read.sum +
# dram__bytes_write.sum over the matrices of the launch): a constant of the build, not a measurement of this run
POTRF_DRAM_BYTES_PER_MATRIX = 10809.8
POTRF_TRAFFIC_SOURCE = "static: ncu --set full of this kernel, profiles/r1h_ncu_chol32.json (only the lower triangle moves)"
SEED_A, SEED_B = 0x5EED0002, 0x5EED0102
# pipe peaks measured on this pool's B200 with scripts/microbench/peaks.cu (profiles/r1_peaks_fp64_fp32.json)
FP64_TFLOPS, FP32_TFLOPS = 37.1, 71.7
METRIC = "batched GEMM/Cholesky/QR matrices/s & %roofline (HBM or FP64 TC), 1-8 B200"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--batch", type=int, default=1_000_000, help="matrices per GPU (BASELINE config 2: 1e6)")
    p.add_argument("--configs", choices=["all", "sweep", "none"], default="all", help="the other BASELINE configs (all: N = 1 only)")
    p.add_argument("--reps", type=int, default=5, help="timed repetitions per entry of the configs block")
    p.add_argument("--cpu-reps", type=int, default=3, help="passes of the CPU baseline over the whole batch")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-sharded-api", action="store_true")
    return p.parse_args()


# ---- synthetic inputs: the counter-based generator of SURVEY.md 8(d) in plain torch, so that BOTH arms get identical buffers
# without the reference arm touching the repo's library (mirror of gpub_u01 in csrc/common.cuh and oracle_np.u01) ---------------
def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, s: int):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix64(z):
    z = z + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def gen_u01(seed: int, start: int, count: int, device):
    import torch
    i = torch.arange(start, start + count, dtype=torch.int64, device=device)
    h = _mix64(_mix64(i) ^ _s64(seed))
    return _lsr(h, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def gen_uniform(out, lo: float, hi: float, seed: int, chunk: int = 1 << 26):
    """out.flatten()[i] = lo + (hi - lo) * u01(seed, i)"""
    flat = out.view(-1)
    for s in range(0, flat.numel(), chunk):
        c = min(chunk, flat.numel() - s)
        flat[s:s + c] = (lo + (hi - lo) * gen_u01(seed, s, c, out.device)).to(out.dtype)
    return out


def gen_spd(out, shift: float, seed: int):
    """out: (k, n, n); A_i = G_i G_i' + shift I with G_i ~ U[-1, 1] (element (r, c) of G_i is draw number i n^2 + r + c n)"""
    import torch
    k, n = out.shape[0], out.shape[1]
    chunk_mats = max(1, (1 << 25) // (n * n))
    eye = shift * torch.eye(n, dtype=torch.float64, device=out.device)
    for s in range(0, k, chunk_mats):
        c = min(chunk_mats, k - s)
        g = (2.0 * gen_u01(seed, s * n * n, c * n * n, out.device) - 1.0).view(c, n, n)     # [mat][col][row]
        out[s:s + c] = (torch.bmm(g.transpose(1, 2), g) + eye).to(out.dtype)                # (G G')[r][c] = sum_l G[r,l] G[c,l]
    return out


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line). The timed region of this
    bench is ~60 ms, shorter than nvidia-smi's fastest loop can resolve reliably, so NVML is polled in-process every 2 ms
    (same counters nvidia-smi prints); `nvidia-smi -lms` is the fallback when the NVML binding is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, uuid: str | None = None):
        self.index = index
        self.uuid = uuid
        self.proc = None
        self.lines = []
        self.samples = []          # (sm_mhz, max_mhz, power_w, reasons bitmask) from NVML
        self.nvml = None
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    pw = float("nan")
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = 0
                self.samples.append((sm, self.max_mhz, pw, rs))
            except Exception:
                break
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            n = self.nvml
            names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4))
            reasons = set()
            for _, _, _, rs in self.samples:
                for name, attr, dflt in names:
                    if rs & int(getattr(n, attr, dflt)):
                        reasons.add(name)
            sm = [x[0] for x in self.samples]
            pw = [x[2] for x in self.samples if x[2] == x[2]]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz if sm else None,
                    "power_w_max": max(pw) if pw else None, "samples": len(sm), "source": "nvml, 2 ms polling during the timed region",
                    "reasons": sorted(reasons)}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "source": "nvidia-smi -lms 20", "reasons": sorted(reasons)}


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def cpu_baseline(A_host, b_host, reps: int):
    """SURVEY.md 8(d) Baseline 2: single-threaded OpenBLAS dpotrf + dpotrs per matrix (scipy's bundled OpenBLAS) under an
    OpenMP loop over the batch, on every host core this process may use. The sample is the WHOLE batch, `reps` passes."""
    sys.path.insert(0, str(REPO / "oracle"))
    import numpy as np
    import cpu_lapack
    threads = cpu_lapack.host_threads()
    k = A_host.shape[0]
    a = np.empty((k, N_MAT, N_MAT), dtype=np.float64); b = np.empty((k, 1, N_MAT), dtype=np.float64)
    info = np.zeros(k, dtype=np.int32)
    secs = []
    for _ in range(max(reps, 1)):
        np.copyto(a, A_host.numpy()); np.copyto(b, b_host.numpy())         # restore, outside the timed loop
        secs.append(cpu_lapack.chol_batch(a, b, info, threads=threads))
    assert min(secs) > 0 and not info.any()
    dt = statistics.median(secs)
    return {"value": k / dt, "unit": "matrices/s", "cores": int(cpu_lapack.lib().cpu_lapack_threads()), "kind": "openblas",
            "sample": (f"all {k} 32x32 fp64 SPD systems of the workload, dpotrf + dpotrs per matrix (scipy's OpenBLAS, 1 LAPACK thread per "
                       f"matrix) under `omp parallel for` over the batch, {len(secs)} passes, median {dt:.3f} s"),
            "seconds_per_pass": secs}


# ---- the two arms: every method returns a list of device milliseconds, one per repetition -----------------------------------
def _ev():
    import torch
    return torch.cuda.Event(enable_timing=True)


class OursArm:
    """The product: libgputils_b200.so through its C ABI (gputils_b200/capi.py). Fails loudly if the library is missing."""
    name = "ours"

    def __init__(self, device_index: int):
        from gputils_b200 import capi
        self.capi = capi
        self.ctx = capi.Context(device_index)
        self.launches = 0

    def timed(self, fn, restore, reps, warm=2):
        import torch
        for _ in range(warm):
            restore(); fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(reps):
            restore()
            e0, e1 = _ev(), _ev()
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def chol_step(self, A, b, info):
        """one factorise + solve, returns the three events around the two launches"""
        e0, e1, e2 = _ev(), _ev(), _ev()
        e0.record(); self.capi.potrf_batched(self.ctx, A, info)
        e1.record(); self.capi.potrs_batched(self.ctx, A, b)
        e2.record()
        self.launches += 2
        return e0, e1, e2

    def gemm(self, A, B, Cm, reps, restore=lambda: None):
        return self.timed(lambda: self.capi.gemm_batched(self.ctx, Cm, A, B), restore, reps)

    def potrf(self, A0, A, info, reps):
        return self.timed(lambda: self.capi.potrf_batched(self.ctx, A, info), lambda: A.copy_(A0), reps)

    def potrs(self, L, b0, b, reps):
        return self.timed(lambda: self.capi.potrs_batched(self.ctx, L, b), lambda: b.copy_(b0), reps)

    def gels(self, A0, b0, A, b, reps):
        def restore():
            A.copy_(A0); b.copy_(b0)
        return self.timed(lambda: self.capi.gels_batched(self.ctx, A, b), restore, reps)

    def geqrf(self, A0, A, tau, reps):
        return self.timed(lambda: self.capi.geqrf_batched(self.ctx, A, tau), lambda: A.copy_(A0), reps)

    def svd(self, A0, A, want_u, reps):
        import torch
        capi, ctx = self.capi, self.ctx
        k, n, m = A.shape
        ws = getattr(ctx.lib, "gpub_gesvd_batched_worksize_f64")(m, n, ord("A") if want_u else ord("N"), k)
        work = torch.empty(ws, dtype=torch.uint8, device=A.device)
        S = torch.empty((k, n), dtype=A.dtype, device=A.device); Vt = torch.empty((k, n, n), dtype=A.dtype, device=A.device)
        U = torch.empty((k, m, m), dtype=A.dtype, device=A.device) if want_u else None
        info = torch.zeros(k, dtype=torch.int32, device=A.device)
        p = capi._p

        def fn():
            ctx.call("gesvd_batched", A, ord("A") if want_u else ord("N"), m, n, p(A), m, m * n, p(S), n, p(U) if want_u else None, m, m * m,
                     p(Vt), n, n * n, p(work), ws, p(info), k)
        ms = self.timed(fn, lambda: A.copy_(A0), reps, warm=1)
        self.last_svd = (S, U, Vt)
        return ms

    def nullspace(self, a0, b0, reps):
        """Nullspace constructor (tr -> gesvd with U -> rank -> pack -> N N') and project(); returns (build ms, project ms)"""
        import torch
        capi, ctx = self.capi, self.ctx
        state = {}

        def build():
            state["N"], state["P"], state["rank"] = capi.nullspace_build(ctx, a0)
        t_build = self.timed(build, lambda: None, reps, warm=1)
        b = b0.clone()
        t_proj = self.timed(lambda: capi.nullspace_project(ctx, state["P"], b), lambda: b.copy_(b0), reps, warm=1)
        self.last_nullspace = (state["N"], state["P"], b)
        return t_build, t_proj


class RefArm:
    """The UNMODIFIED reference header on cuBLAS / cuSOLVER (oracle/_ref/libgputils_ref.so, built by oracle/Makefile from
    /root/reference/include). Nothing of the repo's product is imported or loaded on this arm."""
    name = "reference"

    def __init__(self):
        path = REPO / "oracle" / "_ref" / "libgputils_ref.so"
        if not path.exists():
            raise FileNotFoundError("oracle/_ref/libgputils_ref.so not built (make -C oracle ref)")
        self.lib = C.CDLL(str(path))
        self.lib.ref_chol_batch_host_f64.restype = C.c_double
        self.launches = None

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    @staticmethod
    def _suf(t):
        import torch
        return "f64" if t.dtype == torch.float64 else "f32"

    def chol_step(self, A0, b0, A, b, info):
        msf, mss = C.c_float(), C.c_float()
        p = self._p
        self.lib.ref_chol_batch_f64(C.c_size_t(A0.shape[1]), C.c_size_t(A0.shape[0]), p(A0), p(A), p(b0), p(b), p(info), 1, C.byref(msf), C.byref(mss))
        return msf.value, mss.value

    def gemm(self, A, B, Cm, reps, restore=None):
        k, ka, m = A.shape
        n = B.shape[1]
        ct = C.c_double if self._suf(A) == "f64" else C.c_float
        ms = C.c_float()
        fn = getattr(self.lib, f"ref_addAB_{self._suf(A)}")
        p = self._p
        fn(C.c_size_t(m), C.c_size_t(n), C.c_size_t(ka), C.c_size_t(k), p(A), p(B), p(Cm), ct(1.0), ct(0.0), 1, None)
        fn(C.c_size_t(m), C.c_size_t(n), C.c_size_t(ka), C.c_size_t(k), p(A), p(B), p(Cm), ct(1.0), ct(0.0), reps, C.byref(ms))
        return [ms.value]

    def _chol(self, A0, A, b0, b, info, reps):
        msf, mss = C.c_float(), C.c_float()
        fn = getattr(self.lib, f"ref_chol_batch_{self._suf(A0)}")
        p = self._p
        fn(C.c_size_t(A0.shape[1]), C.c_size_t(A0.shape[0]), p(A0), p(A), p(b0), p(b), p(info), reps, C.byref(msf), C.byref(mss))
        return msf.value, mss.value

    def potrf_potrs(self, A0, A, b0, b, info, reps):
        self._chol(A0, A, b0, b, info, 1)
        f, s = self._chol(A0, A, b0, b, info, reps)
        return [f], [s]

    def gels(self, A0, b0, A, b, reps):
        k, n, m = A0.shape
        ms = C.c_float()
        fn = getattr(self.lib, f"ref_gels_{self._suf(A0)}")
        p = self._p
        fn(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), p(A0), p(A), p(b0), p(b), 1, None)
        fn(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), p(A0), p(A), p(b0), p(b), reps, C.byref(ms))
        return [ms.value]

    def geqrf(self, A0, A, tau, reps):
        k, n, m = A0.shape
        ms = C.c_float()
        p = self._p
        self.lib.ref_qr_f64(C.c_size_t(m), C.c_size_t(n), C.c_size_t(min(k, 4)), p(A0), p(A), None, None, 1, None, None)
        self.lib.ref_qr_f64(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), p(A0), p(A), None, None, reps, C.byref(ms), None)
        return [ms.value]

    def svd(self, A0, A, want_u, reps):
        import torch
        k, n, m = A0.shape
        S = torch.empty((k, n), dtype=A0.dtype, device=A0.device); Vt = torch.empty((k, n, n), dtype=A0.dtype, device=A0.device)
        U = torch.empty((k, m, m), dtype=A0.dtype, device=A0.device) if want_u else None
        ms = C.c_float()
        p = self._p
        self.lib.ref_svd_f64(C.c_size_t(m), C.c_size_t(n), C.c_size_t(min(k, 2)), p(A0), p(S), p(Vt), p(U), None, None, C.c_double(1e-6), 1, None)
        self.lib.ref_svd_f64(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), p(A0), p(S), p(Vt), p(U), None, None, C.c_double(1e-6), reps, C.byref(ms))
        self.last_svd = (S, U, Vt)
        return [ms.value]

    def nullspace(self, a0, b0, reps):
        import torch
        k, n, m = a0.shape
        N = torch.empty((k, n, n), dtype=a0.dtype, device=a0.device); pr = torch.empty_like(b0)
        msb, msp = C.c_float(), C.c_float()
        p = self._p
        self.lib.ref_nullspace_f64(C.c_size_t(m), C.c_size_t(n), C.c_size_t(min(k, 2)), p(a0), p(N), p(b0), p(pr), 1, None, None)
        self.lib.ref_nullspace_f64(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), p(a0), p(N), p(b0), p(pr), reps, C.byref(msb), C.byref(msp))
        self.last_nullspace = (N, None, pr)
        return [msb.value], [msp.value]


def entry(name, shape, dt, k, ms_list, bytes_per, flops_per, bound, hbm_peak, extra=None):
    """one line of the configs block"""
    ms = statistics.median(ms_list)
    e = {"name": name, "shape": shape, "dtype": dt, "matrices": k, "ms": ms, "ms_best": min(ms_list), "reps": len(ms_list),
         "matrices_per_s": k / (ms * 1e-3), "gflops": flops_per * k / (ms * 1e-3) / 1e9}
    if bound == "hbm":
        ach = bytes_per * k / (ms * 1e-3) / 1e9
        e["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                         "algorithmic_bytes_per_matrix": bytes_per}
    else:
        peak = FP64_TFLOPS if bound == "fp64" else FP32_TFLOPS
        ach = flops_per * k / (ms * 1e-3) / 1e12
        e["roofline"] = {"bound": "tensor" if bound == "fp64" else "ffma", "pipe": "FP64 (DMMA / DFMA)" if bound == "fp64" else "FP32 FFMA",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                         "flops_per_matrix": flops_per, "peak_source": "scripts/microbench/peaks.cu on this pool (profiles/r1_peaks_fp64_fp32.json)"}
    if extra:
        e.update(extra)
    return e


def run_configs(arm, args, dev, hbm_peak, which, rank, world, reduce_max):
    """BASELINE configs 1, 3, 4 (N = 1 only) and 5 (every N: each rank sweeps its own shard; max over ranks)."""
    import torch
    out = []
    reps = args.reps
    is_ref = arm.name == "reference"
    f64, f32 = torch.float64, torch.float32

    def agg(ms_list):
        return [reduce_max(x) for x in ms_list] if world > 1 else ms_list

    if which == "all" and world == 1:
        # ---- config 1: addAB 8 x 8 . 8 x 8 fp64, k = 4096 (testTensor.cu sizes) ----
        n, k = 8, 4096
        A = gen_uniform(torch.empty((k, n, n), dtype=f64, device=dev), -1.0, 1.0, 0x5EED0001)
        B = gen_uniform(torch.empty((k, n, n), dtype=f64, device=dev), -1.0, 1.0, 0x5EED0101)
        Cm = torch.zeros_like(A)
        ms = arm.gemm(A, B, Cm, 20)
        out.append(entry("cfg1_addAB_8x8_f64_k4096_single_call", [8, 8, 8], "f64", k, ms, 3 * n * n * 8, 2 * n ** 3, "hbm", hbm_peak,
                         {"note": "one call on a 6.3 MB problem: bound by launch latency, not by any roofline (SURVEY.md 7-6); L2-resident"}))
        if not is_ref:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(device=dev)
            LAUNCHES = 64
            with torch.cuda.stream(s):
                arm.ctx.bind_torch_stream(0)
                arm.capi.gemm_batched(arm.ctx, Cm, A, B)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(LAUNCHES):
                        arm.capi.gemm_batched(arm.ctx, Cm, A, B)
                g.replay(); torch.cuda.synchronize()
                tl = []
                for _ in range(10):
                    e0, e1 = _ev(), _ev()
                    e0.record(s); g.replay(); e1.record(s)
                    torch.cuda.synchronize()
                    tl.append(e0.elapsed_time(e1) / LAUNCHES)
            arm.ctx.bind_torch_stream(0)
            out.append(entry("cfg1_addAB_8x8_f64_k4096_cuda_graph_replay", [8, 8, 8], "f64", k, tl, 3 * n * n * 8, 2 * n ** 3, "hbm", hbm_peak,
                             {"note": f"launch-amortised: CUDA graph of {LAUNCHES} addAB calls through the C ABI, per call; operands stay in L2 "
                                      "(6.3 MB), so the HBM fraction is an L2 figure, reported for the launch cost only"}))
        k = 4_000_000
        A = gen_uniform(torch.empty((k, n, n), dtype=f64, device=dev), -1.0, 1.0, 0x5EED0001)
        B = gen_uniform(torch.empty((k, n, n), dtype=f64, device=dev), -1.0, 1.0, 0x5EED0101)
        Cm = torch.zeros_like(A)
        ms = arm.gemm(A, B, Cm, reps)
        out.append(entry("cfg1_addAB_8x8_f64_k4e6_dram_bound", [8, 8, 8], "f64", k, ms, 3 * n * n * 8, 2 * n ** 3, "hbm", hbm_peak,
                         {"note": "the same call with k scaled until the operands (6.1 GB) are DRAM-resident"}))
        del A, B, Cm

        # ---- config 3: leastSquaresBatched 64 x 16 fp32, k = 2^20 ----
        m, n, k = 64, 16, 1 << 20
        A0 = gen_uniform(torch.empty((k, n, m), dtype=f32, device=dev), -1.0, 1.0, 0x5EED0003)
        b0 = gen_uniform(torch.empty((k, 1, m), dtype=f32, device=dev), -1.0, 1.0, 0x5EED0103)
        A = torch.empty_like(A0); b = torch.empty_like(b0)
        ms = arm.gels(A0, b0, A, b, reps)
        x = b[:2048, 0, :n].double(); Am = A0[:2048].transpose(1, 2).double(); bm = b0[:2048, 0].double()
        grad = torch.bmm(Am.transpose(1, 2), (torch.bmm(Am, x.unsqueeze(2)).squeeze(2) - bm).unsqueeze(2))
        rel = float(grad.norm() / (Am.norm() * bm.norm()))
        assert rel < 1e-5, rel
        out.append(entry("cfg3_gels_64x16_f32_k2^20", [64, 16], "f32", k, ms, 2 * m * n * 4 + 2 * m * 4 + 4,
                         2 * m * n * n - 2 * n ** 3 / 3 + 4 * m * n - n * n, "hbm", hbm_peak, {"normal_equations_residual_rel": rel}))
        del A0, b0, A, b

        # ---- config 4: 1024 x 128 fp64, k = 256 ----
        m, n, k = 1024, 128, 256
        A0 = gen_uniform(torch.empty((k, n, m), dtype=f64, device=dev), -1.0, 1.0, 0x5EED0004)
        A = torch.empty_like(A0); tau = torch.zeros((k, n), dtype=f64, device=dev)
        ms = arm.geqrf(A0, A, tau, reps)
        out.append(entry("cfg4a_geqrf_1024x128_f64_k256", [m, n], "f64", k, ms, 2 * m * n * 8 + n * 8, 2 * m * n * n - 2 * n ** 3 / 3, "fp64", hbm_peak,
                         {"note": "QRFactoriser::factorise; the reference is single-matrix, its arm loops 256 calls (tensor.cuh:1811-1813)"}))
        sv_reps = 1 if is_ref else reps
        ms = arm.svd(A0, A, False, sv_reps)
        S = arm.last_svd[0]
        s_ref = torch.linalg.svdvals(A0[:4].transpose(1, 2))
        rel = float((S[:4] - s_ref).norm() / s_ref.norm())
        assert rel < 1e-10, rel
        out.append(entry("cfg4b_svd_1024x128_f64_k256_noU", [m, n], "f64", k, ms, (m * n + n * n + n) * 8, 2 * m * n * n + 2 * n ** 3, "fp64", hbm_peak,
                         {"flop_model": "2 m n^2 + 2 n^3 (QR-first estimate of SURVEY.md 8d)", "singular_values_rel_err_vs_torch": rel}))
        ms = arm.svd(A0, A, True, sv_reps)
        out.append(entry("cfg4b_svd_1024x128_f64_k256_fullU", [m, n], "f64", k, ms, (m * n + m * m + n * n + n) * 8, 4 * m * m * n + 22 * n ** 3, "fp64",
                         hbm_peak, {"flop_model": "4 m^2 n + 22 n^3 (SURVEY.md 8d)"}))
        arm.last_svd = None
        a0 = A0.transpose(1, 2).contiguous()                       # the fat transposes (128 x 1024) Nullspace takes
        b0 = gen_uniform(torch.empty((k, 1, m), dtype=f64, device=dev), -1.0, 1.0, 0x5EED0104)
        del A, A0
        tb, tp = arm.nullspace(a0, b0, 1 if is_ref else min(reps, 3))
        N = arm.last_nullspace[0]
        an = float(torch.bmm(a0[:8].transpose(1, 2), N[:8].transpose(1, 2)).abs().max())      # a N = 0
        assert an < 1e-9, an
        out.append(entry("cfg4c_nullspace_build_128x1024_f64_k256", [128, 1024], "f64", k, tb, (m * n + 2 * m * m) * 8, 4 * m * m * n + 22 * n ** 3 + 2 * m ** 3,
                         "fp64", hbm_peak, {"flop_model": "the reference formulation: SVD with full U (4 m^2 n + 22 n^3) + N N' (2 m^3)", "max_abs_a_times_N": an,
                                             "frac_note": "not a pipe utilisation for this arm: U is assembled through one block reflector and the projector as "
                                                          "I - U1 U1' (7 x fewer columns), so fewer flops are executed than the model counts"}))
        out.append(entry("cfg4c_nullspace_project_1024_f64_k256", [1024, 1024, 1], "f64", k, tp, (m * m + 2 * m) * 8, 2 * m * m, "hbm", hbm_peak))
        arm.last_nullspace = None
        del a0, b0, N
        torch.cuda.empty_cache()

    if which in ("all", "sweep"):
        # ---- config 5: n in {4 .. 128}, fp32 + fp64, GEMM + Cholesky, >= 1 GiB per operand, this rank's shard ----
        for dt, tdt, s in (("f64", f64, 8), ("f32", f32, 4)):
            for n in (4, 8, 16, 32, 64, 128):
                k = -(-(1 << 30) // (n * n * s))
                A0 = gen_spd(torch.empty((k, n, n), dtype=tdt, device=dev), float(n), 0x5EED0005 + rank)
                B = gen_uniform(torch.empty((k, n, n), dtype=tdt, device=dev), -1.0, 1.0, 0x5EED0105 + rank)
                Cm = torch.empty_like(A0)
                ms = agg(arm.gemm(A0, B, Cm, reps))
                bound = "hbm" if n <= 32 or (n == 64 and dt == "f32") else ("fp64" if dt == "f64" else "fp32")
                if n == 64 and dt == "f64":
                    bound = "hbm"          # AI 5.3 flop/B, below the 5.7 flop/B ridge of 37.1 TFLOP/s over 6.55 TB/s
                out.append(entry(f"cfg5_gemm_n{n}_{dt}", [n, n, n], dt, k * world, ms, 3 * n * n * s, 2 * n ** 3, bound, hbm_peak))
                del B, Cm
                b0 = gen_uniform(torch.empty((k, 1, n), dtype=tdt, device=dev), -1.0, 1.0, 0x5EED0205 + rank)
                A = torch.empty_like(A0); b = torch.empty_like(b0)
                info = torch.zeros(k, dtype=torch.int32, device=dev)
                if is_ref:
                    tf, ts = arm.potrf_potrs(A0, A, b0, b, info, max(reps // 2, 1))
                else:
                    tf = arm.potrf(A0, A, info, reps)
                    ts = arm.potrs(A, b0, b, reps)                  # A holds the factors of the last repetition
                assert int(info.abs().max()) == 0
                sl = slice(0, 512)
                x = b[sl].transpose(1, 2).double()
                r = float((torch.bmm(A0[sl].double(), x) - b0[sl].transpose(1, 2).double()).norm() / b0[sl].double().norm())
                assert r < (1e-11 if dt == "f64" else 1e-3), r
                out.append(entry(f"cfg5_potrf_n{n}_{dt}", [n, n], dt, k * world, agg(tf), 2 * n * n * s + 4, n ** 3 / 3 + n * n / 2 + n / 6, "hbm", hbm_peak))
                out.append(entry(f"cfg5_potrs_n{n}_{dt}", [n, n], dt, k * world, agg(ts), n * n * s + 2 * n * s, 2 * n * n, "hbm", hbm_peak,
                                 {"solve_residual_rel": r}))
                del A0, A, b0, b, info
        torch.cuda.empty_cache()
    return out


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: everything libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    is_ref = args.impl == "reference"
    if is_ref and rank != 0:
        return 0                                         # rank 0 alone runs the reference arm
    torch.cuda.set_device(local_rank)
    use_dist = world > 1 and not is_ref
    cpu_group = None
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")       # host-side barrier while rank 0 runs the one-process multi-GPU API
    eff_world = world if use_dist else 1
    dev = torch.device("cuda", local_rank)

    if is_ref:
        try:
            arm = RefArm()
        except (FileNotFoundError, OSError) as exc:
            real_stdout.write(json.dumps({"impl": "reference", "unavailable": str(exc)}) + "\n")
            real_stdout.flush()
            return 0
    else:
        arm = OursArm(local_rank)                        # fails loudly if libgputils_b200.so is missing

    def reduce_max(x: float) -> float:
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    k, n = args.batch, N_MAT
    # synthetic inputs, generated on the owning device with torch (A = G G' + 32 I, b ~ U[-1, 1]); every
    # rank owns its shard of the mats axis (its own seed offset), shard-resident. Both arms get bit-identical buffers.
    A0 = gen_spd(torch.empty((k, n, n), dtype=torch.float64, device=dev), 32.0, SEED_A + rank)
    b0 = gen_uniform(torch.empty((k, 1, n), dtype=torch.float64, device=dev), -1.0, 1.0, SEED_B + rank)
    A = torch.empty_like(A0); b = torch.empty_like(b0)
    info = torch.zeros(k, dtype=torch.int32, device=dev)

    def restore(cnt=None):
        if cnt is None:
            A.copy_(A0); b.copy_(b0)
        else:
            A[:cnt].copy_(A0[:cnt]); b[:cnt].copy_(b0[:cnt])

    def run_steps(cnt, steps):
        """`steps` x (restore, factorise, solve) on the first cnt matrices; returns (sum factorise ms, sum solve ms)"""
        Ac, bc, ic = A[:cnt], b[:cnt], info[:cnt]
        pending = []
        for _ in range(steps):
            restore(cnt)
            pending.append(arm.chol_step(A0[:cnt], b0[:cnt], Ac, bc, ic) if is_ref else arm.chol_step(Ac, bc, ic))
        torch.cuda.synchronize()
        if is_ref:
            return sum(p[0] for p in pending), sum(p[1] for p in pending)
        return sum(p[0].elapsed_time(p[1]) for p in pending), sum(p[1].elapsed_time(p[2]) for p in pending)

    # ---- warm-up ---------------------------------------------------------------------------------------------
    warmup = max(args.warmup, 3)
    run_steps(k, warmup)
    assert int(info.abs().max()) == 0, "factorisation reported a non-SPD matrix on synthetic SPD input"

    # ---- timed region: exactly K steps ------------------------------------------------------------------------
    sampler = ClockSampler(local_rank, str(torch.cuda.get_device_properties(local_rank).uuid))
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    if not is_ref:
        arm.launches = 0
    f_sum, s_sum = run_steps(k, args.steps)
    if use_dist:
        dist.barrier()
    clocks = sampler.stop()
    gpu_launches = None if is_ref else arm.launches
    total_ms = reduce_max(f_sum + s_sum)
    factor_ms_sum = reduce_max(f_sum)
    ms_per_step = total_ms / args.steps
    value = eff_world * k / (ms_per_step * 1e-3)

    # property check at full size (size-independent): residual of the solve on a slice of the batch
    sl = slice(0, min(k, 4096))
    x = b[sl].transpose(1, 2)                                  # (k, n, 1)
    resid = torch.linalg.norm(torch.bmm(A0[sl].transpose(1, 2), x) - b0[sl].transpose(1, 2)) / torch.linalg.norm(b0[sl])
    assert float(resid) < 1e-12, f"solve residual {float(resid)}"

    # ---- strong scaling: the fixed k-matrix batch sharded over the ranks ----------------------------------------
    k_strong = -(-k // eff_world)
    run_steps(k_strong, 2)
    if use_dist:
        dist.barrier()
    fs, ss = run_steps(k_strong, args.steps)
    strong_ms = reduce_max(fs + ss) / args.steps
    strong = {"scaling": "strong", "total_matrices": k, "matrices_per_gpu": k_strong, "ms_per_step": strong_ms,
              "value": min(k, k_strong * eff_world) / (strong_ms * 1e-3), "unit": "matrices/s",
              "factorise_ms": reduce_max(fs) / args.steps, "solve_ms": reduce_max(ss) / args.steps,
              "note": "shard-resident compute, results left sharded; the all-gather is timed separately below"}

    # ---- all-gather of the result shards over NCCL (one process per GPU) ---------------------------------------
    allgather = None
    if use_dist:
        allgather = {}
        for name, shard in (("x", b[:k_strong]), ("L", A[:k_strong])):
            full = torch.empty((eff_world * shard.shape[0],) + tuple(shard.shape[1:]), dtype=shard.dtype, device=dev)
            dist.all_gather_into_tensor(full, shard.contiguous())
            torch.cuda.synchronize(); dist.barrier()
            ts = []
            for _ in range(5):
                e0, e1 = _ev(), _ev()
                e0.record(); dist.all_gather_into_tensor(full, shard); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = reduce_max(statistics.median(ts))
            recv = shard.numel() * shard.element_size() * (eff_world - 1)
            allgather[name] = {"bytes_into_each_device": recv, "ms": t, "GBps_into_each_device": recv / (t * 1e-3) / 1e9,
                               "transport": "ncclAllGather (torch.distributed, NVLink 5 / NVSwitch)"}
            del full
        allgather["note"] = ("strong-scaling shards; compute per step is %.3f ms, so a gather of L is not on the scaling path "
                             "(SURVEY.md 8e)") % strong_ms

    # ---- end to end through host buffers --------------------------------------------------------------------
    e2e = e2e_dropin = e2e_lower = None
    A_host = b_host = None
    want_host = (not args.no_e2e) or (rank == 0 and not args.no_cpu)
    if want_host:
        A_host = torch.empty((k, n, n), dtype=torch.float64, pin_memory=True)
        b_host = torch.empty((k, 1, n), dtype=torch.float64, pin_memory=True)
        A_host.copy_(A0); b_host.copy_(b0)
    if not args.no_e2e:
        x_host = torch.empty((k, 1, n), dtype=torch.float64, pin_memory=True)
        info_host = torch.empty(k, dtype=torch.int32, pin_memory=True)
        e2e_steps = max(1, min(args.steps, 5))
        h2d = A_host.numel() * 8 + b_host.numel() * 8
        d2h = x_host.numel() * 8 + info_host.numel() * 4
        if not is_ref:
            # the product's host call: three streams, chunks of the batch flow upload -> factorise + solve -> download
            def time_host_call(lower_only):
                def e2e_step():
                    arm.capi.chol_solve_from_host(arm.ctx, A, b, info, A_host, b_host, x_host, info_host, chunks=16, lower_only=lower_only)
                e2e_step(); torch.cuda.synchronize()
                if use_dist:
                    dist.barrier()
                t0 = time.perf_counter()
                s, e = _ev(), _ev()
                s.record()
                for _ in range(e2e_steps):
                    e2e_step()                               # blocking: returns when x and info are on the host
                e.record(); torch.cuda.synchronize()
                wall_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
                ms = reduce_max(max(s.elapsed_time(e) / e2e_steps, wall_ms))
                assert int(info_host.abs().max()) == 0
                assert torch.equal(x_host[:1024], b[:1024].cpu())
                return ms
            e2e_ms = time_host_call(False)
            e2e = {"value": eff_world * k / (e2e_ms * 1e-3), "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_ms, "steps": e2e_steps,
                   "path": "gpub_chol_solve_from_host_f64 (CholeskyBatchFactoriser::factoriseAndSolveFromHost): pinned host buffers, 16 chunks, "
                           "upload / factorise + solve / download on three streams; every byte of A and b is transferred"}
            # the same call with GPUB_LOWER_ONLY: the factorisation reads the lower triangle, so the 16 x 16 block above the diagonal
            # of every matrix is not put on the wire (75 % of A's bytes; h2d_bytes_per_step counts what is actually copied). Reported
            # beside the headline e2e, which moves full matrices like the reference arm.
            x_host.zero_()
            low_ms = time_host_call(True)
            e2e_lower = {"value": eff_world * k / (low_ms * 1e-3), "unit": "matrices/s",
                         "h2d_bytes_per_step": (n * (n // 2) + (n // 2) * (n // 2)) * 8 * k + b_host.numel() * 8, "d2h_bytes_per_step": d2h,
                         "ms_per_step": low_ms, "steps": e2e_steps,
                         "path": "as e2e with lowerTriangleOnly: of every matrix the block wholly above the diagonal is not transferred "
                                 "(strided DMA; the factorisation never reads it)"}
        if eff_world == 1:
            # the reference-shaped sequence from pageable memory: upload(A), upload(b), factorise, solve, download(x), download(info)
            import numpy as np
            A_page = np.empty((k, n, n), dtype=np.float64); b_page = np.empty((k, 1, n), dtype=np.float64)
            np.copyto(A_page, A_host.numpy()); np.copyto(b_page, b_host.numpy())
            x_page = np.empty((k, 1, n), dtype=np.float64); info_page = np.empty(k, dtype=np.int32)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            drop_steps = 2
            if is_ref:
                arm.lib.ref_chol_batch_host_f64(C.c_size_t(n), C.c_size_t(min(k, 1000)), vp(A_page), vp(b_page), vp(x_page), vp(info_page), 1)
                sec = arm.lib.ref_chol_batch_host_f64(C.c_size_t(n), C.c_size_t(k), vp(A_page), vp(b_page), vp(x_page), vp(info_page), drop_steps)
                drop_ms = sec * 1e3
                path = "DTensor(n,n,k) + upload(std::vector) + CholeskyBatchFactoriser::factorise + solve + download, wall clock"
            else:
                lib, h = arm.ctx.lib, arm.ctx.h

                def drop_step():
                    arm.capi.check(lib.gpub_upload(h, 0, C.c_void_p(A.data_ptr()), vp(A_page), A_page.nbytes), "gpub_upload")
                    arm.capi.check(lib.gpub_upload(h, 0, C.c_void_p(b.data_ptr()), vp(b_page), b_page.nbytes), "gpub_upload")
                    arm.capi.potrf_batched(arm.ctx, A, info)
                    arm.capi.potrs_batched(arm.ctx, A, b)
                    arm.capi.check(lib.gpub_download(h, 0, vp(x_page), C.c_void_p(b.data_ptr()), x_page.nbytes), "gpub_download")
                    arm.capi.check(lib.gpub_download(h, 0, vp(info_page), C.c_void_p(info.data_ptr()), info_page.nbytes), "gpub_download")
                drop_step(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(drop_steps):
                    drop_step()
                torch.cuda.synchronize()
                drop_ms = (time.perf_counter() - t0) * 1e3 / drop_steps
                path = "gpub_upload (= DTensor::upload) + potrf + potrs + gpub_download (= DTensor::download), pageable numpy buffers, wall clock"
            assert not info_page.any()
            xs = torch.from_numpy(x_page[:2048]).to(dev).transpose(1, 2)
            r2 = float(torch.linalg.norm(torch.bmm(A0[:2048].transpose(1, 2), xs) - b0[:2048].transpose(1, 2)) / torch.linalg.norm(b0[:2048]))
            assert r2 < 1e-12, r2
            e2e_dropin = {"value": k / (drop_ms * 1e-3), "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "ms_per_step": drop_ms, "steps": drop_steps, "path": path}
            del A_page, b_page, x_page, info_page
            if is_ref:
                e2e = dict(e2e_dropin)                        # the reference's only host path IS the whole-tensor call sequence

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(A_host, b_host, args.cpu_reps)
    del A_host, b_host

    # ---- the other BASELINE configs -----------------------------------------------------------------------------
    del A, b, info, A0, b0
    torch.cuda.empty_cache()
    hbm_peak, peak_src = measured_peak()
    configs = None
    if args.configs != "none":
        which = args.configs if eff_world == 1 else "sweep"
        configs = run_configs(arm, args, dev, hbm_peak, which, rank, eff_world, reduce_max)

    # ---- the product's one-process multi-GPU API on all the devices of the box (rank 0; the others wait on the host) ----
    sharded_api = None
    if use_dist and not args.no_sharded_api:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            sharded_api = run_sharded_api(eff_world)
        dist.barrier(group=cpu_group)

    if rank == 0:
        s = 8
        potrf_bytes = (2 * n * n * s + 4) * k
        potrs_bytes = (n * n * s + 2 * n * s) * k
        potrf_ms = factor_ms_sum / args.steps
        solve_ms = (total_ms - factor_ms_sum) / args.steps
        achieved = potrf_bytes / (potrf_ms * 1e-3) / 1e9
        traffic = POTRF_DRAM_BYTES_PER_MATRIX * k if not is_ref else None
        line = {
            "metric": METRIC,
            "value": value, "unit": "matrices/s", "n_gpus": eff_world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "CholeskyBatchFactoriser factorise+solve 32x32 SPD fp64 (BASELINE config 2)",
                       "matrices_per_gpu": k, "n": n, "rhs": 1,
                       "l2": "inputs (8.4 GB) larger than L2; restored from a pristine device copy between steps, outside the timed sub-regions",
                       "timing": "CUDA events on the launching stream around factorise() and solve() of every step, summed; max over ranks",
                       "inputs": "counter-based generator of SURVEY.md 8(d) evaluated with torch on the device: identical buffers in both arms",
                       "parallelism": f"mats axis sharded over {eff_world} GPU(s), no data-path collective"},
            "gpu_launches": gpu_launches,
            "breakdown": {"factorise_ms": potrf_ms, "solve_ms": solve_ms,
                          "factorise_matrices_per_s": eff_world * k / (potrf_ms * 1e-3),
                          "solve_matrices_per_s": eff_world * k / (solve_ms * 1e-3),
                          "factorise_gflops": eff_world * k * 11440 / (potrf_ms * 1e-3) / 1e9,
                          "solve_hbm_gbs": potrs_bytes / (solve_ms * 1e-3) / 1e9},
            "roofline": {"bound": "hbm", "kernel": "k_potrf_pair<double,32>" if not is_ref else "cusolverDnDpotrfBatched",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_kind": POTRF_TRAFFIC_SOURCE if not is_ref else None,
                         "frac_on_traffic": (traffic / (potrf_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": potrf_bytes,
                         "frac_of_8TBs_nominal": achieved / 8000.0},
            "clocks": clocks,
            "e2e": e2e,
            "e2e_dropin": e2e_dropin,
            "e2e_lower_only": e2e_lower,
            "strong": strong,
            "allgather": allgather,
            "sharded_api": sharded_api,
            "cpu_baseline": cpu,
            "configs": configs,
            "solve_residual_rel": float(resid),
        }
        if is_ref:
            line["impl"] = "reference"
            line["reference"] = "GPUtils include/tensor.cuh (unmodified) + cuBLAS/cuSOLVER 12.9, CholeskyBatchFactoriser::factorise/solve, same GPU"
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if use_dist:
        dist.destroy_process_group()
    return 0


def run_sharded_api(n_devices: int):
    """tests/host_harness/sharded_test (ShardedDTensor, ShardedCholeskyBatchFactoriser, gpub_multi_allgather) on devices
    0..n-1 of the box: every sharded result must be bit-identical to the single-GPU path; reports the all-gather rates."""
    exe = REPO / "build" / "tests" / "sharded_test"
    if not exe.exists():
        return {"devices": n_devices, "passed": False, "error": "build/tests/sharded_test missing (python -c 'import __graft_entry__ as g; g.build()')"}
    devs = ",".join(str(i) for i in range(n_devices))
    res = {"devices": n_devices, "passed": True, "allgather_GBps": {}}
    for transport in ("nccl", "p2p"):
        try:
            r = subprocess.run([str(exe), devs, transport, "200000"], capture_output=True, text=True, timeout=600)
        except subprocess.TimeoutExpired:
            res["passed"] = False
            res[transport] = "timeout"
            continue
        out = r.stdout + r.stderr
        ok = r.returncode == 0 and "ALL PASSED" in out and "FAIL " not in out
        res["passed"] = res["passed"] and ok
        res[transport + "_checks_passed"] = out.count("PASS ")
        for ln in out.splitlines():
            if ln.startswith("INFO allgather"):
                try:
                    res["allgather_GBps"][transport] = float(ln.rsplit("GBps_into_each_device=", 1)[1].split()[0])
                except (IndexError, ValueError):
                    pass
            if ln.startswith("INFO solve_allgather"):
                # potrs followed by an all-gather of x, against the solve kernel that stores x into every device's tensor itself
                try:
                    kv = dict(t.split("=") for t in ln.split()[2:])
                    res.setdefault("solve_allgather_ms", {})[transport] = {"separate": float(kv["separate_ms"]), "fused_kernel": float(kv["fused_ms"]),
                                                                           "systems": int(kv["systems"])}
                except (KeyError, ValueError):
                    pass
        if not ok:
            res[transport + "_tail"] = out[-800:]
    return res


if __name__ == "__main__":
    sys.exit(main())
