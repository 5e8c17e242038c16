#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched linear-algebra hot path (see BASELINE.json / SURVEY.md 8d).

Workload (N = 1): BASELINE config 2 -- CholeskyBatchFactoriser factorise + solve, 32 x 32 SPD fp64,
k = 1,000,000 matrices with one right-hand side each (8.19 GB of A + 256 MB of b, synthetic). A "step" is one
factorise() + one solve() over the whole batch through the C ABI (include/gputils_b200.h), i.e. two kernel
launches. The operation is in place, so A and b are restored from a pristine device copy between steps, outside
the timed sub-regions (SURVEY.md 8d "Timing method"); the 8.4 GB restore also evicts L2 (126 MB).

  value    : matrices / s, whole job, inputs resident in HBM (device events on the launching stream, max over ranks)
  e2e      : same metric through host buffers: pinned host -> device copy of A and b, factorise, solve, device ->
             host copy of x and info, every step inside the timed region
  roofline : dominant kernel = potrf; algorithmic bytes = (2 n^2 s + 4) * k per launch
             over its mean device time; peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the plain-C oracle port (oracle/oracle.c, OpenMP over the batch) on the host cores, bounded sample
  --impl reference : the UNMODIFIED reference header built against cuBLAS/cuSOLVER (oracle/_ref/libgputils_ref.so),
             same config, same timing. The reference has no CPU implementation (it is a cuBLAS/cuSOLVER wrapper), so
             its arm runs on the same GPU; the host-LAPACK-style CPU figure is reported as cpu_baseline in both arms.
  N > 1    : the mats axis is sharded, one process per GPU, k matrices per GPU (weak scaling), no collective on the
             data path; NCCL only for the barrier and the max-over-ranks reduction of the timings.
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
# DRAM bytes per 32 x 32 fp64 matrix of k_potrf_group, measured with ncu (profiles/r1h_ncu_chol32.json):
# (1.550705 + 1.151753) GB over 250,000 matrices (k_potrf_pair<double,32>)
POTRF_DRAM_BYTES_PER_MATRIX = 10809.8
SEED_A, SEED_B = 0x5EED0002, 0x5EED0102


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--batch", type=int, default=1_000_000, help="matrices per GPU (BASELINE config 2: 1e6)")
    p.add_argument("--cpu-sample", type=int, default=400_000, help="matrices in the bounded CPU-baseline sample")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


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


def cpu_baseline(A_host, b_host, sample: int):
    """potrf + potrs of the oracle port (oracle/oracle.c, OpenMP) on `sample` matrices of the same workload."""
    sys.path.insert(0, str(REPO / "oracle"))
    import numpy as np
    import oracle_np as oracle
    lib = oracle.clib()
    k = min(sample, A_host.shape[0])
    a = np.array(A_host[:k].numpy(), copy=True)          # (k, n, n) symmetric: layout is irrelevant
    b = np.array(b_host[:k].numpy(), copy=True).reshape(k, N_MAT)
    info = np.zeros(k, dtype=np.int32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    lib.oracle_potrf_batched_f64(C.c_size_t(N_MAT), p(a), p(info), C.c_size_t(k))
    lib.oracle_potrs_batched_f64(C.c_size_t(N_MAT), p(a), p(b), C.c_size_t(k))
    dt = time.perf_counter() - t0
    return {"value": k / dt, "unit": "matrices/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{k} of the same 32x32 fp64 SPD systems, potrf+potrs once, oracle/oracle.c with OpenMP over the batch, {dt:.2f} s"}


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
    if args.impl == "reference" and rank != 0:
        return 0                                         # rank 0 alone runs the reference arm
    torch.cuda.set_device(local_rank)
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from gputils_b200 import capi
    ctx = capi.Context(local_rank)                       # fails loudly if libgputils_b200.so is missing
    k, n = args.batch, N_MAT
    dev = torch.device("cuda", local_rank)

    # synthetic inputs, generated on the owning device (A = G G^T + 32 I, b ~ U[-1, 1]);
    # every rank gets its own shard of the mats axis (different seed offset), shard-resident
    A0 = torch.empty((k, n, n), dtype=torch.float64, device=dev)
    b0 = torch.empty((k, 1, n), dtype=torch.float64, device=dev)
    capi.fill_spd_batched(ctx, A0, 32.0, SEED_A + rank)
    capi.fill_uniform(ctx, b0, -1.0, 1.0, SEED_B + rank)
    A = torch.empty_like(A0); b = torch.empty_like(b0)
    info = torch.zeros(k, dtype=torch.int32, device=dev)

    if args.impl == "reference":
        ref_path = REPO / "oracle" / "_ref" / "libgputils_ref.so"
        if not ref_path.exists():
            real_stdout.write(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgputils_ref.so not built (make -C oracle ref)"}) + "\n")
            real_stdout.flush()
            return 0
        ref = C.CDLL(str(ref_path))

    def restore():
        A.copy_(A0); b.copy_(b0)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step_ours():
        capi.potrf_batched(ctx, A, info)
        capi.potrs_batched(ctx, A, b)

    t_factor, t_solve = [], []

    def timed_step():
        """returns device ms of (factorise, solve); launches on torch's current stream (bound as stream 0)"""
        if args.impl == "ours":
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(); capi.potrf_batched(ctx, A, info)
            e1.record(); capi.potrs_batched(ctx, A, b)
            e2.record()
            return e0, e1, e2
        msf, mss = C.c_float(), C.c_float()
        ref.ref_chol_batch_f64(C.c_size_t(n), C.c_size_t(k), C.c_void_p(A0.data_ptr()), C.c_void_p(A.data_ptr()),
                               C.c_void_p(b0.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(info.data_ptr()), 1,
                               C.byref(msf), C.byref(mss))
        return msf.value, mss.value

    # ---- warm-up -------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        restore()
        timed_step()
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0, "factorisation reported a non-SPD matrix on synthetic SPD input"

    # ---- timed region: exactly K steps ------------------------------------------------------------------------
    sampler = ClockSampler(local_rank, str(torch.cuda.get_device_properties(local_rank).uuid))
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    pending = []
    for _ in range(args.steps):
        restore()
        pending.append(timed_step())
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    clocks = sampler.stop()
    for p in pending:
        if args.impl == "ours":
            t_factor.append(p[0].elapsed_time(p[1])); t_solve.append(p[1].elapsed_time(p[2]))
        else:
            t_factor.append(p[0]); t_solve.append(p[1])
    total_ms = sum(t_factor) + sum(t_solve)
    if use_dist:
        t = torch.tensor([total_ms, sum(t_factor)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, factor_ms_sum = float(t[0]), float(t[1])
    else:
        factor_ms_sum = sum(t_factor)
    ms_per_step = total_ms / args.steps
    value = world * k / (ms_per_step * 1e-3) if args.impl == "ours" else k / (ms_per_step * 1e-3)

    # property check at full size (size-independent): residual of the solve on a slice of the batch
    sl = slice(0, min(k, 4096))
    x = b[sl].transpose(1, 2)                                  # (k, n, 1)
    resid = torch.linalg.norm(torch.bmm(A0[sl].transpose(1, 2), x) - b0[sl].transpose(1, 2)) / torch.linalg.norm(b0[sl])
    assert float(resid) < 1e-12, f"solve residual {float(resid)}"

    # ---- end to end through host buffers --------------------------------------------------------------------
    e2e = None
    A_host = b_host = None
    if not args.no_e2e or not args.no_cpu:
        A_host = torch.empty((k, n, n), dtype=torch.float64, pin_memory=True)
        b_host = torch.empty((k, 1, n), dtype=torch.float64, pin_memory=True)
        A_host.copy_(A0); b_host.copy_(b0)
    if not args.no_e2e:
        x_host = torch.empty((k, 1, n), dtype=torch.float64, pin_memory=True)
        info_host = torch.empty(k, dtype=torch.int32, pin_memory=True)
        e2e_steps = max(1, min(args.steps, 5))

        # ours: the batch is cut into chunks that flow through three streams (pinned host -> device, factorise + solve,
        # device -> pinned host), so the kernels and the download hide behind the upload of the next chunk. The reference
        # API has no such path: upload (tensor.cuh:1128-1145), factorise, solve, download (1147-1154) are whole-tensor calls.
        NCH = 16
        bounds = [(i * k // NCH, (i + 1) * k // NCH) for i in range(NCH)]
        s_up, s_run, s_down = (torch.cuda.Stream(device=dev) for _ in range(3))
        if args.impl == "ours":
            with torch.cuda.stream(s_run):
                ctx.bind_torch_stream(1)                 # stream index 1 of the context = s_run

        def e2e_step():
            if args.impl == "ours":
                cur = torch.cuda.current_stream(dev)
                s_up.wait_stream(cur)
                last = None
                for lo, hi in bounds:
                    if hi <= lo:
                        continue
                    with torch.cuda.stream(s_up):
                        A[lo:hi].copy_(A_host[lo:hi], non_blocking=True); b[lo:hi].copy_(b_host[lo:hi], non_blocking=True)
                        up = torch.cuda.Event(); up.record(s_up)
                    s_run.wait_event(up)
                    with torch.cuda.stream(s_run):
                        Ac, bc, ic = A[lo:hi], b[lo:hi], info[lo:hi]
                        ctx.call("potrf_batched", Ac, n, capi._p(Ac), n, n * n, capi._p(ic), hi - lo, sidx=1)
                        ctx.call("potrs_batched", Ac, n, capi._p(Ac), n, n * n, capi._p(bc), n, hi - lo, sidx=1)
                        done = torch.cuda.Event(); done.record(s_run)
                    s_down.wait_event(done)
                    with torch.cuda.stream(s_down):
                        x_host[lo:hi].copy_(b[lo:hi], non_blocking=True); info_host[lo:hi].copy_(info[lo:hi], non_blocking=True)
                        last = torch.cuda.Event(); last.record(s_down)
                cur.wait_event(last)
                return
            A.copy_(A_host, non_blocking=True); b.copy_(b_host, non_blocking=True)
            if args.impl == "ours":
                step_ours()
            else:
                ref.ref_chol_batch_f64(C.c_size_t(n), C.c_size_t(k), C.c_void_p(A.data_ptr()), C.c_void_p(A.data_ptr()),
                                       C.c_void_p(b.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(info.data_ptr()), 1, None, None)
            x_host.copy_(b, non_blocking=True); info_host.copy_(info, non_blocking=True)

        e2e_step(); torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        s, e = ev(), ev()
        s.record()
        for _ in range(e2e_steps):
            e2e_step()
        e.record(); torch.cuda.synchronize()
        e2e_ms = s.elapsed_time(e) / e2e_steps
        if use_dist:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t[0])
        assert int(info_host.abs().max()) == 0
        h2d = A_host.numel() * 8 + b_host.numel() * 8
        d2h = x_host.numel() * 8 + info_host.numel() * 4
        e2e = {"value": (world if args.impl == "ours" else 1) * k / (e2e_ms * 1e-3), "unit": "matrices/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps}

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(A_host, b_host, args.cpu_sample)

    if rank == 0:
        peak, peak_src = measured_peak()
        s = 8
        potrf_bytes = (2 * n * n * s + 4) * k
        potrs_bytes = (n * n * s + 2 * n * s) * k
        potrf_ms = factor_ms_sum / args.steps
        achieved = potrf_bytes / (potrf_ms * 1e-3) / 1e9
        line = {
            "metric": "batched GEMM/Cholesky/QR matrices/s & %roofline (HBM or FP64 TC), 1-8 B200",
            "value": value, "unit": "matrices/s", "n_gpus": world if args.impl == "ours" else 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "CholeskyBatchFactoriser factorise+solve 32x32 SPD fp64 (BASELINE config 2)",
                       "matrices_per_gpu": k, "n": n, "rhs": 1,
                       "l2": "inputs (8.4 GB) larger than L2; restored from a pristine device copy between steps, outside the timed sub-regions",
                       "timing": "CUDA events on the launching stream around factorise() and solve() of every step, summed; max over ranks",
                       "parallelism": f"mats axis sharded over {world} GPU(s), no data-path collective"},
            "gpu_launches": 2 * args.steps if args.impl == "ours" else None,
            "breakdown": {"factorise_ms": potrf_ms, "solve_ms": (total_ms - factor_ms_sum) / args.steps,
                          "factorise_matrices_per_s": (world if args.impl == "ours" else 1) * k / (potrf_ms * 1e-3),
                          "solve_matrices_per_s": (world if args.impl == "ours" else 1) * k / ((total_ms - factor_ms_sum) / args.steps * 1e-3),
                          "factorise_gflops": (world if args.impl == "ours" else 1) * k * 11440 / (potrf_ms * 1e-3) / 1e9,
                          "solve_hbm_gbs": potrs_bytes / ((total_ms - factor_ms_sum) / args.steps * 1e-3) / 1e9},
            "roofline": {"bound": "hbm", "kernel": "k_potrf_pair<double,32>" if args.impl == "ours" else "cusolverDnDpotrfBatched",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (POTRF_DRAM_BYTES_PER_MATRIX * k if args.impl == "ours" else None),
                         "traffic_source": ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the potrf kernel: "
                                            "profiles/r1h_ncu_chol32.json (only the lower triangle moves)" if args.impl == "ours" else None),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": potrf_bytes,
                         "frac_of_8TBs_nominal": achieved / 8000.0},
            "clocks": clocks,
            "e2e": e2e,
            "cpu_baseline": cpu,
            "solve_residual_rel": float(resid),
        }
        if args.impl == "reference":
            line["impl"] = "reference"
            line["reference"] = "GPUtils include/tensor.cuh (unmodified) + cuBLAS/cuSOLVER 12.9, CholeskyBatchFactoriser::factorise/solve, same GPU"
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
