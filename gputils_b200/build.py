"""Builds libgputils_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The shared library is the product: it travels to the GPU box with the repo snapshot (it is git-ignored,
not gpurun-ignored). Nothing here falls back to a CPU or library path -- if nvcc is missing the build
fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = LIBDIR / "obj"
# tuning aid: GPUB_VARIANT=name + GPUB_EXTRA_NVCC_FLAGS="-DX=.." builds lib/variants/libgputils_b200_<name>.so
VARIANT = os.environ.get("GPUB_VARIANT", "")
if VARIANT:
    OBJDIR = LIBDIR / "variants" / ("obj_" + VARIANT)
    LIB = LIBDIR / "variants" / f"libgputils_b200_{VARIANT}.so"
else:
    LIB = LIBDIR / "libgputils_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", str(REPO / "include"), "-I", str(CSRC)]
# per-file extra flags: the LAPACK-faithful SVD core must round exactly like its host build
EXTRA = {"svd.cu": ["--fmad=false"]}
SOURCES = ["ctx.cu", "mem.cu", "blas1.cu", "gemm.cu", "chol.cu", "qr.cu", "svd.cu", "multi.cu"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libgputils_b200 cannot be built (there is no CPU fallback)")
    return exe


def _newest_header() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((REPO / "include").glob("*.h"))
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: str, force: bool, verbose: bool) -> Path:
    obj = OBJDIR / (src + ".o")
    srcp = CSRC / src
    if not force and obj.exists() and obj.stat().st_mtime > max(srcp.stat().st_mtime, _newest_header()):
        return obj
    cmd = [nvcc(), *ARCH, *COMMON, *EXTRA.get(src, []), *os.environ.get("GPUB_EXTRA_NVCC_FLAGS", "").split(), "-c", str(srcp), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJDIR.mkdir(parents=True, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    if force or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [nvcc(), *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
