"""GPU: compute-sanitizer over the kernels themselves (the reference's CI runs memcheck on its test binary and
on example_main, ci/script.sh:48-74 -- tests/test_gpu_dropin.py does the first, this file the rest).

  memcheck  : example/main.cu compiled unchanged against the new header (ci/script.sh:56-74), and the additive-API program
              (host pipeline, pool, staged copies, batched QR / Givens, Nullspace on another stream) with --leak-check=full;
  synccheck : scripts/dev_sanitize.py (every kernel family on small batches with ragged last warps and padded sizes). The one
              kernel synccheck cannot model, k_potrs_quad128, reaches a NAMED barrier with an explicit thread count from
              role-dependent branches (legal PTX: bar.sync 1, 128) and is excluded BY NAME -- not skipped silently;
  racecheck : the same script. k_potrf_pipe hands columns from the diagonal warp to the panel warps through mbarrier
              arrive (release) / try_wait (acquire), which racecheck does not model (it reports the stores and loads on both
              sides as hazards); it is excluded by name, and its results are bit-checked against the oracle instead
              (tests/test_gpu_parity.py::test_potrf_potrs on batches larger than the resident grid).
"""
import re
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
SAN = "/usr/local/cuda/bin/compute-sanitizer"


def _errors(out: str) -> int:
    """errors (memcheck / synccheck: `ERROR SUMMARY: N errors`) plus, for racecheck, every hazard it displays, warnings included
    (`RACECHECK SUMMARY: H hazards displayed (E errors, W warnings)`)"""
    m = re.findall(r"ERROR SUMMARY: (\d+) error", out)
    h = re.findall(r"RACECHECK SUMMARY: (\d+) hazard", out)
    assert m or h, out[-3000:]
    return sum(int(x) for x in m) + sum(int(x) for x in h)


def _run(args, timeout=1800):
    r = subprocess.run([SAN, *args], capture_output=True, text=True, timeout=timeout, cwd=str(REPO))
    return r.stdout + r.stderr


def test_memcheck_on_the_reference_example_main():
    exe = REPO / "build" / "dropin" / "example_b200"
    if not exe.exists():
        pytest.fail(f"{exe} missing: run `make -C oracle dropin` in the build container")
    out = _run(["--tool", "memcheck", "--leak-check=full", str(exe)], timeout=600)
    assert _errors(out) == 0, out[-4000:]


def test_memcheck_on_the_additive_api_program():
    exe = REPO / "build" / "tests" / "additive_test"
    if not exe.exists():
        pytest.fail(f"{exe} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    out = _run(["--tool", "memcheck", "--leak-check=full", str(exe)], timeout=1800)
    assert "ALL PASSED" in out and _errors(out) == 0, out[-4000:]


def test_synccheck_on_every_kernel_family():
    out = _run(["--tool", "synccheck", "--kernel-name-exclude", "kernel_substring=k_potrs_quad128", sys.executable, "scripts/dev_sanitize.py"])
    assert "done" in out and _errors(out) == 0, out[-4000:]


def test_racecheck_on_every_kernel_family():
    out = _run(["--tool", "racecheck", "--kernel-name-exclude", "kernel_substring=k_potrf_pipe", sys.executable, "scripts/dev_sanitize.py"],
               timeout=2400)
    assert "done" in out and _errors(out) == 0, out[-4000:]


def test_memcheck_and_racecheck_on_the_chunked_svd_and_the_nullspace_build():
    """The library's own side streams: a batched SVD cut into sub-batches (fork / join by events, sliced workspace) and the Nullspace
    build with the packing beside the projector."""
    out = _run(["--tool", "memcheck", sys.executable, "scripts/dev_sanitize_chunked.py"], timeout=900)
    assert "done" in out and _errors(out) == 0, out[-4000:]
    out = _run(["--tool", "racecheck", sys.executable, "scripts/dev_sanitize_chunked.py"], timeout=900)
    assert "done" in out and _errors(out) == 0, out[-4000:]

