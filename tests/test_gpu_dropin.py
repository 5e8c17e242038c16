"""GPU: the drop-in proof. The reference's own test/testTensor.cu, main.cu and example/main.cu, compiled UNCHANGED
against include/tensor.cuh + libgputils_b200 (oracle/Makefile target `dropin`, built in the build container and
shipped in build/dropin/), must behave exactly as with the reference header: all 58 gtest cases pass."""
import re
import subprocess

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
BIN = REPO / "build" / "dropin"


def _need(p):
    if not p.exists():
        pytest.fail(f"{p} missing: run `make -C oracle dropin` in the build container (needs /root/reference)")


def test_reference_gtest_suite_passes_against_the_new_header():
    _need(BIN / "gputils_test_b200")
    r = subprocess.run([str(REPO / "scripts" / "run_gtest_binary.sh"), str(BIN / "gputils_test_b200")],
                       capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    m = re.search(r"\[==========\] (\d+) tests ran", out)
    assert m and int(m.group(1)) == 58, out[-3000:]
    assert r.returncode == 0 and "[  PASSED  ] 58 tests." in out, out[-3000:]


def test_reference_main_and_example_run():
    _need(BIN / "main_b200")
    _need(BIN / "example_b200")
    r = subprocess.run([str(BIN / "main_b200")], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "max error : 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([str(BIN / "example_b200")], capture_output=True, text=True, timeout=300, cwd="/tmp")
    vals = re.findall(r"^\s*(-?\d+(?:\.\d+)?),", r.stdout, flags=re.M)
    assert r.returncode == 0 and [float(v) for v in vals] == [39.0, 54.0, 69.0], r.stdout   # example/main.cu:8-12


def test_compute_sanitizer_memcheck_on_the_gtest_suite():
    """ci/script.sh:48-54 runs compute-sanitizer memcheck with leak check and greps for '0 errors'."""
    _need(BIN / "gputils_test_b200")
    r = subprocess.run([str(REPO / "scripts" / "run_gtest_binary.sh"), "/usr/local/cuda/bin/compute-sanitizer",
                        "--tool", "memcheck", "--leak-check=full", str(BIN / "gputils_test_b200")],
                       capture_output=True, text=True, timeout=1800)
    out = r.stdout + r.stderr
    assert "ERROR SUMMARY: 0 errors" in out, out[-4000:]
