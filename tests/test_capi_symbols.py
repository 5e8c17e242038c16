"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/gputils_b200.h
declares. No compute call is made (there is no GPU here)."""
import ctypes
import re
import subprocess

from conftest import REPO


def _declared():
    text = (REPO / "include" / "gputils_b200.h").read_text()
    return sorted(set(re.findall(r"\b(gpub_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from gputils_b200 import build, capi
    lib_path = build.build()
    assert lib_path.exists()
    lib = ctypes.CDLL(str(lib_path))
    declared = _declared()
    assert len(declared) > 50
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # the ctypes table in gputils_b200/capi.py covers the same set
    assert sorted(capi.EXPORTED) == declared


def test_library_contains_sm100a_code_and_no_cublas_dependency():
    from gputils_b200 import build
    lib_path = build.build()
    elf = subprocess.run(["cuobjdump", "-lelf", str(lib_path)], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    needed = subprocess.run(["objdump", "-p", str(lib_path)], capture_output=True, text=True).stdout
    assert "cublas" not in needed.lower() and "cusolver" not in needed.lower()


def test_version_string_and_argument_errors_without_gpu():
    from gputils_b200 import capi
    lib = capi.load()
    assert b"sm_100a" in lib.gpub_version()
    # null context is rejected before any CUDA call
    assert lib.gpub_ctx_sync(None, 0) == -1
    assert lib.gpub_ctx_device(None) == -1
    # multi-GPU plumbing: argument errors before any CUDA or NCCL call, nothing to release
    assert lib.gpub_multi_allgather(None, 2, 0, None, None, None, 0, None) == -1
    assert lib.gpub_multi_enable_peer_access(None, 0, None) == -1
    assert lib.gpub_multi_release() == 0


def test_header_cites_reference_lines():
    text = (REPO / "include" / "gputils_b200.h").read_text()
    assert len(re.findall(r"ref: tensor\.cuh:\d+", text)) >= 15
