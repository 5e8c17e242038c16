import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(REPO / "tests" / "golden" / "reference_vectors.json") as f:
        return json.load(f)


def mats(entry, dtype=np.float64):
    """golden entry {dims:[m,n,k], data:[...]} (+ layout) -> numpy batch (k, m, n)."""
    m, n, k = entry["dims"]
    a = np.asarray(entry["data"], dtype=dtype)
    layout = entry.get("layout", "cm")
    if layout.startswith("rm"):
        return a.reshape(k, m, n).copy()
    return a.reshape(k, n, m).transpose(0, 2, 1).copy()


def with_layout(group, key):
    e = dict(group[key])
    e.setdefault("layout", group.get("layout", "cm"))
    return e


@pytest.fixture(scope="session")
def oracle():
    import oracle_np
    oracle_np.build_c_oracle()
    return oracle_np


@pytest.fixture(scope="session")
def gpu_ctx():
    """The C-ABI context on cuda:0. Fails loudly (no fallback) if the CUDA library is not built."""
    import torch
    from gputils_b200 import capi
    assert torch.cuda.is_available(), "-m gpu tests need a GPU"
    capi.load()
    return capi.Context(0)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
