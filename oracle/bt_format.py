"""Reader / writer for GPUtils' `.bt` binary tensor files (test infrastructure).

Format (ref: tensor.cuh:762-782 reader, 806-816 writer; python/gputils_api/gputils_api.py:48-52):
three little-endian uint64 -- rows, cols, mats -- followed by the raw elements in column-major order
with the mats axis slowest. Arrays are exchanged as numpy (rows, cols, mats), the indexing the
reference's cross-language test pins (testTensor.cu:193-202: numpy [i, j, k] == DTensor(i, j, k)).
Unlike the reference's Python reader this one also orders 2-D (mats == 1) arrays correctly
(SURVEY.md section 8c).
"""
from __future__ import annotations

import numpy as np


def write_bt(path: str, x: np.ndarray) -> None:
    x = np.asarray(x)
    if x.ndim > 3:
        raise ValueError("at most 3 dimensions")
    x3 = x.reshape(x.shape + (1,) * (3 - x.ndim))
    with open(path, "wb") as f:
        np.asarray(x3.shape, dtype="<u8").tofile(f)
        # element (i, j, k) at i + rows * (j + cols * k)
        np.ascontiguousarray(x3.transpose(2, 1, 0)).tofile(f)


def read_bt(path: str, dtype=np.float64) -> np.ndarray:
    with open(path, "rb") as f:
        nr, nc, nm = (int(v) for v in np.fromfile(f, dtype="<u8", count=3))
        data = np.fromfile(f, dtype=dtype, count=nr * nc * nm)
    return data.reshape(nm, nc, nr).transpose(2, 1, 0).copy()


def reference_b_d() -> np.ndarray:
    """The (3, 3, 2) array the reference's Python test writes to b_d.bt (python/test/test.py:9-13, 41):
    B[i, j, k] = 1 + 2 j + 6 i + k."""
    i, j, k = np.meshgrid(np.arange(3), np.arange(3), np.arange(2), indexing="ij")
    return (1 + 2 * j + 6 * i + k).astype(np.float64)
