"""gputils_b200 -- B200-native (sm_100a) batched linear-algebra kernels behind the GPUtils DTensor API.

The product is the C ABI shared library (gputils_b200/lib/libgputils_b200.so, built from gputils_b200/csrc by
gputils_b200/build.py) and the drop-in C++ header include/tensor.cuh. This Python package only binds the C
ABI with ctypes for tests and bench.py; torch is used for device memory and streams, nothing else.
"""
from . import capi  # noqa: F401
from .capi import Context, GpubError, load  # noqa: F401

__all__ = ["capi", "Context", "GpubError", "load"]
