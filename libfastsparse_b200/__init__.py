"""libfastsparse_b200 -- B200 (sm_100a) implementation of libfastsparse's sparse x dense
hot path, behind the reference's own C API.

The product is the C-ABI shared library `lib/libfastsparse_b200.so` (include/fsb.h) and
the drop-in C headers under include/fastsparse/.  This Python package is only the thin
ctypes mirror of the reference interface used by tests/ and bench.py: same names,
argument order (output first) and error behaviour as csr.h / sparse.h / dsparse.h /
cbcsr.h / cg.h.  There is no CPU fallback: if the library is missing, or no CUDA device
is present, the multiply/solve calls raise.
"""
from ._lib import LIB_PATH, FsbError, check, lib  # noqa: F401
from .api import *  # noqa: F401,F403
