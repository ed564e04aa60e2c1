"""ropebwt2_b200 -- B200-native engine for ropebwt2's batched multi-string BWT insertion.

The product is ``_build/libropebwt2_b200.so``: hand-written sm_100a CUDA kernels
(``csrc/rb2_engine.cu``) behind a plain-C host layer (``csrc/mrope_b200.c``) that exports the
reference's own ``mrope.h`` / ``rope.h`` API (``include/``).  This Python package is only the
loader used by tests and ``bench.py``: it builds the library in-tree and binds it with
ctypes.  There is no Python or CPU compute path; importing works without a GPU, calling any
engine function without one aborts loudly.
"""
from .binding import load, lib_path, MRope, Engine, Stats  # noqa: F401
