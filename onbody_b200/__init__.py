"""onbody_b200 - B200-native (sm_100a) implementation of onbody's summation hot path.

The product is the C-ABI shared library ``libonbody_b200.so`` (hand-written CUDA, see ``csrc/`` and
``include/onbody_b200.h``) plus the drop-in shim libraries and C++ drivers built from ``csrc/host``.
This Python package is only the ctypes mirror of that ABI, used by ``bench.py`` and the tests.
There is no CPU fallback: importing works anywhere, but creating a session without the built library
or without a B200 raises.
"""
from .api import GpuSession, OnbodyError, lib_path, load_library, PHYSICS, driver_inputs  # noqa: F401
