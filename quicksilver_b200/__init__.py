"""quicksilver_b200 -- B200-native cycle-tracking hot path for Quicksilver-class Monte Carlo transport.

Layout: csrc/ holds the C++ host model, the sm_100a CUDA kernels and the C ABI (include/qsb.h),
built into libqsb.so next to this file; the Python modules are thin ctypes mirrors used by the tests,
bench.py and the multi-GPU driver (torch.distributed is only plumbing).
"""
from ._capi import BAL, BAL_COUNT, BAL_NAMES, EXCHANGE_DTYPE, PARTICLE_DTYPE, QsbError  # noqa: F401
from . import decks  # noqa: F401
