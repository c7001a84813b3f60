"""blst_eip2537_b200 -- B200-native engine for the EIP-2537 MULTIEXP / PAIRING hot path.

The product is the C-ABI library `libblst_eip2537.so` (plain-C host layer + hand-written sm_100a
kernels, see csrc/ and include/eip2537.h).  This package is only its loader and a ctypes mirror of
the reference's Go/Rust wrappers; it contains no arithmetic and no CPU fallback.
"""
from . import _native  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (EIP2537Error, MapBatch, MultiexpBatch, PairingBatch, generator_mul, launch_count, points_check, raw_call, raw_call_into,  # noqa: F401
                  set_checked_msm, set_pairing_coop_max, set_window)
