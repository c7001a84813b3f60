"""Loader for libblst_eip2537.so -- the product's C-ABI library (plain-C host layer + sm_100a engine).

The library is built in-tree by `make -C blst_eip2537_b200/csrc` (nvcc cross-compiles without a
GPU).  There is NO Python or CPU fallback: if the shared library is missing or fails to load,
importing this module's `lib()` raises, and every compute entry point returns
EIP2537_MEMORY_ERROR when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libblst_eip2537.so")
_LIB = None

# every symbol include/eip2537.h and include/eip2537_b200.h declare
ABI_FUNCTIONS = [
    "bls12_g1add", "bls12_g1mul", "bls12_g1multiexp", "bls12_g1multiexp_naive", "bls12_g1multiexp_bc",
    "bls12_g2add", "bls12_g2mul", "bls12_g2multiexp", "bls12_g2multiexp_naive", "bls12_g2multiexp_bc",
    "bls12_pairing", "bls12_map_fp_to_g1", "bls12_map_fp2_to_g2",
    "bls12_g1add_gas", "bls12_g1mul_gas", "bls12_g1multiexp_gas", "bls12_g2add_gas", "bls12_g2mul_gas",
    "bls12_g2multiexp_gas", "bls12_pairing_gas", "bls12_map_fp_to_g1_gas", "bls12_map_fp2_to_g2_gas",
]
ABI_CONSTANTS = [
    "BLS12_G1ADD_GAS", "BLS12_G1MUL_GAS", "BLS12_G2ADD_GAS", "BLS12_G2MUL_GAS", "BLS12_PAIRING_BASE_GAS",
    "BLS12_PAIRING_PAIR_GAS", "BLS12_MAP_FP_TO_G1_GAS", "BLS12_MAP_FP2_TO_G2_GAS",
    "BLS12_MULTIEXP_MULTIPLIER_GAS", "BLS12_MULTIEXP_DISCOUNT_TABLE_LEN", "BLS12_MULTIEXP_DISCOUNT",
]
EXT_FUNCTIONS = [
    "bls12_b200_init", "bls12_b200_shutdown", "bls12_b200_last_error", "bls12_b200_launch_count",
    "bls12_b200_set_window", "bls12_pairing_batch", "bls12_g1multiexp_batch", "bls12_g2multiexp_batch", "bls12_b200_msm_device", "bls12_b200_msm_partial_device", "bls12_b200_msm_partial_host",
    "bls12_b200_msm_combine_device", "bls12_b200_pairing_batch_device", "bls12_b200_g1_generator_mul",
    "bls12_b200_g2_generator_mul", "bls12_b200_fp_microbench", "bls12_b200_selftest", "bls12_b200_points_check", "bls12_b200_points_check_device", "bls12_b200_set_checked_msm", "bls12_b200_set_pairing_coop_max", "bls12_map_fp_to_g1_batch", "bls12_map_fp2_to_g2_batch", "bls12_b200_partial_bytes", "bls12_b200_set_profile", "bls12_b200_last_msm_profile", "bls12_b200_last_msm_work", "bls12_b200_last_pairing_profile",
    "bls12_b200_init_multi", "bls12_b200_multi_gpus", "bls12_b200_device_launch_count", "bls12_b200_comm_unique_id", "bls12_b200_comm_init",
    "bls12_b200_comm_destroy", "bls12_b200_msm_sharded_device", "bls12_b200_msm_sharded_host", "bls12_b200_last_pairing_chunk",
]


def build(verbose: bool = False) -> str:
    """Compile the CUDA engine for sm_100a in-tree (make decides what is stale)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC], stdout=out)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("build finished but %s is missing" % LIB_PATH)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s not found: build it with `make -C %s` (or __graft_entry__.build()). "
            "There is no CPU fallback." % (LIB_PATH, CSRC))
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, u64, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
    for name in ABI_FUNCTIONS[:13]:
        f = getattr(L, name)
        f.restype = i32
        f.argtypes = [vp, vp, sz]
    for name in ABI_FUNCTIONS[13:]:
        f = getattr(L, name)
        f.restype = u64
        f.argtypes = [u64] if name in ("bls12_g1multiexp_gas", "bls12_g2multiexp_gas", "bls12_pairing_gas") else []
    L.bls12_b200_init.restype = i32
    L.bls12_b200_init.argtypes = [i32]
    L.bls12_b200_shutdown.restype = None
    L.bls12_b200_last_error.restype = ctypes.c_char_p
    L.bls12_b200_launch_count.restype = u64
    L.bls12_b200_set_window.argtypes = [i32]
    L.bls12_b200_set_window.restype = None
    L.bls12_pairing_batch.restype = i32
    L.bls12_pairing_batch.argtypes = [vp, vp, vp, vp, sz]
    for nm in ("bls12_g1multiexp_batch", "bls12_g2multiexp_batch"):
        getattr(L, nm).restype = i32
        getattr(L, nm).argtypes = [vp, vp, vp, vp, sz]
    for nm in ("bls12_map_fp_to_g1_batch", "bls12_map_fp2_to_g2_batch"):
        getattr(L, nm).restype = i32
        getattr(L, nm).argtypes = [vp, vp, vp, sz]
    L.bls12_b200_msm_device.restype = i32
    L.bls12_b200_msm_device.argtypes = [i32, vp, sz, vp, vp, vp]
    L.bls12_b200_msm_partial_device.restype = i32
    L.bls12_b200_msm_partial_device.argtypes = [i32, vp, sz, u64, vp, vp, vp]
    L.bls12_b200_msm_combine_device.restype = i32
    L.bls12_b200_msm_combine_device.argtypes = [i32, vp, i32, vp, vp]
    L.bls12_b200_pairing_batch_device.restype = i32
    L.bls12_b200_pairing_batch_device.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.bls12_b200_g1_generator_mul.restype = i32
    L.bls12_b200_g1_generator_mul.argtypes = [vp, vp, sz]
    L.bls12_b200_g2_generator_mul.restype = i32
    L.bls12_b200_g2_generator_mul.argtypes = [vp, vp, sz]
    L.bls12_b200_fp_microbench.restype = i32
    L.bls12_b200_fp_microbench.argtypes = [i32, sz, i32, ctypes.POINTER(ctypes.c_float), vp]
    L.bls12_b200_selftest.restype = i32
    L.bls12_b200_selftest.argtypes = [vp, sz]
    L.bls12_b200_set_profile.restype = None
    L.bls12_b200_set_profile.argtypes = [i32]
    L.bls12_b200_last_msm_profile.restype = i32
    L.bls12_b200_last_msm_profile.argtypes = [vp, vp]
    L.bls12_b200_last_msm_work.restype = i32
    L.bls12_b200_last_msm_work.argtypes = [vp]
    L.bls12_b200_partial_bytes.restype = sz
    L.bls12_b200_partial_bytes.argtypes = [i32]
    L.bls12_b200_last_pairing_profile.restype = i32
    L.bls12_b200_last_pairing_profile.argtypes = [vp]
    L.bls12_b200_points_check.restype = i32
    L.bls12_b200_points_check.argtypes = [i32, vp, sz, sz, i32, vp]
    L.bls12_b200_points_check_device.restype = i32
    L.bls12_b200_points_check_device.argtypes = [i32, vp, sz, sz, i32, vp, vp]
    L.bls12_b200_set_checked_msm.restype = None
    L.bls12_b200_set_checked_msm.argtypes = [i32]
    L.bls12_b200_set_pairing_coop_max.restype = ctypes.c_long
    L.bls12_b200_set_pairing_coop_max.argtypes = [ctypes.c_long]
    L.bls12_b200_msm_partial_host.restype = i32
    L.bls12_b200_msm_partial_host.argtypes = [i32, vp, sz, u64, vp, vp]
    L.bls12_b200_init_multi.restype = i32
    L.bls12_b200_init_multi.argtypes = [i32]
    L.bls12_b200_multi_gpus.restype = i32
    L.bls12_b200_device_launch_count.restype = u64
    L.bls12_b200_device_launch_count.argtypes = [i32]
    L.bls12_b200_comm_unique_id.restype = i32
    L.bls12_b200_comm_unique_id.argtypes = [vp]
    L.bls12_b200_comm_init.restype = i32
    L.bls12_b200_comm_init.argtypes = [i32, i32, vp]
    L.bls12_b200_comm_destroy.restype = None
    L.bls12_b200_msm_sharded_device.restype = i32
    L.bls12_b200_msm_sharded_device.argtypes = [i32, vp, sz, u64, vp, vp, vp]
    L.bls12_b200_msm_sharded_host.restype = i32
    L.bls12_b200_msm_sharded_host.argtypes = [i32, vp, sz, u64, vp]
    L.bls12_b200_last_pairing_chunk.restype = i32
    _LIB = L
    return L
