"""ctypes loader for the C oracle (TEST INFRASTRUCTURE ONLY; see eip2537_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It builds liboracle_eip2537.so with gcc on first use.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_eip2537.so")
    srcs = ["eip2537_oracle.c", "oracle_field.h", "oracle_ec.inc", "oracle_constants.h"]
    stale = force or not os.path.exists(so) or any(
        os.path.exists(os.path.join(_HERE, s)) and os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(so)
        for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle_eip2537.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        for name in ("g1mul", "g1multiexp", "g1multiexp_naive", "g1multiexp_bc",
                     "g2mul", "g2multiexp", "g2multiexp_naive", "g2multiexp_bc", "pairing",
                     "map_fp_to_g1", "map_fp2_to_g2"):
            f = getattr(_LIB, "oracle_bls12_" + name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        _LIB.oracle_pairing_gt.restype = ctypes.c_int
        _LIB.oracle_pairing_gt.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        _LIB.oracle_g1_in_subgroup.argtypes = [ctypes.c_char_p, ctypes.c_int]
        _LIB.oracle_g2_in_subgroup.argtypes = [ctypes.c_char_p, ctypes.c_int]
        _LIB.oracle_g1_gen_mul.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        _LIB.oracle_g2_gen_mul.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        _LIB.oracle_g1_gen_mul.restype = None
        _LIB.oracle_g2_gen_mul.restype = None
        _LIB.oracle_g1_arith_progression.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        _LIB.oracle_g2_arith_progression.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        _LIB.oracle_bls12_pairing_batch.restype = None
        _LIB.oracle_bls12_pairing_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_size_t]
    return _LIB


_OUTLEN = {"g1": 128, "g2": 256, "pairing": 32}


def call(name: str, data: bytes):
    """-> (err, out bytes or None); `name` like 'g1multiexp', 'g2multiexp_bc', 'pairing'."""
    outlen = _OUTLEN["pairing" if name == "pairing" else (name[-2:] if name.startswith("map_") else name[:2])]
    out = ctypes.create_string_buffer(outlen)
    err = getattr(lib(), "oracle_bls12_" + name)(out, bytes(data), len(data))
    return err, (out.raw if err == 0 else None)


def pairing_gt(data: bytes):
    out = ctypes.create_string_buffer(576)
    err = lib().oracle_pairing_gt(out, bytes(data), len(data))
    return err, out.raw


def g1_gen_mul(k: int) -> bytes:
    out = ctypes.create_string_buffer(128)
    lib().oracle_g1_gen_mul(out, int(k).to_bytes(32, "big"))
    return out.raw


def g2_gen_mul(k: int) -> bytes:
    out = ctypes.create_string_buffer(256)
    lib().oracle_g2_gen_mul(out, int(k).to_bytes(32, "big"))
    return out.raw


def g1_progression(a: bytes, d: bytes, n: int) -> bytes:
    out = ctypes.create_string_buffer(128 * n)
    assert lib().oracle_g1_arith_progression(out, a, d, n) == 0
    return out.raw


def g2_progression(a: bytes, d: bytes, n: int) -> bytes:
    out = ctypes.create_string_buffer(256 * n)
    assert lib().oracle_g2_arith_progression(out, a, d, n) == 0
    return out.raw
