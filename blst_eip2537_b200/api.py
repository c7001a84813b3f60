"""Host-side mirror of the reference's language bindings, over the same C ABI.

The reference ships Go (go/blst_eip2537.go:44-211) and Rust (rust/src/lib.rs:98-328) wrappers;
neither toolchain exists in this image, so the same thin wrapper is provided in Python through
ctypes: same function set, same argument meaning (one concatenated byte string in, encoded bytes
out), same error strings as go/blst_eip2537.go:17-42, same "empty input -> invalid length"
pre-check (go:45-47 etc.).  The parity tests call these, and through them the C ABI.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native

SUCCESS, POINT_NOT_ON_CURVE, POINT_NOT_IN_SUBGROUP, INVALID_ELEMENT = 0, 1, 2, 3
ENCODING_ERROR, INVALID_LENGTH, EMPTY_INPUT, MEMORY_ERROR = 4, 5, 6, 7

_ERR_STR = {  # go/blst_eip2537.go:17-42
    0: "Success", 1: "point not on curve", 2: "point not in subgroup", 3: "invalid element",
    4: "encoding error", 5: "invalid length", 6: "empty input", 7: "memory allocation error",
}


class EIP2537Error(Exception):
    def __init__(self, code: int):
        self.code = int(code)
        detail = ""
        if self.code == MEMORY_ERROR:
            try:
                detail = " [%s]" % _native.lib().bls12_b200_last_error().decode()
            except Exception:  # pragma: no cover
                pass
        super().__init__(_ERR_STR.get(self.code, "unknown error condition") + detail)


def _buf(data):
    """bytes / bytearray / numpy uint8 / torch CPU uint8 tensor -> (pointer, length, keepalive)."""
    if isinstance(data, (bytes, bytearray, memoryview)):
        b = bytes(data)
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b), b
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return ctypes.c_void_p(a.ctypes.data), a.size, a
    if hasattr(data, "data_ptr"):  # torch tensor on the CPU (pinned or pageable)
        assert data.device.type == "cpu" and data.is_contiguous()
        return ctypes.c_void_p(data.data_ptr()), data.numel() * data.element_size(), data
    raise TypeError("unsupported buffer type %r" % type(data))


def raw_call(name: str, data, out_len: int):
    """-> (code, out bytes or None): the C ABI call with no wrapper-side checks."""
    ptr, n, keep = _buf(data)
    out = ctypes.create_string_buffer(out_len)
    code = getattr(_native.lib(), name)(out, ptr, n)
    del keep
    return code, (out.raw if code == SUCCESS else None)


def raw_call_into(name: str, data, out_len: int, fill: int = 0xA5):
    """-> (code, out bytes): like raw_call, but `out` is pre-filled with `fill` and returned whatever the
    code, so tests can check the reference's "`out` is written only on success" rule (eip2537.c:613, :701,
    :1072-1078 write last)."""
    ptr, n, keep = _buf(data)
    out = ctypes.create_string_buffer(bytes([fill]) * out_len, out_len)
    code = getattr(_native.lib(), name)(out, ptr, n)
    del keep
    return code, out.raw


def _wrap(name: str, out_len: int):
    def fn(data) -> bytes:
        ptr, n, keep = _buf(data)
        if n == 0:  # the Go wrapper never calls C with an empty slice
            raise EIP2537Error(INVALID_LENGTH)
        out = ctypes.create_string_buffer(out_len)
        code = getattr(_native.lib(), name)(out, ptr, n)
        del keep
        if code != SUCCESS:
            raise EIP2537Error(code)
        return out.raw
    fn.__name__ = name
    return fn


G1Add = _wrap("bls12_g1add", 128)
G1Mul = _wrap("bls12_g1mul", 128)
G1Multiexp = _wrap("bls12_g1multiexp", 128)
G1MultiexpNaive = _wrap("bls12_g1multiexp_naive", 128)
G1MultiexpBosCoster = _wrap("bls12_g1multiexp_bc", 128)
G2Add = _wrap("bls12_g2add", 256)
G2Mul = _wrap("bls12_g2mul", 256)
G2Multiexp = _wrap("bls12_g2multiexp", 256)
G2MultiexpNaive = _wrap("bls12_g2multiexp_naive", 256)
G2MultiexpBosCoster = _wrap("bls12_g2multiexp_bc", 256)
Pairing = _wrap("bls12_pairing", 32)
MapFpToG1 = _wrap("bls12_map_fp_to_g1", 128)
MapFp2ToG2 = _wrap("bls12_map_fp2_to_g2", 256)


# ---- additive batch API (include/eip2537_b200.h) -------------------------------------------
def PairingBatch(data, offsets):
    """n independent PAIRING calls in one submission.

    data: concatenated call inputs (host buffer); offsets: n+1 byte offsets.
    Returns (outs uint8[n,32], errs int32[n]); per-call semantics identical to Pairing().
    """
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offs.size - 1
    outs = np.zeros((n, 32), dtype=np.uint8)
    errs = np.zeros(n, dtype=np.int32)
    ptr, _, keep = _buf(data)
    code = _native.lib().bls12_pairing_batch(outs.ctypes.data, errs.ctypes.data, ptr, offs.ctypes.data, n)
    del keep
    if code != SUCCESS:
        raise EIP2537Error(code)
    return outs, errs


def MultiexpBatch(group: int, data, offsets):
    """n independent G1 (group 1) / G2 (group 2) MULTIEXP calls in one submission.

    Returns (outs uint8[n, 128 or 256], errs int32[n]); per-call semantics identical to G{1,2}Multiexp().
    """
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offs.size - 1
    outs = np.zeros((n, 128 if group == 1 else 256), dtype=np.uint8)
    errs = np.zeros(n, dtype=np.int32)
    ptr, _, keep = _buf(data)
    fn = _native.lib().bls12_g1multiexp_batch if group == 1 else _native.lib().bls12_g2multiexp_batch
    code = fn(outs.ctypes.data, errs.ctypes.data, ptr, offs.ctypes.data, n)
    del keep
    if code != SUCCESS:
        raise EIP2537Error(code)
    return outs, errs


def MapBatch(group: int, data):
    """n independent MAP_FP_TO_G1 (group 1, 64-byte elements) / MAP_FP2_TO_G2 (group 2, 128-byte) calls.

    Returns (outs uint8[n, 128 or 256], errs int32[n]); errs[j] is 0 or INVALID_ELEMENT (3).
    """
    ptr, nbytes, keep = _buf(data)
    width = 64 if group == 1 else 128
    if nbytes % width:
        raise EIP2537Error(INVALID_LENGTH)
    n = nbytes // width
    outs = np.zeros((n, 2 * width), dtype=np.uint8)
    errs = np.zeros(n, dtype=np.int32)
    fn = _native.lib().bls12_map_fp_to_g1_batch if group == 1 else _native.lib().bls12_map_fp2_to_g2_batch
    code = fn(outs.ctypes.data, errs.ctypes.data, ptr, n)
    del keep
    if code != SUCCESS:
        raise EIP2537Error(code)
    return outs, errs


def generator_mul(group: int, scalars) -> bytes:
    """out[i] = encode(k_i * generator) for 32-byte big-endian scalars (synthetic workloads)."""
    ptr, n, keep = _buf(scalars)
    assert n % 32 == 0
    cnt = n // 32
    out = np.zeros(cnt * (128 if group == 1 else 256), dtype=np.uint8)
    fn = _native.lib().bls12_b200_g1_generator_mul if group == 1 else _native.lib().bls12_b200_g2_generator_mul
    code = fn(out.ctypes.data, ptr, cnt)
    del keep
    if code != SUCCESS:
        raise EIP2537Error(code)
    return out.tobytes()


def launch_count() -> int:
    return int(_native.lib().bls12_b200_launch_count())


def set_window(c: int) -> None:
    _native.lib().bls12_b200_set_window(int(c))


def points_check(group: int, data, stride: int = 0, check_subgroup: bool = True):
    """Batched validation of encoded points -> int32 codes (0 ok, 3 invalid element, 1 off curve, 2 not in subgroup)."""
    ptr, nbytes, keep = _buf(data)
    stride = stride or (128 if group == 1 else 256)
    assert nbytes % stride == 0 and stride % 16 == 0
    n = nbytes // stride
    codes = np.zeros(n, dtype=np.int32)
    code = _native.lib().bls12_b200_points_check(group, ptr, n, stride, 1 if check_subgroup else 0, codes.ctypes.data)
    del keep
    if code != SUCCESS:
        raise EIP2537Error(code)
    return codes


def set_pairing_coop_max(n_calls: int) -> int:
    """Batches of <= n_calls PAIRING calls use the warp-per-call kernel (-1: default 128); returns the old value."""
    return int(_native.lib().bls12_b200_set_pairing_coop_max(n_calls))


def set_checked_msm(on: bool) -> None:
    _native.lib().bls12_b200_set_checked_msm(1 if on else 0)
