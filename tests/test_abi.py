"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = re.findall(r"\b(bls12_\w+)\s*\(", text)
    consts = re.findall(r"extern const uint64_t (\w+)", text)
    return sorted(set(funcs)), sorted(set(consts))


def test_every_declared_symbol_is_exported(product):
    L = product._native.lib()
    for header in ("eip2537.h", "eip2537_b200.h"):
        funcs, consts = declared_symbols(header)
        assert funcs, header
        for name in funcs + consts:
            assert hasattr(L, name), "%s declares %s but the library does not export it" % (header, name)
    assert set(declared_symbols("eip2537.h")[0]) == set(product._native.ABI_FUNCTIONS)


def test_enum_values_and_gas_schedule(product):
    L = product._native.lib()
    assert (product.SUCCESS, product.POINT_NOT_ON_CURVE, product.POINT_NOT_IN_SUBGROUP, product.INVALID_ELEMENT,
            product.ENCODING_ERROR, product.INVALID_LENGTH, product.EMPTY_INPUT, product.MEMORY_ERROR) == tuple(range(8))
    tab = (ctypes.c_uint64 * 128).in_dll(L, "BLS12_MULTIEXP_DISCOUNT")
    assert tab[0] == 1200 and tab[127] == 174 and ctypes.c_uint64.in_dll(L, "BLS12_MULTIEXP_DISCOUNT_TABLE_LEN").value == 128
    # src/eip2537.c:1199-1271 semantics
    assert L.bls12_g1multiexp_gas(0) == 0 and L.bls12_g1multiexp_gas(159) == 0
    assert L.bls12_g1multiexp_gas(160) == 12000 * 1200 // 1000
    assert L.bls12_g1multiexp_gas(160 * 2) == 2 * 12000 * 888 // 1000
    assert L.bls12_g1multiexp_gas(160 * 500) == 500 * 12000 * 174 // 1000
    assert L.bls12_g2multiexp_gas(288 * 128) == 128 * 55000 * 174 // 1000
    assert L.bls12_pairing_gas(384 * 3) == 3 * 23000 + 115000 and L.bls12_pairing_gas(383) == 0
    assert (L.bls12_g1add_gas(), L.bls12_g1mul_gas(), L.bls12_g2add_gas(), L.bls12_g2mul_gas()) == (600, 12000, 4500, 55000)
    assert (L.bls12_map_fp_to_g1_gas(), L.bls12_map_fp2_to_g2_gas()) == (5500, 110000)


def test_length_checks_do_not_need_a_gpu(product):
    """eip2537.c:543, :831, :1022 -- answered by the C host layer before any device work."""
    g = bytes(160)
    assert product.raw_call("bls12_g1multiexp", b"", 128)[0] == 5
    assert product.raw_call("bls12_g1multiexp", g[:159], 128)[0] == 5
    assert product.raw_call("bls12_g2multiexp", bytes(289), 256)[0] == 5
    assert product.raw_call("bls12_pairing", bytes(385), 32)[0] == 5
    assert product.raw_call("bls12_g1mul", bytes(161), 128)[0] == 5
    assert product.raw_call("bls12_g1add", bytes(255), 128)[0] == 5
    with pytest.raises(product.EIP2537Error) as ei:
        product.G1Multiexp(b"")          # Go wrapper pre-check, go/blst_eip2537.go:71-73
    assert ei.value.code == 5 and str(ei.value) == "invalid length"


def test_no_cpu_fallback_without_cuda(product):
    """Without a CUDA device compute calls must FAIL (MEMORY_ERROR), never silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code, out = product.raw_call("bls12_g1multiexp", bytes(160), 128)
    assert code == 7 and out is None
    assert b"CUDA" in product._native.lib().bls12_b200_last_error() or product._native.lib().bls12_b200_last_error()


def test_product_sources_do_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "blst_eip2537_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "c_oracle" not in text and "py_oracle" not in text and "liboracle" not in text, f
