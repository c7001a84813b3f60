"""GPU leg of the golden-vector tests and full-size, size-independent properties (through the C ABI)."""
import numpy as np
import pytest

import py_oracle as po
import vectors
import workloads as wl

pytestmark = pytest.mark.gpu

ABI = {"g1multiexp": ("bls12_g1multiexp", 128), "g2multiexp": ("bls12_g2multiexp", 256), "pairing": ("bls12_pairing", 32),
       "g1mul": ("bls12_g1mul", 128), "g2mul": ("bls12_g2mul", 256),
       "map_fp_to_g1": ("bls12_map_fp_to_g1", 128), "map_fp2_to_g2": ("bls12_map_fp2_to_g2", 256)}


def test_golden_fixtures_on_gpu(product):
    for row in vectors.load_golden():
        name, outlen = ABI[row["Function"]]
        code, out = product.raw_call(name, bytes.fromhex(row["Input"]), outlen)
        if "ExpectedErrorCode" in row:
            assert code == row["ExpectedErrorCode"] and out is None, row["Name"]
        else:
            assert code == 0 and out.hex() == row["Expected"], row["Name"]


def test_reference_vectors_on_gpu_if_supplied(product):
    cases = list(vectors.reference_cases())
    if not cases:
        pytest.skip("reference test_vectors/ not supplied (downloaded by the reference's build.sh)")
    for fname, fn, inp, out, code in cases:
        name, outlen = ABI[fn]
        got_code, got = product.raw_call(name, inp, outlen)
        if code == "any":
            assert got_code != 0, fname
        elif code:
            assert got_code == code, fname
        else:
            assert got_code == 0 and got == out, fname


def test_g1_msm_2_20_closed_form_and_linearity(product, oracle_c):
    """Full BASELINE size: points a_i*G, uniform 256-bit scalars; expected = (sum a_i k_i mod r) * G,
    one scalar multiplication in the oracle.  Plus linearity: MSM(A ++ B) = MSM(A) + MSM(B)."""
    rng = np.random.default_rng(0x2537)
    n = 1 << 20
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    pts = np.frombuffer(product.generator_mul(1, a), dtype=np.uint8).reshape(n, 128)
    # spot-check the generator kernel against the oracle
    for i in (0, 1, n // 2, n - 1):
        assert bytes(pts[i]) == oracle_c.g1_gen_mul(int.from_bytes(bytes(a[i]), "big"))
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    data = np.concatenate([pts, k], axis=1).reshape(-1)
    ai = [int.from_bytes(bytes(r), "big") for r in a]
    ki = [int.from_bytes(bytes(r), "big") for r in k]
    total = sum(x * y for x, y in zip(ai, ki)) % po.R
    full = product.G1Multiexp(data)
    assert full == oracle_c.g1_gen_mul(total)
    half = n // 2
    lo = product.G1Multiexp(data[:160 * half])
    hi = product.G1Multiexp(data[160 * half:])
    assert product.G1Add(lo + hi) == full
    assert lo == oracle_c.g1_gen_mul(sum(x * y for x, y in zip(ai[:half], ki[:half])) % po.R)


def test_g2_msm_2_16_closed_form(product, oracle_c):
    rng = np.random.default_rng(0x2538)
    n = 1 << 16
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    pts = np.frombuffer(product.generator_mul(2, a), dtype=np.uint8).reshape(n, 256)
    assert bytes(pts[5]) == oracle_c.g2_gen_mul(int.from_bytes(bytes(a[5]), "big"))
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    data = np.concatenate([pts, k], axis=1).reshape(-1)
    total = sum(int.from_bytes(bytes(x), "big") * int.from_bytes(bytes(y), "big") for x, y in zip(a, k)) % po.R
    assert product.G2Multiexp(data) == oracle_c.g2_gen_mul(total)


def test_g2_msm_2_18_closed_form_and_linearity(product, oracle_c):
    """BASELINE configs[2] at full size: 2^18 points a_i*G2 with uniform 256-bit scalars; expected = (sum a_i k_i mod r) * G2,
    and MSM(A ++ B) = MSM(A) + MSM(B) through G2ADD (the host path streams the two halves in different chunkings)."""
    rng = np.random.default_rng(0x2539)
    n = 1 << 18
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    pts = np.frombuffer(product.generator_mul(2, a), dtype=np.uint8).reshape(n, 256)
    for i in (0, n // 3, n - 1):
        assert bytes(pts[i]) == oracle_c.g2_gen_mul(int.from_bytes(bytes(a[i]), "big"))
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    data = np.concatenate([pts, k], axis=1).reshape(-1)
    ai = [int.from_bytes(bytes(r), "big") for r in a]
    ki = [int.from_bytes(bytes(r), "big") for r in k]
    full = product.G2Multiexp(data)
    assert full == oracle_c.g2_gen_mul(sum(x * y for x, y in zip(ai, ki)) % po.R)
    cut = n // 2 + 12345
    lo = product.G2Multiexp(data[:288 * cut])
    hi = product.G2Multiexp(data[288 * cut:])
    assert product.G2Add(lo + hi) == full
    assert hi == oracle_c.g2_gen_mul(sum(x * y for x, y in zip(ai[cut:], ki[cut:])) % po.R)


def test_skewed_scalars_do_not_break_bucket_accumulation(product, oracle_c):
    """All scalars equal: every point lands in the same bucket of every window."""
    n = 4096
    data, _ = wl.g1_msm_input(n, 0x77)
    arr = np.frombuffer(data, dtype=np.uint8).reshape(n, 160).copy()
    arr[:, 128:] = np.frombuffer((0xC0FFEE << 200 | 0x1234567).to_bytes(32, "big"), dtype=np.uint8)
    blob = arr.tobytes()
    assert product.G1Multiexp(blob) == oracle_c.call("g1multiexp", blob)[1]


def test_pairing_batch_1024_calls_properties(product, oracle_c):
    """Batch built so call j is a true check unless j % 4 == 3; spot-check calls against the oracle."""
    data, offs, truth = wl.pairing_batch(256, 0x2537 + 4, 2, 16)
    outs, errs = product.PairingBatch(data, offs)
    assert not errs.any()
    assert [bool(o[31]) for o in outs] == truth
    for j in (0, 3, 14, 255):
        assert oracle_c.call("pairing", data[offs[j]:offs[j + 1]]) == (0, bytes(outs[j]))


def test_streamed_chunks_uneven_size_and_error_precedence(product, oracle_c):
    """Host-buffer MULTIEXP streams the input in chunks; sizes that do not divide evenly and errors in
    different chunks must behave exactly like one sequential pass (first failing pair wins)."""
    rng = np.random.default_rng(0x77)
    n = (1 << 16) + 3
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    pts = np.frombuffer(product.generator_mul(1, a), dtype=np.uint8).reshape(n, 128)
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    data = np.concatenate([pts, k], axis=1).copy()
    total = sum(int.from_bytes(bytes(x), "big") * int.from_bytes(bytes(y), "big") for x, y in zip(a, k)) % po.R
    assert product.G1Multiexp(data.reshape(-1)) == oracle_c.g1_gen_mul(total)
    g = po.encode_g1(po.G1)
    off_curve = np.frombuffer(g[:64] + po.fp_to_bytes(5), dtype=np.uint8)
    bad_elem = np.frombuffer(bytes(16) + po.P.to_bytes(48, "big") + g[64:], dtype=np.uint8)
    d2 = data.copy()
    d2[60000, :128] = bad_elem      # last chunk: code 3
    d2[20000, :128] = off_curve     # second chunk: code 1, earlier index -> wins
    assert product.raw_call("bls12_g1multiexp", d2.reshape(-1), 128)[0] == 1
    d3 = data.copy()
    d3[60000, :128] = bad_elem
    assert product.raw_call("bls12_g1multiexp", d3.reshape(-1), 128)[0] == 3


def test_external_known_answers_on_gpu(product):
    """The rows of the golden file that reproduce published values (RFC 9380 J.9.1/J.9.2/J.10.2, geth 2*G1 / 2*G2,
    3*G1; tests/kat.py) through the C ABI on the GPU."""
    import kat
    rows = [r for r in vectors.load_golden() if r.get("ExternalKAT")]
    assert len(rows) >= 6
    for row in rows:
        name, outlen = ABI[row["Function"]]
        code, out = product.raw_call(name, bytes.fromhex(row["Input"]), outlen)
        assert code == 0 and out.hex() == row["Expected"], row["Name"]
    assert product.MapFpToG1(po.fp_to_bytes(kat.G1_NU_U)) == po.encode_g1(kat.G1_NU_P)
    assert product.MapFp2ToG2(po.fp_to_bytes(kat.G2_NU_U[0]) + po.fp_to_bytes(kat.G2_NU_U[1])) == po.encode_g2(kat.G2_NU_P)
    g1b, g2b = po.encode_g1(po.G1), po.encode_g2(po.G2)
    assert product.G1Add(g1b + g1b) == po.encode_g1(kat.TWO_G1)
    assert product.G2Add(g2b + g2b) == po.encode_g2(kat.TWO_G2)
    assert product.G1Mul(g1b + (3).to_bytes(32, "big")) == po.encode_g1(kat.THREE_G1)
