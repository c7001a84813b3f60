"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle (bit-exact bytes and codes).

Mirrors the reference's own test semantics: exact-byte success cases,
specific error codes per failure class, dispatch/naive/bc entry points agree.
"""
import random

import numpy as np
import pytest

import py_oracle as po
import workloads as wl

pytestmark = pytest.mark.gpu

G1B = po.encode_g1(po.G1)
G2B = po.encode_g2(po.G2)
ONES = (2 ** 256 - 1).to_bytes(32, "big")


def be32(k):
    return int(k).to_bytes(32, "big")


def test_library_reports_cuda(product):
    assert product._native.lib().bls12_b200_init(-1) == 0, product._native.lib().bls12_b200_last_error()


def test_ptx_field_ops_match_portable_on_device(product):
    m = np.zeros(4, dtype=np.uint64)
    assert product._native.lib().bls12_b200_selftest(m.ctypes.data, 1 << 16) == 0
    assert list(m) == [0, 0, 0, 0], "mismatches mul/add/sub/inv: %s" % list(m)


def test_fp_mul_ptx_chain_matches_bigint(product):
    """k_fp_chain: x0 = (R mod p) ^ tid-bit, x <- x*y (Montgomery) `iters` times; thread 0 digest."""
    import ctypes
    L = product._native.lib()
    ms = ctypes.c_float()
    dig = ctypes.create_string_buffer(48)
    iters = 100
    assert L.bls12_b200_fp_microbench(0, 256, iters, ctypes.byref(ms), dig) == 0
    P = po.P
    Rm = (1 << 384) % P
    x = Rm  # montgomery one, thread 0 xor 0
    y = po.G1_X * Rm % P
    rinv = pow(1 << 384, -1, P)
    for _ in range(iters):
        x = x * y * rinv % P
    assert int.from_bytes(dig.raw, "little") == x


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 17, 128, 1000])
def test_g1_multiexp_random_matches_oracle(product, oracle_c, n):
    data, _ = wl.g1_msm_input(n, 0x2537 + n)
    err, ref = oracle_c.call("g1multiexp", data)
    assert err == 0
    assert product.G1Multiexp(data) == ref
    assert product.G1MultiexpNaive(data) == ref
    assert product.G1MultiexpBosCoster(data) == ref
    if n == 1:
        assert product.G1Mul(data) == ref


@pytest.mark.parametrize("c", [3, 4, 7, 8, 11, 13, 16])
def test_g1_multiexp_all_window_widths(product, oracle_c, c):
    data, _ = wl.g1_msm_input(300, 0x9999)
    _, ref = oracle_c.call("g1multiexp", data)
    product.set_window(c)
    try:
        assert product.G1Multiexp(data) == ref
    finally:
        product.set_window(0)


def test_g1_multiexp_closed_form_2_14(product, oracle_c):
    n = 1 << 14
    data, s = wl.g1_msm_input(n, 0x2537 + 1)
    assert product.G1Multiexp(data) == oracle_c.g1_gen_mul(s)
    err, ref = oracle_c.call("g1multiexp", data)   # Bos-Coster on the CPU, a few seconds
    assert err == 0 and ref == oracle_c.g1_gen_mul(s)


def test_g1_multiexp_adversarial(product, oracle_c):
    neg_g = po.encode_g1(po.ec_neg(po.FP_OPS, po.G1))
    order3 = po.encode_g1((0, 2))   # on E(Fp), NOT in G1: legal for MULTIEXP (eip2537.c:340)
    data = (G1B + ONES + bytes(128) + ONES + G1B + bytes(32) + G1B + be32(5) + neg_g + be32(5)
            + G1B + be32(7) + G1B + be32(7) + order3 + ONES + order3 + be32(po.R + 12345)
            + G1B + be32(po.R) + G1B + be32(po.R - 1) + G1B + be32(1 << 255))
    err, ref = oracle_c.call("g1multiexp", data)
    assert err == 0
    assert po.bls12_g1multiexp(data) == (0, ref)
    for c in (0, 4, 8, 16):
        product.set_window(c)
        try:
            assert product.G1Multiexp(data) == ref
        finally:
            product.set_window(0)
    # all-infinity / zero scalars -> 128 zero bytes
    assert product.G1Multiexp(bytes(160) * 3) == bytes(128)
    assert product.G1Multiexp((G1B + bytes(32)) * 5) == bytes(128)
    # same point many times with the same scalar: one bucket per window takes everything
    rep = (G1B + be32(0xDEADBEEF12345)) * 200
    assert product.G1Multiexp(rep) == oracle_c.call("g1multiexp", rep)[1]
    # P and -P with equal scalars cancel
    assert product.G1Multiexp(G1B + ONES + neg_g + ONES) == bytes(128)


def test_g1_error_codes_and_precedence(product, oracle_c):
    bad_pad = bytearray(G1B); bad_pad[3] = 1
    ge_p = bytes(16) + po.P.to_bytes(48, "big") + G1B[64:]
    off_curve = G1B[:64] + po.fp_to_bytes(5)
    cases = [
        (b"", 5), (G1B + ONES[:31], 5), (G1B + ONES + b"\x00", 5),
        (bytes(bad_pad) + ONES, 3), (ge_p + ONES, 3), (off_curve + ONES, 1),
        # first failing pair decides: off-curve (1) at index 1 beats invalid element (3) at index 2
        (G1B + ONES + off_curve + ONES + ge_p + ONES, 1),
        (G1B + ONES + ge_p + ONES + off_curve + ONES, 3),
        # within a point INVALID_ELEMENT beats NOT_ON_CURVE: x invalid, y makes it off-curve anyway
        (ge_p[:64] + po.fp_to_bytes(5) + ONES, 3),
        # x = 0, y != 0 is not infinity: goes to the curve test
        (po.fp_to_bytes(0) + po.fp_to_bytes(3) + ONES, 1),
    ]
    for data, code in cases:
        for name in ("bls12_g1multiexp", "bls12_g1multiexp_naive", "bls12_g1multiexp_bc"):
            got, out = product.raw_call(name, data, 128)
            assert got == code, (name, len(data), got, code)
            assert out is None
        if data:
            assert oracle_c.call("g1multiexp", data)[0] == code
    # 6 pairs with the error late in a long input
    many = (G1B + ONES) * 50 + off_curve + ONES + (G1B + ONES) * 10
    assert product.raw_call("bls12_g1multiexp", many, 128)[0] == 1
    assert product.raw_call("bls12_g1mul", G1B, 128)[0] == 5


@pytest.mark.parametrize("n", [1, 2, 4, 5, 33, 300])
def test_g2_multiexp_random_matches_oracle(product, oracle_c, n):
    data, _ = wl.g2_msm_input(n, 0x2537 + 100 + n)
    err, ref = oracle_c.call("g2multiexp", data)
    assert err == 0
    assert product.G2Multiexp(data) == ref
    assert product.G2MultiexpNaive(data) == ref
    assert product.G2MultiexpBosCoster(data) == ref
    if n == 1:
        assert product.G2Mul(data) == ref


def test_g2_adversarial_and_errors(product, oracle_c):
    neg = po.encode_g2(po.ec_neg(po.F2_OPS, po.G2))
    data = G2B + ONES + bytes(256) + ONES + G2B + bytes(32) + neg + be32(9) + G2B + be32(9) + G2B + be32(po.R + 5)
    assert product.G2Multiexp(data) == oracle_c.call("g2multiexp", data)[1]
    assert product.G2Multiexp(G2B + ONES + neg + ONES) == bytes(256)
    off = G2B[:192] + po.fp_to_bytes(7)
    bad = bytes(16) + po.P.to_bytes(48, "big") + G2B[64:]
    for d, code in ((b"", 5), (G2B + ONES[:30], 5), (off + ONES, 1), (bad + ONES, 3), (G2B + ONES + off + ONES + bad + ONES, 1)):
        assert product.raw_call("bls12_g2multiexp", d, 256)[0] == code


def test_add_entry_points(product):
    two_g = po.encode_g1(po.ec_add(po.FP_OPS, po.G1, po.G1))
    assert product.G1Add(G1B + G1B) == two_g
    assert product.G1Add(G1B + bytes(128)) == G1B
    assert product.G1Add(G1B + po.encode_g1(po.ec_neg(po.FP_OPS, po.G1))) == bytes(128)
    assert product.G2Add(G2B + G2B) == po.encode_g2(po.ec_add(po.F2_OPS, po.G2, po.G2))
    assert product.raw_call("bls12_g1add", G1B, 128)[0] == 5


def test_pairing_true_false_and_gt_semantics(product, oracle_c):
    rng = wl.SplitMix64(77)
    for k in (1, 2, 3, 5):
        for truth in (True, False):
            data = wl.pairing_call(k, rng, truth) if k > 1 else (G1B + G2B if not truth else bytes(128) + G2B)
            err, ref = oracle_c.call("pairing", data)
            assert err == 0
            assert product.Pairing(data) == ref
    # infinity members contribute 1
    assert product.Pairing(bytes(384)) == bytes(31) + b"\x01"
    assert product.Pairing(G1B + bytes(256)) == bytes(31) + b"\x01"
    e_true = G1B + G2B + po.encode_g1(po.ec_neg(po.FP_OPS, po.G1)) + G2B
    assert product.Pairing(e_true) == bytes(31) + b"\x01"
    assert product.Pairing(e_true + bytes(384)) == bytes(31) + b"\x01"


def test_pairing_error_codes_and_order(product, oracle_c):
    order3 = po.encode_g1((0, 2))
    off1 = G1B[:64] + po.fp_to_bytes(5)
    off2 = G2B[:192] + po.fp_to_bytes(7)
    bad1 = bytes(16) + po.P.to_bytes(48, "big") + G1B[64:]
    cases = [
        (b"", 5), (G1B + G2B[:255], 5),
        (order3 + G2B, 2),            # on curve, not in G1 (invalid_subgroup_for_pairing semantics)
        (off1 + G2B, 1), (bad1 + G2B, 3), (G1B + off2, 1),
        (order3 + off2, 2),           # G1 subgroup failure comes before G2 decode
        (off1 + off2, 1),
        (G1B + G2B + order3 + G2B, 2),
        (G1B + off2 + order3 + G2B, 1),   # first failing pair decides
    ]
    for data, code in cases:
        got, out = product.raw_call("bls12_pairing", data, 32)
        assert got == code, (len(data), got, code)
        assert out is None
        if data:
            assert oracle_c.call("pairing", data)[0] == code


def test_pairing_batch_matches_oracle(product, oracle_c):
    data, offs, truth = wl.pairing_batch(40, 0x2537 + 4, kmin=2, kmax=6)
    outs, errs = product.PairingBatch(data, offs)
    assert list(errs) == [0] * 40
    assert [int(o[31]) for o in outs] == [1 if t else 0 for t in truth]
    for j in range(40):
        assert oracle_c.call("pairing", data[offs[j]:offs[j + 1]]) == (0, bytes(outs[j]))
    # a batch with failing and malformed calls keeps per-call codes
    order3 = po.encode_g1((0, 2))
    calls = [data[offs[0]:offs[1]], order3 + G2B, b"", G1B + G2B[:100], data[offs[1]:offs[2]]]
    blob = b"".join(calls)
    o2 = [0]
    for cdata in calls:
        o2.append(o2[-1] + len(cdata))
    outs, errs = product.PairingBatch(blob, o2)
    assert list(errs) == [0, 2, 5, 5, 0]
    assert bytes(outs[0]) == oracle_c.call("pairing", calls[0])[1]
    assert bytes(outs[4]) == oracle_c.call("pairing", calls[4])[1]
    assert bytes(outs[1]) == bytes(32)


def test_pairing_warp_and_thread_kernels_agree(product, oracle_c):
    """The warp-cooperative (coop12.cuh) and thread-per-chunk pairing kernels give the same bytes and
    codes as the oracle, on a batch mixing true/false products, k = 1..9 pairs, infinity inputs and a
    failing call."""
    data, offs, truth = wl.pairing_batch(24, 0x2537 + 9, kmin=1, kmax=9)
    inf1, inf2 = bytes(128), bytes(256)
    order3 = po.encode_g1((0, 2))
    extra = [inf1 + G2B, G1B + inf2, inf1 + inf2 + G1B + G2B, order3 + G2B,
             G1B + G2B + po.encode_g1(po.ec_neg(po.FP_OPS, po.G1)) + G2B]
    calls = [data[offs[j]:offs[j + 1]] for j in range(24)] + extra
    blob = b"".join(calls)
    o2 = [0]
    for cdata in calls:
        o2.append(o2[-1] + len(cdata))
    want = [oracle_c.call("pairing", cdata) for cdata in calls]
    assert [w[0] for w in want[24:]] == [0, 0, 0, 2, 0]
    assert want[-1][1][31] == 1 and want[26][1][31] == 0
    results = {}
    for name, thr in (("warp", 1 << 40), ("thread", 0)):
        old = product.set_pairing_coop_max(thr)
        try:
            outs, errs = product.PairingBatch(blob, o2)
        finally:
            product.set_pairing_coop_max(old)
        results[name] = (list(errs), [bytes(o) for o in outs])
        for j, (code, out) in enumerate(want):
            assert errs[j] == code, (name, j, errs[j], code)
            assert bytes(outs[j]) == (out if code == 0 else bytes(32)), (name, j)
    assert results["warp"] == results["thread"]
    # small batch of LONG calls: chunks accumulated in parallel, the warp multiplies the chunk values
    rng = wl.SplitMix64(0x2537 + 10)
    long_calls = [wl.pairing_call(30, rng, True), wl.pairing_call(17, rng, False), wl.pairing_call(64, rng, True)]
    blob = b"".join(long_calls)
    o3 = [0]
    for cdata in long_calls:
        o3.append(o3[-1] + len(cdata))
    outs, errs = product.PairingBatch(blob, o3)
    assert list(errs) == [0, 0, 0] and [int(o[31]) for o in outs] == [1, 0, 1]
    for j, cdata in enumerate(long_calls):
        assert oracle_c.call("pairing", cdata) == (0, bytes(outs[j]))
        assert product.Pairing(cdata) == bytes(outs[j])


def test_g2_subgroup_rejection_in_pairing(product, oracle_c):
    """A point on E'(Fp2) outside G2 must give code 2 (found by hashing x until on-curve)."""
    rnd = random.Random(11)
    while True:
        x = (rnd.randrange(po.P), rnd.randrange(po.P))
        rhs = po.f2_add(po.f2_mul(po.f2_sqr(x), x), (4, 4))
        # sqrt in Fp2 via norm trick
        n = (rhs[0] * rhs[0] + rhs[1] * rhs[1]) % po.P
        s = pow(n, (po.P + 1) // 4, po.P)
        if s * s % po.P != n:
            continue
        y = None
        for sg in (s, -s % po.P):
            t = (rhs[0] + sg) * po.INV2 % po.P
            xr = pow(t, (po.P + 1) // 4, po.P)
            if xr * xr % po.P == t and xr:
                yi = rhs[1] * po.fp_inv(2 * xr % po.P) % po.P
                if po.f2_sqr((xr, yi)) == rhs:
                    y = (xr, yi)
                    break
        if y:
            break
    q = po.encode_g2((x, y))
    assert oracle_c.call("pairing", G1B + q)[0] == 2
    assert product.raw_call("bls12_pairing", G1B + q, 32)[0] == 2
    # ... but it is a legal MULTIEXP input
    assert product.G2Multiexp(q + ONES) == oracle_c.call("g2multiexp", q + ONES)[1]


def test_batched_point_validation(product, oracle_c):
    """K3 as a standalone batch: per-point codes incl. the subgroup test, against the oracle's fast and naive tests."""
    rnd = random.Random(21)
    pts = [G1B, bytes(128), po.encode_g1((0, 2)), G1B[:64] + po.fp_to_bytes(5), bytes(16) + po.P.to_bytes(48, "big") + G1B[64:]]
    while len(pts) < 40:
        x = rnd.randrange(po.P)
        y = pow((x ** 3 + 4) % po.P, (po.P + 1) // 4, po.P)
        if y * y % po.P == (x ** 3 + 4) % po.P:
            pts.append(po.encode_g1((x, y)))                      # on curve, outside G1
        pts.append(oracle_c.g1_gen_mul(rnd.randrange(po.R)))      # in G1
    codes = product.points_check(1, b"".join(pts))
    want = []
    for p in pts:
        err, _ = oracle_c.call("g1mul", p + bytes(32))
        if err == 0 and oracle_c.lib().oracle_g1_in_subgroup(p, 1) != 1:
            err = 2
        want.append(err)
    assert list(codes) == want
    assert list(product.points_check(1, b"".join(pts), check_subgroup=False)) == [0 if w == 2 else w for w in want]
    q = [G2B, bytes(256), G2B[:192] + po.fp_to_bytes(7), oracle_c.g2_gen_mul(12345)]
    assert list(product.points_check(2, b"".join(q))) == [0, 0, 1, 0]


def test_checked_msm_opt_in(product, oracle_c):
    """Off by default (reference semantics); when enabled, non-subgroup points give code 2."""
    order3 = po.encode_g1((0, 2))
    data = G1B + ONES + order3 + ONES
    assert product.raw_call("bls12_g1multiexp", data, 128)[0] == 0
    product.set_checked_msm(True)
    try:
        assert product.raw_call("bls12_g1multiexp", data, 128)[0] == 2
        ok = G1B + ONES + oracle_c.g1_gen_mul(77) + ONES
        assert product.G1Multiexp(ok) == oracle_c.call("g1multiexp", ok)[1]
    finally:
        product.set_checked_msm(False)
    assert product.raw_call("bls12_g1multiexp", data, 128)[0] == 0


def test_concurrent_callers_share_one_gpu(product, oracle_c):
    """The reference is re-entrant (no globals); the replacement must be callable from many threads."""
    import threading
    jobs = []
    for n in (3, 50, 700, 9):
        d, _ = wl.g1_msm_input(n, 0x5151 + n)
        jobs.append(("g1", d, oracle_c.call("g1multiexp", d)[1]))
    d2, _ = wl.g2_msm_input(20, 7)
    jobs.append(("g2", d2, oracle_c.call("g2multiexp", d2)[1]))
    rng = wl.SplitMix64(9)
    pd = wl.pairing_call(3, rng, True)
    jobs.append(("pairing", pd, oracle_c.call("pairing", pd)[1]))
    results = [None] * (len(jobs) * 3)

    def work(slot, kind, data):
        fn = {"g1": product.G1Multiexp, "g2": product.G2Multiexp, "pairing": product.Pairing}[kind]
        results[slot] = fn(data)

    threads = []
    for rep in range(3):
        for j, (kind, data, _) in enumerate(jobs):
            threads.append(threading.Thread(target=work, args=(rep * len(jobs) + j, kind, data)))
    [t.start() for t in threads]
    [t.join() for t in threads]
    for rep in range(3):
        for j, (_, _, want) in enumerate(jobs):
            assert results[rep * len(jobs) + j] == want


def test_multiexp_batch_matches_single_calls(product, oracle_c):
    """Many small independent MULTIEXP calls in one submission: bytes and codes per call as the single-call ABI."""
    calls, want = [], []
    for n in (1, 2, 3, 5, 17, 40, 1, 128):
        d, _ = wl.g1_msm_input(n, 0x6000 + n)
        calls.append(d)
    off_curve = G1B[:64] + po.fp_to_bytes(5)
    bad = bytes(16) + po.P.to_bytes(48, "big") + G1B[64:]
    calls += [G1B + ONES + off_curve + ONES + bad + ONES, G1B + ONES + bad + ONES + off_curve + ONES, b"", G1B + ONES[:20],
              bytes(160) * 3, po.encode_g1((0, 2)) + ONES + G1B + be32(po.R)]
    offs = [0]
    for cdata in calls:
        offs.append(offs[-1] + len(cdata))
    outs, errs = product.MultiexpBatch(1, b"".join(calls), offs)
    for i, cdata in enumerate(calls):
        code, out = (5, None) if len(cdata) == 0 or len(cdata) % 160 else oracle_c.call("g1multiexp", cdata)
        assert errs[i] == code, i
        assert bytes(outs[i]) == (out if code == 0 else bytes(128)), i
    calls2 = []
    for n in (1, 4, 9):
        d, _ = wl.g2_msm_input(n, 0x7000 + n)
        calls2.append(d)
    calls2.append(G2B[:192] + po.fp_to_bytes(7) + ONES)
    offs2 = [0]
    for cdata in calls2:
        offs2.append(offs2[-1] + len(cdata))
    outs, errs = product.MultiexpBatch(2, b"".join(calls2), offs2)
    for i, cdata in enumerate(calls2):
        code, out = oracle_c.call("g2multiexp", cdata)
        assert errs[i] == code and bytes(outs[i]) == (out if code == 0 else bytes(256)), i


def test_random_shapes_against_oracle(product, oracle_c):
    """Randomised shapes: sizes, repeated points, special scalars, window widths (seeded, bit-exact vs the oracle)."""
    rnd = random.Random(0xB200)
    base_pts = [oracle_c.g1_gen_mul(rnd.randrange(po.R)) for _ in range(12)] + [bytes(128), po.encode_g1((0, 2)),
                po.encode_g1(po.ec_neg(po.FP_OPS, po.G1)), G1B]
    special = [0, 1, 2, po.R - 1, po.R, po.R + 1, (1 << 256) - 1, 1 << 255, (1 << 128) - 1, 0xFFFF, 1 << 16, (1 << 16) - 1]
    for trial in range(12):
        n = rnd.choice([1, 2, 3, 7, 33, 100, 257])
        data = b"".join(rnd.choice(base_pts) + be32(rnd.choice(special) if rnd.random() < 0.4 else rnd.randrange(1 << 256))
                        for _ in range(n))
        want = oracle_c.call("g1multiexp", data)
        product.set_window(rnd.choice([0, 0, 3, 5, 9, 12, 15, 16]))
        try:
            assert product.raw_call("bls12_g1multiexp", data, 128) == want, trial
        finally:
            product.set_window(0)


def test_map_to_curve_single_and_batch(product, oracle_c):
    """MAP_FP_TO_G1 / MAP_FP2_TO_G2 through the reference ABI and the batch entry point, vs the oracle."""
    rnd = random.Random(0x2537 + 40)
    exc = po.fp_sqrt(-pow(11, -1, po.P) % po.P)
    us = [0, 1, 2, po.P - 1, exc, po.P - exc] + [rnd.randrange(po.P) for _ in range(300)]
    blob = b"".join(po.fp_to_bytes(u) for u in us)
    outs, errs = product.MapBatch(1, blob)
    assert list(errs) == [0] * len(us)
    for j in range(len(us)):
        assert oracle_c.call("map_fp_to_g1", blob[64 * j:64 * j + 64]) == (0, bytes(outs[j])), hex(us[j])
    for j in (0, 4, 17):
        assert product.MapFpToG1(blob[64 * j:64 * j + 64]) == bytes(outs[j])
    us2 = [(0, 0), (1, 0), (0, 1), (0, po.P - 2), (po.P - 1, po.P - 1)] + [(rnd.randrange(po.P), rnd.randrange(po.P)) for _ in range(100)]
    blob2 = b"".join(po.fp_to_bytes(a) + po.fp_to_bytes(b) for a, b in us2)
    outs2, errs2 = product.MapBatch(2, blob2)
    assert list(errs2) == [0] * len(us2)
    for j in range(len(us2)):
        assert oracle_c.call("map_fp2_to_g2", blob2[128 * j:128 * j + 128]) == (0, bytes(outs2[j])), us2[j]
    assert product.MapFp2ToG2(blob2[128 * 7:128 * 8]) == bytes(outs2[7])
    # error codes: length first, then invalid element (pad byte / value >= p); out untouched on error
    ge_p = bytes(16) + po.P.to_bytes(48, "big")
    pad = b"\x80" + po.fp_to_bytes(3)[1:]
    for name, outlen, data, code in (("bls12_map_fp_to_g1", 128, b"", 5), ("bls12_map_fp_to_g1", 128, bytes(63), 5),
                                     ("bls12_map_fp_to_g1", 128, bytes(128), 5), ("bls12_map_fp_to_g1", 128, ge_p, 3),
                                     ("bls12_map_fp_to_g1", 128, pad, 3), ("bls12_map_fp2_to_g2", 256, bytes(64), 5),
                                     ("bls12_map_fp2_to_g2", 256, po.fp_to_bytes(1) + ge_p, 3),
                                     ("bls12_map_fp2_to_g2", 256, pad + po.fp_to_bytes(1), 3)):
        got, out = product.raw_call(name, data, outlen)
        assert got == code and out is None, (name, len(data), got)
    # a batch keeps per-element codes
    outs, errs = product.MapBatch(1, po.fp_to_bytes(9) + ge_p + po.fp_to_bytes(10))
    assert list(errs) == [0, 3, 0]
    assert bytes(outs[2]) == oracle_c.call("map_fp_to_g1", po.fp_to_bytes(10))[1]
    # mapped points are valid PAIRING inputs (in G1 / G2): e(P, Q) * e(-P, Q) == 1
    p1, q2 = bytes(outs[0]), bytes(outs2[5])
    err, pt = po.decode_g1(p1)
    neg = po.encode_g1(po.ec_neg(po.FP_OPS, pt))
    assert product.Pairing(p1 + q2 + neg + q2)[31] == 1
