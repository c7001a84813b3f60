#!/usr/bin/env python3
"""Generate tests/golden/eip2537_golden.json from the big-int model (oracle/py_oracle.py).

The reference vendors no vectors (build.sh downloads them), so these fixtures are OUR pins: inputs
chosen to cover every accept/reject rule, outputs computed by the independent Python big-int model
(textbook group law; textbook pairing for the PAIRING rows).  Format follows the geth precompile
JSON the reference's Go tests read (go/blst_eip2537_test.go:18-29): Input / Expected / Name, plus
ExpectedErrorCode for failure rows.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import py_oracle as o  # noqa: E402

rnd = random.Random(0x2537)
R, P = o.R, o.P
G1B, G2B = o.encode_g1(o.G1), o.encode_g2(o.G2)


def be32(k):
    return int(k).to_bytes(32, "big")


def rg1():
    return o.encode_g1(o.ec_mul(o.FP_OPS, rnd.randrange(1, R), o.G1))


def rg2():
    return o.encode_g2(o.ec_mul(o.F2_OPS, rnd.randrange(1, R), o.G2))


rows = []


def add(fn, name, data, textbook=False):
    if fn == "g1multiexp":
        code, out = o.bls12_g1multiexp(data)
    elif fn == "g2multiexp":
        code, out = o.bls12_g2multiexp(data)
    elif fn == "map_fp_to_g1":
        code, out = o.bls12_map_fp_to_g1(data)
    elif fn == "map_fp2_to_g2":
        code, out = o.bls12_map_fp2_to_g2(data)
    else:
        code, out = o.bls12_pairing(data, textbook=textbook)
    row = {"Name": name, "Function": fn, "Input": data.hex()}
    if code:
        row["ExpectedErrorCode"] = code
    else:
        row["Expected"] = out.hex()
    rows.append(row)


ONES = be32(2 ** 256 - 1)
neg_g1 = o.encode_g1(o.ec_neg(o.FP_OPS, o.G1))
neg_g2 = o.encode_g2(o.ec_neg(o.F2_OPS, o.G2))
order3 = o.encode_g1((0, 2))
off1 = G1B[:64] + o.fp_to_bytes(5)
off2 = G2B[:192] + o.fp_to_bytes(7)
ge_p = bytes(16) + P.to_bytes(48, "big") + G1B[64:]
pad = bytearray(G1B); pad[7] = 0x80; pad = bytes(pad)

add("g1multiexp", "g1_mul_generator_by_2", G1B + be32(2))
add("g1multiexp", "g1_mul_generator_by_r", G1B + be32(R))
add("g1multiexp", "g1_mul_generator_by_max", G1B + ONES)
add("g1multiexp", "g1_mul_infinity", bytes(128) + ONES)
add("g1multiexp", "g1_mul_order3_point_scalar_ge_r", order3 + be32(R + 1))
for k in (2, 3, 4, 5, 9):
    add("g1multiexp", "g1_multiexp_random_k%d" % k, b"".join(rg1() + be32(rnd.randrange(1 << 256)) for _ in range(k)))
add("g1multiexp", "g1_multiexp_cancel", G1B + be32(77) + neg_g1 + be32(77))
add("g1multiexp", "g1_multiexp_duplicates_and_zero", G1B + be32(3) + G1B + be32(3) + G1B + bytes(32) + bytes(128) + be32(9))
add("g1multiexp", "g1_fail_empty", b"")
add("g1multiexp", "g1_fail_short", G1B + ONES[:31])
add("g1multiexp", "g1_fail_not_on_curve", off1 + ONES)
add("g1multiexp", "g1_fail_ge_modulus", ge_p + ONES)
add("g1multiexp", "g1_fail_pad", pad + ONES)
add("g1multiexp", "g1_fail_first_error_wins", G1B + ONES + off1 + ONES + ge_p + ONES)
add("g2multiexp", "g2_mul_generator_by_2", G2B + be32(2))
add("g2multiexp", "g2_mul_generator_by_max", G2B + ONES)
for k in (2, 5):
    add("g2multiexp", "g2_multiexp_random_k%d" % k, b"".join(rg2() + be32(rnd.randrange(1 << 256)) for _ in range(k)))
add("g2multiexp", "g2_multiexp_cancel", G2B + be32(5) + neg_g2 + be32(5))
add("g2multiexp", "g2_fail_not_on_curve", off2 + ONES)
add("g2multiexp", "g2_fail_short", G2B + ONES[:1])
a, b = rnd.randrange(1, R), rnd.randrange(1, R)
aP = o.encode_g1(o.ec_mul(o.FP_OPS, a, o.G1)); bQ = o.encode_g2(o.ec_mul(o.F2_OPS, b, o.G2))
nab = o.encode_g1(o.ec_neg(o.FP_OPS, o.ec_mul(o.FP_OPS, a * b % R, o.G1)))
add("pairing", "pairing_bilinear_true_textbook", aP + bQ + nab + G2B, textbook=True)
add("pairing", "pairing_single_false_textbook", aP + bQ, textbook=True)
add("pairing", "pairing_e(g,q)e(-g,q)", G1B + G2B + neg_g1 + G2B)
add("pairing", "pairing_infinity_pairs", bytes(384) + G1B + bytes(256))
add("pairing", "pairing_three_pairs_true", aP + bQ + nab + G2B + bytes(384))
add("pairing", "pairing_fail_g1_subgroup", order3 + G2B)
add("pairing", "pairing_fail_g1_off_curve", off1 + G2B)
add("pairing", "pairing_fail_g2_off_curve", G1B + off2)
add("pairing", "pairing_fail_order", order3 + off2)
add("pairing", "pairing_fail_length", G1B + G2B[:200])
# MAP rows (appended last so the rows above keep their random stream)
fpb = o.fp_to_bytes
exc = o.fp_sqrt(-pow(11, -1, P) % P)              # u^2 = -1/Z: the exceptional case of simplified SWU
for name, u in (("zero", 0), ("one", 1), ("p_minus_1", P - 1), ("sswu_exceptional", exc),
                ("random_a", rnd.randrange(P)), ("random_b", rnd.randrange(P)), ("random_c", rnd.randrange(P))):
    add("map_fp_to_g1", "map_fp_to_g1_" + name, fpb(u))
add("map_fp_to_g1", "map_fp_to_g1_fail_length", fpb(1)[:63])
add("map_fp_to_g1", "map_fp_to_g1_fail_ge_modulus", bytes(16) + P.to_bytes(48, "big"))
add("map_fp_to_g1", "map_fp_to_g1_fail_pad", b"\x01" + fpb(1)[1:])
for name, u in (("zero", (0, 0)), ("one", (1, 0)), ("i", (0, 1)), ("c0_zero_c1_odd", (0, P - 2)),
                ("random_a", (rnd.randrange(P), rnd.randrange(P))), ("random_b", (rnd.randrange(P), rnd.randrange(P)))):
    add("map_fp2_to_g2", "map_fp2_to_g2_" + name, fpb(u[0]) + fpb(u[1]))
add("map_fp2_to_g2", "map_fp2_to_g2_fail_length", fpb(1) + fpb(2) + b"\x00")
add("map_fp2_to_g2", "map_fp2_to_g2_fail_c1_ge_modulus", fpb(1) + bytes(16) + P.to_bytes(48, "big"))
# externally published known answers (tests/kat.py): the rows below must reproduce them exactly
sys.path.insert(0, os.path.join(HERE, ".."))
import kat  # noqa: E402
add("map_fp_to_g1", "map_fp_to_g1_rfc9380_J.9.2_msg_empty", fpb(kat.G1_NU_U))
assert rows[-1]["Expected"] == o.encode_g1(kat.G1_NU_P).hex()
add("map_fp_to_g1", "map_fp_to_g1_rfc9380_J.9.1_u0", fpb(kat.G1_RO_U0))
add("map_fp_to_g1", "map_fp_to_g1_rfc9380_J.9.1_u1", fpb(kat.G1_RO_U1))
# P of J.9.1 = sum of the two single-element maps (clear_cofactor is linear): G1MULTIEXP row with scalars 1, 1
add("g1multiexp", "g1_multiexp_rfc9380_J.9.1_P_from_two_maps",
    bytes.fromhex(rows[-2]["Expected"]) + be32(1) + bytes.fromhex(rows[-1]["Expected"]) + be32(1))
assert rows[-1]["Expected"] == o.encode_g1(kat.G1_RO_P).hex()
add("map_fp2_to_g2", "map_fp2_to_g2_rfc9380_J.10.2_msg_empty", fpb(kat.G2_NU_U[0]) + fpb(kat.G2_NU_U[1]))
assert rows[-1]["Expected"] == o.encode_g2(kat.G2_NU_P).hex()
add("g1multiexp", "g1_mul_generator_by_3_kat", G1B + be32(3))
assert rows[-1]["Expected"] == o.encode_g1(kat.THREE_G1).hex()
by_name = {r["Name"]: r for r in rows}
assert by_name["g1_mul_generator_by_2"]["Expected"] == o.encode_g1(kat.TWO_G1).hex()
assert by_name["g2_mul_generator_by_2"]["Expected"] == o.encode_g2(kat.TWO_G2).hex()
for nm in ("g1_mul_generator_by_2", "g2_mul_generator_by_2", "g1_mul_generator_by_3_kat", "map_fp_to_g1_rfc9380_J.9.2_msg_empty",
           "g1_multiexp_rfc9380_J.9.1_P_from_two_maps", "map_fp2_to_g2_rfc9380_J.10.2_msg_empty"):
    by_name[nm]["ExternalKAT"] = True
json.dump(rows, open(os.path.join(HERE, "eip2537_golden.json"), "w"), indent=1)
print("wrote", len(rows), "rows")
