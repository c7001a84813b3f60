"""Pins the big-int model: public constants, known answers, algebraic laws."""
import random

import py_oracle as o

P, R, Z = o.P, o.R, o.Z


def test_curve_parameter_relations():
    assert R == Z ** 4 - Z ** 2 + 1
    assert P == (Z - 1) ** 2 * R // 3 + Z
    assert o.H1 == (Z - 1) ** 2 // 3 and o.H1 % 2 == 1
    assert P % 4 == 3
    assert (-pow(P, -1, 1 << 32)) % (1 << 32) == 0xFFFCFFFD
    assert ((Z - 1) ** 2 * (Z + P) * (Z * Z + P * P - 1) + 3) == 3 * (P ** 4 - P ** 2 + 1) // R


def test_generators_and_known_answer_2g1():
    assert o.ec_on_curve(o.FP_OPS, o.G1) and o.ec_on_curve(o.F2_OPS, o.G2)
    assert o.ec_mul(o.FP_OPS, R, o.G1) is None and o.ec_mul(o.F2_OPS, R, o.G2) is None
    d = o.ec_add(o.FP_OPS, o.G1, o.G1)   # geth bls_g1add_(g1+g1=2*g1) expected value
    assert d[0] == 0x0572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e
    assert d[1] == 0x166a9d8cabc673a322fda673779d8e3822ba3ecb8670e461f73bb9021d5fd76a4c56d9d4cd16bd1bba86881979749d28
    assert o.ec_on_curve(o.FP_OPS, (0, 2)) and o.ec_mul(o.FP_OPS, 3, (0, 2)) is None


def test_tower_and_frobenius():
    rnd = random.Random(1)
    f2 = lambda: (rnd.randrange(P), rnd.randrange(P))
    a = ((f2(), f2(), f2()), (f2(), f2(), f2()))
    assert o.f12_frob(a, 1) == o.f12_pow(a, P)
    assert o.f12_frob(a, 2) == o.f12_frob(o.f12_frob(a, 1), 1)
    assert o.f12_mul(a, o.f12_inv(a)) == o.F12_ONE


def test_pairing_two_formulations_and_bilinearity():
    rnd = random.Random(2)
    ea = o.pairing_textbook(o.G1, o.G2)
    eb = o.final_exp(o.miller_loop_pairs([(o.G1, o.G2)]))
    assert ea != o.F12_ONE
    assert eb == o.f12_mul(o.f12_sqr(ea), ea)          # formulation B = A^3
    assert o.f12_pow(ea, R) == o.F12_ONE
    a, b = rnd.randrange(R), rnd.randrange(R)
    ap, bq = o.ec_mul(o.FP_OPS, a, o.G1), o.ec_mul(o.F2_OPS, b, o.G2)
    assert o.final_exp(o.miller_loop_pairs([(ap, bq)])) == o.f12_pow(eb, a * b % R)
    nab = o.ec_neg(o.FP_OPS, o.ec_mul(o.FP_OPS, a * b % R, o.G1))
    assert o.pairing_check([(ap, bq), (nab, o.G2)])
    assert not o.pairing_check([(ap, bq), (o.ec_neg(o.FP_OPS, o.ec_mul(o.FP_OPS, (a * b + 1) % R, o.G1)), o.G2)])
    assert o.pairing_check([(o.G1, o.G2), (o.ec_neg(o.FP_OPS, o.G1), o.G2)])


def test_fast_subgroup_criteria_agree_with_r_times_p():
    rnd = random.Random(7)
    # a point on E(Fp) outside G1, the cleared-cofactor image, and small-order torsion points
    while True:
        x = rnd.randrange(P)
        y = pow((x * x * x + 4) % P, (P + 1) // 4, P)
        if y * y % P == (x * x * x + 4) % P:
            break
    pt = (x, y)
    assert not o.g1_in_subgroup(pt) and not o.g1_in_subgroup_fast(pt)
    cleared = o.ec_mul(o.FP_OPS, o.H1, pt)
    assert o.g1_in_subgroup(cleared) and o.g1_in_subgroup_fast(cleared)
    tors = o.ec_mul(o.FP_OPS, R, pt)
    assert tors is not None and not o.g1_in_subgroup_fast(tors)
    assert not o.g1_in_subgroup_fast((0, 2)) and not o.g1_in_subgroup_fast(o.ec_add(o.FP_OPS, (0, 2), o.G1))
    for q in (11, 10177):
        t = o.ec_mul(o.FP_OPS, o.H1 * R // (q * q), pt)
        while t is not None and o.ec_mul(o.FP_OPS, q, t) is not None:
            t = o.ec_mul(o.FP_OPS, q, t)
        if t is not None:
            assert not o.g1_in_subgroup_fast(t)
    assert o.g2_in_subgroup_fast(o.G2) and o.g2_in_subgroup_fast(o.ec_mul(o.F2_OPS, 12345, o.G2))


def test_codec_rules():
    g = o.encode_g1(o.G1)
    assert o.decode_g1(g) == (0, o.G1)
    assert o.decode_g1(bytes(128)) == (0, None)
    bad = bytearray(g); bad[0] = 1
    assert o.decode_g1(bytes(bad))[0] == o.INVALID_ELEMENT
    assert o.decode_g1(bytes(16) + P.to_bytes(48, "big") + g[64:])[0] == o.INVALID_ELEMENT
    assert o.decode_g1(g[:64] + o.fp_to_bytes(5))[0] == o.POINT_NOT_ON_CURVE
    assert o.decode_g1(o.fp_to_bytes(0) + o.fp_to_bytes(3))[0] == o.POINT_NOT_ON_CURVE   # x=0,y!=0 is not infinity
    assert o.bls12_g1multiexp(b"")[0] == o.INVALID_LENGTH and o.bls12_pairing(b"")[0] == o.INVALID_LENGTH
    assert o.bls12_g1multiexp(g + bytes(32)) == (0, bytes(128))


# ---- MAP_FP_TO_G1 / MAP_FP2_TO_G2 ------------------------------------------------------------------
def _on_iso_curve(ops, iso, pt):
    x, y = pt
    return ops.sqr(y) == ops.add(ops.add(ops.mul(ops.sqr(x), x), ops.mul(iso["A"], x)), iso["B"])


def test_isogeny_constants_are_rederivable():
    """oracle/derive_isogeny.py recomputes both isogenies from the division polynomial (Kohel) and checks the
    recalled RFC 9380 coefficients; the committed JSON (the single copy, oracle/isogeny_constants.json, which
    tools/gen_device_constants.py also reads) must equal that derivation."""
    import json
    import os
    import derive_isogeny as d
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    assert os.path.exists(os.path.join(root, "oracle", "isogeny_constants.json"))
    assert not os.path.exists(os.path.join(root, "tools", "isogeny_constants.json"))
    g1, g2 = d.g1_isogeny(), d.g2_isogeny()
    for key in ("x_num", "x_den", "y_num", "y_den"):
        assert g1[key] == o.ISO["g1"][key]
        assert [tuple(v) for v in g2[key]] == [tuple(v) for v in o.ISO["g2"][key]]
    assert (g1["A"], g1["B"], g1["Z"], g1["scale"]) == (o.ISO["g1"]["A"], o.ISO["g1"]["B"], 11, 11)
    assert g2["scale"] == ((-3) % o.P, 0)


def test_isogenous_curve_order_and_isogeny_is_a_homomorphism():
    rnd = random.Random(0x2537 + 30)
    iso = o.ISO["g1"]
    A = iso["A"]
    def add(p, q):            # group law on E1' (a != 0)
        if p is None: return q
        if q is None: return p
        if p[0] == q[0]:
            if (p[1] + q[1]) % o.P == 0: return None
            lam = (3 * p[0] * p[0] + A) * pow(2 * p[1], -1, o.P) % o.P
        else:
            lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, o.P) % o.P
        x = (lam * lam - p[0] - q[0]) % o.P
        return (x, (lam * (p[0] - x) - p[1]) % o.P)
    def mul(k, p):
        r = None
        while k:
            if k & 1: r = add(r, p)
            p = add(p, p); k >>= 1
        return r
    pts = []
    while len(pts) < 2:
        x = rnd.randrange(o.P)
        y = o.fp_sqrt((x * x * x + A * x + iso["B"]) % o.P)
        if y is not None:
            pts.append((x, y))
    p, q = pts
    assert mul(o.H1 * o.R, p) is None                      # #E1'(Fp) = #E1(Fp): E1' is isogenous to E1
    ip, iq, ipq = (o._iso_map(o.FP_OPS, iso, t) for t in (p, q, add(p, q)))
    assert o.ec_on_curve(o.FP_OPS, ip) and o.ec_on_curve(o.FP_OPS, iq)
    assert o.ec_add(o.FP_OPS, ip, iq) == ipq               # iso(P + Q) = iso(P) + iso(Q)


def test_map_outputs_in_subgroup_and_g2_cofactor_two_ways():
    rnd = random.Random(0x2537 + 31)
    for u in (0, 1, rnd.randrange(o.P)):
        q = o._sswu(o.FP_OPS, o.fp_sqrt, o.fp_sgn0, o.ISO["g1"]["A"], o.ISO["g1"]["B"], 11, u)
        assert _on_iso_curve(o.FP_OPS, o.ISO["g1"], q) and o.fp_sgn0(q[1]) == o.fp_sgn0(u)
        pt = o.map_fp_to_g1(u)
        assert o.ec_on_curve(o.FP_OPS, pt) and o.g1_in_subgroup(pt)
    assert o.map_fp_to_g1(5) == o.ec_neg(o.FP_OPS, o.map_fp_to_g1(o.P - 5))     # map(-u) = -map(u)
    iso = o.ISO["g2"]
    for u in ((0, 0), (0, 1), (rnd.randrange(o.P), rnd.randrange(o.P))):
        q = o._sswu(o.F2_OPS, o.f2_sqrt, o.f2_sgn0, iso["A"], iso["B"], iso["Z"], u)
        assert _on_iso_curve(o.F2_OPS, iso, q) and o.f2_sgn0(q[1]) == o.f2_sgn0(u)
        e2 = o._iso_map(o.F2_OPS, iso, q)
        assert o.ec_on_curve(o.F2_OPS, e2)
        cleared = o.ec_mul(o.F2_OPS, o.H_EFF_G2, e2)
        assert cleared == o.g2_clear_cofactor_bp(e2)       # recalled h_eff == Budroni-Pintore (RFC 9380 G.4)
        assert o.g2_in_subgroup(cleared) and cleared == o.map_fp2_to_g2(u)


def test_external_known_answers():
    """tests/kat.py: RFC 9380 J.9.1 / J.9.2 / J.10.2 (msg = "") and the geth 2*G1 / 2*G2 vectors."""
    import kat
    iso = o.ISO["g1"]
    def q_of(u):
        return o._iso_map(o.FP_OPS, iso, o._sswu(o.FP_OPS, o.fp_sqrt, o.fp_sgn0, iso["A"], iso["B"], iso["Z"], u))
    assert q_of(kat.G1_NU_U) == kat.G1_NU_Q
    assert o.map_fp_to_g1(kat.G1_NU_U) == kat.G1_NU_P
    q0, q1 = q_of(kat.G1_RO_U0), q_of(kat.G1_RO_U1)
    assert (q0, q1) == (kat.G1_RO_Q0, kat.G1_RO_Q1)
    assert o.ec_mul(o.FP_OPS, o.H_EFF_G1, o.ec_add(o.FP_OPS, q0, q1)) == kat.G1_RO_P
    assert o.map_fp2_to_g2(kat.G2_NU_U) == kat.G2_NU_P
    assert o.ec_add(o.FP_OPS, o.G1, o.G1) == kat.TWO_G1
    assert o.ec_mul(o.FP_OPS, 3, o.G1) == kat.THREE_G1
    assert o.ec_add(o.F2_OPS, o.G2, o.G2) == kat.TWO_G2
