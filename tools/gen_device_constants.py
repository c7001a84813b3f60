#!/usr/bin/env python3
"""Emit blst_eip2537_b200/csrc/constants.cuh: BLS12-381 constants as 12x32-bit Montgomery limbs.

All values are derived here from first principles (p, r, z, the tower xi = 1+u); nothing is
taken from blst or the reference.  Deliberately independent of oracle/ (the product must not
depend on test infrastructure): the handful of derived constants are recomputed inline.
Run: python tools/gen_device_constants.py > blst_eip2537_b200/csrc/constants.cuh
"""
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
Z = -0xd201000000010000
RM = (1 << 384) % P

G1 = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
      0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2 = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
      (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))


def f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2pow(a, e):
    out = (1, 0)
    while e:
        if e & 1:
            out = f2mul(out, a)
        a = f2mul(a, a)
        e >>= 1
    return out


def f2inv(a):
    t = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * t % P, -a[1] * t % P)


XI = (1, 1)


def frob(k):
    g = f2pow(XI, (P ** k - 1) // 6)
    out, acc = [], (1, 0)
    for _ in range(6):
        out.append(acc)
        acc = f2mul(acc, g)
    return out


def jac_mul(k, pt):
    # tiny affine double-and-add over Fp, only used to pick the right cube root of unity
    def add(a, b):
        if a is None:
            return b
        if b is None:
            return a
        if a[0] == b[0]:
            if (a[1] + b[1]) % P == 0:
                return None
            lam = 3 * a[0] * a[0] * pow(2 * a[1], P - 2, P) % P
        else:
            lam = (b[1] - a[1]) * pow(b[0] - a[0], P - 2, P) % P
        x = (lam * lam - a[0] - b[0]) % P
        return (x, (lam * (a[0] - x) - a[1]) % P)
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def find_beta():
    g = pow(2, (P - 1) // 3, P)
    tgt = jac_mul((-Z * Z) % R, G1)
    for b in (g, g * g % P):
        if (b * G1[0] % P, G1[1]) == tgt:
            return b
    raise SystemExit("beta not found")


def words(v):
    return ", ".join("0x%08xu" % ((v >> (32 * i)) & 0xFFFFFFFF) for i in range(12))


decls = []


def emit(name, vals):
    """vals: list of integers (already in the representation to store)."""
    n = 12 * len(vals)
    body = ",\n    ".join(words(v) for v in vals)
    decls.append((name, n, body))


def mont(v):
    return v * RM % P


def fp2m(v):
    return [mont(v[0]), mont(v[1])]


emit("P", [P])
emit("RR", [mont(RM)])
emit("ONE", [mont(1)])
emit("R3", [pow(RM, 3, P)])   # R^3 mod p, raw: mont_mul(x, R3) = x*R^2
emit("B1", [mont(4)])
emit("B2", fp2m((4, 4)))
emit("B2X3", fp2m((12, 12)))
emit("INV2", [mont(pow(2, P - 2, P))])
emit("BETA", [mont(find_beta())])
emit("PSI_CX", fp2m(f2inv(f2pow(XI, (P - 1) // 3))))
emit("PSI_CY", fp2m(f2inv(f2pow(XI, (P - 1) // 2))))
f1, f2 = frob(1), frob(2)
emit("FROB1", [x for v in f1 for x in fp2m(v)])
emit("FROB2", [x for v in f2 for x in fp2m(v)])
emit("G1GEN", [mont(G1[0]), mont(G1[1])])
emit("G2GEN", fp2m(G2[0]) + fp2m(G2[1]))

# MAP_FP_TO_G1 / MAP_FP2_TO_G2 (RFC 9380 8.8): SSWU curve parameters and isogeny coefficients from
# oracle/isogeny_constants.json (derived and checked by oracle/derive_isogeny.py; this generator is a developer
# tool -- the shipped library only sees the constants.cuh it writes), plus the fixed exponents (raw integers).
import json
import os
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "isogeny_constants.json")) as fh:
    ISO = json.load(fh)


def cv(v):
    return int(v, 16) if isinstance(v, str) else (int(v[0], 16), int(v[1], 16))


i1, i2 = ISO["g1"], ISO["g2"]
A1, B1c, Z1 = cv(i1["A"]), cv(i1["B"]), cv(i1["Z"])
emit("ISO1_A", [mont(A1)])
emit("ISO1_B", [mont(B1c)])
emit("ISO1_Z", [mont(Z1)])
sq = pow((-Z1) % P, (P + 1) // 4, P)
assert sq * sq % P == (-Z1) % P
emit("ISO1_SQRT_MZ", [mont(sq)])                       # sqrt(-Z) for sqrt_ratio (p = 3 mod 4)
for key in ("x_num", "x_den", "y_num", "y_den"):
    emit("ISO1_" + key.upper().replace("_", ""), [mont(cv(v)) for v in i1[key]])
emit("ISO2_A", fp2m(cv(i2["A"])))
emit("ISO2_B", fp2m(cv(i2["B"])))
emit("ISO2_Z", fp2m(cv(i2["Z"])))
for key in ("x_num", "x_den", "y_num", "y_den"):
    emit("ISO2_" + key.upper().replace("_", ""), [x for v in i2[key] for x in fp2m(cv(v))])
emit("EXP_PM3D4", [(P - 3) // 4])
emit("EXP_PP1D4", [(P + 1) // 4])

print("// GENERATED by tools/gen_device_constants.py -- do not edit.")
print("// BLS12-381 constants, 12 x 32-bit little-endian limbs, Montgomery form (R = 2^384) unless noted.")
print("#pragma once")
print("#include <stdint.h>")
print("namespace b200 {")
print("#define B200_M0 0x%08xu  /* -p^-1 mod 2^32 */" % ((-pow(P, -1, 1 << 32)) % (1 << 32)))
print("#define B200_Z_ABS 0x%016xull" % (-Z))
for i in range(12):
    print("#define B200_P%d 0x%08x" % (i, (P >> (32 * i)) & 0xFFFFFFFF))
for name, n, body in decls:
    print("#ifdef __CUDACC__")
    print("static __device__ __constant__ uint32_t DC_%s[%d] = {\n    %s};" % (name, n, body))
    print("#endif")
    print("static const uint32_t HC_%s[%d] = {\n    %s};" % (name, n, body))
    print("B200_HD const uint32_t* C_%s() {\n#ifdef __CUDA_ARCH__\n  return DC_%s;\n#else\n  return HC_%s;\n#endif\n}" % (name, name, name))
print("}  // namespace b200")
