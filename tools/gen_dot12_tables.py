#!/usr/bin/env python3
"""Generate blst_eip2537_b200/csrc/dot12_tables.cuh: the term tables of the dot-product Fp12 engine (dot12.cuh).

An Fp12 value is held as six Fp2 coefficients a_k of w^k (w^6 = xi = 1 + u).  Every Fp12-level operation the
pairing needs is bilinear, so each OUTPUT Fp component (coefficient k, real/imaginary part) is a sum of a few
Fp products  sum_t A_t * B_t  with
    A_t = alpha * x.re + beta * x.im    (alpha, beta in -2..2), x = a coefficient of f, a line element or a constant
    B_t = one plain Fp component of a coefficient of the right-hand operand
The device accumulates the double-width products of one output and does ONE Montgomery reduction for it.

Term encoding (uint16):  bits 0-3  B component index c = 2*k + (0 re | 1 im), 0..11
                         bits 4-6  A coefficient index (0..5)        bit 7  A source (0 = f, 1 = second operand g)
                         bits 8-10 alpha + 2                         bits 11-13 beta + 2
                         bit 14    B source (0 = f, 1 = second operand g)
Tables are checked on the CPU against the tower formulas of pairing.cuh (tests/test_host_emul.py).
"""
import os

XI_WRAP = 6


def enc(a_src, a_idx, alpha, beta, b_c, b_src=0):
    assert -2 <= alpha <= 2 and -2 <= beta <= 2 and 0 <= b_c < 12 and 0 <= a_idx < 6
    return b_c | (a_idx << 4) | (a_src << 7) | ((alpha + 2) << 8) | ((beta + 2) << 11) | (b_src << 14)


def product_terms(a_src, i, b_src, j, xi, scale):
    """scale * [xi] * X_i * Y_j with X on the A side (transformed) and Y on the B side (plain).
    Returns (re_terms, im_terms).  X = (u, v); L = scale*(u, v) or scale*xi*(u, v) = scale*(u - v, u + v);
    re = L.re*y.re - L.im*y.im, im = L.re*y.im + L.im*y.re."""
    if not xi:
        lre, lim = (scale, 0), (0, scale)
    else:
        lre, lim = (scale, -scale), (scale, scale)
    neg = lambda ab: (-ab[0], -ab[1])
    re = [enc(a_src, i, lre[0], lre[1], 2 * j, b_src), enc(a_src, i, *neg(lim), 2 * j + 1, b_src)]
    im = [enc(a_src, i, lre[0], lre[1], 2 * j + 1, b_src), enc(a_src, i, lim[0], lim[1], 2 * j, b_src)]
    return re, im


def square_terms(i, xi):
    """[xi] * a_i^2, a_i = (u, v): plain  re = u*u - v*v, im = 2u*v;
    with xi  re = u*u + (-2u - v)*v,  im = u*u + (2u - v)*v."""
    if not xi:
        return [enc(0, i, 1, 0, 2 * i), enc(0, i, 0, -1, 2 * i + 1)], [enc(0, i, 2, 0, 2 * i + 1)]
    return ([enc(0, i, 1, 0, 2 * i), enc(0, i, -2, -1, 2 * i + 1)],
            [enc(0, i, 1, 0, 2 * i), enc(0, i, 2, -1, 2 * i + 1)])


def table_sqr():
    out = [[] for _ in range(12)]
    for i in range(6):
        for j in range(i, 6):
            k, xi = (i + j) % 6, (i + j) >= XI_WRAP
            re, im = square_terms(i, xi) if i == j else product_terms(0, i, 0, j, xi, 2)
            out[2 * k] += re
            out[2 * k + 1] += im
    return out


def table_mul014():
    """f * (l0 + l1 w^2 + l4 w^3); g = line, element e at w-power (0, 2, 3)[e]; A side = line, B side = f."""
    out = [[] for _ in range(12)]
    for e, pw in enumerate((0, 2, 3)):
        for j in range(6):
            k, xi = (pw + j) % 6, (pw + j) >= XI_WRAP
            re, im = product_terms(1, e, 0, j, xi, 1)
            out[2 * k] += re
            out[2 * k + 1] += im
    return out


def table_mul():
    """f * g, A side = f (transformed), B side = g (plain)."""
    out = [[] for _ in range(12)]
    for i in range(6):
        for j in range(6):
            k, xi = (i + j) % 6, (i + j) >= XI_WRAP
            re, im = product_terms(0, i, 1, j, xi, 1)
            out[2 * k] += re
            out[2 * k + 1] += im
    return out


def table_cyc():
    """Bilinear part Q_k of the Granger-Scott cyclotomic squaring (pairing.cuh fp12_cyclotomic_sqr) in powers of w:
         Q_0 = a_0^2 + xi a_3^2   Q_3 = 2 a_0 a_3      Q_2 = a_1^2 + xi a_4^2   Q_5 = 2 a_1 a_4
         Q_4 = a_2^2 + xi a_5^2   Q_1 = 2 xi a_2 a_5
    the new coefficient is a_k' = 3 Q_k - 2 a_k (k even) or 3 Q_k + 2 a_k (k odd): the linear epilogue is done by the caller."""
    out = [[] for _ in range(12)]

    def put(k, terms):
        out[2 * k] += terms[0]
        out[2 * k + 1] += terms[1]
    for k, (i, j) in ((0, (0, 3)), (2, (1, 4)), (4, (2, 5))):
        put(k, square_terms(i, False))
        put(k, square_terms(j, True))
    put(3, product_terms(0, 0, 0, 3, False, 2))
    put(5, product_terms(0, 1, 0, 4, False, 2))
    put(1, product_terms(0, 2, 0, 5, True, 2))
    return out


def emit(name, tab, width):
    lines = ["__device__ __constant__ static const uint16_t D12_%s_DEV[12][%d] = {" % (name, width)]
    host = ["static const uint16_t D12_%s_HOST[12][%d] = {" % (name, width)]
    for row in tab:
        assert len(row) <= width, (name, len(row))
        body = "  {" + ", ".join("0x%04x" % t for t in row + [0] * (width - len(row))) + "},"
        lines.append(body)
        host.append(body)
    lines.append("};")
    host.append("};")
    vals = ", ".join(str(len(r)) for r in tab)
    lines.append("__device__ __constant__ static const unsigned char D12_%s_COUNT_DEV[12] = {%s};" % (name, vals))
    host.append("static const unsigned char D12_%s_COUNT_HOST[12] = {%s};" % (name, vals))
    return "#ifdef __CUDACC__\n" + "\n".join(lines) + "\n#endif\n" + "\n".join(host) + "\n"


def main():
    parts = ["// GENERATED by tools/gen_dot12_tables.py -- do not edit.  Term tables of the dot-product Fp12 engine (dot12.cuh).",
             "#pragma once", "#include <stdint.h>", "namespace b200 { namespace dot {"]
    parts.append(emit("SQR", table_sqr(), 8))
    parts.append(emit("MUL014", table_mul014(), 6))
    parts.append(emit("MUL", table_mul(), 12))
    parts.append(emit("CYC", table_cyc(), 4))
    parts.append("} }  // namespace b200::dot")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blst_eip2537_b200", "csrc", "dot12_tables.cuh")
    with open(out, "w") as fh:
        fh.write("\n".join(parts) + "\n")
    for nm, t in (("sqr", table_sqr()), ("mul014", table_mul014()), ("mul", table_mul()), ("cyc", table_cyc())):
        print(nm, [len(r) for r in t], "products", sum(len(r) for r in t))


if __name__ == "__main__":
    main()
