#!/usr/bin/env python3
"""Derive the hash-to-curve isogeny maps used by MAP_FP_TO_G1 / MAP_FP2_TO_G2 from first principles.

TEST INFRASTRUCTURE (oracle/): not used by the product at run time; it generates constants.

The reference calls blst_map_to_g1 / blst_map_to_g2 (/root/reference/src/eip2537.c:1113,1155) -- blst is
absent, so the published algorithm (RFC 9380 section 8.8: simplified SWU on an isogenous curve, the
11-/3-isogeny to E1/E2, cofactor clearing by h_eff) is restated.  The only recalled inputs are the curve
parameters of the isogenous curves E1' (A', B') and E2' (240i, 1012(1+i)) and SSWU's Z (11, -(2+i)); E1'
is checked by its group order (#E1' = #E1).  The isogeny itself is DERIVED here: kernel polynomial = the
rational degree-(l-1)/2 factor of the l-division polynomial, map by Kohel's formula, codomain scaled to
y^2 = x^3 + 4 (resp. 4(1+i)).  The codomain has j = 0, so six scalings (x by a cube root of unity, y by
+-1) are valid isogenies; the RFC's choice is selected by matching recalled low-order coefficients
k_(1,0), k_(3,0) of the RFC tables bit-for-bit (a 381-bit coincidence is impossible; a wrong recollection
would match none of the six and the script fails).
"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import py_oracle as po

P = po.P

# ---------------- generic field helpers: Fp as ints, Fp2 as tuples ----------------------------------
class FpF:
    zero, one = 0, 1
    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return a * b % P
    @staticmethod
    def inv(a): return pow(a, -1, P)
    @staticmethod
    def const(n): return n % P
    @staticmethod
    def is_zero(a): return a % P == 0
    order = P

class Fp2F:
    zero, one = (0, 0), (1, 0)
    add = staticmethod(po.f2_add)
    sub = staticmethod(po.f2_sub)
    mul = staticmethod(po.f2_mul)
    inv = staticmethod(po.f2_inv)
    @staticmethod
    def const(n): return (n % P, 0)
    @staticmethod
    def is_zero(a): return a[0] % P == 0 and a[1] % P == 0
    order = P * P

# ---------------- dense polynomials, low degree first -----------------------------------------------
def ptrim(F, a):
    while a and F.is_zero(a[-1]): a = a[:-1]
    return a
def padd(F, a, b):
    n = max(len(a), len(b))
    return ptrim(F, [F.add(a[i] if i < len(a) else F.zero, b[i] if i < len(b) else F.zero) for i in range(n)])
def psub(F, a, b):
    n = max(len(a), len(b))
    return ptrim(F, [F.sub(a[i] if i < len(a) else F.zero, b[i] if i < len(b) else F.zero) for i in range(n)])
def pmul(F, a, b):
    if not a or not b: return []
    r = [F.zero] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if F.is_zero(x): continue
        for j, y in enumerate(b):
            r[i + j] = F.add(r[i + j], F.mul(x, y))
    return ptrim(F, r)
def pscale(F, a, c): return ptrim(F, [F.mul(x, c) for x in a])
def pdivmod(F, a, b):
    a = list(a); q = [F.zero] * max(0, len(a) - len(b) + 1)
    ib = F.inv(b[-1])
    while len(a) >= len(b) and a:
        c = F.mul(a[-1], ib); d = len(a) - len(b)
        q[d] = c
        for i, y in enumerate(b):
            a[d + i] = F.sub(a[d + i], F.mul(c, y))
        a = ptrim(F, a[:-1]) if True else a
    return ptrim(F, q), ptrim(F, a)
def pmod(F, a, m): return pdivmod(F, a, m)[1]
def pmonic(F, a): return pscale(F, a, F.inv(a[-1]))
def pgcd(F, a, b):
    while b:
        a, b = b, pmod(F, a, b)
    return pmonic(F, a) if a else a
def pderiv(F, a): return ptrim(F, [F.mul(F.const(i), a[i]) for i in range(1, len(a))])
def peval(F, a, x):
    r = F.zero
    for c in reversed(a): r = F.add(F.mul(r, x), c)
    return r
def ppowmod(F, base, e, m):
    r = [F.one]; b = pmod(F, base, m)
    while e:
        if e & 1: r = pmod(F, pmul(F, r, b), m)
        b = pmod(F, pmul(F, b, b), m); e >>= 1
    return r

# ---------------- division polynomials (f_n: psi_n for odd n, psi_n/(2y) for even n) ------------------
def division_poly(F, A, B, n):
    c = F.const
    Fx = [F.mul(c(4), B), F.mul(c(4), A), F.zero, c(4)]                  # (2y)^2 = 4(x^3+Ax+B)
    F2 = pmul(F, Fx, Fx)
    f = {0: [], 1: [F.one], 2: [F.one]}
    f[3] = ptrim(F, [F.sub(F.zero, F.mul(A, A)), F.mul(c(12), B), F.mul(c(6), A), F.zero, c(3)])
    A2, A3, B2 = F.mul(A, A), F.mul(F.mul(A, A), A), F.mul(B, B)
    f[4] = pscale(F, [F.sub(F.sub(F.zero, F.mul(c(8), B2)), A3), F.sub(F.zero, F.mul(c(4), F.mul(A, B))),
                      F.sub(F.zero, F.mul(c(5), A2)), F.mul(c(20), B), F.mul(c(5), A), F.zero, F.one], c(2))
    def get(k):
        if k in f: return f[k]
        m = k // 2
        if k % 2 == 0:
            r = pmul(F, get(m), psub(F, pmul(F, get(m + 2), pmul(F, get(m - 1), get(m - 1))),
                                     pmul(F, get(m - 2), pmul(F, get(m + 1), get(m + 1)))))
        else:
            t1 = pmul(F, get(m + 2), pmul(F, get(m), pmul(F, get(m), get(m))))
            t2 = pmul(F, get(m - 1), pmul(F, get(m + 1), pmul(F, get(m + 1), get(m + 1))))
            r = psub(F, pmul(F, F2, t1), t2) if m % 2 == 0 else psub(F, t1, pmul(F, F2, t2))
        f[k] = r
        return r
    return get(n)

# ---------------- Kohel: normalised isogeny with monic kernel polynomial h (odd degree 2n+1) ---------
def kohel(F, A, B, h):
    c = F.const
    n = len(h) - 1
    s1 = F.sub(F.zero, h[n - 1])
    s2 = h[n - 2] if n >= 2 else F.zero
    s3 = F.sub(F.zero, h[n - 3]) if n >= 3 else F.zero
    p1 = s1
    p2 = F.sub(F.mul(s1, s1), F.mul(c(2), s2))
    p3 = F.add(F.sub(F.mul(F.mul(s1, s1), s1), F.mul(c(3), F.mul(s1, s2))), F.mul(c(3), s3))
    t = F.add(F.mul(c(6), p2), F.mul(c(2 * n), A))
    w = F.add(F.add(F.mul(c(10), p3), F.mul(F.mul(c(6), A), p1)), F.mul(c(4 * n), B))
    A2 = F.sub(A, F.mul(c(5), t))
    B2 = F.sub(B, F.mul(c(7), w))
    h1 = pderiv(F, h); h2 = pderiv(F, h1)
    cubic = [F.mul(c(4), B), F.mul(c(4), A), F.zero, c(4)]
    N = pmul(F, [F.sub(F.zero, F.mul(c(2), s1)), c(2 * n + 1)], pmul(F, h, h))
    N = psub(F, N, pmul(F, [F.mul(c(2), A), F.zero, c(6)], pmul(F, h1, h)))
    N = padd(F, N, pmul(F, cubic, psub(F, pmul(F, h1, h1), pmul(F, h, h2))))
    M = psub(F, pmul(F, pderiv(F, N), h), pscale(F, pmul(F, N, h1), c(2)))       # y' = y * M / h^3
    return A2, B2, N, pmul(F, h, h), M, pmul(F, h, pmul(F, h, h))

def derive_g1():
    F = FpF
    A = 0x144698a3b8e9433d693a02c96d4982b0ea985383ee66a8d8e8981aefd881ac98936f8da0e0f97f5cf428082d584c1d
    B = 0x12e2908d11688030018b12e8753eee3b2016c1f0f24f4070a0b9c14fcef35ef55a23215a316ceaa5d1cc48e98e172be0
    f11 = pmonic(F, division_poly(F, A, B, 11))
    assert len(f11) - 1 == 60
    x = [F.zero, F.one]
    cands = []
    xp = ppowmod(F, x, P, f11)
    g1 = pgcd(F, f11, psub(F, xp, x))
    if len(g1) - 1 >= 1: cands.append(("linear", g1))
    # x^(p^5) by repeated modular composition is heavy; use exponentiation by p five times
    xq = xp
    for _ in range(4):
        xq = ppowmod(F, xq, P, f11)
    g5 = pgcd(F, f11, psub(F, xq, x))
    cands.append(("deg|5", g5))
    return A, B, f11, cands


def sixth_roots_fp(target):
    """All s in Fp with s^6 = target (p = 1 mod 6)."""
    def sqrt(a):
        r = pow(a, (P + 1) // 4, P)
        return r if r * r % P == a % P else None
    def cube_roots(a):
        if pow(a, (P - 1) // 3, P) != 1: return []
        m = P - 1; k = 0
        while m % 3 == 0: m //= 3; k += 1
        g = 2
        while pow(g, (P - 1) // 3, P) == 1: g += 1
        x = pow(a, pow(3, -1, m), P)            # correct up to an element of the 3-Sylow subgroup (order 3^k)
        cgen = pow(g, m, P)
        return [x * pow(cgen, j, P) % P for j in range(3 ** k) if pow(x * pow(cgen, j, P) % P, 3, P) == a % P]
    out = set()
    r2 = sqrt(target)
    for sq in ([r2, P - r2] if r2 is not None else []):
        for cr in cube_roots(sq):
            if pow(cr, 6, P) == target % P: out.add(cr)
    return sorted(out)


def g1_isogeny():
    """-> dict with A', B', Z and the four coefficient lists (low degree first) of the 11-isogeny E1' -> E1."""
    F = FpF
    A, B, f11, cands = derive_g1()
    lin = [g for name, g in cands if name == "linear"]
    assert lin and len(lin[0]) - 1 == 5, "expected exactly one rational 11-subgroup with x-coordinates in Fp"
    h = lin[0]
    A2, B2, N, D, M, D3 = kohel(F, A, B, h)
    assert A2 == 0, "codomain must have j = 0"
    K10 = 0x11a05f2b1e833340b809101dd99815856b303e88a2d7005ff2627b56cdb4e2c85610c2d5f2e62d6eaeac1662734649b7
    K20 = 0x08ca8d548cff19ae18b2e62f4bd3fa6f01d5ef4ba35b48ba9c9588617fc8ac62b558d681be343df8993cf9fa40d21b1c
    K30 = 0x090d97c81ba24ee0259d1f094980dcfa11ad138e48a869522b52af6c956543d3cd0c7aee9b3ba3c2be9845719707bb33
    K40 = 0x16112c4c3a9c98b252181140fad0eae9601a6de578980be6eec3232b5be72e7a07f3688ef60c206d01479253b03663c1
    assert D[0] == K20 and D3[0] == K40, "derived denominators do not match the recalled RFC 9380 E.2 constants"
    chosen = None
    for s in sixth_roots_fp(B2 * pow(4, -1, P) % P):
        is2 = pow(s * s, -1, P); is3 = pow(s * s * s, -1, P)
        if N[0] * is2 % P == K10 and M[0] * is3 % P == K30:
            chosen = (s, is2, is3)
    assert chosen, "none of the six codomain scalings reproduces the recalled k_(1,0), k_(3,0)"
    s, is2, is3 = chosen
    return {"A": A, "B": B, "Z": 11, "scale": s,
            "x_num": [v * is2 % P for v in N], "x_den": list(D),
            "y_num": [v * is3 % P for v in M], "y_den": list(D3)}


def g2_isogeny():
    """3-isogeny E2' : y^2 = x^3 + 240i x + 1012(1+i)  ->  E2 : y^2 = x^3 + 4(1+i)."""
    F = Fp2F
    A, B = (0, 240), (1012, 1012)
    f3 = division_poly(F, A, B, 3)
    x0 = ((-6) % P, 6)                                   # RFC x_den = x^2 + (12-12i)x - 72i = (x + 6 - 6i)^2
    assert F.is_zero(peval(F, f3, x0)), "x0 = -6+6i must be a 3-torsion x-coordinate of E2'"
    h = [F.sub(F.zero, x0), F.one]
    A2, B2, N, D, M, D3 = kohel(F, A, B, h)
    assert F.is_zero(A2)
    K10 = 0x5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6
    K30 = 0x1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706
    # s^6 = B*/(4(1+i)); try the small rational candidates first, then their products with 6th roots of unity in Fp2
    target = F.mul(B2, F.inv((4, 4)))
    # sixth roots of unity: +-1, +-w, +-w^2 with w a primitive cube root of unity in Fp
    w = pow(2, (P - 1) // 3, P)
    g = 2
    while w == 1:
        g += 1; w = pow(g, (P - 1) // 3, P)
    units = [(1, 0), (P - 1, 0), (w, 0), (P - w, 0), (w * w % P, 0), (P - w * w % P, 0)]
    base = None
    for cand in range(1, 64):
        c6 = F.const(cand ** 6)
        if c6 == target: base = F.const(cand)
    assert base is not None, "scale is expected to be a small integer (3)"
    chosen = None
    for u in units:
        s = F.mul(base, u)
        is2 = F.inv(F.mul(s, s)); is3 = F.inv(F.mul(F.mul(s, s), s))
        if F.mul(N[0], is2) == (K10, K10) and F.mul(M[0], is3) == (K30, K30):
            chosen = (s, is2, is3)
    assert chosen, "no codomain scaling reproduces the recalled k_(1,0), k_(3,0) of RFC 9380 E.3"
    s, is2, is3 = chosen
    return {"A": A, "B": B, "Z": ((-2) % P, (-1) % P), "scale": s,
            "x_num": [F.mul(v, is2) for v in N], "x_den": list(D),
            "y_num": [F.mul(v, is3) for v in M], "y_den": list(D3)}


if __name__ == "__main__":
    import json
    g1 = g1_isogeny()
    g2 = g2_isogeny()
    print("G1: scale", g1["scale"], "degrees", [len(g1[k]) - 1 for k in ("x_num", "x_den", "y_num", "y_den")])
    print("G2: scale", g2["scale"], "degrees", [len(g2[k]) - 1 for k in ("x_num", "x_den", "y_num", "y_den")])
    hx = lambda v: hex(v) if isinstance(v, int) else [hex(v[0]), hex(v[1])]
    doc = {"_comment": "generated by oracle/derive_isogeny.py (derived, then matched against recalled RFC 9380 coefficients)",
           "g1": {k: ([hx(c) for c in v] if isinstance(v, list) else hx(v)) for k, v in g1.items()},
           "g2": {k: ([hx(c) for c in v] if isinstance(v, list) else hx(v)) for k, v in g2.items()}}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "isogeny_constants.json")
    with open(out, "w") as fh:
        json.dump(doc, fh, indent=1)
    print("wrote", out)
