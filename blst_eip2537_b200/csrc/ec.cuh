// ec.cuh -- K2: field-generic short-Weierstrass (a = 0) point arithmetic for E(Fp) and E'(Fp2).
//
// Replaces blst_p{1,2}_add_or_double (/root/reference/src/eip2537.c:183,197,605,693 and
// :237,251,894,983), blst_p{1,2}_from_affine / _to_affine (:598,610,698 / :887,899,988) and
// blst_p{1,2}_affine_on_curve (:336, :397).  Accumulators use extended Jacobian "XYZZ"
// coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed add 8M+2S, full add 12M+2S,
// double 6M+3S.  All exceptional cases are handled exactly (P == Q doubles, P == -Q gives
// infinity, either operand infinite) because MULTIEXP inputs are adversarial: no subgroup
// check is applied to them (eip2537.c:340, :401), repeated points and P/-P pairs are legal.
// E(Fp) and E'(Fp2) have odd order, so y = 0 never occurs on-curve and doubling needs no
// 2-torsion case.
#pragma once
#include "fp.cuh"

namespace b200 {

template <class F>
struct Affine {   // (0,0) encodes infinity, as the reference's blst_p*_affine does
  F x, y;
};
template <class F>
struct XYZZ {     // infinity <=> zz == 0
  F x, y, zz, zzz;
};

template <class F>
B200_HD bool is_inf(const Affine<F>& a) { return is_zero(a.x) && is_zero(a.y); }
template <class F>
B200_HD bool is_inf(const XYZZ<F>& a) { return is_zero(a.zz); }

template <class F>
B200_HD XYZZ<F> xyzz_inf() {
  XYZZ<F> r;
  r.x = FieldOps<F>::zero(); r.y = FieldOps<F>::zero(); r.zz = FieldOps<F>::zero(); r.zzz = FieldOps<F>::zero();
  return r;
}
template <class F>
B200_HD XYZZ<F> xyzz_from_affine(const Affine<F>& a) {
  XYZZ<F> r;
  r.x = a.x; r.y = a.y;
  if (is_inf(a)) { r.zz = FieldOps<F>::zero(); r.zzz = FieldOps<F>::zero(); }
  else           { r.zz = FieldOps<F>::one();  r.zzz = FieldOps<F>::one(); }
  return r;
}
template <class F>
B200_HD Affine<F> affine_neg(const Affine<F>& a) { Affine<F> r; r.x = a.x; r.y = neg(a.y); return r; }
template <class F>
B200_HD XYZZ<F> xyzz_neg(const XYZZ<F>& a) { XYZZ<F> r = a; r.y = neg(a.y); return r; }

// y^2 == x^3 + b ; caller has already excluded the (0,0) infinity encoding
template <class F>
B200_HD bool affine_on_curve(const Affine<F>& a) {
  F lhs = sqr(a.y);
  F rhs = add(mul(sqr(a.x), a.x), FieldOps<F>::curve_b());
  return eq(lhs, rhs);
}

// 2*(x,y) for a finite affine point -> XYZZ   (mdbl-2008-s-1)
template <class F>
B200_HD_NI XYZZ<F> xyzz_dbl_affine(const Affine<F>& a) {
  XYZZ<F> r;
  F u = dbl(a.y);
  F v = sqr(u);
  F w = mul(u, v);
  F s = mul(a.x, v);
  F x2 = sqr(a.x);
  F m = add(dbl(x2), x2);
  r.x = sub(sqr(m), dbl(s));
  r.y = mul_diff(m, sub(s, r.x), w, a.y);
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2*P   (dbl-2008-s-1)
template <class F>
B200_HD_NI XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
  if (is_inf(p)) return p;
  XYZZ<F> r;
  F u = dbl(p.y);
  F v = sqr(u);
  F w = mul(u, v);
  F s = mul(p.x, v);
  F x2 = sqr(p.x);
  F m = add(dbl(x2), x2);
  r.x = sub(sqr(m), dbl(s));
  r.y = mul_diff(m, sub(s, r.x), w, p.y);
  r.zz = mul(v, p.zz);
  r.zzz = mul(w, p.zzz);
  return r;
}

// acc += q (q affine, may be infinity)   (madd-2008-s)
template <class F>
B200_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q) {
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = xyzz_from_affine(q); return; }
  F u2 = mul(q.x, acc.zz);
  F s2 = mul(q.y, acc.zzz);
  F p = sub(u2, acc.x);
  F r = sub(s2, acc.y);
  if (is_zero(p)) {
    if (is_zero(r)) acc = xyzz_dbl_affine(q);   // same point: double
    else            acc = xyzz_inf<F>();        // opposite points cancel
    return;
  }
  F pp = sqr(p);
  F ppp = mul(p, pp);
  F qq = mul(acc.x, pp);
  F x3 = sub(sub(sqr(r), ppp), dbl(qq));
  acc.y = mul_diff(r, sub(qq, x3), acc.y, ppp);
  acc.x = x3;
  acc.zz = mul(acc.zz, pp);
  acc.zzz = mul(acc.zzz, ppp);
}

#ifdef __CUDACC__
// the same mixed addition over Fp with the multiplication loop unrolled ROWS rows per iteration (fp.cuh mul_unrolled:
// fewer register moves on the multiply pipe; measured in profiles/r02_k1_probes.md)
template <int ROWS>
__device__ __forceinline__ void xyzz_madd_unrolled(XYZZ<Fp>& acc, const Affine<Fp>& q) {
#ifdef __CUDA_ARCH__
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = xyzz_from_affine(q); return; }
  Fp u2 = mul_unrolled<ROWS>(q.x, acc.zz);
  Fp s2 = mul_unrolled<ROWS>(q.y, acc.zzz);
  Fp p = sub(u2, acc.x);
  Fp r = sub(s2, acc.y);
  if (is_zero(p)) {
    if (is_zero(r)) acc = xyzz_dbl_affine(q);
    else            acc = xyzz_inf<Fp>();
    return;
  }
  Fp pp = mul_unrolled<ROWS>(p, p);
  Fp ppp = mul_unrolled<ROWS>(p, pp);
  Fp qq = mul_unrolled<ROWS>(acc.x, pp);
  Fp x3 = sub(sub(mul_unrolled<ROWS>(r, r), ppp), dbl(qq));
  acc.y = mul_diff(r, sub(qq, x3), acc.y, ppp);
  acc.x = x3;
  acc.zz = mul_unrolled<ROWS>(acc.zz, pp);
  acc.zzz = mul_unrolled<ROWS>(acc.zzz, ppp);
#endif
}
#endif

// acc += q (both XYZZ)   (add-2008-s)
template <class F>
B200_HD_NI void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = q; return; }
  F u1 = mul(acc.x, q.zz);
  F u2 = mul(q.x, acc.zz);
  F s1 = mul(acc.y, q.zzz);
  F s2 = mul(q.y, acc.zzz);
  F p = sub(u2, u1);
  F r = sub(s2, s1);
  if (is_zero(p)) {
    if (is_zero(r)) acc = xyzz_dbl(acc);
    else            acc = xyzz_inf<F>();
    return;
  }
  F pp = sqr(p);
  F ppp = mul(p, pp);
  F qq = mul(u1, pp);
  F x3 = sub(sub(sqr(r), ppp), dbl(qq));
  acc.y = mul_diff(r, sub(qq, x3), s1, ppp);
  acc.x = x3;
  acc.zz = mul(mul(acc.zz, q.zz), pp);
  acc.zzz = mul(mul(acc.zzz, q.zzz), ppp);
}

// ---- affine + affine with a SHARED inversion (batched-affine bucket accumulation, msm.cuh k_pair_round) ------------------
// The sum of two affine points costs one inversion, and inversions can be shared (Montgomery's trick): with the inverse of
// the denominator handed in, an addition is 3 multiplications (lambda, lambda^2, y3) + 3 for the trick itself, against 10
// for the XYZZ mixed addition.  pair_prepare classifies the pair and returns the denominator to be inverted (1 when none is
// needed, so that the shared product never becomes zero); pair_finish completes it.  Inputs are adversarial (repeated
// points, P / -P pairs, infinity): every case is exact.
enum PairKind : int { PAIR_ADD = 0, PAIR_TAKE_A = 1, PAIR_TAKE_B = 2, PAIR_INF = 3, PAIR_DBL = 4 };
template <class F>
B200_HD int pair_prepare(const Affine<F>& a, const Affine<F>& b, F& den) {
  den = FieldOps<F>::one();
  if (is_inf(b)) return is_inf(a) ? PAIR_INF : PAIR_TAKE_A;
  if (is_inf(a)) return PAIR_TAKE_B;
  const F dx = sub(b.x, a.x);
  if (is_zero(dx)) {
    if (!eq(a.y, b.y)) return PAIR_INF;     // opposite points
    den = dbl(a.y);                         // the same point: tangent (y != 0: the curves have odd order)
    return PAIR_DBL;
  }
  den = dx;
  return PAIR_ADD;
}
template <class F>
B200_HD Affine<F> pair_finish(int kind, const Affine<F>& a, const Affine<F>& b, const F& inv_den) {
  if (kind == PAIR_TAKE_A) return a;
  if (kind == PAIR_TAKE_B) return b;
  Affine<F> r;
  if (kind == PAIR_INF) { r.x = FieldOps<F>::zero(); r.y = FieldOps<F>::zero(); return r; }
  F num = sub(b.y, a.y);
  if (kind == PAIR_DBL) { const F xx = sqr(a.x); num = add(dbl(xx), xx); }
  const F lam = mul(num, inv_den);
  r.x = sub(sub(sqr(lam), a.x), b.x);
  r.y = sub(mul(lam, sub(a.x, r.x)), a.y);
  return r;
}

// ---- homogeneous projective coordinates with the COMPLETE formulas of Renes-Costello-Batina (a = 0) -------------------
// x = X/Z, y = Y/Z, infinity = (0:1:0).  Used by the latency-bound tail of the MSM (window walk, window combine): both the
// addition (Alg. 7: 12 products in TWO dependency levels) and the doubling (Alg. 9: 8 products in two levels) are
// shallower than the XYZZ forms (4 and 3 levels), and they have no exceptional cases on E(Fp) and E'(Fp2), whose orders
// are odd (the only exceptional pairs of the underlying Bosma-Lenstra law differ by a point of order two).
template <class F>
struct Hom { F x, y, z; };
B200_HD Fp mul_b3(const Fp& t) {            // 3b = 12 on E: y^2 = x^3 + 4
  const Fp t4 = dbl(dbl(t));
  return add(dbl(t4), t4);
}
B200_HD Fp2 mul_b3(const Fp2& t) {          // 3b' = 12 (1 + u) on E': y^2 = x^3 + 4 (1 + u)
  const Fp2 t4 = dbl(dbl(mul_xi(t)));
  return add(dbl(t4), t4);
}
template <class F>
B200_HD Hom<F> hom_inf() { Hom<F> r; r.x = FieldOps<F>::zero(); r.y = FieldOps<F>::one(); r.z = FieldOps<F>::zero(); return r; }
template <class F>
B200_HD_NI Hom<F> xyzz_to_hom(const XYZZ<F>& p) {      // X ZZZ : Y ZZ : ZZ ZZZ
  if (is_inf(p)) return hom_inf<F>();
  Hom<F> r;
  r.x = mul(p.x, p.zzz); r.y = mul(p.y, p.zz); r.z = mul(p.zz, p.zzz);
  return r;
}
template <class F>
B200_HD_NI XYZZ<F> hom_to_xyzz(const Hom<F>& p) {      // X Z, Y Z^2, Z^2, Z^3
  XYZZ<F> r;
  r.zz = sqr(p.z);
  r.zzz = mul(r.zz, p.z);
  r.x = mul(p.x, p.z);
  r.y = mul(p.y, r.zz);
  return r;
}
template <class F>
B200_HD_NI Hom<F> hom_dbl(const Hom<F>& p) {
  const F t0 = sqr(p.y), t1 = mul(p.y, p.z), t3 = mul(p.x, p.y);
  const F z8 = dbl(dbl(dbl(t0)));
  const F t2 = mul_b3(sqr(p.z));
  const F t0b = sub(t0, add(dbl(t2), t2));
  Hom<F> r;
  r.x = dbl(mul(t0b, t3));
  r.y = add(mul(t2, z8), mul(t0b, add(t0, t2)));
  r.z = mul(t1, z8);
  return r;
}
template <class F>
B200_HD_NI Hom<F> hom_add(const Hom<F>& p, const Hom<F>& q) {
  F t0 = mul(p.x, q.x), t1 = mul(p.y, q.y), t2 = mul(p.z, q.z);
  const F t3 = sub(sub(mul(add(p.x, p.y), add(q.x, q.y)), t0), t1);
  const F t4 = sub(sub(mul(add(p.y, p.z), add(q.y, q.z)), t1), t2);
  const F y3 = mul_b3(sub(sub(mul(add(p.x, p.z), add(q.x, q.z)), t0), t2));
  t0 = add(dbl(t0), t0);
  t2 = mul_b3(t2);
  const F z3 = add(t1, t2);
  t1 = sub(t1, t2);
  Hom<F> r;
  r.x = sub(mul(t3, t1), mul(t4, y3));
  r.y = add(mul(t1, z3), mul(y3, t0));
  r.z = add(mul(z3, t4), mul(t0, t3));
  return r;
}

// x = X/ZZ, y = Y/ZZZ with one inversion; infinity -> (0,0)
template <class F>
B200_HD_NI Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
  Affine<F> r;
  if (is_inf(p)) { r.x = FieldOps<F>::zero(); r.y = FieldOps<F>::zero(); return r; }
  F i = inv(mul(p.zz, p.zzz));
  r.x = mul(p.x, mul(i, p.zzz));
  r.y = mul(p.y, mul(i, p.zz));
  return r;
}

// k * P by double-and-add, k given as little-endian 32-bit words, `nbits` significant bits
template <class F>
B200_HD_NI XYZZ<F> xyzz_scalar_mul(const Affine<F>& p, const uint32_t* k, int nbits) {
  XYZZ<F> acc = xyzz_inf<F>();
  for (int i = nbits - 1; i >= 0; i--) {
    acc = xyzz_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_madd(acc, p);
  }
  return acc;
}

// ---- Jacobian coordinates (x = X/Z^2, y = Y/Z^3), used where doublings dominate: the fixed-scalar
// multiplications of the subgroup checks.  dbl-2009-l costs 2M+5S (XYZZ: 6M+3S), madd-2007-bl 7M+4S.
template <class F>
struct Jac {     // infinity <=> z == 0
  F x, y, z;
};
template <class F>
B200_HD_NI void jac_dbl(Jac<F>& p) {
  if (is_zero(p.z)) return;
  if constexpr (sizeof(F) == sizeof(Fp)) {
    // Fp (a squaring costs a multiplication): D = 4*X*B directly, and Y3 = E*(D - X3) - 8*B*B as ONE fused
    // difference of products -- 5 multiplications + 1 fused pair instead of 7 multiplications
    F a = sqr(p.x), b = sqr(p.y);
    F d = dbl(dbl(mul(p.x, b)));
    F e = add(dbl(a), a);
    F z3 = dbl(mul(p.y, p.z));
    F x3 = sub(sqr(e), dbl(d));
    p.y = mul_diff(e, sub(d, x3), dbl(dbl(dbl(b))), b);
    p.x = x3;
    p.z = z3;
  } else {
    F a = sqr(p.x), b = sqr(p.y), c = sqr(b);
    F d = dbl(sub(sub(sqr(add(p.x, b)), a), c));
    F e = add(dbl(a), a);
    F f = sqr(e);
    F z3 = dbl(mul(p.y, p.z));
    F x3 = sub(f, dbl(d));
    F c8 = dbl(dbl(dbl(c)));
    p.y = sub(mul(e, sub(d, x3)), c8);
    p.x = x3;
    p.z = z3;
  }
}
// p += q, q affine and finite; complete (p infinite, p == q, p == -q handled exactly)
template <class F>
B200_HD_NI void jac_madd(Jac<F>& p, const Affine<F>& q) {
  if (is_zero(p.z)) { p.x = q.x; p.y = q.y; p.z = FieldOps<F>::one(); return; }
  F z1z1 = sqr(p.z);
  F u2 = mul(q.x, z1z1);
  F s2 = mul(mul(q.y, p.z), z1z1);
  F h = sub(u2, p.x);
  F r = sub(s2, p.y);
  if (is_zero(h)) {
    if (is_zero(r)) { p.x = q.x; p.y = q.y; p.z = FieldOps<F>::one(); jac_dbl(p); }
    else            { p.x = FieldOps<F>::zero(); p.y = FieldOps<F>::zero(); p.z = FieldOps<F>::zero(); }
    return;
  }
  r = dbl(r);
  F hh = sqr(h);
  F i = dbl(dbl(hh));
  F j = mul(h, i);
  F v = mul(p.x, i);
  F x3 = sub(sub(sqr(r), j), dbl(v));
  F y3 = mul_diff(r, sub(v, x3), dbl(p.y), j);
  F z3 = sub(sub(sqr(add(p.z, h)), z1z1), hh);
  p.x = x3; p.y = y3; p.z = z3;
}
// k * P for a finite affine P, k = `nbits` significant bits of little-endian words
template <class F>
B200_HD_NI Jac<F> jac_scalar_mul(const Affine<F>& p, const uint32_t* k, int nbits) {
  Jac<F> acc;
  acc.x = FieldOps<F>::zero(); acc.y = FieldOps<F>::zero(); acc.z = FieldOps<F>::zero();
  for (int i = nbits - 1; i >= 0; i--) {
    jac_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) jac_madd(acc, p);
  }
  return acc;
}

// Jacobian (X, Y, Z) -> XYZZ (X, Y, Z^2, Z^3): the same point, x = X/Z^2, y = Y/Z^3
template <class F>
B200_HD XYZZ<F> jac_to_xyzz(const Jac<F>& j) {
  XYZZ<F> r;
  r.x = j.x; r.y = j.y; r.zz = sqr(j.z); r.zzz = mul(r.zz, j.z);
  return r;
}

using G1Affine = Affine<Fp>;
using G2Affine = Affine<Fp2>;
using G1XYZZ = XYZZ<Fp>;
using G2XYZZ = XYZZ<Fp2>;

}  // namespace b200
