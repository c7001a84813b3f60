// map.cuh -- MAP_FP_TO_G1 / MAP_FP2_TO_G2 (reference: src/eip2537.c:1093-1121, :1135-1163, which call
// blst_map_to_g1 / blst_map_to_g2 with v = NULL).  blst is not part of the reference tree; the published
// algorithm is RFC 9380 section 8.8:
//   1. simplified SWU onto the isogenous curve E' : y^2 = x^3 + A'x + B'   (6.6.2; straight-line form F.2)
//   2. the 11-isogeny E1' -> E1  /  3-isogeny E2' -> E2                    (E.2 / E.3)
//   3. cofactor clearing: [1 - z]P on G1, Budroni-Pintore psi-based h_eff on G2 (8.8.1 / 8.8.2, G.4)
// The isogeny coefficients in constants.cuh are DERIVED (division polynomial + Kohel's formula,
// oracle/derive_isogeny.py) and matched against the RFC's published low-order coefficients.
//
// Formulation notes (deliberately different from the CPU oracle, which uses the textbook 6.6.2 form with
// inversions and the "complex method" square root, so that the parity tests compare two derivations):
//   * Fp:  inversion-free SSWU with sqrt_ratio for p = 3 mod 4 (one exponentiation by (p-3)/4).
//   * Fp2: the same straight-line SSWU; sqrt_ratio(n, d) via a square root of n*d (or Z*n*d) computed with
//          the norm method: sqrt(a+bi) = (r, b/2r), r^2 = (a +- sqrt(a^2+b^2))/2  -- Fp exponentiations only.
// One thread maps one field element; everything is B200_HD so tests/host_emul runs it on the CPU.
#pragma once
#include "ec.cuh"
#include "codec.cuh"

namespace b200 {

// a^e, e = 12 raw little-endian words (constants.cuh EXP_*), fixed 4-bit windows
B200_HD_NI Fp fp_pow_const(const Fp& a, const uint32_t* e) {
  Fp tab[16];
  tab[0] = fp_one();
  tab[1] = a;
  for (int i = 2; i < 16; i++) tab[i] = mul(tab[i - 1], a);
  Fp acc = fp_one();
  bool started = false;
  for (int w = 95; w >= 0; w--) {
    uint32_t d = (e[w >> 3] >> ((w & 7) * 4)) & 0xF;
    if (started) { acc = sqr(acc); acc = sqr(acc); acc = sqr(acc); acc = sqr(acc); }
    if (d) { acc = started ? mul(acc, tab[d]) : tab[d]; started = true; }
  }
  return acc;
}

// square root in Fp (p = 3 mod 4): returns true and r with r^2 == a iff a is a square
B200_HD bool fp_sqrt(Fp& r, const Fp& a) {
  r = fp_pow_const(a, C_EXP_PP1D4());
  return eq(sqr(r), a);
}

// square root in Fp2 by the norm method; true iff a is a square
B200_HD_NI bool fp2_sqrt(Fp2& r, const Fp2& a) {
  if (is_zero(a.c1)) {                       // a in Fp: sqrt(a) or i*sqrt(-a)
    Fp s;
    if (fp_sqrt(s, a.c0)) { r.c0 = s; r.c1 = fp_zero(); return true; }
    fp_sqrt(s, neg(a.c0));
    r.c0 = fp_zero(); r.c1 = s;
    return true;
  }
  Fp n = add(sqr(a.c0), sqr(a.c1)), s;
  if (!fp_sqrt(s, n)) return false;          // a square in Fp2 <=> its norm is a square in Fp
  Fp t = half(add(a.c0, s)), x;
  if (!fp_sqrt(x, t)) fp_sqrt(x, sub(t, s)); // exactly one of (a0 +- s)/2 is a square when b != 0
  r.c0 = x;
  r.c1 = mul(a.c1, inv(dbl(x)));
  return true;
}

B200_HD int sgn0(const Fp& a) { return (int)(fp_from_mont(a).v[0] & 1); }
B200_HD int sgn0(const Fp2& a) {             // RFC 9380 4.1, m = 2
  Fp c0 = fp_from_mont(a.c0);
  int s0 = (int)(c0.v[0] & 1), z0 = is_zero(c0) ? 1 : 0, s1 = (int)(fp_from_mont(a.c1).v[0] & 1);
  return s0 | (z0 & s1);
}

template <class F> struct MapConsts;
template <> struct MapConsts<Fp> {
  static B200_HD Fp A() { return fp_load_const(C_ISO1_A()); }
  static B200_HD Fp B() { return fp_load_const(C_ISO1_B()); }
  static B200_HD Fp Z() { return fp_load_const(C_ISO1_Z()); }
  static B200_HD Fp coef(const uint32_t* t, int i) { return fp_load_const(t + 12 * i); }
  static B200_HD const uint32_t* xnum() { return C_ISO1_XNUM(); }
  static B200_HD const uint32_t* xden() { return C_ISO1_XDEN(); }
  static B200_HD const uint32_t* ynum() { return C_ISO1_YNUM(); }
  static B200_HD const uint32_t* yden() { return C_ISO1_YDEN(); }
  static constexpr int NXN = 12, NXD = 11, NYN = 16, NYD = 16;
};
template <> struct MapConsts<Fp2> {
  static B200_HD Fp2 A() { return fp2_load_const(C_ISO2_A()); }
  static B200_HD Fp2 B() { return fp2_load_const(C_ISO2_B()); }
  static B200_HD Fp2 Z() { return fp2_load_const(C_ISO2_Z()); }
  static B200_HD Fp2 coef(const uint32_t* t, int i) { return fp2_load_const(t + 24 * i); }
  static B200_HD const uint32_t* xnum() { return C_ISO2_XNUM(); }
  static B200_HD const uint32_t* xden() { return C_ISO2_XDEN(); }
  static B200_HD const uint32_t* ynum() { return C_ISO2_YNUM(); }
  static B200_HD const uint32_t* yden() { return C_ISO2_YDEN(); }
  static constexpr int NXN = 4, NXD = 3, NYN = 4, NYD = 4;
};

// sqrt_ratio(n, d) of RFC 9380 F.2.1: (true, sqrt(n/d)) if n/d is a square, else (false, sqrt(Z*n/d)).  d != 0.
B200_HD_NI bool sqrt_ratio(Fp& y, const Fp& n, const Fp& d) {      // F.2.1.2, q = 3 mod 4
  Fp tv1 = sqr(d), tv2 = mul(n, d);
  tv1 = mul(tv1, tv2);
  Fp y1 = mul(fp_pow_const(tv1, C_EXP_PM3D4()), tv2);
  bool is_qr = eq(mul(sqr(y1), d), n);
  y = is_qr ? y1 : mul(y1, fp_load_const(C_ISO1_SQRT_MZ()));
  return is_qr;
}
B200_HD_NI bool sqrt_ratio(Fp2& y, const Fp2& n, const Fp2& d) {   // sqrt(n/d) = sqrt(n*d)/d
  Fp2 nd = mul(n, d), s;
  bool is_qr = fp2_sqrt(s, nd);
  if (!is_qr) fp2_sqrt(s, mul(nd, MapConsts<Fp2>::Z()));
  y = mul(s, inv(d));
  return is_qr;
}

// simplified SWU, straight-line (RFC 9380 F.2): affine point on E'
template <class F>
B200_HD_NI Affine<F> sswu(const F& u) {
  using C = MapConsts<F>;
  const F A = C::A(), B = C::B(), Zc = C::Z();
  F tv1 = mul(Zc, sqr(u));
  F tv2 = add(sqr(tv1), tv1);
  F tv3 = mul(B, add(tv2, FieldOps<F>::one()));
  F tv4 = mul(A, is_zero(tv2) ? Zc : neg(tv2));
  tv2 = sqr(tv3);
  F tv6 = sqr(tv4);
  F tv5 = mul(A, tv6);
  tv2 = mul(add(tv2, tv5), tv3);
  tv6 = mul(tv6, tv4);
  tv5 = mul(B, tv6);
  tv2 = add(tv2, tv5);                         // gx1 = tv2 / tv6,  x1 = tv3 / tv4
  F x = mul(tv1, tv3);                         // x2 numerator
  F y1;
  bool gx1_square = sqrt_ratio(y1, tv2, tv6);
  F y = mul(mul(tv1, u), y1);
  if (gx1_square) { x = tv3; y = y1; }
  if (sgn0(u) != sgn0(y)) y = neg(y);
  Affine<F> r;
  r.x = mul(x, inv(tv4));
  r.y = y;
  return r;
}

template <class F>
B200_HD F iso_horner(const uint32_t* tab, int n, const F& x) {
  F acc = MapConsts<F>::coef(tab, n - 1);
  for (int i = n - 2; i >= 0; i--) acc = add(mul(acc, x), MapConsts<F>::coef(tab, i));
  return acc;
}

// isogeny E' -> E; a kernel point (denominator 0) maps to infinity, encoded (0,0)
template <class F>
B200_HD_NI Affine<F> iso_map(const Affine<F>& p) {
  using C = MapConsts<F>;
  F xn = iso_horner<F>(C::xnum(), C::NXN, p.x), xd = iso_horner<F>(C::xden(), C::NXD, p.x);
  F yn = iso_horner<F>(C::ynum(), C::NYN, p.x), yd = iso_horner<F>(C::yden(), C::NYD, p.x);
  Affine<F> r;
  F den = mul(xd, yd);
  if (is_zero(den)) { r.x = FieldOps<F>::zero(); r.y = FieldOps<F>::zero(); return r; }
  F i = inv(den);
  r.x = mul(xn, mul(i, yd));
  r.y = mul(p.y, mul(yn, mul(i, xd)));
  return r;
}

// G1: h_eff = 1 - z = 0xd201000000010001
B200_HD_NI G1Affine clear_cofactor(const G1Affine& p) {
  if (is_inf(p)) return p;
  const uint32_t k[2] = {(uint32_t)((B200_Z_ABS + 1) & 0xffffffffu), (uint32_t)((B200_Z_ABS + 1) >> 32)};
  return xyzz_to_affine(jac_to_xyzz(jac_scalar_mul(p, k, 64)));   // doubling-dominated: Jacobian ladder (2M+5S per doubling)
}

// psi on the twist: (conj(x) * cx, conj(y) * cy); conj is a field automorphism so it passes through XYZZ
B200_HD XYZZ<Fp2> g2_psi(const XYZZ<Fp2>& p) {
  XYZZ<Fp2> r;
  r.x = mul(conj(p.x), fp2_load_const(C_PSI_CX()));
  r.y = mul(conj(p.y), fp2_load_const(C_PSI_CY()));
  r.zz = conj(p.zz);
  r.zzz = conj(p.zzz);
  return r;
}
// G2: Budroni-Pintore, h_eff * P = [z^2 - z - 1]P + [z - 1]psi(P) + psi^2(2P)   (z < 0)
B200_HD_NI G2Affine clear_cofactor(const G2Affine& p) {
  if (is_inf(p)) return p;
  const uint32_t zabs[2] = {(uint32_t)(B200_Z_ABS & 0xffffffffu), (uint32_t)(B200_Z_ABS >> 32)};
  XYZZ<Fp2> P0 = xyzz_from_affine(p);
  XYZZ<Fp2> t1 = xyzz_neg(jac_to_xyzz(jac_scalar_mul(p, zabs, 64)));  // [z]P
  XYZZ<Fp2> t2 = g2_psi(P0);                                          // psi(P)
  XYZZ<Fp2> t3 = g2_psi(g2_psi(xyzz_dbl_affine(p)));                  // psi^2(2P)
  xyzz_add(t3, xyzz_neg(t2));                                         // psi^2(2P) - psi(P)
  xyzz_add(t2, t1);                                                   // [z]P + psi(P)
  G2Affine t2a = xyzz_to_affine(t2);
  XYZZ<Fp2> t4 = is_inf(t2a) ? xyzz_inf<Fp2>() : xyzz_neg(jac_to_xyzz(jac_scalar_mul(t2a, zabs, 64)));   // [z]([z]P + psi(P))
  xyzz_add(t3, t4);
  xyzz_add(t3, xyzz_neg(t1));
  xyzz_add(t3, xyzz_neg(P0));
  return xyzz_to_affine(t3);
}

// field element slot(s) -> point in G1 / G2, wire to wire.  Returns E_SUCCESS or E_INVALID_ELEMENT.
B200_HD_NI int map_to_group(uint32_t* out_words, const uint32_t* in_words, Fp*) {
  Fp u;
  if (fp_from_slot(u, in_words) < 0) return E_INVALID_ELEMENT;
  encode_point(out_words, clear_cofactor(iso_map<Fp>(sswu<Fp>(u))));
  return E_SUCCESS;
}
B200_HD_NI int map_to_group(uint32_t* out_words, const uint32_t* in_words, Fp2*) {
  Fp2 u;
  int s0 = fp_from_slot(u.c0, in_words), s1 = fp_from_slot(u.c1, in_words + 16);
  if (s0 < 0 || s1 < 0) return E_INVALID_ELEMENT;
  encode_point(out_words, clear_cofactor(iso_map<Fp2>(sswu<Fp2>(u))));
  return E_SUCCESS;
}

#ifdef __CUDACC__
// one thread per field element; in: n x (64 | 128) bytes, out: n x (128 | 256) bytes, codes[n]
template <class F>
__global__ void __launch_bounds__(64) k_map_to_group(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ out,
                                                      int* __restrict__ codes) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int IW = Wire<F>::POINT_WORDS / 2, OW = Wire<F>::POINT_WORDS;
  uint32_t wi[IW], wo[OW];
#pragma unroll
  for (int k = 0; k < IW; k++) wi[k] = in[i * IW + k];
#pragma unroll
  for (int k = 0; k < OW; k++) wo[k] = 0;
  int code = map_to_group(wo, wi, (F*)nullptr);
#pragma unroll
  for (int k = 0; k < OW; k++) out[i * OW + k] = wo[k];
  codes[i] = code;
}
#endif

}  // namespace b200
