// coop.cuh -- lane-cooperative point arithmetic for the latency-bound tail of the MSM.
//
// The tail of a Pippenger MSM (upper levels of the bucket-reduction tree, the Horner combine of
// the window sums with its ~240 sequential doublings) has almost no data parallelism; on one
// thread each point operation is a chain of 9-14 dependent Fp multiplications (~1 us each when a
// warp runs alone).  Here a group of 8 lanes holds the SAME point in registers; at every
// dependency level of the formula each lane computes one of the independent products, and the
// products are exchanged with warp shuffles.  An XYZZ addition becomes 4 multiplication latencies
// instead of 14, a doubling 3 instead of 9.  Results are bit-identical to ec.cuh's
// xyzz_add / xyzz_dbl (same formulas, same exceptional-case handling).
#pragma once
#include "ec.cuh"

namespace b200 {
#ifdef __CUDACC__

// Group geometry: 8 product slots.  Over Fp a slot is one lane (group = 8 lanes).  Over Fp2 a slot
// is 4 lanes, three of which compute the three Fp products of a Karatsuba Fp2 multiplication in
// parallel (group = a full warp), so a G2 point operation also costs ~1 Fp-multiplication latency
// per formula level instead of 3.
template <class F> struct Coop;
template <> struct Coop<Fp>  { static constexpr int LPS = 1, LANES = 8; };
template <> struct Coop<Fp2> { static constexpr int LPS = 4, LANES = 32; };

struct CoopGroup {
  int lane;        // product slot 0..7 inside the group
  int sub;         // lane inside the slot (Fp2: 0..3)
  int base;        // warp lane of the group's first lane
  int lps;         // lanes per slot
  unsigned mask;   // participating warp lanes
};
template <class F>
__device__ __forceinline__ CoopGroup coop_group() {
  CoopGroup g;
  const int wl = threadIdx.x & 31;
  g.base = wl & ~(Coop<F>::LANES - 1);
  const int in = wl - g.base;
  g.lps = Coop<F>::LPS;
  g.lane = in / Coop<F>::LPS;
  g.sub = in % Coop<F>::LPS;
  g.mask = Coop<F>::LANES == 32 ? 0xffffffffu : (0xFFu << g.base);
  return g;
}

// broadcast the value held by product slot `src` to every lane of the group
template <class F>
__device__ __forceinline__ F coop_bcast(const F& v, int src, const CoopGroup& g) {
  F r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_sync(g.mask, s[i], g.base + src * g.lps);
  return r;
}
template <class F>
__device__ __forceinline__ void coop_pick(F& dst, bool take, const F& v) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = take ? s[i] : d[i];
}
// the slot's product a*b, available in every lane of the slot
// The cooperative kernels are latency-bound on a handful of warps, so their multiplication is the fully unrolled CIOS form
// (1.34 us per dependent multiplication on a lone warp against 1.60 us rolled, profiles/r02_k1_probes.md); the code size
// that stops the throughput kernels from using it does not matter in these small kernels.
__device__ __forceinline__ Fp coop_mul(const Fp& a, const Fp& b) {
#ifdef __CUDA_ARCH__
  return mul_unrolled<12>(a, b);
#else
  return mul(a, b);
#endif
}
__device__ __forceinline__ Fp coop_product(const Fp& a, const Fp& b, const CoopGroup&) { return coop_mul(a, b); }
__device__ __forceinline__ Fp2 coop_product(const Fp2& a, const Fp2& b, const CoopGroup& g) {
  // sub-lane 0: a0*b0, 1: a1*b1, 2 (and the spare 3): (a0+a1)*(b0+b1)
  Fp x = add(a.c0, a.c1), y = add(b.c0, b.c1);
  coop_pick(x, g.sub == 0, a.c0); coop_pick(y, g.sub == 0, b.c0);
  coop_pick(x, g.sub == 1, a.c1); coop_pick(y, g.sub == 1, b.c1);
  Fp t = coop_mul(x, y);
  const int slot0 = g.base + g.lane * 4;
  Fp t0, t1, t2;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    t0.v[i] = __shfl_sync(g.mask, t.v[i], slot0);
    t1.v[i] = __shfl_sync(g.mask, t.v[i], slot0 + 1);
    t2.v[i] = __shfl_sync(g.mask, t.v[i], slot0 + 2);
  }
  Fp2 r;
  r.c0 = sub(t0, t1);
  r.c1 = sub(sub(t2, t0), t1);
  return r;
}

// 2*p ; every lane of the group passes the same p and receives the same result
template <class F>
__device__ __noinline__ void coop_dbl(XYZZ<F>& p, const CoopGroup g) {
  if (is_inf(p)) return;
  const int l = g.lane;
  F u = dbl(p.y);
  // level 1: v = u^2, x2 = x^2
  F a = u;
  coop_pick(a, l == 1, p.x);
  F r1 = coop_product(a, a, g);
  F v = coop_bcast(r1, 0, g), x2 = coop_bcast(r1, 1, g);
  F m = add(dbl(x2), x2);
  // level 2: w = u*v, s = x*v, mm = m*m, zz3 = v*zz
  a = u; F b = v;
  coop_pick(a, l == 1, p.x);
  coop_pick(a, l == 2, m); coop_pick(b, l == 2, m);
  coop_pick(a, l == 3, p.zz);
  F r2 = coop_product(a, b, g);
  F w = coop_bcast(r2, 0, g), s = coop_bcast(r2, 1, g), mm = coop_bcast(r2, 2, g), zz3 = coop_bcast(r2, 3, g);
  F x3 = sub(mm, dbl(s));
  // level 3: ya = m*(s - x3), yb = w*y, zzz3 = w*zzz
  a = m; b = sub(s, x3);
  coop_pick(a, l == 1, w); coop_pick(b, l == 1, p.y);
  coop_pick(a, l == 2, w); coop_pick(b, l == 2, p.zzz);
  F r3 = coop_product(a, b, g);
  F ya = coop_bcast(r3, 0, g), yb = coop_bcast(r3, 1, g), zzz3 = coop_bcast(r3, 2, g);
  p.x = x3; p.y = sub(ya, yb); p.zz = zz3; p.zzz = zzz3;
}

// acc += q ; same contract
template <class F>
__device__ __noinline__ void coop_add(XYZZ<F>& acc, const XYZZ<F>& q, const CoopGroup g) {
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = q; return; }
  const int l = g.lane;
  // level 1: u1 = X1*ZZ2, u2 = X2*ZZ1, s1 = Y1*ZZZ2, s2 = Y2*ZZZ1, zz12 = ZZ1*ZZ2, zzz12 = ZZZ1*ZZZ2
  F a = acc.x, b = q.zz;
  coop_pick(a, l == 1, q.x);     coop_pick(b, l == 1, acc.zz);
  coop_pick(a, l == 2, acc.y);   coop_pick(b, l == 2, q.zzz);
  coop_pick(a, l == 3, q.y);     coop_pick(b, l == 3, acc.zzz);
  coop_pick(a, l == 4, acc.zz);  coop_pick(b, l == 4, q.zz);
  coop_pick(a, l == 5, acc.zzz); coop_pick(b, l == 5, q.zzz);
  F r1 = coop_product(a, b, g);
  F u1 = coop_bcast(r1, 0, g), u2 = coop_bcast(r1, 1, g), s1 = coop_bcast(r1, 2, g), s2 = coop_bcast(r1, 3, g);
  F zz12 = coop_bcast(r1, 4, g), zzz12 = coop_bcast(r1, 5, g);
  F p = sub(u2, u1), r = sub(s2, s1);
  if (is_zero(p)) {                       // uniform across the group
    if (is_zero(r)) coop_dbl(acc, g);     // same point
    else            acc = xyzz_inf<F>();  // opposite points
    return;
  }
  // level 2: pp = p^2, rr = r^2
  a = p; coop_pick(a, l == 1, r);
  F r2 = coop_product(a, a, g);
  F pp = coop_bcast(r2, 0, g), rr = coop_bcast(r2, 1, g);
  // level 3: ppp = p*pp, qq = u1*pp, zz3 = zz12*pp
  a = p; coop_pick(a, l == 1, u1); coop_pick(a, l == 2, zz12);
  F r3 = coop_product(a, pp, g);
  F ppp = coop_bcast(r3, 0, g), qq = coop_bcast(r3, 1, g), zz3 = coop_bcast(r3, 2, g);
  F x3 = sub(sub(rr, ppp), dbl(qq));
  // level 4: ya = r*(qq - x3), yb = s1*ppp, zzz3 = zzz12*ppp
  a = r; b = sub(qq, x3);
  coop_pick(a, l == 1, s1);    coop_pick(b, l == 1, ppp);
  coop_pick(a, l == 2, zzz12); coop_pick(b, l == 2, ppp);
  F r4 = coop_product(a, b, g);
  F ya = coop_bcast(r4, 0, g), yb = coop_bcast(r4, 1, g), zzz3 = coop_bcast(r4, 2, g);
  acc.x = x3; acc.y = sub(ya, yb); acc.zz = zz3; acc.zzz = zzz3;
}

// ---- homogeneous coordinates (ec.cuh Hom): two dependency levels per operation, no exceptional cases --------------------
template <class F>
__device__ __noinline__ void coop_hdbl(Hom<F>& p, const CoopGroup g) {
  const int l = g.lane;
  // level 1: t0 = Y^2, t1 = Y Z, t2 = Z^2, t3 = X Y
  F a = p.y, b = p.y;
  coop_pick(b, l == 1, p.z);
  coop_pick(a, l == 2, p.z); coop_pick(b, l == 2, p.z);
  coop_pick(a, l == 3, p.x);
  const F r1 = coop_product(a, b, g);
  const F t0 = coop_bcast(r1, 0, g), t1 = coop_bcast(r1, 1, g), t3 = coop_bcast(r1, 3, g);
  const F t2 = mul_b3(coop_bcast(r1, 2, g));
  const F z8 = dbl(dbl(dbl(t0)));
  const F t0b = sub(t0, add(dbl(t2), t2));
  // level 2: t2 z8, t1 z8, t0b (t0 + t2), t0b t3
  a = t2; b = z8;
  coop_pick(a, l == 1, t1);
  coop_pick(a, l == 2, t0b); coop_pick(b, l == 2, add(t0, t2));
  coop_pick(a, l == 3, t0b); coop_pick(b, l == 3, t3);
  const F r2 = coop_product(a, b, g);
  p.x = dbl(coop_bcast(r2, 3, g));
  p.y = add(coop_bcast(r2, 0, g), coop_bcast(r2, 2, g));
  p.z = coop_bcast(r2, 1, g);
}
template <class F>
__device__ __noinline__ void coop_hadd(Hom<F>& acc, const Hom<F>& q, const CoopGroup g) {
  const int l = g.lane;
  // level 1: X1X2, Y1Y2, Z1Z2, (X1+Y1)(X2+Y2), (Y1+Z1)(Y2+Z2), (X1+Z1)(X2+Z2)
  F a = acc.x, b = q.x;
  coop_pick(a, l == 1, acc.y); coop_pick(b, l == 1, q.y);
  coop_pick(a, l == 2, acc.z); coop_pick(b, l == 2, q.z);
  coop_pick(a, l == 3, add(acc.x, acc.y)); coop_pick(b, l == 3, add(q.x, q.y));
  coop_pick(a, l == 4, add(acc.y, acc.z)); coop_pick(b, l == 4, add(q.y, q.z));
  coop_pick(a, l == 5, add(acc.x, acc.z)); coop_pick(b, l == 5, add(q.x, q.z));
  const F r1 = coop_product(a, b, g);
  F t0 = coop_bcast(r1, 0, g), t1 = coop_bcast(r1, 1, g), t2 = coop_bcast(r1, 2, g);
  const F t3 = sub(sub(coop_bcast(r1, 3, g), t0), t1);
  const F t4 = sub(sub(coop_bcast(r1, 4, g), t1), t2);
  const F y3 = mul_b3(sub(sub(coop_bcast(r1, 5, g), t0), t2));
  t0 = add(dbl(t0), t0);
  t2 = mul_b3(t2);
  const F z3 = add(t1, t2);
  t1 = sub(t1, t2);
  // level 2: t3 t1, t4 y3, t1 z3, y3 t0, z3 t4, t0 t3
  a = t3; b = t1;
  coop_pick(a, l == 1, t4); coop_pick(b, l == 1, y3);
  coop_pick(a, l == 2, t1); coop_pick(b, l == 2, z3);
  coop_pick(a, l == 3, y3); coop_pick(b, l == 3, t0);
  coop_pick(a, l == 4, z3); coop_pick(b, l == 4, t4);
  coop_pick(a, l == 5, t0); coop_pick(b, l == 5, t3);
  const F r2 = coop_product(a, b, g);
  acc.x = sub(coop_bcast(r2, 0, g), coop_bcast(r2, 1, g));
  acc.y = add(coop_bcast(r2, 2, g), coop_bcast(r2, 3, g));
  acc.z = add(coop_bcast(r2, 4, g), coop_bcast(r2, 5, g));
}
// conversions: one and two levels
template <class F>
__device__ __noinline__ Hom<F> coop_xyzz_to_hom(const XYZZ<F>& p, const CoopGroup g) {
  if (is_inf(p)) return hom_inf<F>();      // uniform across the group
  const int l = g.lane;
  F a = p.x, b = p.zzz;
  coop_pick(a, l == 1, p.y); coop_pick(b, l == 1, p.zz);
  coop_pick(a, l == 2, p.zz);
  const F r1 = coop_product(a, b, g);
  Hom<F> r;
  r.x = coop_bcast(r1, 0, g); r.y = coop_bcast(r1, 1, g); r.z = coop_bcast(r1, 2, g);
  return r;
}
template <class F>
__device__ __noinline__ XYZZ<F> coop_hom_to_xyzz(const Hom<F>& p, const CoopGroup g) {
  const int l = g.lane;
  F a = p.z, b = p.z;
  coop_pick(a, l == 1, p.x);
  const F r1 = coop_product(a, b, g);
  XYZZ<F> r;
  r.zz = coop_bcast(r1, 0, g); r.x = coop_bcast(r1, 1, g);
  a = r.zz; b = p.z;
  coop_pick(b, l == 1, p.y);
  const F r2 = coop_product(a, b, g);
  r.zzz = coop_bcast(r2, 0, g); r.y = coop_bcast(r2, 1, g);
  return r;
}

#endif  // __CUDACC__
}  // namespace b200
