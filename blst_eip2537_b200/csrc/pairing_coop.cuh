// pairing_coop.cuh -- latency path of PAIRING for SMALL batches: one block of two warps per PAIR.
//
// A single bls12_pairing call (the legacy ABI: one call per EVM precompile invocation) spends most of its time in two
// thread-serial chains before the warp-cooperative Miller accumulation (coop12.cuh) even starts: the G1 membership
// ladder (~1,000 dependent Fp multiplications on one thread, 1.6 ms) and the 68 line steps of the pair (~1,800, 2.9 ms).
// Here both are walked by lane groups, the independent products of every formula level on different lanes:
//   warp 0, lanes 0-7   decode P; [z^2]P with the complete homogeneous formulas (coop.cuh: two levels per doubling);
//                       phi(P) == -[z^2]P                                  (blst_p1_affine_in_g1, eip2537.c:1041)
//   warp 1 (8 x 4 lanes) decode Q; the Miller-loop walk T = [|z|]Q: two levels per doubling step, four per addition
//                       step, the lines scaled by xP / yP inside the same levels; psi(Q) == -T is the deferred G2
//                       membership test of pairing_dot.cuh                 (blst_p2_affine_in_g2, :1051; blst_miller_loop, :1060)
// Error codes and their precedence are those of k_pairing_decode (G1 decode, G1 subgroup, G2 decode, G2 subgroup).
#pragma once
#include "coop.cuh"
#include "pairing.cuh"

namespace b200 {
#ifdef __CUDACC__

__device__ __forceinline__ Fp2 fp2_from_fp(const Fp& a) { Fp2 r; r.c0 = a; r.c1 = fp_zero(); return r; }

// one doubling step of the Miller loop (ml_dbl_step) and its line, scaled: (l0, l1 * xP, l4 * yP)
__device__ __noinline__ void coop_line_dbl(G2Proj& t, Line& ln, const Fp2& pxe, const Fp2& pye, const CoopGroup g) {
  const int l = g.lane;
  // level 1: X Y, Y^2, Z^2, (Y + Z)^2, X^2
  Fp2 a = t.x, b = t.y;
  coop_pick(a, l == 1, t.y);
  coop_pick(a, l == 2, t.z); coop_pick(b, l == 2, t.z);
  const Fp2 yz = add(t.y, t.z);
  coop_pick(a, l == 3, yz); coop_pick(b, l == 3, yz);
  coop_pick(b, l == 4, t.x);
  const Fp2 r1 = coop_product(a, b, g);
  const Fp2 A = half(coop_bcast(r1, 0, g)), B = coop_bcast(r1, 1, g), C = coop_bcast(r1, 2, g), J = coop_bcast(r1, 4, g);
  const Fp2 H = sub(coop_bcast(r1, 3, g), add(B, C));
  const Fp2 c4 = dbl(dbl(mul_xi(C)));
  const Fp2 E = add(dbl(c4), c4);
  const Fp2 Fv = add(dbl(E), E);
  const Fp2 G = half(add(B, Fv));
  // level 2: A (B - F), G^2, E^2, B H, (3 J) xP, (-H) yP
  a = A; b = sub(B, Fv);
  coop_pick(a, l == 1, G); coop_pick(b, l == 1, G);
  coop_pick(a, l == 2, E); coop_pick(b, l == 2, E);
  coop_pick(a, l == 3, B); coop_pick(b, l == 3, H);
  coop_pick(a, l == 4, add(dbl(J), J)); coop_pick(b, l == 4, pxe);
  coop_pick(a, l == 5, neg(H)); coop_pick(b, l == 5, pye);
  const Fp2 r2 = coop_product(a, b, g);
  const Fp2 E2 = coop_bcast(r2, 2, g);
  t.x = coop_bcast(r2, 0, g);
  t.y = sub(coop_bcast(r2, 1, g), add(dbl(E2), E2));
  t.z = coop_bcast(r2, 3, g);
  ln.l0 = sub(E, B); ln.l1 = coop_bcast(r2, 4, g); ln.l4 = coop_bcast(r2, 5, g);
}
// one addition step (ml_add_step)
__device__ __noinline__ void coop_line_add(G2Proj& t, const G2Affine& q, Line& ln, const Fp2& pxe, const Fp2& pye, const CoopGroup g) {
  const int l = g.lane;
  // level 1: yQ Z, xQ Z
  Fp2 a = q.y, b = t.z;
  coop_pick(a, l == 1, q.x);
  const Fp2 r1 = coop_product(a, b, g);
  const Fp2 theta = sub(t.y, coop_bcast(r1, 0, g)), lam = sub(t.x, coop_bcast(r1, 1, g));
  // level 2: theta^2, lam^2, theta xQ, lam yQ, (-theta) xP, lam yP
  a = theta; b = theta;
  coop_pick(a, l == 1, lam); coop_pick(b, l == 1, lam);
  coop_pick(b, l == 2, q.x);
  coop_pick(a, l == 3, lam); coop_pick(b, l == 3, q.y);
  coop_pick(a, l == 4, neg(theta)); coop_pick(b, l == 4, pxe);
  coop_pick(a, l == 5, lam); coop_pick(b, l == 5, pye);
  const Fp2 r2 = coop_product(a, b, g);
  const Fp2 c = coop_bcast(r2, 0, g), d = coop_bcast(r2, 1, g);
  ln.l0 = sub(coop_bcast(r2, 2, g), coop_bcast(r2, 3, g)); ln.l1 = coop_bcast(r2, 4, g); ln.l4 = coop_bcast(r2, 5, g);
  // level 3: lam d, Z c, X d
  a = lam; b = d;
  coop_pick(a, l == 1, t.z); coop_pick(b, l == 1, c);
  coop_pick(a, l == 2, t.x);
  const Fp2 r3 = coop_product(a, b, g);
  const Fp2 e = coop_bcast(r3, 0, g), f = coop_bcast(r3, 1, g), gg = coop_bcast(r3, 2, g);
  const Fp2 h = sub(add(e, f), dbl(gg));
  // level 4: lam h, theta (g - h), e Y, Z e
  a = lam; b = h;
  coop_pick(a, l == 1, theta); coop_pick(b, l == 1, sub(gg, h));
  coop_pick(a, l == 2, e); coop_pick(b, l == 2, t.y);
  coop_pick(a, l == 3, t.z); coop_pick(b, l == 3, e);
  const Fp2 r4 = coop_product(a, b, g);
  t.x = coop_bcast(r4, 0, g);
  t.y = sub(coop_bcast(r4, 1, g), coop_bcast(r4, 2, g));
  t.z = coop_bcast(r4, 3, g);
}

__global__ void __launch_bounds__(64) k_pairing_pair_coop(const uint32_t* __restrict__ raw, size_t total_pairs, Line* __restrict__ lines,
                                                          unsigned char* __restrict__ skip, int* __restrict__ status) {
  __shared__ G1Affine sp;
  __shared__ int code1, code2, inf1, inf2, dec1;      // dec1: decode-level verdict on P, never written after the first barrier
  const size_t j = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* src = reinterpret_cast<const uint4*>(raw + j * 96);
  G1Affine p;
  G2Affine q;
  int code;
  if (warp == 0) {
    uint32_t w[32];
#pragma unroll
    for (int k = 0; k < 8; k++) { uint4 v = __ldg(src + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    code = decode_point(p, w);
    if (lane == 0) { sp = p; code1 = code; dec1 = code; inf1 = (code == E_SUCCESS && is_inf(p)) ? 1 : 0; }
  } else {
    uint32_t w[64];
#pragma unroll
    for (int k = 0; k < 16; k++) { uint4 v = __ldg(src + 8 + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    code = decode_point(q, w);
    if (lane == 0) { code2 = code; inf2 = (code == E_SUCCESS && is_inf(q)) ? 1 : 0; }
  }
  __syncthreads();
  if (warp == 0) {
    if (code == E_SUCCESS && !is_inf(p) && lane < 8) {
      // T = [z^2]P, most significant bit first; the formulas are complete, so P may be any curve point
      const CoopGroup g = coop_group<Fp>();
      const uint32_t zsq[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};
      Hom<Fp> base;
      base.x = p.x; base.y = p.y; base.z = fp_one();
      Hom<Fp> t = base;
#pragma unroll 1
      for (int i = 126; i >= 0; i--) {
        coop_hdbl(t, g);
        if ((zsq[i >> 5] >> (i & 31)) & 1u) coop_hadd(t, base, g);
      }
      // phi(P) == -T:  beta x Tz == Tx  and  y Tz == -Ty   (T = infinity: not a member)
      const int l = g.lane;
      Fp a = mul(p.x, fp_load_const(C_BETA())), b = t.z;
      coop_pick(a, l == 1, p.y);
      const Fp r = coop_product(a, b, g);
      const bool member = !is_zero(t.z) && eq(coop_bcast(r, 0, g), t.x) && eq(coop_bcast(r, 1, g), neg(t.y));
      if (lane == 0 && !member) code1 = E_NOT_IN_SUBGROUP;
    }
  } else if (code == E_SUCCESS && !is_inf(q)) {
    const int c1 = dec1;             // decode-level verdict on P (its membership ladder is still running on warp 0)
    if (c1 == E_SUCCESS && inf1) {
      // P = infinity: the pair contributes 1, but Q must still be in G2 -- no walk to piggyback on: exact ladder
      if (lane == 0 && !g2_in_subgroup(q)) code2 = E_NOT_IN_SUBGROUP;
    } else if (c1 == E_SUCCESS) {
      const CoopGroup g = coop_group<Fp2>();
      const Fp2 pxe = fp2_from_fp(sp.x), pye = fp2_from_fp(sp.y);
      G2Proj t;
      t.x = q.x; t.y = q.y; t.z = fp2_one();
      int s = 0;
#pragma unroll 1
      for (int i = 62; i >= 0; i--) {
        const int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
        for (int a = 0; a < nsteps; a++, s++) {
          Line ln;
          if (a == 0) coop_line_dbl(t, ln, pxe, pye, g); else coop_line_add(t, q, ln, pxe, pye, g);
          // every lane holds the line: 18 lanes store 16 bytes each
          if (lane < 18) reinterpret_cast<uint4*>(&lines[(size_t)s * total_pairs + j])[lane] = reinterpret_cast<const uint4*>(&ln)[lane];
        }
      }
      // psi(Q) == -T (see k_pairing_lines_slots); Z = 0 marks an exceptional chain: the exact ladder decides
      bool member;
      if (is_zero(t.z)) {
        member = g2_in_subgroup(q);
      } else {
        const int l = g.lane;
        Fp2 a = conj(q.x), b = fp2_load_const(C_PSI_CX());
        coop_pick(a, l == 1, conj(q.y)); coop_pick(b, l == 1, fp2_load_const(C_PSI_CY()));
        const Fp2 r1 = coop_product(a, b, g);
        a = coop_bcast(r1, 0, g);
        coop_pick(a, l == 1, coop_bcast(r1, 1, g));
        const Fp2 r2 = coop_product(a, t.z, g);
        member = eq(coop_bcast(r2, 0, g), t.x) && eq(coop_bcast(r2, 1, g), neg(t.y));
      }
      if (lane == 0 && !member) code2 = E_NOT_IN_SUBGROUP;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int c = code1 != E_SUCCESS ? code1 : code2;
    status[j] = c;
    skip[j] = (c != E_SUCCESS || inf1 || inf2) ? 1 : 0;
  }
}

#endif  // __CUDACC__
}  // namespace b200
