// pairing.cuh -- K5: batched optimal-ate pairing checks over BLS12-381 (+ K3 subgroup checks).
//
// Replaces, for many independent calls at once, the body of bls12_pairing
// (/root/reference/src/eip2537.c:1020-1081): per pair decode_g1_point (:1036),
// blst_p1_affine_in_g1 (:1041), decode_g2_point (:1046), blst_p2_affine_in_g2 (:1051),
// blst_miller_loop (:1060/:1065), blst_fp12_mul (:1061); per call blst_final_exp (:1070) and
// blst_fp12_is_one (:1076).  Only the boolean leaves the function, so the Miller-loop
// variant and the final-exponent multiple (here 3*(p^12-1)/r) are free choices
// (SURVEY.md Appendix D-8); a pair with an infinite member contributes 1 (Appendix D-2).
//
// Tower: Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - (1+u)), Fp12 = Fp6[w]/(w^2 - v).
// Lines are sparse elements l0 + l1*w^2 + l4*w^3 ("014").
#pragma once
#include "codec.cuh"

namespace b200 {

struct Fp6 { Fp2 c0, c1, c2; };
struct Fp12 { Fp6 c0, c1; };

// out-of-line Fp2 kernels keep the (very large) pairing code compact
// Fp2 mul / sqr / mul_fp and Fp inv are out-of-line on the device (fp.cuh), which keeps the
// (very large) pairing code compact
B200_HD Fp2 mulo(const Fp2& a, const Fp2& b) { return mul(a, b); }
B200_HD Fp2 sqro(const Fp2& a) { return sqr(a); }
B200_HD Fp2 mulfpo(const Fp2& a, const Fp& k) { return mul_fp(a, k); }
B200_HD Fp2 fp2_inv_o(const Fp2& a) { return inv(a); }

// ---------------------------------------------------------------- Fp6
B200_HD void fp6_add(Fp6& r, const Fp6& a, const Fp6& b) { r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); r.c2 = add(a.c2, b.c2); }
B200_HD void fp6_sub(Fp6& r, const Fp6& a, const Fp6& b) { r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); r.c2 = sub(a.c2, b.c2); }
B200_HD void fp6_neg(Fp6& r, const Fp6& a) { r.c0 = neg(a.c0); r.c1 = neg(a.c1); r.c2 = neg(a.c2); }
B200_HD void fp6_mul_v(Fp6& r, const Fp6& a) {
  Fp2 t = mul_xi(a.c2);
  r.c2 = a.c1; r.c1 = a.c0; r.c0 = t;
}
B200_HD_NI void fp6_mul(Fp6& r, const Fp6& a, const Fp6& b) {
  Fp2 t0 = mulo(a.c0, b.c0), t1 = mulo(a.c1, b.c1), t2 = mulo(a.c2, b.c2);
  Fp2 c0 = add(mul_xi(sub(sub(mul_sum2(a.c1, a.c2, b.c1, b.c2), t1), t2)), t0);
  Fp2 c1 = add(sub(sub(mul_sum2(a.c0, a.c1, b.c0, b.c1), t0), t1), mul_xi(t2));
  Fp2 c2 = add(sub(sub(mul_sum2(a.c0, a.c2, b.c0, b.c2), t0), t2), t1);
  r.c0 = c0; r.c1 = c1; r.c2 = c2;
}
// a * (b0 + b1 v)
B200_HD_NI void fp6_mul_by_01(Fp6& r, const Fp6& a, const Fp2& b0, const Fp2& b1) {
  Fp2 aa = mulo(a.c0, b0), bb = mulo(a.c1, b1);
  Fp2 c0 = add(mul_xi(sub(mul_sum1(a.c1, a.c2, b1), bb)), aa);
  Fp2 c1 = sub(sub(mul_sum2(b0, b1, a.c0, a.c1), aa), bb);
  Fp2 c2 = add(sub(mul_sum1(a.c0, a.c2, b0), aa), bb);
  r.c0 = c0; r.c1 = c1; r.c2 = c2;
}
// a * (b1 v)
B200_HD_NI void fp6_mul_by_1(Fp6& r, const Fp6& a, const Fp2& b1) {
  Fp2 c0 = mul_xi(mulo(a.c2, b1)), c1 = mulo(a.c0, b1), c2 = mulo(a.c1, b1);
  r.c0 = c0; r.c1 = c1; r.c2 = c2;
}
B200_HD_NI void fp6_inv(Fp6& r, const Fp6& a) {
  Fp2 t0 = sub(sqro(a.c0), mul_xi(mulo(a.c1, a.c2)));
  Fp2 t1 = sub(mul_xi(sqro(a.c2)), mulo(a.c0, a.c1));
  Fp2 t2 = sub(sqro(a.c1), mulo(a.c0, a.c2));
  Fp2 d = add(mul_xi(add(mulo(a.c2, t1), mulo(a.c1, t2))), mulo(a.c0, t0));
  d = fp2_inv_o(d);
  r.c0 = mulo(t0, d); r.c1 = mulo(t1, d); r.c2 = mulo(t2, d);
}

// ---------------------------------------------------------------- Fp12
B200_HD void fp12_set_one(Fp12& r) {
  r.c0.c0 = fp2_one(); r.c0.c1 = fp2_zero(); r.c0.c2 = fp2_zero();
  r.c1.c0 = fp2_zero(); r.c1.c1 = fp2_zero(); r.c1.c2 = fp2_zero();
}
B200_HD bool fp12_is_one(const Fp12& a) {
  return eq(a.c0.c0, fp2_one()) && is_zero(a.c0.c1) && is_zero(a.c0.c2) && is_zero(a.c1.c0) && is_zero(a.c1.c1) && is_zero(a.c1.c2);
}
B200_HD_NI void fp12_mul(Fp12& r, const Fp12& a, const Fp12& b) {
  Fp6 t0, t1, s, u, c1;
  fp6_mul(t0, a.c0, b.c0);
  fp6_mul(t1, a.c1, b.c1);
  fp6_add(s, a.c0, a.c1); fp6_add(u, b.c0, b.c1);
  fp6_mul(c1, s, u);
  fp6_sub(c1, c1, t0); fp6_sub(c1, c1, t1);
  fp6_mul_v(t1, t1);
  fp6_add(r.c0, t0, t1);
  r.c1 = c1;
}
// complex squaring with two Fp6 temporaries (thread-local memory per thread is what limits the pairing
// kernels: the resident threads' frames must fit in L2): ab = c0*c1, s = (c0+c1)(c0+v*c1);
// c0' = s - ab - v*ab, c1' = 2ab
B200_HD_NI void fp12_sqr_inplace(Fp12& f) {
  Fp6 ab, s;
  fp6_mul(ab, f.c0, f.c1);
  fp6_add(s, f.c0, f.c1);
  f.c0.c0 = add(f.c0.c0, mul_xi(f.c1.c2)); f.c0.c1 = add(f.c0.c1, f.c1.c0); f.c0.c2 = add(f.c0.c2, f.c1.c1);   // c0 + v*c1
  fp6_mul(s, s, f.c0);
  f.c0.c0 = sub(sub(s.c0, ab.c0), mul_xi(ab.c2));
  f.c0.c1 = sub(sub(s.c1, ab.c1), ab.c0);
  f.c0.c2 = sub(sub(s.c2, ab.c2), ab.c1);
  fp6_add(f.c1, ab, ab);
}
B200_HD void fp12_sqr(Fp12& r, const Fp12& a) {
  if (&r != &a) r = a;
  fp12_sqr_inplace(r);
}
B200_HD void fp12_conj(Fp12& r, const Fp12& a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }
B200_HD_NI void fp12_inv(Fp12& r, const Fp12& a) {
  Fp6 t0, t1;
  fp6_mul(t0, a.c0, a.c0); fp6_mul(t1, a.c1, a.c1); fp6_mul_v(t1, t1); fp6_sub(t0, t0, t1);
  fp6_inv(t0, t0);
  fp6_mul(r.c0, a.c0, t0);
  fp6_mul(t1, a.c1, t0); fp6_neg(r.c1, t1);
}
// f *= l0 + l1 w^2 + l4 w^3   (13 Fp2 products)
B200_HD_NI void fp12_mul_by_014(Fp12& f, const Fp2& l0, const Fp2& l1, const Fp2& l4) {
  Fp6 s;                                    // one temporary: aa and bb are formed in place in f.c0 / f.c1
  fp6_add(s, f.c0, f.c1);
  fp6_mul_by_01(s, s, l0, add(l1, l4));     // (c0 + c1) * (l0, l1 + l4)
  fp6_mul_by_01(f.c0, f.c0, l0, l1);        // aa
  fp6_mul_by_1(f.c1, f.c1, l4);             // bb
  fp6_sub(s, s, f.c0); fp6_sub(s, s, f.c1);
  fp6_mul_v(f.c1, f.c1);
  fp6_add(f.c0, f.c0, f.c1);                // aa + v*bb
  f.c1 = s;
}
B200_HD Fp2& fp12_wcoef(Fp12& a, int i) {   // coefficient of w^i
  Fp6& h = (i & 1) ? a.c1 : a.c0;
  return (i >> 1) == 0 ? h.c0 : ((i >> 1) == 1 ? h.c1 : h.c2);
}
B200_HD_NI void fp12_frob(Fp12& r, const Fp12& a, int k) {   // k = 1 or 2
  Fp12 t = a;
  const uint32_t* tab = (k == 1) ? C_FROB1() : C_FROB2();
  for (int i = 0; i < 6; i++) {
    Fp2& c = fp12_wcoef(t, i);
    if (k == 1) c = conj(c);
    c = mulo(c, fp2_load_const(tab + 24 * i));
  }
  r = t;
}
// Granger-Scott squaring in the cyclotomic subgroup
B200_HD void fp4_sqr(Fp2& r0, Fp2& r1, const Fp2& a, const Fp2& b) {
  Fp2 t0 = sqro(a), t1 = sqro(b);
  r1 = sub(sub(sqr_sum(a, b), t0), t1);
  r0 = add(mul_xi(t1), t0);
}
B200_HD_NI void fp12_cyclotomic_sqr(Fp12& r, const Fp12& f) {
  Fp2 z0 = f.c0.c0, z4 = f.c0.c1, z3 = f.c0.c2, z2 = f.c1.c0, z1 = f.c1.c1, z5 = f.c1.c2;
  Fp2 t0, t1, t2, t3, u0, u1;
  fp4_sqr(t0, t1, z0, z1);
  z0 = add(dbl(sub(t0, z0)), t0);
  z1 = add(dbl(add(t1, z1)), t1);
  fp4_sqr(u0, u1, z2, z3);
  fp4_sqr(t2, t3, z4, z5);
  z4 = add(dbl(sub(u0, z4)), u0);
  z5 = add(dbl(add(u1, z5)), u1);
  t0 = mul_xi(t3);
  z2 = add(dbl(add(t0, z2)), t0);
  z3 = add(dbl(sub(t2, z3)), t2);
  r.c0.c0 = z0; r.c0.c1 = z4; r.c0.c2 = z3;
  r.c1.c0 = z2; r.c1.c1 = z1; r.c1.c2 = z5;
}

// ---------------------------------------------------------------- Miller loop
struct G2Proj { Fp2 x, y, z; };   // homogeneous projective point on the twist

B200_HD_NI void ml_dbl_step(G2Proj& t, Fp2& l0, Fp2& l1, Fp2& l4) {
  Fp2 a = half(mulo(t.x, t.y));
  Fp2 b = sqro(t.y), c = sqro(t.z);
  // e = 3b' * c with 3b' = 12(1+u): multiplication by xi, then 12x by doublings (no field multiplication)
  Fp2 c4 = dbl(dbl(mul_xi(c)));
  Fp2 e = add(dbl(c4), c4);
  Fp2 f = add(dbl(e), e);
  Fp2 g = half(add(b, f));
  Fp2 h = sub(sqr_sum(t.y, t.z), add(b, c));
  Fp2 j = sqro(t.x);
  Fp2 e2 = sqro(e);
  l0 = sub(e, b);
  t.x = mulo(a, sub(b, f));
  t.y = sub(sqro(g), add(dbl(e2), e2));
  t.z = mulo(b, h);
  l1 = add(dbl(j), j);
  l4 = neg(h);
}
B200_HD_NI void ml_add_step(G2Proj& t, const G2Affine& q, Fp2& l0, Fp2& l1, Fp2& l4) {
  Fp2 theta = sub(t.y, mulo(q.y, t.z));
  Fp2 lam = sub(t.x, mulo(q.x, t.z));
  Fp2 c = sqro(theta), d = sqro(lam);
  Fp2 e = mulo(lam, d), f = mulo(t.z, c), g = mulo(t.x, d);
  Fp2 h = sub(add(e, f), dbl(g));
  t.x = mulo(lam, h);
  t.y = sub(mulo(theta, sub(g, h)), mulo(e, t.y));
  t.z = mulo(t.z, e);
  l0 = sub(mulo(theta, q.x), mulo(lam, q.y));
  l1 = neg(theta);
  l4 = lam;
}
// f = conj( f_{|z|,Q}(P) ) ; infinity on either side -> 1
B200_HD_NI void miller_loop(Fp12& f, const G1Affine& p, const G2Affine& q) {
  fp12_set_one(f);
  if (is_inf(p) || is_inf(q)) return;
  G2Proj t;
  t.x = q.x; t.y = q.y; t.z = fp2_one();
  Fp2 l0, l1, l4;
  for (int i = 62; i >= 0; i--) {
    if (i != 62) fp12_sqr(f, f);
    ml_dbl_step(t, l0, l1, l4);
    fp12_mul_by_014(f, l0, mulfpo(l1, p.x), mulfpo(l4, p.y));
    if ((B200_Z_ABS >> i) & 1) {
      ml_add_step(t, q, l0, l1, l4);
      fp12_mul_by_014(f, l0, mulfpo(l1, p.x), mulfpo(l4, p.y));
    }
  }
  fp12_conj(f, f);
}

// ---------------------------------------------------------------- final exponentiation
B200_HD_NI void cyc_exp_z(Fp12& r, const Fp12& a) {   // a^z, z < 0, a in the cyclotomic subgroup
  Fp12 acc = a;
  for (int i = 62; i >= 0; i--) {
    fp12_cyclotomic_sqr(acc, acc);
    if ((B200_Z_ABS >> i) & 1) fp12_mul(acc, acc, a);
  }
  fp12_conj(r, acc);
}
// f^((p^6-1)(p^2+1)) then ^((z-1)^2 (z+p)(z^2+p^2-1)) * ^3  =  f^(3 (p^12-1)/r)
B200_HD_NI void final_exp(Fp12& r, const Fp12& fin) {
  Fp12 f, t0, t1, t2, u;
  fp12_inv(t0, fin); fp12_conj(f, fin); fp12_mul(f, f, t0);
  fp12_frob(t0, f, 2); fp12_mul(f, t0, f);
  cyc_exp_z(t0, f); fp12_conj(u, f); fp12_mul(t0, t0, u);
  cyc_exp_z(t1, t0); fp12_conj(u, t0); fp12_mul(t0, t1, u);
  cyc_exp_z(t1, t0); fp12_frob(u, t0, 1); fp12_mul(t1, t1, u);
  cyc_exp_z(t2, t1); cyc_exp_z(u, t2);
  fp12_frob(t2, t1, 2); fp12_mul(t2, u, t2);
  fp12_conj(u, t1); fp12_mul(t2, t2, u);
  fp12_cyclotomic_sqr(u, f); fp12_mul(u, u, f);
  fp12_mul(r, t2, u);
}

// ---------------------------------------------------------------- subgroup membership (Scott 2021)
// G1: phi(P) = (beta x, y) == [-z^2]P ;  G2: psi(Q) == [z]Q.  Infinity is a member.
B200_HD_NI bool g1_in_subgroup(const G1Affine& p) {
  if (is_inf(p)) return true;
  const uint32_t zsq[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};   // z^2, little-endian words
  Jac<Fp> q = jac_scalar_mul(p, zsq, 128);
  if (is_zero(q.z)) return false;
  // phi(P) == -q  without leaving Jacobian coordinates:  beta*x*Z^2 == X  and  y*Z^3 == -Y
  Fp z2 = sqr(q.z);
  Fp bx = mul(p.x, fp_load_const(C_BETA()));
  return eq(mul(bx, z2), q.x) && eq(mul(p.y, mul(z2, q.z)), neg(q.y));
}
B200_HD_NI bool g2_in_subgroup(const G2Affine& p) {
  if (is_inf(p)) return true;
  const uint32_t zabs[2] = {(uint32_t)(B200_Z_ABS & 0xffffffffu), (uint32_t)(B200_Z_ABS >> 32)};
  Jac<Fp2> q = jac_scalar_mul(p, zabs, 64);
  if (is_zero(q.z)) return false;
  Fp2 z2 = sqro(q.z);
  Fp2 px = mulo(conj(p.x), fp2_load_const(C_PSI_CX()));
  Fp2 py = mulo(conj(p.y), fp2_load_const(C_PSI_CY()));
  return eq(mulo(px, z2), q.x) && eq(mulo(py, mulo(z2, q.z)), neg(q.y));   // [z]Q = -[|z|]Q
}

static constexpr int PAIRING_CHUNK = 3;       // default pairs per chunk (the engine adapts it to the batch)
static constexpr int ML_STEPS = 68;          // 63 doubling steps + 5 addition steps (|z| has weight 6)

struct Line { Fp2 l0, l1, l4; };             // l0 + l1*w^2 + l4*w^3, already multiplied by xP / yP
struct PairingTask { uint32_t first_pair, npairs, slot; };   // slot: index of the chunk's Miller value in fchunk

// Planning state on the device (one instance per batch, zeroed before k_pairing_count):
//   tasks_for[c]  number of chunks the batch has when cut into chunks of <= c pairs, c = 1..max_chunk
//   len_count[c][l]  how many of those chunks hold exactly l pairs
//   chunk  the chunk size k_pairing_plan settles on;  ntasks = tasks_for[chunk];  nslots, cursor[l]: allocation cursors
//   slot_base[k]  (dot engine) tasks are laid out by decreasing length, so the tasks that have a k-th pair are a prefix;
//                 the k-th pair of task t owns pair slot slot_base[k] + t; npair_slots = pairs of all valid calls
constexpr int PAIRING_MAX_CHUNK = 16;       // dot-engine kernels (three lanes per chunk); the thread-per-chunk kernel uses <= 6
constexpr int PAIRING_MAX_CHUNK_THREAD = 6;
struct PairingPlanState {
  uint32_t tasks_for[PAIRING_MAX_CHUNK + 1];
  uint32_t len_count[PAIRING_MAX_CHUNK + 1][PAIRING_MAX_CHUNK + 1];
  uint32_t cursor[PAIRING_MAX_CHUNK + 1];
  uint32_t chunk, ntasks, nslots, npair_slots;
  uint32_t slot_base[PAIRING_MAX_CHUNK + 1];
};

// Chunk size rule (measured on B200, profiles/r01_bench.md, profiles/r02_pairing.md): the accumulate kernel holds
// `wave` chunks at once and is latency-bound below that, so the best cut is the one whose task count just fits
// `wave` -- more tasks start another round, fewer leave each chunk a longer serial chain.  Smallest c with
// tasks <= wave (max_chunk when the batch is many waves long: least total work), then larger chunks while the
// batch still fills extend_pct % of it (they share more squarings).  The thread-per-chunk kernel passes its
// resident thread count and 85; the dot-engine kernel passes 1.8 x its resident chunk slots and 97 (blocks are
// scheduled longest-first, so ~1.8 rounds of 32-chunk blocks balance the SMs best: chunk 8 at the BASELINE mix).
B200_HD uint32_t pairing_choose_chunk(const uint32_t* tasks_for, uint32_t wave, uint32_t forced, uint32_t max_chunk,
                                      uint32_t extend_pct = 85) {
  if (forced >= 1 && forced <= max_chunk) return forced;
  uint32_t c = max_chunk;
  for (uint32_t t = 1; t <= max_chunk; t++)
    if (tasks_for[t] <= wave) { c = t; break; }
  while (c < max_chunk && (uint64_t)tasks_for[c + 1] * 100 >= (uint64_t)wave * extend_pct) c++;
  return c;
}

B200_HD bool in_subgroup(const G1Affine& p) { return g1_in_subgroup(p); }
B200_HD bool in_subgroup(const G2Affine& p) { return g2_in_subgroup(p); }

#ifdef __CUDACC__
// ---------------------------------------------------------------- kernels
// Batched point validation (K3): one thread per encoded point (stride = `stride_words` 32-bit words, so
// the same kernel walks bare point arrays and the point fields of MULTIEXP / PAIRING inputs).
// codes[i] = 0 ok, 3 invalid field element, 1 not on curve, 2 not in the r-torsion subgroup
// (only when check_subgroup != 0: the check the reference applies in PAIRING, eip2537.c:1041/:1051,
// and leaves as a TODO for MULTIEXP, :340/:401).
template <class F>
__global__ void __launch_bounds__(64) k_points_check(const uint32_t* __restrict__ raw, size_t n, int stride_words,
                                                      int check_subgroup, int* __restrict__ codes) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int PW = Wire<F>::POINT_WORDS;
  uint32_t w[PW];
  const uint4* src = reinterpret_cast<const uint4*>(raw + i * (size_t)stride_words);
#pragma unroll
  for (int k = 0; k < PW / 4; k++) { uint4 v = __ldg(src + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
  Affine<F> pt;
  int code = decode_point(pt, w);
  if (code == E_SUCCESS && check_subgroup && !in_subgroup(pt)) code = E_NOT_IN_SUBGROUP;
  codes[i] = code;
}
// fold per-point codes into the MSM status key (first failing pair wins)
__global__ void k_codes_to_status(const int* __restrict__ codes, size_t n, size_t index_base, unsigned long long* status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (codes[i] != E_SUCCESS) atomicMin(status, ((unsigned long long)(index_base + i) << 8) | (unsigned)codes[i]);
}

// one thread per pair: decode + subgroup checks in the reference's order
// (G1 decode, G1 subgroup, G2 decode, G2 subgroup; eip2537.c:1036-1053); status[j] = first failing code
// CHECK_G2 = false: the G2 membership test is left to k_pairing_lines_slots, whose Miller-loop walk T = [|z|]Q IS the
// 64-bit ladder of that test (the dot-engine pipeline; saves ~1,170 of the 2,256 Fp-mul per pair of this stage).
template <bool CHECK_G2>
__global__ void __launch_bounds__(64, 6) k_pairing_decode(const uint32_t* __restrict__ raw, size_t total_pairs,
                                                       G1Affine* __restrict__ g1, G2Affine* __restrict__ g2, int* __restrict__ status) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total_pairs) return;
  const uint4* src = reinterpret_cast<const uint4*>(raw + j * 96);
  uint32_t w[64];
  int code;
  G1Affine p;
  G2Affine q;
#pragma unroll
  for (int k = 0; k < 8; k++) { uint4 v = __ldg(src + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
  code = decode_point(p, w);
  if (code == E_SUCCESS && !g1_in_subgroup(p)) code = E_NOT_IN_SUBGROUP;
  if (code == E_SUCCESS) {
#pragma unroll
    for (int k = 0; k < 16; k++) { uint4 v = __ldg(src + 8 + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    code = decode_point(q, w);
    if (CHECK_G2 && code == E_SUCCESS && !g2_in_subgroup(q)) code = E_NOT_IN_SUBGROUP;
  }
  status[j] = code;
  if (code == E_SUCCESS) { g1[j] = p; g2[j] = q; }
}

// Deferred G2 membership (dot-engine pipeline), two small per-call passes that keep the reference's precedence
// "first failing pair in input order, G1 checks before G2 checks inside a pair" (eip2537.c:1033-1053):
//  k_pairing_early_fix  a call that already failed (errs != 0 from k_pairing_count) may hold an EARLIER pair whose only
//                       fault is G2 membership: run the exact ladder on the pairs before the failing one (rare path)
//  k_pairing_late_errs  after the line kernel has written the deferred G2 results into status[]: first failing pair
__global__ void __launch_bounds__(64) k_pairing_early_fix(const unsigned long long* __restrict__ offsets, size_t n_calls,
                                                          const G2Affine* __restrict__ g2, int* __restrict__ status, int* __restrict__ errs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_calls || errs[i] == E_SUCCESS || errs[i] == E_INVALID_LENGTH) return;
  size_t first = (size_t)(offsets[i] / 384), last = (size_t)(offsets[i + 1] / 384);
  for (size_t j = first; j < last && status[j] == E_SUCCESS; j++) {
    G2Affine q = g2[j];
    if (!g2_in_subgroup(q)) { status[j] = E_NOT_IN_SUBGROUP; errs[i] = E_NOT_IN_SUBGROUP; return; }
  }
}
__global__ void __launch_bounds__(128) k_pairing_late_errs(const unsigned long long* __restrict__ offsets, size_t n_calls,
                                                           const int* __restrict__ status, int* __restrict__ errs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_calls || errs[i] != E_SUCCESS) return;
  size_t first = (size_t)(offsets[i] / 384), last = (size_t)(offsets[i + 1] / 384);
  for (size_t j = first; j < last; j++)
    if (status[j] != E_SUCCESS) { errs[i] = status[j]; return; }
}

// Latency variant for small batches: 32 pairs per block of two warps, warp 0 takes the G1 member of each pair, warp 1
// the G2 member -- the two ladders are independent, so a pair's checks run side by side instead of one after the
// other (single two-pair call 12.3 -> 11.4 ms).  Large batches keep the kernel above (1-4 % faster there: one warp
// role per block finishes early).  Code of the pair = G1's if it failed, else G2's: the same precedence.
__global__ void __launch_bounds__(64, 6) k_pairing_decode_split(const uint32_t* __restrict__ raw, size_t total_pairs,
                                                       G1Affine* __restrict__ g1, G2Affine* __restrict__ g2, int* __restrict__ status) {
  __shared__ int code_g1[32];
  const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t j = (size_t)blockIdx.x * 32 + lane;
  const bool live = j < total_pairs;
  int code = E_SUCCESS;
  if (live) {
    const uint4* src = reinterpret_cast<const uint4*>(raw + j * 96);
    if (role == 0) {
      uint32_t w[32];
      G1Affine p;
#pragma unroll
      for (int k = 0; k < 8; k++) { uint4 v = __ldg(src + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
      code = decode_point(p, w);
      if (code == E_SUCCESS && !g1_in_subgroup(p)) code = E_NOT_IN_SUBGROUP;
      if (code == E_SUCCESS) g1[j] = p;
      code_g1[lane] = code;
    } else {
      uint32_t w[64];
      G2Affine q;
#pragma unroll
      for (int k = 0; k < 16; k++) { uint4 v = __ldg(src + 8 + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
      code = decode_point(q, w);
      if (code == E_SUCCESS && !g2_in_subgroup(q)) code = E_NOT_IN_SUBGROUP;
      if (code == E_SUCCESS) g2[j] = q;
    }
  }
  __syncthreads();
  if (role == 1 && live) {
    const int c1 = code_g1[lane];
    status[j] = c1 != E_SUCCESS ? c1 : code;
  }
}

// ---- Miller loop, split in two kernels -------------------------------------------------------
// The reference runs k independent single-pair Miller loops per call and multiplies them
// (eip2537.c:1060-1065).  Only the boolean is observable, so here the pairs of a call share the
// Fp12 squarings (a multi-Miller loop), and the work is cut along its natural seam:
//   k_pairing_lines       one thread per PAIR: walks T = [.]Q on the twist and writes the 68 line
//                         functions, already evaluated at P (3 Fp2 = 288 B each), to HBM.  Small
//                         state (T, Q, P), no Fp12.
//   k_pairing_accumulate  one thread per CHUNK of <= PAIRING_CHUNK pairs of one call:
//                         f <- f^2 * prod(lines) step by step: one Fp12 squaring per step per chunk
//                         instead of per pair, 13 Fp2 products per line.
//   k_pairing_calls       one thread per call: product of its chunks, final exponentiation, is-one.

__global__ void __launch_bounds__(64, 6) k_pairing_lines(const G1Affine* __restrict__ g1, const G2Affine* __restrict__ g2,
                                                      const int* __restrict__ status, size_t total_pairs,
                                                      Line* __restrict__ lines, unsigned char* __restrict__ skip) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total_pairs) return;
  if (status[j] != E_SUCCESS) { skip[j] = 1; return; }
  G1Affine p = g1[j];
  G2Affine q = g2[j];
  if (is_inf(p) || is_inf(q)) { skip[j] = 1; return; }   // contributes 1
  skip[j] = 0;
  G2Proj t;
  t.x = q.x; t.y = q.y; t.z = fp2_one();
  int s = 0;
  for (int i = 62; i >= 0; i--) {
    Line ln;
    ml_dbl_step(t, ln.l0, ln.l1, ln.l4);
    ln.l1 = mulfpo(ln.l1, p.x); ln.l4 = mulfpo(ln.l4, p.y);
    lines[(size_t)s * total_pairs + j] = ln;
    s++;
    if ((B200_Z_ABS >> i) & 1) {
      ml_add_step(t, q, ln.l0, ln.l1, ln.l4);
      ln.l1 = mulfpo(ln.l1, p.x); ln.l4 = mulfpo(ln.l4, p.y);
      lines[(size_t)s * total_pairs + j] = ln;
      s++;
    }
  }
}

// one thread per call: decide the call's error code (first failing pair, eip2537.c:1033-1053) and count the
// chunks it would contribute for every candidate chunk size (block-local shared-memory histograms, one global
// atomic per non-empty counter per block)
__global__ void __launch_bounds__(128) k_pairing_count(const unsigned long long* __restrict__ offsets, size_t n_calls,
                                                       const int* __restrict__ status, PairingPlanState* st, int* __restrict__ errs,
                                                       uint32_t max_chunk) {
  __shared__ uint32_t h_tasks[PAIRING_MAX_CHUNK + 1];
  __shared__ uint32_t h_len[PAIRING_MAX_CHUNK + 1][PAIRING_MAX_CHUNK + 1];
  for (int x = threadIdx.x; x < (PAIRING_MAX_CHUNK + 1) * (PAIRING_MAX_CHUNK + 1); x += blockDim.x) (&h_len[0][0])[x] = 0;
  if (threadIdx.x <= PAIRING_MAX_CHUNK) h_tasks[threadIdx.x] = 0;
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_calls) {
    size_t first = (size_t)(offsets[i] / 384), last = (size_t)(offsets[i + 1] / 384);
    int code = first == last ? E_INVALID_LENGTH : E_SUCCESS;
    for (size_t j = first; j < last && code == E_SUCCESS; j++) code = status[j];
    errs[i] = code;
    if (code == E_SUCCESS) {
      uint32_t k = (uint32_t)(last - first);
      for (uint32_t c = 1; c <= max_chunk; c++) {
        // ceil(k/c) chunks of BALANCED length (13 pairs at c = 12 are 7 + 6, not 12 + 1): the longest chain bounds a wave
        uint32_t nch = (k + c - 1) / c, base = k / nch, rem = k % nch;
        atomicAdd(&h_tasks[c], nch);
        if (rem) atomicAdd(&h_len[c][base + 1], rem);
        atomicAdd(&h_len[c][base], nch - rem);
      }
    }
  }
  __syncthreads();
  for (int x = threadIdx.x; x < (PAIRING_MAX_CHUNK + 1) * (PAIRING_MAX_CHUNK + 1); x += blockDim.x) {
    uint32_t v = (&h_len[0][0])[x];
    if (v) atomicAdd(&st->len_count[0][0] + x, v);
  }
  if (threadIdx.x <= PAIRING_MAX_CHUNK && h_tasks[threadIdx.x]) atomicAdd(&st->tasks_for[threadIdx.x], h_tasks[threadIdx.x]);
}

// one thread per call: cut the call into chunks of the chosen size; tasks are laid out by DECREASING length so
// that the 32 lanes of a warp walk equally long chains (a warp runs as long as its longest lane).
// slot_pair (dot engine, may be null): slot_pair[slot_base[k] + t] = index of the k-th pair of task t.
__global__ void __launch_bounds__(128) k_pairing_plan(const unsigned long long* __restrict__ offsets, size_t n_calls,
                                                      const int* __restrict__ errs, uint32_t wave, uint32_t forced_chunk, uint32_t max_chunk,
                                                      uint32_t extend_pct, PairingPlanState* st, PairingTask* __restrict__ tasks,
                                                      uint32_t* __restrict__ call_first_task, uint32_t* __restrict__ slot_pair) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_calls) return;
  const uint32_t chunk = pairing_choose_chunk(st->tasks_for, wave, forced_chunk, max_chunk, extend_pct);
  uint32_t base_of[PAIRING_MAX_CHUNK + 1], slot_base[PAIRING_MAX_CHUNK + 1];
  uint32_t run = 0;
  for (int l = PAIRING_MAX_CHUNK; l >= 1; l--) { base_of[l] = run; run += st->len_count[chunk][l]; }
  // tasks holding a k-th pair (length > k): the first cnt_k of the length-sorted order
  // (each position's slot range starts at a multiple of 32, so a block's 32 lanes read 128-byte aligned rows)
  uint32_t slots = 0;
  for (int k = 0; k < PAIRING_MAX_CHUNK; k++) {
    slot_base[k] = slots;
    uint32_t cnt = 0;
    for (int l = k + 1; l <= PAIRING_MAX_CHUNK; l++) cnt += st->len_count[chunk][l];
    slots += (cnt + 31u) & ~31u;
  }
  if (i == 0) {
    st->chunk = chunk; st->ntasks = st->tasks_for[chunk]; st->npair_slots = slots;
    for (int k = 0; k < PAIRING_MAX_CHUNK; k++) st->slot_base[k] = slot_base[k];
  }
  if (errs[i] != E_SUCCESS) return;
  size_t first = (size_t)(offsets[i] / 384), last = (size_t)(offsets[i + 1] / 384);
  uint32_t k = (uint32_t)(last - first), nch = (k + chunk - 1) / chunk;
  uint32_t slot0 = atomicAdd(&st->nslots, nch);
  call_first_task[i] = slot0;
  const uint32_t blen = k / nch, rem = k % nch;      // the first `rem` chunks hold blen + 1 pairs (see k_pairing_count)
  for (uint32_t cidx = 0; cidx < nch; cidx++) {
    uint32_t lo = cidx * blen + (cidx < rem ? cidx : rem), len = blen + (cidx < rem ? 1 : 0);
    uint32_t pos = base_of[len] + atomicAdd(&st->cursor[len], 1);
    tasks[pos] = PairingTask{(uint32_t)first + lo, len, slot0 + cidx};
    if (slot_pair)
      for (uint32_t q = 0; q < len; q++) slot_pair[slot_base[q] + pos] = (uint32_t)first + lo + q;
  }
}

__global__ void __launch_bounds__(64, 6) k_pairing_accumulate(const PairingTask* __restrict__ tasks, const PairingPlanState* __restrict__ st,
                                                           const Line* __restrict__ lines, const unsigned char* __restrict__ skip,
                                                           size_t total_pairs, Fp12* __restrict__ fchunk) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= st->ntasks) return;
  PairingTask task = tasks[t];
  Fp12 f;
  fp12_set_one(f);
  bool started = false;
  int s = 0;
  for (int i = 62; i >= 0; i--) {
    if (started) fp12_sqr(f, f);
    const int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
    for (int a = 0; a < nsteps; a++, s++) {
      for (uint32_t k = 0; k < task.npairs; k++) {
        size_t j = task.first_pair + k;
        if (skip[j]) continue;
        const Line& ln = lines[(size_t)s * total_pairs + j];   // read in place: no thread-local copy
        fp12_mul_by_014(f, ln.l0, ln.l1, ln.l4);
        started = true;
      }
    }
  }
  fp12_conj(f, f);
  fchunk[task.slot] = f;
}

// one thread per call: product of the chunks' Miller values (blst_fp12_mul, :1061), one final
// exponentiation (:1070), is-one -> out[31] (:1076)
__global__ void __launch_bounds__(64) k_pairing_calls(size_t n_calls, const unsigned long long* __restrict__ offsets,
                                                      const PairingPlanState* __restrict__ st, const uint32_t* __restrict__ call_first_task, const Fp12* __restrict__ fchunk,
                                                      uint32_t* __restrict__ outs, const int* __restrict__ errs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_calls) return;
  uint32_t* out = outs + 8 * i;
  for (int k = 0; k < 8; k++) out[k] = 0;
  if (errs[i] != E_SUCCESS) return;
  const uint32_t chunk = st->chunk;
  uint32_t npairs = (uint32_t)(offsets[i + 1] / 384 - offsets[i] / 384), nch = (npairs + chunk - 1) / chunk;
  uint32_t base = call_first_task[i];
  Fp12 acc = fchunk[base], cur;
  for (uint32_t c = 1; c < nch; c++) {
    cur = fchunk[base + c];
    fp12_mul(acc, acc, cur);
  }
  final_exp(acc, acc);
  if (fp12_is_one(acc)) out[7] = 0x01000000u;   // byte 31 of the 32-byte big-endian word
}
#endif  // __CUDACC__

}  // namespace b200
