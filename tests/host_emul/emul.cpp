// tests/host_emul/emul.cpp -- TEST ONLY.  Compiles the product's device functions (fp.cuh,
// ec.cuh, codec.cuh, pairing.cuh: all B200_HD) for the HOST with g++, so their algorithms can
// be checked against the oracle without a GPU.  The shipped library never contains this path
// (engine.cu launches kernels only); on the host the portable 64-bit Fp multiply stands in for
// the PTX carry-chain one, everything above it is the same source the GPU runs.
#include <cstring>
#define B200_COUNT_MULS 1
namespace b200 { unsigned long long g_fp_mul_count = 0; }
#include "../../blst_eip2537_b200/csrc/msm.cuh"
#include "../../blst_eip2537_b200/csrc/pairing.cuh"
#include "../../blst_eip2537_b200/csrc/coop12.cuh"
#include "../../blst_eip2537_b200/csrc/dot12.cuh"
#include "../../blst_eip2537_b200/csrc/map.cuh"

using namespace b200;

static void load_words(uint32_t* w, const unsigned char* in, int nwords) { memcpy(w, in, 4 * nwords); }

template <class F> struct Node { XYZZ<F> s, w; };
extern "C" {
// r = a*b (canonical 48-byte big-endian in/out, via Montgomery form)
void emul_fp_mul(unsigned char* out64, const unsigned char* a64, const unsigned char* b64) {
  uint32_t w[16];
  Fp a, b;
  load_words(w, a64, 16); fp_from_slot(a, w);
  load_words(w, b64, 16); fp_from_slot(b, w);
  Fp r = mul(a, b);
  fp_to_slot(w, r);
  memcpy(out64, w, 64);
}
void emul_fp_inv(unsigned char* out64, const unsigned char* a64) {
  uint32_t w[16];
  Fp a;
  load_words(w, a64, 16); fp_from_slot(a, w);
  Fp r = inv(a);
  fp_to_slot(w, r);
  memcpy(out64, w, 64);
}
}  // extern C
// straightforward MSM through the product's point code: sum k_i * P_i with per-point double-and-add
template <class F>
static int msm_simple(unsigned char* out, const unsigned char* in, size_t n) {
  constexpr int PW = Wire<F>::POINT_WORDS, SW = Wire<F>::PAIR_WORDS;
  XYZZ<F> acc = xyzz_inf<F>();
  for (size_t i = 0; i < n; i++) {
    uint32_t w[SW];
    load_words(w, in + 4 * SW * i, SW);
    Affine<F> p;
    int code = decode_point(p, w);
    if (code) return code;
    uint32_t k[8];
    scalar_from_slot(k, w + PW);
    XYZZ<F> t = xyzz_scalar_mul(p, k, 256);
    xyzz_add(acc, t);
  }
  Affine<F> a = xyzz_to_affine(acc);
  uint32_t w[PW];
  encode_point(w, a);
  memcpy(out, w, 4 * PW);
  return 0;
}
extern "C" {
int emul_g1_msm(unsigned char* out, const unsigned char* in, size_t n) { return msm_simple<Fp>(out, in, n); }
int emul_g2_msm(unsigned char* out, const unsigned char* in, size_t n) { return msm_simple<Fp2>(out, in, n); }

}  // extern C
// host re-enactment of the Pippenger pipeline (same digit recoding, bucket layout, reduction tree
// shapes and window combine as engine.cu's msm_pipeline), single-threaded
template <class F>
static int msm_pippenger(unsigned char* out, const unsigned char* in, size_t n, int c) {
  constexpr int PW = Wire<F>::POINT_WORDS, SW = Wire<F>::PAIR_WORDS;
  MsmPlan plan = make_plan(c);
  size_t nbt = (size_t)plan.nwin * plan.nb;
  XYZZ<F>* buckets = new XYZZ<F>[nbt];
  for (size_t b = 0; b < nbt; b++) buckets[b] = xyzz_inf<F>();
  for (size_t i = 0; i < n; i++) {
    uint32_t w[SW];
    load_words(w, in + 4 * SW * i, SW);
    Affine<F> p;
    int code = decode_point(p, w);
    if (code) { delete[] buckets; return code; }
    if (is_inf(p)) continue;
    uint32_t k[8];
    scalar_from_slot(k, w + PW);
    uint32_t carry = 0;
    for (int win = 0; win < plan.nwin; win++) {
      int d;
      auto bits = [&](int bit, int cc) {
        int word = bit >> 5, sh = bit & 31;
        uint32_t v = k[word] >> sh;
        if (sh + cc > 32 && word + 1 < 8) v |= k[word + 1] << (32 - sh);
        return v & ((1u << cc) - 1);
      };
      const int wd = plan.width[win];
      uint32_t raw_d = bits(plan.bitpos[win], wd) + carry;
      if (win < plan.nwin - 1) {
        if (raw_d > (1u << (wd - 1))) { d = (int)raw_d - (int)(1u << wd); carry = 1; } else { d = (int)raw_d; carry = 0; }
      } else d = (int)raw_d;
      if (!d) continue;
      uint32_t mag = d < 0 ? -d : d;
      Affine<F> q = p;
      if (d < 0) q.y = neg(q.y);
      xyzz_madd(buckets[(size_t)win * plan.nb + mag - 1], q);
    }
  }
  // reduction tree
  int L0_log = plan.log_nb < 4 ? plan.log_nb : 4;
  size_t npw = plan.nb >> L0_log;
  Node<F>* cur = new Node<F>[plan.nwin * npw];
  for (size_t t = 0; t < plan.nwin * npw; t++) {
    int L = 1 << L0_log;
    XYZZ<F> run = xyzz_inf<F>(), acc = xyzz_inf<F>();
    for (int j = L - 1; j >= 0; j--) { xyzz_add(run, buckets[t * L + j]); xyzz_add(acc, run); }
    cur[t].s = run; cur[t].w = acc;
  }
  // upper levels as k_reduce_level / k_window_finish do them: vectors of plain sums, scaled once at the root
  const ReduceLevels lv = make_reduce_levels(plan.log_nb, L0_log, 1 + (plan.c & 1));   // both level shapes get exercised
  int K = 2;
  XYZZ<F>* vec = new XYZZ<F>[plan.nwin * npw * 2];
  for (size_t t = 0; t < plan.nwin * npw; t++) { vec[2 * t] = cur[t].s; vec[2 * t + 1] = cur[t].w; }
  for (int i = 0; i < lv.n; i++) {
    const int L = 1 << lv.l_log[i], K_out = K + L - 1;
    const size_t opw = npw >> lv.l_log[i];
    XYZZ<F>* nxt = new XYZZ<F>[plan.nwin * opw * K_out];
    for (size_t t = 0; t < plan.nwin * opw; t++)
      for (int k = 0; k < K_out; k++) {
        const XYZZ<F>* ch = vec + t * L * K;
        XYZZ<F> a;
        if (k < K) { a = ch[k]; for (int j = 1; j < L; j++) xyzz_add(a, ch[(size_t)j * K + k]); }
        else a = ch[(size_t)(k - K + 1) * K];
        nxt[t * K_out + k] = a;
      }
    delete[] vec; vec = nxt; npw = opw; K = K_out;
  }
  if (K != reduce_root_width(lv) || npw != 1) { delete[] vec; delete[] cur; delete[] buckets; return -100; }
  Hom<F>* tw = new Hom<F>[plan.nwin];
  for (int win = 0; win < plan.nwin; win++) {
    const XYZZ<F>* r = vec + (size_t)win * K;
    Hom<F> tacc = hom_inf<F>();
    int base = 2;
    XYZZ<F> U[REDUCE_MAX_LEVELS];
    for (int i = 0; i < lv.n; i++) {
      const int L = 1 << lv.l_log[i];
      XYZZ<F> run = xyzz_inf<F>(), u = xyzz_inf<F>();
      for (int t = L - 1; t >= 1; t--) { xyzz_add(run, r[base + t - 1]); xyzz_add(u, run); }
      U[i] = u; base += L - 1;
    }
    for (int i = lv.n - 1; i >= 0; i--) {
      tacc = hom_add(tacc, xyzz_to_hom(U[i]));
      const int nd = lv.cov[i] - (i > 0 ? lv.cov[i - 1] : 0);
      for (int q = 0; q < nd; q++) tacc = hom_dbl(tacc);
    }
    tw[win] = hom_add(tacc, xyzz_to_hom(r[1]));
  }
  delete[] vec;
  Hom<F> hacc = hom_inf<F>();
  for (int win = plan.nwin - 1; win >= 0; win--) {
    for (int q = 0; q < plan.width[win]; q++) hacc = hom_dbl(hacc);
    hacc = hom_add(hacc, tw[win]);
  }
  delete[] tw;
  XYZZ<F> acc = hom_to_xyzz(hacc);
  Affine<F> a = xyzz_to_affine(acc);
  uint32_t w[PW];
  encode_point(w, a);
  memcpy(out, w, 4 * PW);
  delete[] cur; delete[] buckets;
  return 0;
}
// Homogeneous complete formulas (ec.cuh Hom) against the XYZZ ones: P + Q, 2P, and every case the XYZZ code branches on.
// in = two encoded points (curve points, any subgroup); returns a bit mask of mismatches.
template <class F>
static int hom_check(const unsigned char* in) {
  const int PW = Wire<F>::POINT_WORDS;
  uint32_t w[2 * 64];
  load_words(w, in, 2 * PW);
  Affine<F> p, q;
  if (decode_point(p, w) || decode_point(q, w + PW)) return -1;
  auto aff_eq = [](const Affine<F>& a, const Affine<F>& b) { return eq(a.x, b.x) && eq(a.y, b.y); };
  auto H = [](const Affine<F>& a) { return xyzz_to_hom(xyzz_from_affine(a)); };
  auto A = [](const Hom<F>& h) { return xyzz_to_affine(hom_to_xyzz(h)); };
  int bad = 0;
  XYZZ<F> s = xyzz_from_affine(p);
  xyzz_add(s, xyzz_from_affine(q));
  if (!aff_eq(xyzz_to_affine(s), A(hom_add(H(p), H(q))))) bad |= 1;
  XYZZ<F> d = xyzz_from_affine(p);
  xyzz_add(d, xyzz_from_affine(p));
  if (!aff_eq(xyzz_to_affine(d), A(hom_dbl(H(p))))) bad |= 2;
  if (!aff_eq(xyzz_to_affine(d), A(hom_add(H(p), H(p))))) bad |= 4;                       // addition law on equal points
  const Affine<F> inf = xyzz_to_affine(xyzz_inf<F>());
  if (!aff_eq(inf, A(hom_add(H(p), H(affine_neg(p)))))) bad |= 8;                          // opposite points
  if (!aff_eq(p, A(hom_add(hom_inf<F>(), H(p)))) || !aff_eq(p, A(hom_add(H(p), hom_inf<F>())))) bad |= 16;
  if (!aff_eq(inf, A(hom_add(hom_inf<F>(), hom_inf<F>()))) || !aff_eq(inf, A(hom_dbl(hom_inf<F>())))) bad |= 32;
  // a non-trivial projective representative: 2P + Q with 2P kept projective on both sides
  XYZZ<F> e = d;
  xyzz_add(e, xyzz_from_affine(q));
  if (!aff_eq(xyzz_to_affine(e), A(hom_add(xyzz_to_hom(d), H(q))))) bad |= 64;
  if (!aff_eq(xyzz_to_affine(xyzz_dbl(d)), A(hom_dbl(hom_dbl(H(p)))))) bad |= 128;
  // affine + affine with the inverse handed in (ec.cuh pair_prepare / pair_finish), every case k_pair_round can meet
  auto P = [](const Affine<F>& a, const Affine<F>& b) { F den; const int kind = pair_prepare(a, b, den); return pair_finish(kind, a, b, inv(den)); };
  if (!aff_eq(xyzz_to_affine(s), P(p, q))) bad |= 256;
  if (!aff_eq(xyzz_to_affine(d), P(p, p))) bad |= 512;
  if (!aff_eq(inf, P(p, affine_neg(p)))) bad |= 1024;
  if (!aff_eq(p, P(inf, p)) || !aff_eq(p, P(p, inf)) || !aff_eq(inf, P(inf, inf))) bad |= 2048;
  return bad;
}
extern "C" {
int emul_hom_check(int group, const unsigned char* in) { return group == 1 ? hom_check<Fp>(in) : hom_check<Fp2>(in); }
int emul_g1_pippenger(unsigned char* out, const unsigned char* in, size_t n, int c) { return msm_pippenger<Fp>(out, in, n, c); }
int emul_g2_pippenger(unsigned char* out, const unsigned char* in, size_t n, int c) { return msm_pippenger<Fp2>(out, in, n, c); }

// pairing call exactly as k_pairing_decode + k_pairing_calls do it; also dumps the GT element
int emul_pairing(unsigned char* out32, unsigned char* gt576, const unsigned char* in, size_t k) {
  Fp12 acc, cur;
  for (size_t j = 0; j < k; j++) {
    uint32_t w[96];
    load_words(w, in + 384 * j, 96);
    G1Affine p; G2Affine q;
    int code = decode_point(p, w);
    if (code == 0 && !g1_in_subgroup(p)) code = E_NOT_IN_SUBGROUP;
    if (code) return code;
    code = decode_point(q, w + 32);
    if (code == 0 && !g2_in_subgroup(q)) code = E_NOT_IN_SUBGROUP;
    if (code) return code;
    if (j == 0) miller_loop(acc, p, q);
    else { miller_loop(cur, p, q); fp12_mul(acc, acc, cur); }
  }
  final_exp(acc, acc);
  memset(out32, 0, 32);
  if (fp12_is_one(acc)) out32[31] = 1;
  if (gt576) {
    const Fp* e = reinterpret_cast<const Fp*>(&acc);
    for (int i = 0; i < 12; i++) { uint32_t w[16]; fp_to_slot(w, e[i]); memcpy(gt576 + 48 * i, (unsigned char*)w + 16, 48); }
  }
  return 0;
}
// Fp-multiplication counts of the GPU's own pairing formulas, phase by phase, for one call of k pairs
// (same source as k_pairing_decode / k_pairing_lines / k_pairing_accumulate / k_pairing_calls):
// counts[0] decode + subgroup checks, [1] line functions, [2] chunked accumulate, [3] product + final exp
int emul_pairing_fme_counts(unsigned long long* counts, unsigned char* out32, const unsigned char* in, size_t k) {
  G1Affine* ps = new G1Affine[k];
  G2Affine* qs = new G2Affine[k];
  unsigned long long c0 = g_fp_mul_count;
  for (size_t j = 0; j < k; j++) {
    uint32_t w[96];
    load_words(w, in + 384 * j, 96);
    int code = decode_point(ps[j], w);
    if (code == 0 && !g1_in_subgroup(ps[j])) code = E_NOT_IN_SUBGROUP;
    if (code == 0) code = decode_point(qs[j], w + 32);
    if (code == 0 && !g2_in_subgroup(qs[j])) code = E_NOT_IN_SUBGROUP;
    if (code) { delete[] ps; delete[] qs; return code; }
  }
  counts[0] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  Line* lines = new Line[ML_STEPS * k];
  bool* skip = new bool[k];
  for (size_t j = 0; j < k; j++) {
    skip[j] = is_inf(ps[j]) || is_inf(qs[j]);
    if (skip[j]) continue;
    G2Proj t; t.x = qs[j].x; t.y = qs[j].y; t.z = fp2_one();
    int s = 0;
    for (int i = 62; i >= 0; i--) {
      Line ln;
      ml_dbl_step(t, ln.l0, ln.l1, ln.l4);
      ln.l1 = mulfpo(ln.l1, ps[j].x); ln.l4 = mulfpo(ln.l4, ps[j].y);
      lines[s++ * k + j] = ln;
      if ((B200_Z_ABS >> i) & 1) {
        ml_add_step(t, qs[j], ln.l0, ln.l1, ln.l4);
        ln.l1 = mulfpo(ln.l1, ps[j].x); ln.l4 = mulfpo(ln.l4, ps[j].y);
        lines[s++ * k + j] = ln;
      }
    }
  }
  counts[1] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  size_t nch = (k + PAIRING_CHUNK - 1) / PAIRING_CHUNK;
  Fp12* fch = new Fp12[nch];
  for (size_t c = 0; c < nch; c++) {
    size_t first = c * PAIRING_CHUNK, np = k - first < (size_t)PAIRING_CHUNK ? k - first : PAIRING_CHUNK;
    Fp12 f; fp12_set_one(f);
    bool started = false;
    int s = 0;
    for (int i = 62; i >= 0; i--) {
      if (started) fp12_sqr(f, f);
      int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
      for (int a = 0; a < nsteps; a++, s++)
        for (size_t q = 0; q < np; q++) {
          size_t j = first + q;
          if (skip[j]) continue;
          const Line& ln = lines[s * k + j];
          fp12_mul_by_014(f, ln.l0, ln.l1, ln.l4);
          started = true;
        }
    }
    fp12_conj(f, f);
    fch[c] = f;
  }
  counts[2] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  Fp12 acc = fch[0];
  for (size_t c = 1; c < nch; c++) fp12_mul(acc, acc, fch[c]);
  final_exp(acc, acc);
  counts[3] = g_fp_mul_count - c0;
  memset(out32, 0, 32);
  if (fp12_is_one(acc)) out32[31] = 1;
  delete[] ps; delete[] qs; delete[] lines; delete[] skip; delete[] fch;
  return 0;
}
// Fp multiplications of one XYZZ mixed addition / full addition / doubling over Fp and Fp2
void emul_point_op_fme(unsigned long long* c6, const unsigned char* g1_128, const unsigned char* g2_256) {
  uint32_t w[64];
  load_words(w, g1_128, 32); G1Affine p; decode_point(p, w);
  load_words(w, g2_256, 64); G2Affine q; decode_point(q, w);
  G1XYZZ a = xyzz_dbl_affine(p); G2XYZZ b = xyzz_dbl_affine(q);
  unsigned long long c0 = g_fp_mul_count;
  xyzz_madd(a, p); c6[0] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  { G1XYZZ t = xyzz_dbl(a); xyzz_add(a, t); } c6[1] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  a = xyzz_dbl(a); c6[2] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  xyzz_madd(b, q); c6[3] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  { G2XYZZ t = xyzz_dbl(b); xyzz_add(b, t); } c6[4] = g_fp_mul_count - c0; c0 = g_fp_mul_count;
  b = xyzz_dbl(b); c6[5] = g_fp_mul_count - c0;
}
// coop12.cuh operation tables, run sequentially, against the thread-level Fp12 functions.
// in: 2 x 12 field slots (64 bytes each): f then g.  Returns a bit mask of mismatching operations.
int emul_coop12_check(const unsigned char* in) {
  Fp12 f, g;
  Fp* fe = reinterpret_cast<Fp*>(&f);
  Fp* ge = reinterpret_cast<Fp*>(&g);
  for (int i = 0; i < 12; i++) {
    uint32_t w[16];
    load_words(w, in + 64 * i, 16); if (fp_from_slot(fe[i], w) < 0) return -1;
    load_words(w, in + 64 * (12 + i), 16); if (fp_from_slot(ge[i], w) < 0) return -1;
  }
  auto same = [](const Fp12& a, const Fp12& b) { return memcmp(&a, &b, sizeof(Fp12)) == 0; };
  int bad = 0;
  Fp12 want, got;
  fp12_sqr(want, f); got = f; coop12_exec_seq<OpSqr>(reinterpret_cast<Fp2*>(&got), nullptr); if (!same(want, got)) bad |= 1;
  fp12_mul(want, f, g); got = f; coop12_exec_seq<OpMul>(reinterpret_cast<Fp2*>(&got), reinterpret_cast<const Fp2*>(&g)); if (!same(want, got)) bad |= 2;
  Fp2 line[3] = {g.c0.c0, g.c0.c1, g.c1.c2};
  want = f; fp12_mul_by_014(want, line[0], line[1], line[2]); got = f; coop12_exec_seq<OpMul014>(reinterpret_cast<Fp2*>(&got), line); if (!same(want, got)) bad |= 4;
  fp12_cyclotomic_sqr(want, f); got = f; coop12_exec_seq<OpCycSqr>(reinterpret_cast<Fp2*>(&got), nullptr); if (!same(want, got)) bad |= 8;
  fp12_frob(want, f, 1); got = f; coop12_exec_seq<OpFrob<1>>(reinterpret_cast<Fp2*>(&got), nullptr); if (!same(want, got)) bad |= 16;
  fp12_frob(want, f, 2); got = f; coop12_exec_seq<OpFrob<2>>(reinterpret_cast<Fp2*>(&got), nullptr); if (!same(want, got)) bad |= 32;
  return bad;
}
// dot12.cuh: term tables, double-width accumulator indexing and separated Montgomery reduction, run sequentially,
// against the tower functions of pairing.cuh.  Same input as emul_coop12_check; bit mask of mismatches.
int emul_dot12_check(const unsigned char* in) {
  Fp12 f, g;
  Fp* fe = reinterpret_cast<Fp*>(&f);
  Fp* ge = reinterpret_cast<Fp*>(&g);
  for (int i = 0; i < 12; i++) {
    uint32_t w[16];
    load_words(w, in + 64 * i, 16); if (fp_from_slot(fe[i], w) < 0) return -1;
    load_words(w, in + 64 * (12 + i), 16); if (fp_from_slot(ge[i], w) < 0) return -1;
  }
  auto same = [](const Fp12& a, const Fp12& b) { return memcmp(&a, &b, sizeof(Fp12)) == 0; };
  int bad = 0;
  Fp12 want, got;
  uint32_t fw[144], gw[144], ow[144];
  dot::fp12_to_words(fw, f);
  dot::fp12_to_words(gw, g);
  fp12_sqr(want, f); dot::exec_seq(dot::OP_SQR, ow, fw, nullptr); dot::words_to_fp12(got, ow); if (!same(want, got)) bad |= 1;
  fp12_mul(want, f, g); dot::exec_seq(dot::OP_MUL, ow, fw, gw); dot::words_to_fp12(got, ow); if (!same(want, got)) bad |= 2;
  Line ln = {g.c0.c0, g.c0.c1, g.c1.c2};
  want = f; fp12_mul_by_014(want, ln.l0, ln.l1, ln.l4);
  dot::exec_seq(dot::OP_MUL014, ow, fw, reinterpret_cast<const uint32_t*>(&ln)); dot::words_to_fp12(got, ow); if (!same(want, got)) bad |= 4;
  dot::exec_seq(dot::OP_MUL014, ow, fw, reinterpret_cast<const uint32_t*>(&ln), true); dot::words_to_fp12(got, ow); if (!same(want, got)) bad |= 16;
  fp12_sqr(want, f); dot::exec_seq(dot::OP_SQR, ow, fw, nullptr, true); dot::words_to_fp12(got, ow); if (!same(want, got)) bad |= 32;
  // the Granger-Scott formulas are plain algebra in the coefficients, so they can be compared on ANY element
  fp12_cyclotomic_sqr(want, f); dot::exec_seq(dot::OP_CYC, ow, fw, nullptr, true); dot::words_to_fp12(got, ow); if (!same(want, got)) bad |= 64;
  // a single extreme accumulation: 12 products of (p-1)*(p-1) must still reduce correctly
  {
    Fp pm1 = fp_load_const(C_P()); pm1.v[0] -= 1;
    dot::Acc A; dot::acc_zero(A);
    Fp sum = fp_zero();
    for (int t = 0; t < 12; t++) { dot::acc_product(A, pm1, pm1); sum = add(sum, mul(pm1, pm1)); }
    Fp r = dot::acc_reduce(A, 2);
    if (!eq(r, sum)) bad |= 8;
  }
  return bad;
}
// Deferred G2 membership (pairing_dot.cuh k_pairing_lines_slots): walk T = [|z|]Q with the Miller-loop step formulas
// and test psi(Q) == -T; Z = 0 at the end flags an exceptional chain (exact ladder decides).
// returns bit 0: member according to the walk (or the ladder after an exceptional chain), bit 1: chain was exceptional;
// counts[0] = Fp-mul of the exact ladder test, counts[1] = Fp-mul of the deferred comparison alone
int emul_g2_membership_deferred(unsigned long long* counts, const unsigned char* in256) {
  uint32_t w[64];
  load_words(w, in256, 64);
  G2Affine q;
  if (decode_point(q, w) != 0) return -1;
  if (is_inf(q)) return 1;
  unsigned long long c0 = g_fp_mul_count;
  const bool exact = g2_in_subgroup(q);
  counts[0] = g_fp_mul_count - c0;
  G2Proj t; t.x = q.x; t.y = q.y; t.z = fp2_one();
  Fp2 l0, l1, l4;
  for (int i = 62; i >= 0; i--) {
    ml_dbl_step(t, l0, l1, l4);
    if ((B200_Z_ABS >> i) & 1) ml_add_step(t, q, l0, l1, l4);
  }
  c0 = g_fp_mul_count;
  bool member;
  int exceptional = 0;
  if (is_zero(t.z)) { member = exact; exceptional = 2; }
  else {
    const Fp2 px = mulo(conj(q.x), fp2_load_const(C_PSI_CX()));
    const Fp2 py = mulo(conj(q.y), fp2_load_const(C_PSI_CY()));
    member = eq(mulo(px, t.z), t.x) && eq(mulo(py, t.z), neg(t.y));
  }
  counts[1] = g_fp_mul_count - c0;
  if (member != exact) return -2;          // the two tests must agree on EVERY point of E'(Fp2)
  return (member ? 1 : 0) | exceptional;
}
int emul_g1_in_subgroup(const unsigned char* in128) {
  uint32_t w[32]; load_words(w, in128, 32);
  G1Affine p; int code = decode_point(p, w);
  if (code) return -code;
  return g1_in_subgroup(p) ? 1 : 0;
}
int emul_g2_in_subgroup(const unsigned char* in256) {
  uint32_t w[64]; load_words(w, in256, 64);
  G2Affine p; int code = decode_point(p, w);
  if (code) return -code;
  return g2_in_subgroup(p) ? 1 : 0;
}
int emul_g1_add(unsigned char* out, const unsigned char* in) {
  uint32_t w[64]; load_words(w, in, 64);
  G1Affine a, b;
  int ca = decode_point(a, w), cb = decode_point(b, w + 32);
  if (ca) return ca;
  if (cb) return cb;
  G1XYZZ acc = xyzz_from_affine(b);
  xyzz_madd(acc, a);
  G1Affine r = xyzz_to_affine(acc);
  uint32_t o[32]; encode_point(o, r); memcpy(out, o, 128);
  return 0;
}
// MAP_FP_TO_G1 / MAP_FP2_TO_G2 through the product's map.cuh (straight-line SSWU, isogeny, cofactor clearing)
int emul_map_fp_to_g1(unsigned char* out128, const unsigned char* in64) {
  uint32_t wi[16], wo[32]; load_words(wi, in64, 16);
  int code = map_to_group(wo, wi, (Fp*)nullptr);
  if (code == 0) memcpy(out128, wo, 128);
  return code;
}
int emul_map_fp2_to_g2(unsigned char* out256, const unsigned char* in128) {
  uint32_t wi[32], wo[64]; load_words(wi, in128, 32);
  int code = map_to_group(wo, wi, (Fp2*)nullptr);
  if (code == 0) memcpy(out256, wo, 256);
  return code;
}
// chunk-size rule of the pairing batch planner (pairing.cuh)
unsigned emul_pairing_choose_chunk(const unsigned* tasks_for7, unsigned wave, unsigned forced) {
  return pairing_choose_chunk(tasks_for7, wave, forced, PAIRING_MAX_CHUNK_THREAD);
}
}
