/*
 * eip2537_oracle.c -- TEST INFRASTRUCTURE ONLY: CPU oracle + timed CPU baseline ("port").
 *
 * A portable-C restatement of the reference's hot path, following ITS control flow:
 *   codec            /root/reference/src/eip2537.c:263-420
 *   G1 MUL/MULTIEXP  :487-524 (mul), :541-561 (dispatch), :564-616 (naive), :619-708 (Bos-Coster)
 *   G2 MUL/MULTIEXP  :775-812, :829-849, :851-905, :908-998
 *   heap helpers     :57-151 (compare/sub/num_bits/sift/heapify), :154-259 (heapreplace)
 *   PAIRING          :1020-1081 (per pair: G1 decode, G1 subgroup, G2 decode, G2 subgroup,
 *                    single-pair Miller loop, running Fp12 product; one final exp)
 *   error codes      /root/reference/src/eip2537.h:31-40
 * The arithmetic the reference delegates to blst (github.com/supranational/blst, unpinned
 * HEAD, cloned by build.sh:3-11, ABSENT here) is restated in oracle_field.h/oracle_ec.inc.
 *
 * PARITY UNPINNED: the reference vendors no golden vectors (build.sh:13-52 downloads them).
 * This oracle is pinned instead against oracle/py_oracle.py (big-int model: textbook
 * pairing, naive r*P subgroup tests), public constants and algebraic laws, and by published
 * known answers reproduced exactly (tests/kat.py: RFC 9380 J.9.1/J.9.2/J.10.2 for the map functions,
 * the geth 2*G1 / 2*G2 vectors, 3*G1) -- see tests/test_oracle_c.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_field.h"

typedef unsigned char byte;
enum { OK = 0, NOT_ON_CURVE = 1, NOT_IN_SUBGROUP = 2, INVALID_ELEMENT = 3, INVALID_LENGTH = 5, MEMORY_ERROR = 7 };

unsigned long long oracle_fp_mul_count = 0;

/* ------------------------------------------------------------------ 256-bit heap records */
typedef struct { byte k[32]; uint32_t base_index; } msm_entry;   /* k little-endian */

static uint64_t ld64(const byte *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static void st64(byte *p, uint64_t v) { memcpy(p, &v, 8); }

static int scalar_less(const msm_entry *a, const msm_entry *b) {   /* a < b */
  for (int i = 3; i >= 0; i--) {
    uint64_t x = ld64(a->k + 8 * i), y = ld64(b->k + 8 * i);
    if (x != y) return x < y;
  }
  return 0;
}
static void scalar_sub(msm_entry *a, const msm_entry *b) {          /* a -= b */
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)ld64(a->k + 8 * i) - ld64(b->k + 8 * i) - borrow;
    st64(a->k + 8 * i, (uint64_t)d);
    borrow = (uint64_t)(d >> 64) & 1;
  }
}
static int scalar_bits(const msm_entry *a) {
  for (int i = 3; i >= 0; i--) {
    uint64_t x = ld64(a->k + 8 * i);
    if (x) return 64 * (i + 1) - __builtin_clzll(x);
  }
  return 0;
}
/* move the hole at `pos` towards the root while parents are smaller (":109-124") */
static void heap_siftdown(msm_entry *h, int start, int pos) {
  msm_entry e = h[pos];
  while (pos > start) {
    int parent = (pos - 1) >> 1;
    if (!scalar_less(&h[parent], &e)) break;
    h[pos] = h[parent];
    pos = parent;
  }
  h[pos] = e;
}
/* push `start` to a leaf following larger children, then bubble back up (":127-144") */
static void heap_siftup(msm_entry *h, int size, int start) {
  msm_entry e = h[start];
  int pos = start, child = 2 * start + 1;
  while (child < size) {
    int right = child + 1;
    if (right < size && !scalar_less(&h[right], &h[child])) child = right;
    h[pos] = h[child];
    pos = child;
    child = 2 * pos + 1;
  }
  h[pos] = e;
  heap_siftdown(h, start, pos);
}
static void heap_build(msm_entry *h, int size) {
  for (int i = (size - 1) / 2; i >= 0; --i) heap_siftup(h, size, i);
}

/* ------------------------------------------------------------------ G1 / G2 via the template */
#define FE ofp
#define FF(x) fp_##x
#define PT op1
#define AFF op1_affine
#define PF(x) p1_##x
#define CURVE_B (&OC_B1)
#define FIELD_ONE OC_ONE
#include "oracle_ec.inc"
#undef FE
#undef FF
#undef PT
#undef AFF
#undef PF
#undef CURVE_B
#undef FIELD_ONE


#define FE ofp2
#define FF(x) fp2_##x
#define PT op2
#define AFF op2_affine
#define PF(x) p2_##x
#define CURVE_B (&OC_B2)
#define FIELD_ONE OC_ONE2
#include "oracle_ec.inc"

/* ------------------------------------------------------------------ codec (:263-420) */
/* returns -1 invalid, 0 zero, 1 non-zero; fp left in Montgomery form */
static int fp_from_bytes(ofp *fp, const byte *in) {
  byte pad = 0;
  for (int i = 0; i < 16; i++) pad |= in[i];
  if (pad) return -1;
  for (int i = 0; i < 6; i++) {
    uint64_t limb = 0;
    const byte *src = in + 16 + 8 * (5 - i);
    for (int j = 0; j < 8; j++) limb = (limb << 8) | src[j];
    fp->l[i] = limb;
  }
  /* value must be < p: the reference tests (x + 0 mod p) == x; equivalently x - p borrows */
  uint64_t borrow = 0, nz = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)fp->l[i] - OC_P.l[i] - borrow;
    borrow = (uint64_t)(d >> 64) & 1;
    nz |= fp->l[i];
  }
  if (!borrow) return -1;
  fp_to_mont(fp, fp);
  return nz ? 1 : 0;
}
static void fp_to_bytes(byte *out, const ofp *fp) {
  ofp c;
  fp_from_mont(&c, fp);
  memset(out, 0, 16);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 8; j++) out[16 + 8 * (5 - i) + j] = (byte)(c.l[i] >> (8 * (7 - j)));
}
static int decode_g1_point(op1_affine *out, const byte *in) {
  int sx = fp_from_bytes(&out->x, in), sy = fp_from_bytes(&out->y, in + 64);
  if (sx < 0 || sy < 0) return INVALID_ELEMENT;
  if (sx == 0 && sy == 0) return OK;
  return p1_affine_on_curve(out) ? OK : NOT_ON_CURVE;
}
static void encode_g1_point(byte *out, const op1_affine *a) {
  fp_to_bytes(out, &a->x);
  fp_to_bytes(out + 64, &a->y);
}
static int fp2_from_bytes(ofp2 *f, const byte *in) {
  int s0 = fp_from_bytes(&f->c0, in), s1 = fp_from_bytes(&f->c1, in + 64);
  if (s0 < 0 || s1 < 0) return -1;
  return s0 | s1;
}
static int decode_g2_point(op2_affine *out, const byte *in) {
  int sx = fp2_from_bytes(&out->x, in), sy = fp2_from_bytes(&out->y, in + 128);
  if (sx < 0 || sy < 0) return INVALID_ELEMENT;
  if (sx == 0 && sy == 0) return OK;
  return p2_affine_on_curve(out) ? OK : NOT_ON_CURVE;
}
static void encode_g2_point(byte *out, const op2_affine *a) {
  fp_to_bytes(out, &a->x.c0); fp_to_bytes(out + 64, &a->x.c1);
  fp_to_bytes(out + 128, &a->y.c0); fp_to_bytes(out + 192, &a->y.c1);
}
static void decode_scalar(byte k_le[32], const byte *in) {   /* BE -> LE, never fails */
  for (int i = 0; i < 32; i++) k_le[i] = in[31 - i];
}

/* ------------------------------------------------------------------ MULTIEXP, field-generic */
#define DEFINE_MSM(G, PT, AFF, STRIDE, PTLEN, OUTLEN)                                          \
  int oracle_bls12_##G##mul(byte *out, const byte *in, size_t in_len) {                        \
    if (in_len != STRIDE) return INVALID_LENGTH;                                               \
    AFF a; int ret = decode_##G##_point(&a, in);                                               \
    if (ret) return ret;                                                                       \
    byte k[32]; decode_scalar(k, in + PTLEN);                                                  \
    PT p, q; G##_from_affine(&p, &a);                                                          \
    G##_mult(&q, &p, k, 256);                                                                  \
    AFF r; G##_to_affine(&r, &q); encode_##G##_point(out, &r);                                 \
    return OK;                                                                                 \
  }                                                                                            \
  int oracle_bls12_##G##multiexp_naive(byte *out, const byte *in, size_t in_len) {             \
    if (in_len == 0 || in_len % STRIDE) return INVALID_LENGTH;                                 \
    size_t n = in_len / STRIDE;                                                                \
    if (n == 1) return oracle_bls12_##G##mul(out, in, in_len);                                 \
    PT acc; memset(&acc, 0, sizeof acc);                                                       \
    for (size_t i = 0; i < n; i++, in += STRIDE) {                                             \
      AFF a; int ret = decode_##G##_point(&a, in);                                             \
      if (ret) return ret;                                                                     \
      byte k[32]; decode_scalar(k, in + PTLEN);                                                \
      PT p, q; G##_from_affine(&p, &a);                                                        \
      G##_mult(&q, &p, k, 256);                                                                \
      G##_add_or_double(&acc, &acc, &q);                                                       \
    }                                                                                          \
    AFF r; G##_to_affine(&r, &acc); encode_##G##_point(out, &r);                               \
    return OK;                                                                                 \
  }                                                                                            \
  int oracle_bls12_##G##multiexp_bc(byte *out, const byte *in, size_t in_len) {                \
    if (in_len == 0 || in_len % STRIDE) return INVALID_LENGTH;                                 \
    size_t n = in_len / STRIDE;                                                                \
    if (n == 1) return oracle_bls12_##G##mul(out, in, in_len);                                 \
    PT *bases = (PT *)malloc(n * sizeof(PT));                                                  \
    if (!bases) return MEMORY_ERROR;                                                           \
    msm_entry *heap = (msm_entry *)malloc(n * sizeof(msm_entry));                              \
    if (!heap) { free(bases); return MEMORY_ERROR; }                                           \
    for (size_t i = 0; i < n; i++, in += STRIDE) {                                             \
      AFF a; int ret = decode_##G##_point(&a, in);                                             \
      if (ret) { free(bases); free(heap); return ret; }                                        \
      G##_from_affine(&bases[i], &a);                                                          \
      decode_scalar(heap[i].k, in + PTLEN);                                                    \
      heap[i].base_index = (uint32_t)i;                                                        \
    }                                                                                          \
    heap_build(heap, (int)n);                                                                  \
    PT skipped; memset(&skipped, 0, sizeof skipped);                                           \
    while (G##_heapreplace(&skipped, bases, heap, (int)n)) {}                                  \
    PT res;                                                                                    \
    G##_mult(&res, &bases[heap[0].base_index], heap[0].k, (size_t)scalar_bits(&heap[0]));      \
    if (!G##_is_inf(&skipped)) G##_add_or_double(&res, &res, &skipped);                        \
    AFF r; G##_to_affine(&r, &res); encode_##G##_point(out, &r);                               \
    free(bases); free(heap);                                                                   \
    return OK;                                                                                 \
  }                                                                                            \
  int oracle_bls12_##G##multiexp(byte *out, const byte *in, size_t in_len) {                   \
    if (in_len == 0 || in_len % STRIDE) return INVALID_LENGTH;                                 \
    size_t n = in_len / STRIDE;                                                                \
    if (n == 1) return oracle_bls12_##G##mul(out, in, in_len);                                 \
    if (n <= 4) return oracle_bls12_##G##multiexp_naive(out, in, in_len);                      \
    return oracle_bls12_##G##multiexp_bc(out, in, in_len);                                     \
  }

#define g1_from_affine p1_from_affine
#define g1_mult p1_mult
#define g1_to_affine p1_to_affine
#define g1_add_or_double p1_add_or_double
#define g1_heapreplace p1_heapreplace
#define g1_is_inf p1_is_inf
#define g2_from_affine p2_from_affine
#define g2_mult p2_mult
#define g2_to_affine p2_to_affine
#define g2_add_or_double p2_add_or_double
#define g2_heapreplace p2_heapreplace
#define g2_is_inf p2_is_inf
DEFINE_MSM(g1, op1, op1_affine, 160, 128, 128)
DEFINE_MSM(g2, op2, op2_affine, 288, 256, 256)

/* ------------------------------------------------------------------ subgroup membership */
/* naive: r*P == O  (the definition; blst_p1_affine_in_g1 semantics incl. infinity -> true) */
int oracle_g1_in_subgroup_naive_aff(const op1_affine *a) {
  op1 p, q; p1_from_affine(&p, a); p1_mult(&q, &p, OC_R_LE, 255); return p1_is_inf(&q);
}
int oracle_g2_in_subgroup_naive_aff(const op2_affine *a) {
  op2 p, q; p2_from_affine(&p, a); p2_mult(&q, &p, OC_R_LE, 255); return p2_is_inf(&q);
}
/* fast (Scott 2021): phi(P) == [-z^2]P on E(Fp);  psi(Q) == [z]Q on E'(Fp2) */
static int p1_affine_in_g1(const op1_affine *a) {
  if (fp_is_zero(&a->x) && fp_is_zero(&a->y)) return 1;
  op1 p, q; op1_affine r;
  p1_from_affine(&p, a);
  p1_mult(&q, &p, OC_ZSQ_LE, 128);
  if (p1_is_inf(&q)) return 0;
  p1_to_affine(&r, &q);
  ofp bx, ny;
  fp_mul(&bx, &a->x, &OC_BETA);
  fp_neg(&ny, &r.y);
  return fp_eq(&bx, &r.x) && fp_eq(&a->y, &ny);
}
static int p2_affine_in_g2(const op2_affine *a) {
  if (fp2_is_zero(&a->x) && fp2_is_zero(&a->y)) return 1;
  op2 p, q; op2_affine r;
  byte zle[8];
  st64(zle, OC_Z_ABS);
  p2_from_affine(&p, a);
  p2_mult(&q, &p, zle, 64);
  if (p2_is_inf(&q)) return 0;
  p2_to_affine(&r, &q);
  ofp2 px, py, ny;
  fp2_conj(&px, &a->x); fp2_mul(&px, &px, &OC_PSI_CX);
  fp2_conj(&py, &a->y); fp2_mul(&py, &py, &OC_PSI_CY);
  fp2_neg(&ny, &r.y);                       /* z < 0: [z]Q = -[|z|]Q */
  return fp2_eq(&px, &r.x) && fp2_eq(&py, &ny);
}

/* ------------------------------------------------------------------ pairing */
typedef struct { ofp2 x, y, z; } g2proj;  /* homogeneous projective point on the twist */

static void ml_dbl_step(g2proj *t, ofp2 *l0, ofp2 *l1, ofp2 *l4) {
  ofp2 a, b, c, e, f, g, h, i, j, e2, s;
  fp2_mul(&a, &t->x, &t->y); fp2_mul_fp(&a, &a, &OC_INV2);
  fp2_sqr(&b, &t->y); fp2_sqr(&c, &t->z);
  fp2_mul(&e, &OC_B2x3, &c);
  fp2_dbl(&f, &e); fp2_add(&f, &f, &e);
  fp2_add(&g, &b, &f); fp2_mul_fp(&g, &g, &OC_INV2);
  fp2_add(&h, &t->y, &t->z); fp2_sqr(&h, &h); fp2_add(&s, &b, &c); fp2_sub(&h, &h, &s);
  fp2_sub(&i, &e, &b);
  fp2_sqr(&j, &t->x);
  fp2_sqr(&e2, &e);
  fp2_sub(&s, &b, &f); fp2_mul(&t->x, &a, &s);
  fp2_sqr(&g, &g); fp2_dbl(&s, &e2); fp2_add(&s, &s, &e2); fp2_sub(&t->y, &g, &s);
  fp2_mul(&t->z, &b, &h);
  *l0 = i;
  fp2_dbl(l1, &j); fp2_add(l1, l1, &j);
  fp2_neg(l4, &h);
}
static void ml_add_step(g2proj *t, const op2_affine *q, ofp2 *l0, ofp2 *l1, ofp2 *l4) {
  ofp2 theta, lam, c, d, e, f, g, h, s, u;
  fp2_mul(&s, &q->y, &t->z); fp2_sub(&theta, &t->y, &s);
  fp2_mul(&s, &q->x, &t->z); fp2_sub(&lam, &t->x, &s);
  fp2_sqr(&c, &theta); fp2_sqr(&d, &lam);
  fp2_mul(&e, &lam, &d); fp2_mul(&f, &t->z, &c); fp2_mul(&g, &t->x, &d);
  fp2_add(&h, &e, &f); fp2_dbl(&s, &g); fp2_sub(&h, &h, &s);
  fp2_mul(&t->x, &lam, &h);
  fp2_sub(&s, &g, &h); fp2_mul(&s, &theta, &s); fp2_mul(&u, &e, &t->y); fp2_sub(&t->y, &s, &u);
  fp2_mul(&t->z, &t->z, &e);
  fp2_mul(&s, &theta, &q->x); fp2_mul(&u, &lam, &q->y); fp2_sub(l0, &s, &u);
  fp2_neg(l1, &theta);
  *l4 = lam;
}
/* single-pair Miller loop, argument order (ret, Q, P) as blst_miller_loop (Appendix B);
 * a pair with an infinite member contributes 1 (Appendix D-2) */
static void miller_loop(ofp12 *ret, const op2_affine *q, const op1_affine *p) {
  ofp12 f;
  fp12_set_one(&f);
  int p_inf = fp_is_zero(&p->x) && fp_is_zero(&p->y);
  int q_inf = fp2_is_zero(&q->x) && fp2_is_zero(&q->y);
  if (p_inf || q_inf) { *ret = f; return; }
  g2proj t; t.x = q->x; t.y = q->y; t.z = OC_ONE2;
  ofp2 l0, l1, l4;
  for (int i = 62; i >= 0; i--) {
    fp12_sqr(&f, &f);
    ml_dbl_step(&t, &l0, &l1, &l4);
    fp2_mul_fp(&l1, &l1, &p->x); fp2_mul_fp(&l4, &l4, &p->y);
    fp12_mul_by_014(&f, &l0, &l1, &l4);
    if ((OC_Z_ABS >> i) & 1) {
      ml_add_step(&t, q, &l0, &l1, &l4);
      fp2_mul_fp(&l1, &l1, &p->x); fp2_mul_fp(&l4, &l4, &p->y);
      fp12_mul_by_014(&f, &l0, &l1, &l4);
    }
  }
  fp12_conj(ret, &f);   /* z < 0 */
}
/* a^z in the cyclotomic subgroup */
static void cyc_exp_z(ofp12 *r, const ofp12 *a) {
  ofp12 acc = *a;
  for (int i = 62; i >= 0; i--) {
    fp12_cyclotomic_sqr(&acc, &acc);
    if ((OC_Z_ABS >> i) & 1) fp12_mul(&acc, &acc, a);
  }
  fp12_conj(r, &acc);
}
/* f^((p^6-1)(p^2+1)) then ^(3(p^4-p^2+1)/r) = ^((z-1)^2 (z+p)(z^2+p^2-1)) * ^3 */
static void final_exp(ofp12 *r, const ofp12 *fin) {
  ofp12 f, t0, t1, t2, u;
  fp12_inv(&t0, fin); fp12_conj(&f, fin); fp12_mul(&f, &f, &t0);
  fp12_frob(&t0, &f, 2); fp12_mul(&f, &t0, &f);
  cyc_exp_z(&t0, &f); fp12_conj(&u, &f); fp12_mul(&t0, &t0, &u);          /* f^(z-1)     */
  cyc_exp_z(&t1, &t0); fp12_conj(&u, &t0); fp12_mul(&t0, &t1, &u);        /* f^((z-1)^2) */
  cyc_exp_z(&t1, &t0); fp12_frob(&u, &t0, 1); fp12_mul(&t1, &t1, &u);     /* ^(z+p)      */
  cyc_exp_z(&t2, &t1); cyc_exp_z(&t2, &t2);
  fp12_frob(&u, &t1, 2); fp12_mul(&t2, &t2, &u);
  fp12_conj(&u, &t1); fp12_mul(&t2, &t2, &u);                             /* ^(z^2+p^2-1) */
  fp12_cyclotomic_sqr(&u, &f); fp12_mul(&u, &u, &f);                      /* f^3 */
  fp12_mul(r, &t2, &u);
}

static int pairing_core(ofp12 *gt, const byte *in, size_t in_len) {
  if (in_len == 0 || in_len % 384) return INVALID_LENGTH;
  size_t k = in_len / 384;
  ofp12 acc;
  for (size_t i = 0; i < k; i++, in += 384) {
    op1_affine p; op2_affine q;
    int ret = decode_g1_point(&p, in);
    if (ret) return ret;
    if (!p1_affine_in_g1(&p)) return NOT_IN_SUBGROUP;
    ret = decode_g2_point(&q, in + 128);
    if (ret) return ret;
    if (!p2_affine_in_g2(&q)) return NOT_IN_SUBGROUP;
    if (i > 0) { ofp12 cur; miller_loop(&cur, &q, &p); fp12_mul(&acc, &acc, &cur); }
    else miller_loop(&acc, &q, &p);
  }
  final_exp(gt, &acc);
  return OK;
}
int oracle_bls12_pairing(byte *out, const byte *in, size_t in_len) {
  ofp12 gt;
  int ret = pairing_core(&gt, in, in_len);
  if (ret) return ret;
  memset(out, 0, 32);
  if (fp12_is_one(&gt)) out[31] = 1;
  return OK;
}
/* ------------------------------------------------------------------ MAP_FP_TO_G1 / MAP_FP2_TO_G2
 * eip2537.c:1093-1121 and :1135-1163: length check, field decode (invalid -> 3), blst_map_to_g1/_g2(p, u, NULL)
 * [blst-upstream], to_affine, encode.  blst's map is RFC 9380 8.8: simplified SWU onto the isogenous curve
 * (6.6.2, textbook form with inversions here), isogeny (E.2/E.3, coefficients derived by
 * oracle/derive_isogeny.py), multiplication by the effective cofactor h_eff.                       */
static void fp_pow_le(ofp *r, const ofp *a, const unsigned char *e, int nbytes) {
  ofp acc = OC_ONE, b = *a;
  for (int i = 8 * nbytes - 1; i >= 0; i--) {
    fp_sqr(&acc, &acc);
    if ((e[i >> 3] >> (i & 7)) & 1) fp_mul(&acc, &acc, &b);
  }
  *r = acc;
}
static void fp2_pow_le(ofp2 *r, const ofp2 *a, const unsigned char *e, int nbytes) {
  ofp2 acc = OC_ONE2, b = *a;
  for (int i = 8 * nbytes - 1; i >= 0; i--) {
    fp2_sqr(&acc, &acc);
    if ((e[i >> 3] >> (i & 7)) & 1) fp2_mul(&acc, &acc, &b);
  }
  *r = acc;
}
static int fp_sqrt(ofp *r, const ofp *a) {         /* 1 if a is a square (r = a root) */
  ofp t, s;
  fp_pow_le(&t, a, OC_EXP_SQRT_LE, 48);
  fp_sqr(&s, &t);
  *r = t;
  return fp_eq(&s, a);
}
static int fp2_sqrt(ofp2 *r, const ofp2 *a) {      /* p = 3 mod 4 "complex method" */
  if (fp2_is_zero(a)) { *r = *a; return 1; }
  ofp2 a1, alpha, x0, x, chk, m1 = OC_ONE2;
  fp2_neg(&m1, &m1);
  fp2_pow_le(&a1, a, OC_EXP_PM3D4_LE, 48);
  fp2_sqr(&alpha, &a1); fp2_mul(&alpha, &alpha, a);
  fp2_mul(&x0, &a1, a);
  if (fp2_eq(&alpha, &m1)) {                        /* x = i * x0 */
    fp_neg(&x.c0, &x0.c1); x.c1 = x0.c0;
  } else {
    ofp2 b;
    fp2_add(&b, &alpha, &OC_ONE2);
    fp2_pow_le(&b, &b, OC_EXP_PM1D2_LE, 48);
    fp2_mul(&x, &b, &x0);
  }
  fp2_sqr(&chk, &x);
  *r = x;
  return fp2_eq(&chk, a);
}
static int fp_sgn0(const ofp *a) { ofp c; fp_from_mont(&c, a); return (int)(c.l[0] & 1); }
static int fp2_sgn0(const ofp2 *a) {               /* RFC 9380 4.1 for m = 2 */
  ofp c0, c1;
  fp_from_mont(&c0, &a->c0); fp_from_mont(&c1, &a->c1);
  int s0 = (int)(c0.l[0] & 1), z0 = fp_is_zero(&c0), s1 = (int)(c1.l[0] & 1);
  return s0 | (z0 & s1);
}

#define DEFINE_MAP(FE, FF, NXN, NXD, NYN, NYD, ISO, AFF, PT, PF, HEFF, HEFFBITS)                 \
  static void FF##_curve_rhs(FE *g, const FE *x) {                                                \
    FE t;                                                                                         \
    FF##_sqr(&t, x); FF##_add(&t, &t, &ISO##_A); FF##_mul(&t, &t, x); FF##_add(g, &t, &ISO##_B);  \
  }                                                                                               \
  static void FF##_horner(FE *r, const FE *cf, int n, const FE *x) {                              \
    FE acc = cf[n - 1];                                                                           \
    for (int i = n - 2; i >= 0; i--) { FF##_mul(&acc, &acc, x); FF##_add(&acc, &acc, &cf[i]); }   \
    *r = acc;                                                                                     \
  }                                                                                               \
  static void FF##_map_to_group(AFF *out, const FE *u) {                                          \
    FE u2, zu2, den, tv1, x1, x, y, g, t;                                                         \
    FF##_sqr(&u2, u); FF##_mul(&zu2, &ISO##_Z, &u2);                                              \
    FF##_sqr(&den, &zu2); FF##_add(&den, &den, &zu2);                                             \
    if (FF##_is_zero(&den)) {                            /* x1 = B / (Z A) */                     \
      FF##_mul(&t, &ISO##_Z, &ISO##_A); FF##_inv(&t, &t); FF##_mul(&x1, &ISO##_B, &t);            \
    } else {                                             /* x1 = (-B/A)(1 + 1/den) */             \
      FE one; memset(&one, 0, sizeof one); memcpy(&one, &OC_ONE, sizeof OC_ONE);                  \
      FF##_inv(&tv1, &den); FF##_add(&tv1, &tv1, &one);                                           \
      FF##_inv(&t, &ISO##_A); FF##_mul(&t, &t, &ISO##_B); FF##_neg(&t, &t);                       \
      FF##_mul(&x1, &t, &tv1);                                                                    \
    }                                                                                             \
    FF##_curve_rhs(&g, &x1);                                                                      \
    if (FF##_sqrt(&y, &g)) x = x1;                                                                \
    else { FF##_mul(&x, &zu2, &x1); FF##_curve_rhs(&g, &x); FF##_sqrt(&y, &g); }                  \
    if (FF##_sgn0(u) != FF##_sgn0(&y)) FF##_neg(&y, &y);                                          \
    /* isogeny E' -> E */                                                                         \
    FE xn, xd, yn, yd;                                                                            \
    FF##_horner(&xn, ISO##_XNUM, NXN, &x); FF##_horner(&xd, ISO##_XDEN, NXD, &x);                 \
    FF##_horner(&yn, ISO##_YNUM, NYN, &x); FF##_horner(&yd, ISO##_YDEN, NYD, &x);                 \
    AFF e;                                                                                        \
    memset(&e, 0, sizeof e);                             /* kernel point -> infinity (0,0) */     \
    if (!FF##_is_zero(&xd) && !FF##_is_zero(&yd)) {                                               \
      FF##_inv(&xd, &xd); FF##_mul(&e.x, &xn, &xd);                                               \
      FF##_inv(&yd, &yd); FF##_mul(&e.y, &yn, &yd); FF##_mul(&e.y, &e.y, &y);                     \
    }                                                                                             \
    PT p, q;                                                                                      \
    PF##_from_affine(&p, &e);                                                                     \
    PF##_mult(&q, &p, HEFF, HEFFBITS);                                                            \
    PF##_to_affine(out, &q);                                                                      \
  }

DEFINE_MAP(ofp, fp, 12, 11, 16, 16, OC_ISO1, op1_affine, op1, p1, OC_HEFF1_LE, OC_HEFF1_BITS)
DEFINE_MAP(ofp2, fp2, 4, 3, 4, 4, OC_ISO2, op2_affine, op2, p2, OC_HEFF2_LE, OC_HEFF2_BITS)

int oracle_bls12_map_fp_to_g1(byte *out, const byte *in, size_t in_len) {
  if (in_len != 64) return INVALID_LENGTH;
  ofp u;
  if (fp_from_bytes(&u, in) < 0) return INVALID_ELEMENT;
  op1_affine a;
  fp_map_to_group(&a, &u);
  encode_g1_point(out, &a);
  return OK;
}
int oracle_bls12_map_fp2_to_g2(byte *out, const byte *in, size_t in_len) {
  if (in_len != 128) return INVALID_LENGTH;
  ofp2 u;
  if (fp2_from_bytes(&u, in) < 0) return INVALID_ELEMENT;
  op2_affine a;
  fp2_map_to_group(&a, &u);
  encode_g2_point(out, &a);
  return OK;
}

/* debug/pinning: the GT element as 12 x 48 big-endian bytes, order c0.c0.c0, c0.c0.c1, c0.c1.c0 ... */
int oracle_pairing_gt(byte *out576, const byte *in, size_t in_len) {
  ofp12 gt;
  int ret = pairing_core(&gt, in, in_len);
  if (ret) return ret;
  const ofp *e = (const ofp *)&gt;
  for (int i = 0; i < 12; i++) { byte tmp[64]; fp_to_bytes(tmp, &e[i]); memcpy(out576 + 48 * i, tmp + 16, 48); }
  return OK;
}

/* ------------------------------------------------------------------ helpers for tests / bench */
/* many independent calls; offsets[n+1] byte offsets into `in`; errs[n]; outs n*32 */
void oracle_bls12_pairing_batch(byte *outs, int *errs, const byte *in, const uint64_t *offsets, size_t n) {
  for (size_t i = 0; i < n; i++)
    errs[i] = oracle_bls12_pairing(outs + 32 * i, in + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
}
int oracle_g1_in_subgroup(const byte *in128, int naive) {   /* -1 on decode error */
  op1_affine a;
  if (decode_g1_point(&a, in128)) return -1;
  return naive ? oracle_g1_in_subgroup_naive_aff(&a) : p1_affine_in_g1(&a);
}
int oracle_g2_in_subgroup(const byte *in256, int naive) {
  op2_affine a;
  if (decode_g2_point(&a, in256)) return -1;
  return naive ? oracle_g2_in_subgroup_naive_aff(&a) : p2_affine_in_g2(&a);
}
/* out = k * G1 generator / G2 generator (k big-endian 32 bytes): workload generators */
void oracle_g1_gen_mul(byte *out128, const byte *k_be) {
  op1_affine g = { OC_G1X, OC_G1Y }, r; op1 p, q; byte k[32];
  decode_scalar(k, k_be); p1_from_affine(&p, &g); p1_mult(&q, &p, k, 256); p1_to_affine(&r, &q); encode_g1_point(out128, &r);
}
void oracle_g2_gen_mul(byte *out256, const byte *k_be) {
  op2_affine g = { OC_G2X, OC_G2Y }, r; op2 p, q; byte k[32];
  decode_scalar(k, k_be); p2_from_affine(&p, &g); p2_mult(&q, &p, k, 256); p2_to_affine(&r, &q); encode_g2_point(out256, &r);
}
/* out[i] = a + i*d for i < n (encoded affine), one batched inversion: cheap 2^20-point workloads */
int oracle_g1_arith_progression(byte *out, const byte *a128, const byte *d128, size_t n) {
  op1_affine a, d;
  if (decode_g1_point(&a, a128) || decode_g1_point(&d, d128)) return -1;
  op1 *pts = (op1 *)malloc(n * sizeof(op1));
  ofp *pre = (ofp *)malloc(n * sizeof(ofp));
  if (!pts || !pre) { free(pts); free(pre); return MEMORY_ERROR; }
  op1 dj; p1_from_affine(&dj, &d);
  p1_from_affine(&pts[0], &a);
  for (size_t i = 1; i < n; i++) p1_add_or_double(&pts[i], &pts[i - 1], &dj);
  ofp acc = OC_ONE;
  for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!p1_is_inf(&pts[i])) fp_mul(&acc, &acc, &pts[i].z); }
  ofp inv; fp_inv(&inv, &acc);
  for (size_t i = n; i-- > 0;) {
    op1_affine r; memset(&r, 0, sizeof r);
    if (!p1_is_inf(&pts[i])) {
      ofp zi, zi2, zi3;
      fp_mul(&zi, &inv, &pre[i]); fp_mul(&inv, &inv, &pts[i].z);
      fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
      fp_mul(&r.x, &pts[i].x, &zi2); fp_mul(&r.y, &pts[i].y, &zi3);
    }
    encode_g1_point(out + 128 * i, &r);
  }
  free(pts); free(pre);
  return OK;
}
int oracle_g2_arith_progression(byte *out, const byte *a256, const byte *d256, size_t n) {
  op2_affine a, d;
  if (decode_g2_point(&a, a256) || decode_g2_point(&d, d256)) return -1;
  op2 *pts = (op2 *)malloc(n * sizeof(op2));
  ofp2 *pre = (ofp2 *)malloc(n * sizeof(ofp2));
  if (!pts || !pre) { free(pts); free(pre); return MEMORY_ERROR; }
  op2 dj; p2_from_affine(&dj, &d);
  p2_from_affine(&pts[0], &a);
  for (size_t i = 1; i < n; i++) p2_add_or_double(&pts[i], &pts[i - 1], &dj);
  ofp2 acc = OC_ONE2;
  for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!p2_is_inf(&pts[i])) fp2_mul(&acc, &acc, &pts[i].z); }
  ofp2 inv; fp2_inv(&inv, &acc);
  for (size_t i = n; i-- > 0;) {
    op2_affine r; memset(&r, 0, sizeof r);
    if (!p2_is_inf(&pts[i])) {
      ofp2 zi, zi2, zi3;
      fp2_mul(&zi, &inv, &pre[i]); fp2_mul(&inv, &inv, &pts[i].z);
      fp2_sqr(&zi2, &zi); fp2_mul(&zi3, &zi2, &zi);
      fp2_mul(&r.x, &pts[i].x, &zi2); fp2_mul(&r.y, &pts[i].y, &zi3);
    }
    encode_g2_point(out + 256 * i, &r);
  }
  free(pts); free(pre);
  return OK;
}
unsigned long long oracle_get_fp_mul_count(void) { return oracle_fp_mul_count; }
void oracle_reset_fp_mul_count(void) { oracle_fp_mul_count = 0; }
