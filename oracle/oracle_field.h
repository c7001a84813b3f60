/*
 * oracle_field.h -- TEST INFRASTRUCTURE ONLY (CPU checker + timed CPU baseline).
 *
 * Portable-C restatement of the field tower the reference gets from blst
 * (Fp, Fp2, Fp6, Fp12 over BLS12-381; 6x64-bit Montgomery limbs, R = 2^384, as
 * `blst_fp` is described in SURVEY.md 8(a) row a14).  blst itself is NOT in
 * /root/reference (build.sh:3-11 clones it), so these are the published algorithms
 * (CIOS Montgomery, Karatsuba towers, Granger-Scott cyclotomic squaring), pinned
 * against oracle/py_oracle.py.  PARITY UNPINNED by reference vectors (none vendored).
 *
 * Nothing in blst_eip2537_b200/ may include or link this.
 */
#ifndef ORACLE_FIELD_H
#define ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t l[6]; } ofp;
typedef struct { ofp c0, c1; } ofp2;
typedef struct { ofp2 c0, c1, c2; } ofp6;
typedef struct { ofp6 c0, c1; } ofp12;
typedef unsigned __int128 u128;

#include "oracle_constants.h"

extern unsigned long long oracle_fp_mul_count; /* instrumented FME counter */

/* ---------------------------------------------------------------- Fp */
static inline int fp_is_zero(const ofp *a) {
  uint64_t acc = 0;
  for (int i = 0; i < 6; i++) acc |= a->l[i];
  return acc == 0;
}
static inline int fp_eq(const ofp *a, const ofp *b) {
  uint64_t acc = 0;
  for (int i = 0; i < 6; i++) acc |= a->l[i] ^ b->l[i];
  return acc == 0;
}
/* r = a - p if a >= p (a < 2p, optional carry bit on top) */
static inline void fp_cond_sub_p(ofp *r, const uint64_t a[6], uint64_t top) {
  uint64_t t[6], borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)a[i] - OC_P.l[i] - borrow;
    t[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  int use = (top != 0) || (borrow == 0);
  for (int i = 0; i < 6; i++) r->l[i] = use ? t[i] : a[i];
}
static inline void fp_add(ofp *r, const ofp *a, const ofp *b) {
  uint64_t t[6], carry = 0;
  for (int i = 0; i < 6; i++) {
    u128 s = (u128)a->l[i] + b->l[i] + carry;
    t[i] = (uint64_t)s;
    carry = (uint64_t)(s >> 64);
  }
  fp_cond_sub_p(r, t, carry);
}
static inline void fp_sub(ofp *r, const ofp *a, const ofp *b) {
  uint64_t t[6], borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - borrow;
    t[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  if (borrow) {
    uint64_t carry = 0;
    for (int i = 0; i < 6; i++) {
      u128 s = (u128)t[i] + OC_P.l[i] + carry;
      t[i] = (uint64_t)s;
      carry = (uint64_t)(s >> 64);
    }
  }
  memcpy(r->l, t, sizeof t);
}
static inline void fp_neg(ofp *r, const ofp *a) {
  ofp z = {{0}};
  fp_sub(r, &z, a);
}
static inline void fp_dbl(ofp *r, const ofp *a) { fp_add(r, a, a); }

/* CIOS Montgomery product a*b/R mod p */
static inline void fp_mul(ofp *r, const ofp *a, const ofp *b) {
  uint64_t t[8] = {0};
#ifdef ORACLE_COUNT
  oracle_fp_mul_count++;
#endif
  for (int i = 0; i < 6; i++) {
    uint64_t carry = 0, bi = b->l[i];
    for (int j = 0; j < 6; j++) {
      u128 acc = (u128)a->l[j] * bi + t[j] + carry;
      t[j] = (uint64_t)acc;
      carry = (uint64_t)(acc >> 64);
    }
    u128 acc = (u128)t[6] + carry;
    t[6] = (uint64_t)acc;
    t[7] = (uint64_t)(acc >> 64);
    uint64_t m = t[0] * OC_N0;
    acc = (u128)m * OC_P.l[0] + t[0];
    carry = (uint64_t)(acc >> 64);
    for (int j = 1; j < 6; j++) {
      acc = (u128)m * OC_P.l[j] + t[j] + carry;
      t[j - 1] = (uint64_t)acc;
      carry = (uint64_t)(acc >> 64);
    }
    acc = (u128)t[6] + carry;
    t[5] = (uint64_t)acc;
    t[6] = t[7] + (uint64_t)(acc >> 64);
  }
  fp_cond_sub_p(r, t, t[6]);
}
static inline void fp_sqr(ofp *r, const ofp *a) { fp_mul(r, a, a); }
static inline void fp_to_mont(ofp *r, const ofp *a) { fp_mul(r, a, &OC_RR); }
static inline void fp_from_mont(ofp *r, const ofp *a) {
  ofp one = {{1, 0, 0, 0, 0, 0}};
  fp_mul(r, a, &one);
}
/* a^(p-2) by square-and-multiply; inverse of 0 is 0 */
static inline void fp_inv(ofp *r, const ofp *a) {
  ofp acc = OC_ONE, base = *a;
  uint64_t e[6];
  memcpy(e, OC_P.l, sizeof e);
  e[0] -= 2; /* p is odd and its low limb is far from 0/1 */
  for (int i = 0; i < 384; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(&acc, &acc, &base);
    fp_sqr(&base, &base);
  }
  *r = acc;
}

/* ---------------------------------------------------------------- Fp2 = Fp[u]/(u^2+1) */
static inline int fp2_is_zero(const ofp2 *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static inline int fp2_eq(const ofp2 *a, const ofp2 *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static inline void fp2_add(ofp2 *r, const ofp2 *a, const ofp2 *b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void fp2_sub(ofp2 *r, const ofp2 *a, const ofp2 *b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void fp2_neg(ofp2 *r, const ofp2 *a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void fp2_dbl(ofp2 *r, const ofp2 *a) { fp2_add(r, a, a); }
static inline void fp2_conj(ofp2 *r, const ofp2 *a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }
static inline void fp2_mul(ofp2 *r, const ofp2 *a, const ofp2 *b) {
  ofp t0, t1, s0, s1, t2;
  fp_mul(&t0, &a->c0, &b->c0);
  fp_mul(&t1, &a->c1, &b->c1);
  fp_add(&s0, &a->c0, &a->c1);
  fp_add(&s1, &b->c0, &b->c1);
  fp_mul(&t2, &s0, &s1);
  fp_sub(&r->c0, &t0, &t1);
  fp_sub(&t2, &t2, &t0);
  fp_sub(&r->c1, &t2, &t1);
}
static inline void fp2_sqr(ofp2 *r, const ofp2 *a) {
  ofp s, d, m;
  fp_add(&s, &a->c0, &a->c1);
  fp_sub(&d, &a->c0, &a->c1);
  fp_mul(&m, &a->c0, &a->c1);
  fp_mul(&r->c0, &s, &d);
  fp_dbl(&r->c1, &m);
}
static inline void fp2_mul_fp(ofp2 *r, const ofp2 *a, const ofp *k) { fp_mul(&r->c0, &a->c0, k); fp_mul(&r->c1, &a->c1, k); }
/* multiply by xi = 1+u */
static inline void fp2_mul_xi(ofp2 *r, const ofp2 *a) {
  ofp t;
  fp_sub(&t, &a->c0, &a->c1);
  fp_add(&r->c1, &a->c0, &a->c1);
  r->c0 = t;
}
static inline void fp2_inv(ofp2 *r, const ofp2 *a) {
  ofp t0, t1;
  fp_sqr(&t0, &a->c0);
  fp_sqr(&t1, &a->c1);
  fp_add(&t0, &t0, &t1);
  fp_inv(&t0, &t0);
  fp_mul(&r->c0, &a->c0, &t0);
  fp_mul(&t1, &a->c1, &t0);
  fp_neg(&r->c1, &t1);
}

/* ---------------------------------------------------------------- Fp6 = Fp2[v]/(v^3 - xi) */
static inline void fp6_add(ofp6 *r, const ofp6 *a, const ofp6 *b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static inline void fp6_sub(ofp6 *r, const ofp6 *a, const ofp6 *b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static inline void fp6_neg(ofp6 *r, const ofp6 *a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static inline void fp6_mul_v(ofp6 *r, const ofp6 *a) {
  ofp2 t;
  fp2_mul_xi(&t, &a->c2);
  r->c2 = a->c1;
  r->c1 = a->c0;
  r->c0 = t;
}
static inline void fp6_mul(ofp6 *r, const ofp6 *a, const ofp6 *b) {
  ofp2 t0, t1, t2, s, u, c0, c1, c2;
  fp2_mul(&t0, &a->c0, &b->c0);
  fp2_mul(&t1, &a->c1, &b->c1);
  fp2_mul(&t2, &a->c2, &b->c2);
  fp2_add(&s, &a->c1, &a->c2); fp2_add(&u, &b->c1, &b->c2); fp2_mul(&c0, &s, &u);
  fp2_sub(&c0, &c0, &t1); fp2_sub(&c0, &c0, &t2); fp2_mul_xi(&c0, &c0); fp2_add(&c0, &c0, &t0);
  fp2_add(&s, &a->c0, &a->c1); fp2_add(&u, &b->c0, &b->c1); fp2_mul(&c1, &s, &u);
  fp2_sub(&c1, &c1, &t0); fp2_sub(&c1, &c1, &t1); fp2_mul_xi(&s, &t2); fp2_add(&c1, &c1, &s);
  fp2_add(&s, &a->c0, &a->c2); fp2_add(&u, &b->c0, &b->c2); fp2_mul(&c2, &s, &u);
  fp2_sub(&c2, &c2, &t0); fp2_sub(&c2, &c2, &t2); fp2_add(&c2, &c2, &t1);
  r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static inline void fp6_sqr(ofp6 *r, const ofp6 *a) { fp6_mul(r, a, a); }
/* a * (b0 + b1 v) */
static inline void fp6_mul_by_01(ofp6 *r, const ofp6 *a, const ofp2 *b0, const ofp2 *b1) {
  ofp2 aa, bb, s, t, c0, c1, c2;
  fp2_mul(&aa, &a->c0, b0);
  fp2_mul(&bb, &a->c1, b1);
  fp2_add(&s, &a->c1, &a->c2); fp2_mul(&c0, &s, b1); fp2_sub(&c0, &c0, &bb); fp2_mul_xi(&c0, &c0); fp2_add(&c0, &c0, &aa);
  fp2_add(&s, b0, b1); fp2_add(&t, &a->c0, &a->c1); fp2_mul(&c1, &s, &t); fp2_sub(&c1, &c1, &aa); fp2_sub(&c1, &c1, &bb);
  fp2_add(&s, &a->c0, &a->c2); fp2_mul(&c2, &s, b0); fp2_sub(&c2, &c2, &aa); fp2_add(&c2, &c2, &bb);
  r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
/* a * (b1 v) */
static inline void fp6_mul_by_1(ofp6 *r, const ofp6 *a, const ofp2 *b1) {
  ofp2 c0, c1, c2;
  fp2_mul(&c0, &a->c2, b1); fp2_mul_xi(&c0, &c0);
  fp2_mul(&c1, &a->c0, b1);
  fp2_mul(&c2, &a->c1, b1);
  r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static inline void fp6_inv(ofp6 *r, const ofp6 *a) {
  ofp2 t0, t1, t2, s, d;
  fp2_sqr(&t0, &a->c0); fp2_mul(&s, &a->c1, &a->c2); fp2_mul_xi(&s, &s); fp2_sub(&t0, &t0, &s);
  fp2_sqr(&t1, &a->c2); fp2_mul_xi(&t1, &t1); fp2_mul(&s, &a->c0, &a->c1); fp2_sub(&t1, &t1, &s);
  fp2_sqr(&t2, &a->c1); fp2_mul(&s, &a->c0, &a->c2); fp2_sub(&t2, &t2, &s);
  fp2_mul(&d, &a->c2, &t1); fp2_mul(&s, &a->c1, &t2); fp2_add(&d, &d, &s); fp2_mul_xi(&d, &d);
  fp2_mul(&s, &a->c0, &t0); fp2_add(&d, &d, &s);
  fp2_inv(&d, &d);
  fp2_mul(&r->c0, &t0, &d); fp2_mul(&r->c1, &t1, &d); fp2_mul(&r->c2, &t2, &d);
}

/* ---------------------------------------------------------------- Fp12 = Fp6[w]/(w^2 - v) */
static inline void fp12_set_one(ofp12 *r) { memset(r, 0, sizeof *r); r->c0.c0.c0 = OC_ONE; }
static inline int fp12_is_one(const ofp12 *a) {
  ofp12 one;
  fp12_set_one(&one);
  return memcmp(a, &one, sizeof one) == 0; /* limbs are canonical (fully reduced) */
}
static inline void fp12_mul(ofp12 *r, const ofp12 *a, const ofp12 *b) {
  ofp6 t0, t1, s, u, c1;
  fp6_mul(&t0, &a->c0, &b->c0);
  fp6_mul(&t1, &a->c1, &b->c1);
  fp6_add(&s, &a->c0, &a->c1); fp6_add(&u, &b->c0, &b->c1); fp6_mul(&c1, &s, &u);
  fp6_sub(&c1, &c1, &t0); fp6_sub(&c1, &c1, &t1);
  fp6_mul_v(&t1, &t1);
  fp6_add(&r->c0, &t0, &t1);
  r->c1 = c1;
}
/* complex squaring: 2 Fp6 products */
static inline void fp12_sqr(ofp12 *r, const ofp12 *a) {
  ofp6 ab, s, t, vab;
  fp6_mul(&ab, &a->c0, &a->c1);
  fp6_add(&s, &a->c0, &a->c1);
  fp6_mul_v(&t, &a->c1); fp6_add(&t, &t, &a->c0);
  fp6_mul(&s, &s, &t);
  fp6_mul_v(&vab, &ab);
  fp6_sub(&s, &s, &ab); fp6_sub(&r->c0, &s, &vab);
  fp6_add(&r->c1, &ab, &ab);
}
static inline void fp12_conj(ofp12 *r, const ofp12 *a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
static inline void fp12_inv(ofp12 *r, const ofp12 *a) {
  ofp6 t0, t1;
  fp6_sqr(&t0, &a->c0); fp6_sqr(&t1, &a->c1); fp6_mul_v(&t1, &t1); fp6_sub(&t0, &t0, &t1);
  fp6_inv(&t0, &t0);
  fp6_mul(&r->c0, &a->c0, &t0);
  fp6_mul(&t1, &a->c1, &t0); fp6_neg(&r->c1, &t1);
}
/* f * (l0 + l1 w^2 + l4 w^3): 13 Fp2 products */
static inline void fp12_mul_by_014(ofp12 *f, const ofp2 *l0, const ofp2 *l1, const ofp2 *l4) {
  ofp6 aa, bb, s, c1;
  ofp2 o;
  fp6_mul_by_01(&aa, &f->c0, l0, l1);
  fp6_mul_by_1(&bb, &f->c1, l4);
  fp2_add(&o, l1, l4);
  fp6_add(&s, &f->c0, &f->c1);
  fp6_mul_by_01(&c1, &s, l0, &o);
  fp6_sub(&c1, &c1, &aa); fp6_sub(&c1, &c1, &bb);
  fp6_mul_v(&bb, &bb);
  fp6_add(&f->c0, &aa, &bb);
  f->c1 = c1;
}
/* coefficient i of w^i, i = 0..5 */
static inline ofp2 *fp12_wcoef(ofp12 *a, int i) {
  ofp6 *h = (i & 1) ? &a->c1 : &a->c0;
  return (i >> 1) == 0 ? &h->c0 : ((i >> 1) == 1 ? &h->c1 : &h->c2);
}
static inline void fp12_frob(ofp12 *r, const ofp12 *a, int k) { /* k = 1 or 2 */
  ofp12 t = *a;
  for (int i = 0; i < 6; i++) {
    ofp2 *c = fp12_wcoef(&t, i);
    if (k == 1) { fp2_conj(c, c); fp2_mul(c, c, &OC_FROB1[i]); }
    else        { fp2_mul(c, c, &OC_FROB2[i]); }
  }
  *r = t;
}
/* Granger-Scott squaring in the cyclotomic subgroup */
static inline void fp4_sqr(ofp2 *r0, ofp2 *r1, const ofp2 *a, const ofp2 *b) {
  ofp2 t0, t1, t2;
  fp2_sqr(&t0, a); fp2_sqr(&t1, b);
  fp2_add(&t2, a, b); fp2_sqr(&t2, &t2); fp2_sub(&t2, &t2, &t0); fp2_sub(r1, &t2, &t1);
  fp2_mul_xi(&t1, &t1); fp2_add(r0, &t1, &t0);
}
static inline void fp12_cyclotomic_sqr(ofp12 *r, const ofp12 *f) {
  ofp2 z0 = f->c0.c0, z4 = f->c0.c1, z3 = f->c0.c2, z2 = f->c1.c0, z1 = f->c1.c1, z5 = f->c1.c2;
  ofp2 t0, t1, t2, t3, u0, u1;
  fp4_sqr(&t0, &t1, &z0, &z1);
  fp2_sub(&z0, &t0, &z0); fp2_dbl(&z0, &z0); fp2_add(&z0, &z0, &t0);
  fp2_add(&z1, &t1, &z1); fp2_dbl(&z1, &z1); fp2_add(&z1, &z1, &t1);
  fp4_sqr(&u0, &u1, &z2, &z3);
  fp4_sqr(&t2, &t3, &z4, &z5);
  fp2_sub(&z4, &u0, &z4); fp2_dbl(&z4, &z4); fp2_add(&z4, &z4, &u0);
  fp2_add(&z5, &u1, &z5); fp2_dbl(&z5, &z5); fp2_add(&z5, &z5, &u1);
  fp2_mul_xi(&t0, &t3);
  fp2_add(&z2, &t0, &z2); fp2_dbl(&z2, &z2); fp2_add(&z2, &z2, &t0);
  fp2_sub(&z3, &t2, &z3); fp2_dbl(&z3, &z3); fp2_add(&z3, &z3, &t2);
  r->c0.c0 = z0; r->c0.c1 = z4; r->c0.c2 = z3;
  r->c1.c0 = z2; r->c1.c1 = z1; r->c1.c2 = z5;
}
#endif
