// fp.cuh -- K1: BLS12-381 base field Fp and Fp2 = Fp[u]/(u^2+1) on 12 x 32-bit limbs.
//
// Replaces every blst_fp_* / blst_fp2 operation the reference reaches through blst
// (call sites /root/reference/src/eip2537.c:287,301,316 and everything
// under blst_p1_*/blst_p2_*/pairing).  Montgomery form, R = 2^384, values kept fully reduced
// in [0, p) so equality tests (needed by the exact exceptional-case handling in ec.cuh) are
// plain limb compares.
//
// Device path: CIOS Montgomery multiplication written as PTX mad.lo.cc/madc.hi.cc carry
// chains over an even/odd column split (ptxas fuses each lo/hi pair into one
// IMAD.WIDE.U32 with a predicate carry); add/sub are add.cc/sub.cc chains.
// Host path (B200_HD functions compiled by g++ for tests/host_emul only): the same
// algorithms in portable 64-bit C++.  The shipped library never runs the host path.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define B200_HD __host__ __device__ __forceinline__
#define B200_D __device__ __forceinline__
// out-of-line on the device: keeps code size (and ptxas time) bounded where a call's cost is
// noise next to the >= 3 Fp multiplications inside
#define B200_HD_NI __host__ __device__ __noinline__
#else
#define B200_HD inline
#define B200_D inline
#define B200_HD_NI inline
#endif

#include "constants.cuh"

namespace b200 {

struct alignas(16) Fp {
  uint32_t v[12];
};
struct Fp2 {
  Fp c0, c1;
};

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
B200_HD Fp fp_zero() {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = 0;
  return r;
}
B200_HD Fp fp_load_const(const uint32_t* c) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = c[i];
  return r;
}
B200_HD Fp fp_one() { return fp_load_const(C_ONE()); }
B200_HD bool is_zero(const Fp& a) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) acc |= a.v[i];
  return acc == 0;
}
B200_HD bool eq(const Fp& a, const Fp& b) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) acc |= a.v[i] ^ b.v[i];
  return acc == 0;
}

// r = a - p if a >= p else a        (a < 2p; `top` is an extra carry word of a)
B200_HD Fp fp_reduce_once(const Fp& a, uint32_t top) {
  Fp t;
  uint32_t borrow;
#if defined(__CUDA_ARCH__)
  // one asm statement per carry chain: the CC flag must not be separated from its consumers
  asm(
      "sub.cc.u32 %0, %13, %26;\n\t"
      "subc.cc.u32 %1, %14, %27;\n\t"
      "subc.cc.u32 %2, %15, %28;\n\t"
      "subc.cc.u32 %3, %16, %29;\n\t"
      "subc.cc.u32 %4, %17, %30;\n\t"
      "subc.cc.u32 %5, %18, %31;\n\t"
      "subc.cc.u32 %6, %19, %32;\n\t"
      "subc.cc.u32 %7, %20, %33;\n\t"
      "subc.cc.u32 %8, %21, %34;\n\t"
      "subc.cc.u32 %9, %22, %35;\n\t"
      "subc.cc.u32 %10, %23, %36;\n\t"
      "subc.cc.u32 %11, %24, %37;\n\t"
      "subc.u32 %12, %25, 0;"
      : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]), "=r"(t.v[8]), "=r"(t.v[9]), "=r"(t.v[10]), "=r"(t.v[11]), "=r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(top), "n"(B200_P0), "n"(B200_P1), "n"(B200_P2), "n"(B200_P3), "n"(B200_P4), "n"(B200_P5), "n"(B200_P6), "n"(B200_P7), "n"(B200_P8), "n"(B200_P9), "n"(B200_P10), "n"(B200_P11));
  bool keep_a = (borrow >> 31) != 0;                        // went negative => a < p
#else
  const uint32_t* p = C_P();
  uint64_t b = 0;
  for (int i = 0; i < 12; i++) {
    uint64_t d = (uint64_t)a.v[i] - p[i] - b;
    t.v[i] = (uint32_t)d;
    b = (d >> 32) & 1;
  }
  borrow = (uint32_t)b;
  bool keep_a = (borrow != 0) && (top == 0);
#endif
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = keep_a ? a.v[i] : t.v[i];
  return r;
}

B200_HD Fp add(const Fp& a, const Fp& b) {
  Fp t;
#if defined(__CUDA_ARCH__)
  asm(
      "add.cc.u32 %0, %12, %24;\n\t"
      "addc.cc.u32 %1, %13, %25;\n\t"
      "addc.cc.u32 %2, %14, %26;\n\t"
      "addc.cc.u32 %3, %15, %27;\n\t"
      "addc.cc.u32 %4, %16, %28;\n\t"
      "addc.cc.u32 %5, %17, %29;\n\t"
      "addc.cc.u32 %6, %18, %30;\n\t"
      "addc.cc.u32 %7, %19, %31;\n\t"
      "addc.cc.u32 %8, %20, %32;\n\t"
      "addc.cc.u32 %9, %21, %33;\n\t"
      "addc.cc.u32 %10, %22, %34;\n\t"
      "addc.u32 %11, %23, %35;"
      : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]), "=r"(t.v[8]), "=r"(t.v[9]), "=r"(t.v[10]), "=r"(t.v[11])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
  // p < 2^381: a + b < 2^382 never carries out of 12 limbs
#else
  uint64_t c = 0;
  for (int i = 0; i < 12; i++) {
    uint64_t s = (uint64_t)a.v[i] + b.v[i] + c;
    t.v[i] = (uint32_t)s;
    c = s >> 32;
  }
#endif
  return fp_reduce_once(t, 0);
}

B200_HD Fp sub(const Fp& a, const Fp& b) {
  Fp t;
#if defined(__CUDA_ARCH__)
  uint32_t mask;
  asm(
      "sub.cc.u32 %0, %13, %25;\n\t"
      "subc.cc.u32 %1, %14, %26;\n\t"
      "subc.cc.u32 %2, %15, %27;\n\t"
      "subc.cc.u32 %3, %16, %28;\n\t"
      "subc.cc.u32 %4, %17, %29;\n\t"
      "subc.cc.u32 %5, %18, %30;\n\t"
      "subc.cc.u32 %6, %19, %31;\n\t"
      "subc.cc.u32 %7, %20, %32;\n\t"
      "subc.cc.u32 %8, %21, %33;\n\t"
      "subc.cc.u32 %9, %22, %34;\n\t"
      "subc.cc.u32 %10, %23, %35;\n\t"
      "subc.cc.u32 %11, %24, %36;\n\t"
      "subc.u32 %12, 0, 0;"
      : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]), "=r"(t.v[8]), "=r"(t.v[9]), "=r"(t.v[10]), "=r"(t.v[11]), "=r"(mask)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
  asm(
      "add.cc.u32 %0, %0, %12;\n\t"
      "addc.cc.u32 %1, %1, %13;\n\t"
      "addc.cc.u32 %2, %2, %14;\n\t"
      "addc.cc.u32 %3, %3, %15;\n\t"
      "addc.cc.u32 %4, %4, %16;\n\t"
      "addc.cc.u32 %5, %5, %17;\n\t"
      "addc.cc.u32 %6, %6, %18;\n\t"
      "addc.cc.u32 %7, %7, %19;\n\t"
      "addc.cc.u32 %8, %8, %20;\n\t"
      "addc.cc.u32 %9, %9, %21;\n\t"
      "addc.cc.u32 %10, %10, %22;\n\t"
      "addc.u32 %11, %11, %23;"
      : "+r"(t.v[0]), "+r"(t.v[1]), "+r"(t.v[2]), "+r"(t.v[3]), "+r"(t.v[4]), "+r"(t.v[5]), "+r"(t.v[6]), "+r"(t.v[7]), "+r"(t.v[8]), "+r"(t.v[9]), "+r"(t.v[10]), "+r"(t.v[11])
      : "r"(B200_P0 & mask), "r"(B200_P1 & mask), "r"(B200_P2 & mask), "r"(B200_P3 & mask), "r"(B200_P4 & mask), "r"(B200_P5 & mask), "r"(B200_P6 & mask), "r"(B200_P7 & mask), "r"(B200_P8 & mask), "r"(B200_P9 & mask), "r"(B200_P10 & mask), "r"(B200_P11 & mask));
#else
  uint64_t bw = 0;
  for (int i = 0; i < 12; i++) {
    uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw;
    t.v[i] = (uint32_t)d;
    bw = (d >> 32) & 1;
  }
  if (bw) {
    const uint32_t* p = C_P();
    uint64_t c = 0;
    for (int i = 0; i < 12; i++) {
      uint64_t s = (uint64_t)t.v[i] + p[i] + c;
      t.v[i] = (uint32_t)s;
      c = s >> 32;
    }
  }
#endif
  return t;
}
B200_HD Fp neg(const Fp& a) { return sub(fp_zero(), a); }
B200_HD Fp dbl(const Fp& a) { return add(a, a); }
// a/2 mod p: add p when odd, shift right (works in Montgomery form as well: halving commutes with *R)
B200_HD Fp half(const Fp& a) {
  const uint32_t* p = C_P();
  const uint32_t mask = 0u - (a.v[0] & 1u);
  uint32_t t[13];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t s = (uint64_t)a.v[i] + (p[i] & mask) + c;
    t[i] = (uint32_t)s;
    c = s >> 32;
  }
  t[12] = (uint32_t)c;
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return r;
}

// ------------------------------------------------------------------------------------------
// Montgomery multiplication  r = a*b/2^384 mod p
// ------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && !defined(B200_FP_PORTABLE)
// Column-split CIOS.  Products a[j]*bi are 64-bit; products with even j land on word pairs
// (j, j+1) and chain without overlap, products with odd j likewise one word higher.  Two
// accumulators (`e` aligned to even columns, `o` to odd columns) therefore take pure
// lo/hi carry chains, which ptxas maps to IMAD.WIDE.U32 with predicate carries.  After the
// reduction row the window slides one word, which swaps the roles of the two accumulators.
namespace detail {
// acc[0..12) = a[0,2,..,10] * b  (fresh accumulator: plain wide products, no carries)
B200_D void mul_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(acc[j]) : "r"(a[j]), "r"(b));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(acc[j + 1]) : "r"(a[j]), "r"(b));
  }
}
// acc[0..13) += a[0,2,..,10] * b  -- one carry chain, one asm statement
B200_D void mad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %13, %19, %0;\n\t"
      "madc.hi.cc.u32 %1, %13, %19, %1;\n\t"
      "madc.lo.cc.u32 %2, %14, %19, %2;\n\t"
      "madc.hi.cc.u32 %3, %14, %19, %3;\n\t"
      "madc.lo.cc.u32 %4, %15, %19, %4;\n\t"
      "madc.hi.cc.u32 %5, %15, %19, %5;\n\t"
      "madc.lo.cc.u32 %6, %16, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %16, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %17, %19, %8;\n\t"
      "madc.hi.cc.u32 %9, %17, %19, %9;\n\t"
      "madc.lo.cc.u32 %10, %18, %19, %10;\n\t"
      "madc.hi.cc.u32 %11, %18, %19, %11;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(acc[12])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(b));
}
// acc[0..13) += {P0,P2,..,P10} * m   (modulus limbs as immediates)
#define B200_MADP_ROW(acc, m, Q0, Q1, Q2, Q3, Q4, Q5)                                           \
  asm("mad.lo.cc.u32 %0, %14, %13, %0;\n\t"                                                     \
      "madc.hi.cc.u32 %1, %14, %13, %1;\n\t"                                                    \
      "madc.lo.cc.u32 %2, %15, %13, %2;\n\t"                                                    \
      "madc.hi.cc.u32 %3, %15, %13, %3;\n\t"                                                    \
      "madc.lo.cc.u32 %4, %16, %13, %4;\n\t"                                                    \
      "madc.hi.cc.u32 %5, %16, %13, %5;\n\t"                                                    \
      "madc.lo.cc.u32 %6, %17, %13, %6;\n\t"                                                    \
      "madc.hi.cc.u32 %7, %17, %13, %7;\n\t"                                                    \
      "madc.lo.cc.u32 %8, %18, %13, %8;\n\t"                                                    \
      "madc.hi.cc.u32 %9, %18, %13, %9;\n\t"                                                    \
      "madc.lo.cc.u32 %10, %19, %13, %10;\n\t"                                                  \
      "madc.hi.cc.u32 %11, %19, %13, %11;\n\t"                                                  \
      "addc.u32 %12, %12, 0;"                                                                    \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),      \
        "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]),    \
        "+r"(acc[12])                                                                            \
      : "r"(m), "n"(Q0), "n"(Q1), "n"(Q2), "n"(Q3), "n"(Q4), "n"(Q5))
// dst[0..12) += src[1..13), carry into dst[12]
B200_D void fold(uint32_t* dst, const uint32_t* src) {
  asm("add.cc.u32 %0, %0, %13;\n\t"
      "addc.cc.u32 %1, %1, %14;\n\t"
      "addc.cc.u32 %2, %2, %15;\n\t"
      "addc.cc.u32 %3, %3, %16;\n\t"
      "addc.cc.u32 %4, %4, %17;\n\t"
      "addc.cc.u32 %5, %5, %18;\n\t"
      "addc.cc.u32 %6, %6, %19;\n\t"
      "addc.cc.u32 %7, %7, %20;\n\t"
      "addc.cc.u32 %8, %8, %21;\n\t"
      "addc.cc.u32 %9, %9, %22;\n\t"
      "addc.cc.u32 %10, %10, %23;\n\t"
      "addc.cc.u32 %11, %11, %24;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(dst[0]), "+r"(dst[1]), "+r"(dst[2]), "+r"(dst[3]), "+r"(dst[4]), "+r"(dst[5]), "+r"(dst[6]),
        "+r"(dst[7]), "+r"(dst[8]), "+r"(dst[9]), "+r"(dst[10]), "+r"(dst[11]), "+r"(dst[12])
      : "r"(src[1]), "r"(src[2]), "r"(src[3]), "r"(src[4]), "r"(src[5]), "r"(src[6]), "r"(src[7]),
        "r"(src[8]), "r"(src[9]), "r"(src[10]), "r"(src[11]), "r"(src[12]));
}
}  // namespace detail

// The 12 rows are processed two per iteration of a ROLLED loop (the multiplier limbs shift down
// by two each time), so one multiplication is ~150 instructions of code instead of ~340: the
// bucket-accumulate kernel inlines ten of them per point addition and was instruction-fetch bound
// (ncu: stall_no_instruction dominant, 150 KB kernel) with the fully unrolled form.
B200_D Fp mul(const Fp& a, const Fp& b_in) {
  using namespace detail;
  // x: accumulator aligned to the current lowest column, 13 words; y: one column higher
  uint32_t x[13], y[13], b[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { b[i] = b_in.v[i]; x[i] = 0; }
  x[12] = 0;
#pragma unroll 1
  for (int it = 0; it < 6; it++) {
    uint32_t m;
    // ---- even row: x is the low-aligned accumulator, y is fresh
    mad_row(x, a.v, b[0]);
    mul_row(y, a.v + 1, b[0]); y[12] = 0;
    m = x[0] * B200_M0;
    B200_MADP_ROW(x, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
    B200_MADP_ROW(y, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
    fold(y, x);          // slide one column: y is now the low-aligned accumulator
    // ---- odd row: roles swapped
    mad_row(y, a.v, b[1]);
    mul_row(x, a.v + 1, b[1]); x[12] = 0;
    m = y[0] * B200_M0;
    B200_MADP_ROW(y, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
    B200_MADP_ROW(x, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
    fold(x, y);
#pragma unroll
    for (int i = 0; i < 10; i++) b[i] = b[i + 2];
  }
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = x[i];
  return fp_reduce_once(r, x[12]);
}
// ---- experimental variants of `mul` (bls12_b200_fp_microbench modes 6..8; profiles/r02_k1_probes.md) -----------------
// The rolled loop above shifts the 12 multiplier limbs down by two after every iteration: 60 register moves per
// multiplication, which ptxas emits as IMAD.MOV.U32 -- and those occupy the SAME multiply pipe as the 300 wide products.
// Variant A: three iterations of four rows (shift by four: 16 moves).  Variant B: two iterations of six rows (6 moves).
// Variant C: the multiplier limbs are parked in shared memory and fetched by index (LSU pipe, no moves).
// register move that ptxas must leave on the ALU pipe (a byte permute with the identity selector), for the shifts
B200_D uint32_t alu_move(uint32_t v) {
  uint32_t r;
  asm volatile("prmt.b32 %0, %1, 0, 0x3210;" : "=r"(r) : "r"(v));
  return r;
}
template <int ROWS_PER_ITER, bool ALU_SHIFT = false>
B200_D Fp mul_unrolled(const Fp& a, const Fp& b_in) {
  using namespace detail;
  static_assert(ROWS_PER_ITER == 2 || ROWS_PER_ITER == 4 || ROWS_PER_ITER == 6 || ROWS_PER_ITER == 12, "rows per iteration");
  uint32_t x[13], y[13], b[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { b[i] = b_in.v[i]; x[i] = 0; }
  x[12] = 0;
#pragma unroll 1
  for (int it = 0; it < 12 / ROWS_PER_ITER; it++) {
#pragma unroll
    for (int r = 0; r < ROWS_PER_ITER; r += 2) {
      uint32_t m;
      mad_row(x, a.v, b[r]);
      mul_row(y, a.v + 1, b[r]); y[12] = 0;
      m = x[0] * B200_M0;
      B200_MADP_ROW(x, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
      B200_MADP_ROW(y, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
      fold(y, x);
      mad_row(y, a.v, b[r + 1]);
      mul_row(x, a.v + 1, b[r + 1]); x[12] = 0;
      m = y[0] * B200_M0;
      B200_MADP_ROW(y, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
      B200_MADP_ROW(x, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
      fold(x, y);
    }
    if (ROWS_PER_ITER < 12) {
#pragma unroll
      for (int i = 0; i < 12 - ROWS_PER_ITER; i++) b[i] = ALU_SHIFT ? alu_move(b[i + ROWS_PER_ITER]) : b[i + ROWS_PER_ITER];
    }
  }
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = x[i];
  return fp_reduce_once(r, x[12]);
}
#ifdef __CUDACC__
// multiplier limbs in shared memory: slot = 12 words, word i of thread t at slot[i * stride]
B200_D Fp mul_bsmem(const Fp& a, const Fp& b_in, uint32_t* slot, int stride) {
  using namespace detail;
  uint32_t x[13], y[13];
#pragma unroll
  for (int i = 0; i < 12; i++) { slot[i * stride] = b_in.v[i]; x[i] = 0; }
  x[12] = 0;
  uint32_t b0 = slot[0], b1 = slot[stride];
#pragma unroll 1
  for (int it = 0; it < 6; it++) {
    uint32_t m;
    const uint32_t n0 = slot[((2 * it + 2) % 12) * stride], n1 = slot[((2 * it + 3) % 12) * stride];   // next pair, in flight
    mad_row(x, a.v, b0);
    mul_row(y, a.v + 1, b0); y[12] = 0;
    m = x[0] * B200_M0;
    B200_MADP_ROW(x, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
    B200_MADP_ROW(y, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
    fold(y, x);
    mad_row(y, a.v, b1);
    mul_row(x, a.v + 1, b1); x[12] = 0;
    m = y[0] * B200_M0;
    B200_MADP_ROW(y, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
    B200_MADP_ROW(x, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
    fold(x, y);
    b0 = n0; b1 = n1;
  }
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = x[i];
  return fp_reduce_once(r, x[12]);
}
#endif

// a*b - c*d with ONE Montgomery reduction: the rows of both products are accumulated before each reduction
// row, so the pair costs 2*144 + 156 multiply-accumulates instead of 2*300.  The point formulas end in such a
// difference (Y3 = R*(Q - X3) - Y1*PPP).  c is negated first, so the accumulator holds a*b + (p-c)*d; it stays
// below 3p (T' <= (T + (2^32-1)(3p-2)) / 2^32 <= 3p-2 whenever T <= 3p-2), hence two conditional subtractions.
B200_D Fp mul_sum(const Fp& a, const Fp& b_in, const Fp& c, const Fp& d_in) {   // a*b + c*d, one reduction
  using namespace detail;
  uint32_t x[13], y[13], b[12], d[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { b[i] = b_in.v[i]; d[i] = d_in.v[i]; x[i] = 0; }
  x[12] = 0;
#pragma unroll 1
  for (int it = 0; it < 6; it++) {
    uint32_t m;
    mad_row(x, a.v, b[0]);
    mul_row(y, a.v + 1, b[0]); y[12] = 0;
    mad_row(x, c.v, d[0]);
    mad_row(y, c.v + 1, d[0]);
    m = x[0] * B200_M0;
    B200_MADP_ROW(x, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
    B200_MADP_ROW(y, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
    fold(y, x);
    mad_row(y, a.v, b[1]);
    mul_row(x, a.v + 1, b[1]); x[12] = 0;
    mad_row(y, c.v, d[1]);
    mad_row(x, c.v + 1, d[1]);
    m = y[0] * B200_M0;
    B200_MADP_ROW(y, m, B200_P0, B200_P2, B200_P4, B200_P6, B200_P8, B200_P10);
    B200_MADP_ROW(x, m, B200_P1, B200_P3, B200_P5, B200_P7, B200_P9, B200_P11);
    fold(x, y);
#pragma unroll
    for (int i = 0; i < 10; i++) { b[i] = b[i + 2]; d[i] = d[i + 2]; }
  }
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = x[i];
  return fp_reduce_once(fp_reduce_once(r, x[12]), 0);
}
B200_D Fp mul_diff(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return mul_sum(a, b, sub(fp_zero(), c), d); }
#else
B200_HD Fp mul(const Fp& a, const Fp& b);
B200_HD Fp mul_diff(const Fp& a, const Fp& b, const Fp& c, const Fp& d);
B200_HD Fp mul_sum(const Fp& a, const Fp& b, const Fp& c, const Fp& d);
#endif
#ifdef B200_COUNT_MULS
extern unsigned long long g_fp_mul_count;   // tests/host_emul only: counts Fp multiplications per phase
#endif
B200_HD Fp mul_portable(const Fp& a, const Fp& b) {
#ifdef B200_COUNT_MULS
  g_fp_mul_count++;
#endif
  const uint32_t* p = C_P();
  uint32_t t[14];
#pragma unroll
  for (int i = 0; i < 14; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t carry = 0, acc;
#pragma unroll
    for (int j = 0; j < 12; j++) {
      acc = (uint64_t)a.v[j] * b.v[i] + t[j] + carry;
      t[j] = (uint32_t)acc;
      carry = acc >> 32;
    }
    acc = (uint64_t)t[12] + carry;
    t[12] = (uint32_t)acc;
    t[13] = (uint32_t)(acc >> 32);
    uint32_t m = t[0] * B200_M0;
    acc = (uint64_t)m * p[0] + t[0];
    carry = acc >> 32;
#pragma unroll
    for (int j = 1; j < 12; j++) {
      acc = (uint64_t)m * p[j] + t[j] + carry;
      t[j - 1] = (uint32_t)acc;
      carry = acc >> 32;
    }
    acc = (uint64_t)t[12] + carry;
    t[11] = (uint32_t)acc;
    t[12] = t[13] + (uint32_t)(acc >> 32);
  }
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = t[i];
  return fp_reduce_once(r, t[12]);
}
#if !(defined(__CUDA_ARCH__) && !defined(B200_FP_PORTABLE))
B200_HD Fp mul(const Fp& a, const Fp& b) { return mul_portable(a, b); }
B200_HD Fp mul_diff(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return sub(mul_portable(a, b), mul_portable(c, d)); }
B200_HD Fp mul_sum(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return add(mul_portable(a, b), mul_portable(c, d)); }
#endif

B200_HD Fp sqr(const Fp& a) { return mul(a, a); }

B200_HD Fp fp_to_mont(const Fp& a) { return mul(a, fp_load_const(C_RR())); }
B200_HD Fp fp_from_mont(const Fp& a) {
  Fp one = fp_zero();
  one.v[0] = 1;
  return mul(a, one);
}

// Inverse by the binary extended Euclidean algorithm (~760 shift/subtract steps on 12 limbs, an
// order of magnitude cheaper than the a^(p-2) ladder of ~480 dependent multiplications).
// inverse of 0 is 0 (the convention the reference relies on in blst_p1_to_affine).
// Input/output in Montgomery form: binary_inv(aR) = a^-1 R^-1, then one multiplication by R^3.
namespace detail {
B200_HD bool limbs_is_one(const Fp& a) {
  uint32_t acc = a.v[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < 12; i++) acc |= a.v[i];
  return acc == 0;
}
B200_HD void limbs_shr1(Fp& a, uint32_t top) {
#pragma unroll
  for (int i = 0; i < 11; i++) a.v[i] = (a.v[i] >> 1) | (a.v[i + 1] << 31);
  a.v[11] = (a.v[11] >> 1) | (top << 31);
}
// x = x/2 mod p  (x < p)
B200_HD void halve_mod_p(Fp& x) {
  const uint32_t* p = C_P();
  uint32_t carry = 0;
  if (x.v[0] & 1) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { uint64_t s = (uint64_t)x.v[i] + p[i] + c; x.v[i] = (uint32_t)s; c = s >> 32; }
    carry = (uint32_t)c;
  }
  limbs_shr1(x, carry);
}
// a >= b as 384-bit integers
B200_HD bool limbs_geq(const Fp& a, const Fp& b) {
  uint64_t bw = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw; bw = (d >> 32) & 1; }
  return bw == 0;
}
B200_HD void limbs_sub(Fp& a, const Fp& b) {   // a -= b, a >= b
  uint64_t bw = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw; a.v[i] = (uint32_t)d; bw = (d >> 32) & 1; }
}
}  // namespace detail

B200_HD_NI Fp inv(const Fp& a) {
  using namespace detail;
  if (is_zero(a)) return fp_zero();
  Fp u = a, v = fp_load_const(C_P()), x1 = fp_zero(), x2 = fp_zero();
  x1.v[0] = 1;
  while (!limbs_is_one(u) && !limbs_is_one(v)) {
    while (!(u.v[0] & 1)) { limbs_shr1(u, 0); halve_mod_p(x1); }
    while (!(v.v[0] & 1)) { limbs_shr1(v, 0); halve_mod_p(x2); }
    if (limbs_geq(u, v)) { limbs_sub(u, v); x1 = sub(x1, x2); }
    else                 { limbs_sub(v, u); x2 = sub(x2, x1); }
  }
  Fp r = limbs_is_one(u) ? x1 : x2;
  return mul(r, fp_load_const(C_R3()));
}

// ------------------------------------------------------------------------------------------
// Fp2
// ------------------------------------------------------------------------------------------
B200_HD Fp2 fp2_zero() { Fp2 r; r.c0 = fp_zero(); r.c1 = fp_zero(); return r; }
B200_HD Fp2 fp2_one() { Fp2 r; r.c0 = fp_one(); r.c1 = fp_zero(); return r; }
B200_HD Fp2 fp2_load_const(const uint32_t* c) { Fp2 r; r.c0 = fp_load_const(c); r.c1 = fp_load_const(c + 12); return r; }
B200_HD bool is_zero(const Fp2& a) { return is_zero(a.c0) && is_zero(a.c1); }
B200_HD bool eq(const Fp2& a, const Fp2& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
B200_HD Fp2 add(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); return r; }
B200_HD Fp2 sub(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); return r; }
B200_HD Fp2 neg(const Fp2& a) { Fp2 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
B200_HD Fp2 dbl(const Fp2& a) { return add(a, a); }
B200_HD Fp2 conj(const Fp2& a) { Fp2 r; r.c0 = a.c0; r.c1 = neg(a.c1); return r; }
B200_HD_NI Fp2 mul(const Fp2& a, const Fp2& b) {
#if defined(__CUDA_ARCH__) && !defined(B200_FP_PORTABLE)
  // device: schoolbook with one Montgomery reduction per component, 2 x (288 + 156) multiply-accumulates and one
  // negation -- against Karatsuba's 3 x 300 plus five additions; measured +3.6 % on the G2 MSM and +4.9 % on the
  // pairing batch.  (The host build keeps Karatsuba: its multiplication counter defines the algorithmic work.)
  Fp2 r;
  r.c0 = mul_diff(a.c0, b.c0, a.c1, b.c1);
  r.c1 = mul_sum(a.c0, b.c1, a.c1, b.c0);
  return r;
#else
  Fp t0 = mul(a.c0, b.c0), t1 = mul(a.c1, b.c1);
  Fp t2 = mul(add(a.c0, a.c1), add(b.c0, b.c1));
  Fp2 r;
  r.c0 = sub(t0, t1);
  r.c1 = sub(sub(t2, t0), t1);
  return r;
#endif
}
B200_HD_NI Fp2 sqr(const Fp2& a) {
  Fp2 r;
  Fp m = mul(a.c0, a.c1);
  r.c0 = mul(add(a.c0, a.c1), sub(a.c0, a.c1));
  r.c1 = dbl(m);
  return r;
}
// Karatsuba cross terms with the operand sums formed INSIDE the out-of-line function: the callers
// (Fp6 / Fp12 code) otherwise have to park every sum in thread-local memory just to pass it by reference
B200_HD_NI Fp2 mul_sum2(const Fp2& a1, const Fp2& a2, const Fp2& b1, const Fp2& b2) { return mul(add(a1, a2), add(b1, b2)); }
B200_HD_NI Fp2 mul_sum1(const Fp2& a1, const Fp2& a2, const Fp2& b) { return mul(add(a1, a2), b); }
B200_HD_NI Fp2 sqr_sum(const Fp2& a1, const Fp2& a2) { return sqr(add(a1, a2)); }
B200_HD Fp2 mul_diff(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return sub(mul(a, b), mul(c, d)); }
B200_HD_NI Fp2 mul_fp(const Fp2& a, const Fp& k) { Fp2 r; r.c0 = mul(a.c0, k); r.c1 = mul(a.c1, k); return r; }
B200_HD Fp2 half(const Fp2& a) { Fp2 r; r.c0 = half(a.c0); r.c1 = half(a.c1); return r; }
B200_HD Fp2 mul_xi(const Fp2& a) { Fp2 r; r.c0 = sub(a.c0, a.c1); r.c1 = add(a.c0, a.c1); return r; }
B200_HD Fp2 inv(const Fp2& a) {
  Fp t = inv(add(sqr(a.c0), sqr(a.c1)));
  Fp2 r;
  r.c0 = mul(a.c0, t);
  r.c1 = neg(mul(a.c1, t));
  return r;
}

// field "traits" used by the field-generic point code
template <class F> struct FieldOps;
template <> struct FieldOps<Fp> {
  static B200_HD Fp zero() { return fp_zero(); }
  static B200_HD Fp one() { return fp_one(); }
  static B200_HD Fp curve_b() { return fp_load_const(C_B1()); }
  static constexpr int WORDS = 12;
};
template <> struct FieldOps<Fp2> {
  static B200_HD Fp2 zero() { return fp2_zero(); }
  static B200_HD Fp2 one() { return fp2_one(); }
  static B200_HD Fp2 curve_b() { return fp2_load_const(C_B2()); }
  static constexpr int WORDS = 24;
};

}  // namespace b200
