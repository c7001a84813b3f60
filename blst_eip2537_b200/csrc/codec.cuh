// codec.cuh -- K3 building blocks: EIP-2537 wire format <-> Montgomery limbs.
//
// Replaces fp_from_bytes (/root/reference/src/eip2537.c:263-309), fp_to_bytes (:312-317),
// decode_g1_point (:320-343), decode_g2_point (:381-404), encode_g{1,2}_point (:346-350,
// :407-411) and decode_scalar (:417-420).  Accept/reject rules are the reference's:
//   * bytes [0,16) of every 64-byte field slot must be zero, value must be < p, else
//     INVALID_ELEMENT (3) -- both coordinates are always examined (:322-328)
//   * all-zero point = infinity, accepted (:331-333)
//   * otherwise y^2 = x^3 + b must hold, else POINT_NOT_ON_CURVE (1)
//   * NO subgroup check here (:340, :401); PAIRING adds it separately (:1041, :1051)
//   * scalars are raw 256-bit big-endian integers, never reduced (:417-420)
#pragma once
#include "ec.cuh"

namespace b200 {

enum : int {
  E_SUCCESS = 0, E_NOT_ON_CURVE = 1, E_NOT_IN_SUBGROUP = 2, E_INVALID_ELEMENT = 3,
  E_ENCODING = 4, E_INVALID_LENGTH = 5, E_EMPTY_INPUT = 6, E_MEMORY = 7
};

B200_HD uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __byte_perm(x, 0, 0x0123);
#else
  return __builtin_bswap32(x);
#endif
}

// w[0..16): the 64-byte slot loaded as little-endian 32-bit words.
// returns -1 invalid, 0 zero, 1 non-zero; `out` in Montgomery form when valid
B200_HD int fp_from_slot(Fp& out, const uint32_t* w) {
  if ((w[0] | w[1] | w[2] | w[3]) != 0) return -1;
  Fp t;
  uint32_t nz = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    t.v[11 - k] = bswap32(w[4 + k]);
    nz |= w[4 + k];
  }
  // t < p  <=>  t - p borrows
  const uint32_t* p = C_P();
  uint32_t borrow = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t d = (uint64_t)t.v[i] - p[i] - borrow;
    borrow = (uint32_t)(d >> 32) & 1;
  }
  if (!borrow) return -1;
  out = fp_to_mont(t);
  return nz ? 1 : 0;
}

// Montgomery -> 16 little-endian words of the 64-byte big-endian slot
B200_HD void fp_to_slot(uint32_t* w, const Fp& a) {
  Fp c = fp_from_mont(a);
  w[0] = w[1] = w[2] = w[3] = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) w[4 + k] = bswap32(c.v[11 - k]);
}

// point decoders: `w` = the 128 / 256 byte encoding as LE words. Return an EIP2537 code.
B200_HD int decode_point(G1Affine& out, const uint32_t* w) {
  int sx = fp_from_slot(out.x, w), sy = fp_from_slot(out.y, w + 16);
  if (sx < 0 || sy < 0) { out.x = fp_zero(); out.y = fp_zero(); return E_INVALID_ELEMENT; }
  if (sx == 0 && sy == 0) return E_SUCCESS;
  if (!affine_on_curve(out)) { out.x = fp_zero(); out.y = fp_zero(); return E_NOT_ON_CURVE; }
  return E_SUCCESS;
}
B200_HD int decode_point(G2Affine& out, const uint32_t* w) {
  int s0 = fp_from_slot(out.x.c0, w), s1 = fp_from_slot(out.x.c1, w + 16);
  int s2 = fp_from_slot(out.y.c0, w + 32), s3 = fp_from_slot(out.y.c1, w + 48);
  if (s0 < 0 || s1 < 0 || s2 < 0 || s3 < 0) { out.x = fp2_zero(); out.y = fp2_zero(); return E_INVALID_ELEMENT; }
  if ((s0 | s1) == 0 && (s2 | s3) == 0) return E_SUCCESS;
  if (!affine_on_curve(out)) { out.x = fp2_zero(); out.y = fp2_zero(); return E_NOT_ON_CURVE; }
  return E_SUCCESS;
}
B200_HD void encode_point(uint32_t* w, const G1Affine& a) {
  fp_to_slot(w, a.x);
  fp_to_slot(w + 16, a.y);
}
B200_HD void encode_point(uint32_t* w, const G2Affine& a) {
  fp_to_slot(w, a.x.c0); fp_to_slot(w + 16, a.x.c1);
  fp_to_slot(w + 32, a.y.c0); fp_to_slot(w + 48, a.y.c1);
}

// 32 big-endian bytes (as 8 LE-loaded words) -> 8 little-endian 32-bit limbs
B200_HD void scalar_from_slot(uint32_t* k, const uint32_t* w) {
#pragma unroll
  for (int i = 0; i < 8; i++) k[7 - i] = bswap32(w[i]);
}

template <class F> struct Wire;
template <> struct Wire<Fp>  { static constexpr int POINT_WORDS = 32, PAIR_WORDS = 40; };   // 128 B, 160 B
template <> struct Wire<Fp2> { static constexpr int POINT_WORDS = 64, PAIR_WORDS = 72; };   // 256 B, 288 B

}  // namespace b200
