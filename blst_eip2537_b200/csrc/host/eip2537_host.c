/*
 * eip2537_host.c -- plain-C host layer: the 13 bls12_* entry points and the gas schedule of
 * /root/reference/src/eip2537.h, with the reference's argument checks and error conventions,
 * delegating all arithmetic to the CUDA engine (engine.h).  There is no CPU arithmetic here
 * and no fallback: without a working CUDA device every call returns EIP2537_MEMORY_ERROR.
 *
 * Mirrors (does not copy) the control flow of /root/reference/src/eip2537.c:
 *   length checks           :436, :489, :543, :724, :777, :831, :1022
 *   MULTIEXP dispatch       :541-561, :829-849  (k=1 -> MUL; the naive / Bos-Coster variants
 *                           all produce the same bytes, so all three names share the GPU MSM)
 *   `out` untouched on error (:613, :701, :1072-1078 write last)
 */
#include "../../../include/eip2537.h"
#include "../engine.h"

EIP2537_ERROR bls12_g1add(byte out[128], const byte in[256], size_t in_len) {
  if (in_len != 256) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_add_host(1, in, out);
}
EIP2537_ERROR bls12_g2add(byte out[256], const byte in[512], size_t in_len) {
  if (in_len != 512) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_add_host(2, in, out);
}
EIP2537_ERROR bls12_g1mul(byte out[128], const byte in[160], size_t in_len) {
  if (in_len != 160) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_msm_host(1, in, 1, out);
}
EIP2537_ERROR bls12_g2mul(byte out[256], const byte in[288], size_t in_len) {
  if (in_len != 288) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_msm_host(2, in, 1, out);
}

static EIP2537_ERROR multiexp(int group, size_t stride, byte* out, const byte* in, size_t in_len) {
  if (in_len == 0 || (in_len % stride) != 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_msm_host(group, in, in_len / stride, out);
}
EIP2537_ERROR bls12_g1multiexp(byte out[128], byte* in, size_t in_len) { return multiexp(1, 160, out, in, in_len); }
EIP2537_ERROR bls12_g1multiexp_naive(byte out[128], byte* in, size_t in_len) { return multiexp(1, 160, out, in, in_len); }
EIP2537_ERROR bls12_g1multiexp_bc(byte out[128], byte* in, size_t in_len) { return multiexp(1, 160, out, in, in_len); }
EIP2537_ERROR bls12_g2multiexp(byte out[256], byte* in, size_t in_len) { return multiexp(2, 288, out, in, in_len); }
EIP2537_ERROR bls12_g2multiexp_naive(byte out[256], byte* in, size_t in_len) { return multiexp(2, 288, out, in, in_len); }
EIP2537_ERROR bls12_g2multiexp_bc(byte out[256], byte* in, size_t in_len) { return multiexp(2, 288, out, in, in_len); }

EIP2537_ERROR bls12_pairing(byte out[32], byte* in, size_t in_len) {
  if (in_len == 0 || (in_len % 384) != 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_pairing_host(in, in_len / 384, out);
}

/* MAP_FP_TO_G1 / MAP_FP2_TO_G2 (src/eip2537.c:1093-1121, :1135-1163): length check, then the GPU map
 * (csrc/map.cuh: RFC 9380 SSWU + isogeny + cofactor clearing); an invalid field element gives 3. */
EIP2537_ERROR bls12_map_fp_to_g1(byte out[128], const byte in[64], size_t in_len) {
  if (in_len != 64) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_map_host(1, in, out);
}
EIP2537_ERROR bls12_map_fp2_to_g2(byte out[256], const byte in[128], size_t in_len) {
  if (in_len != 128) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)b200_map_host(2, in, out);
}

/* ---- gas schedule (src/eip2537.c:1169-1271): EIP-2537 constants, integer arithmetic only */
const uint64_t BLS12_G1ADD_GAS = 600;
const uint64_t BLS12_G1MUL_GAS = 12000;
const uint64_t BLS12_G2ADD_GAS = 4500;
const uint64_t BLS12_G2MUL_GAS = 55000;
const uint64_t BLS12_PAIRING_BASE_GAS = 115000;
const uint64_t BLS12_PAIRING_PAIR_GAS = 23000;
const uint64_t BLS12_MAP_FP_TO_G1_GAS = 5500;
const uint64_t BLS12_MAP_FP2_TO_G2_GAS = 110000;
const uint64_t BLS12_MULTIEXP_MULTIPLIER_GAS = 1000;
const uint64_t BLS12_MULTIEXP_DISCOUNT_TABLE_LEN = 128;
const uint64_t BLS12_MULTIEXP_DISCOUNT[128] = {
    1200, 888, 764, 641, 594, 547, 500, 453, 438, 423, 408, 394, 379, 364, 349, 334, 330, 326, 322, 318, 314, 310,
    306,  302, 298, 294, 289, 285, 281, 277, 273, 269, 268, 266, 265, 263, 262, 260, 259, 257, 256, 254, 253, 251,
    250,  248, 247, 245, 244, 242, 241, 239, 238, 236, 235, 233, 232, 231, 229, 228, 226, 225, 223, 222, 221, 220,
    219,  219, 218, 217, 216, 216, 215, 214, 213, 213, 212, 211, 211, 210, 209, 208, 208, 207, 206, 205, 205, 204,
    203,  202, 202, 201, 200, 199, 199, 198, 197, 196, 196, 195, 194, 193, 193, 192, 191, 191, 190, 189, 188, 188,
    187,  186, 185, 185, 184, 183, 182, 182, 181, 180, 179, 179, 178, 177, 176, 176, 175, 174};

static uint64_t multiexp_gas(uint64_t input_len, uint64_t stride, uint64_t mul_gas) {
  uint64_t k = input_len / stride;
  if (k == 0) return 0;
  uint64_t idx = k < BLS12_MULTIEXP_DISCOUNT_TABLE_LEN ? k - 1 : BLS12_MULTIEXP_DISCOUNT_TABLE_LEN - 1;
  return (k * mul_gas * BLS12_MULTIEXP_DISCOUNT[idx]) / BLS12_MULTIEXP_MULTIPLIER_GAS;
}
uint64_t bls12_g1add_gas(void) { return BLS12_G1ADD_GAS; }
uint64_t bls12_g1mul_gas(void) { return BLS12_G1MUL_GAS; }
uint64_t bls12_g1multiexp_gas(uint64_t input_len) { return multiexp_gas(input_len, 160, BLS12_G1MUL_GAS); }
uint64_t bls12_g2add_gas(void) { return BLS12_G2ADD_GAS; }
uint64_t bls12_g2mul_gas(void) { return BLS12_G2MUL_GAS; }
uint64_t bls12_g2multiexp_gas(uint64_t input_len) { return multiexp_gas(input_len, 288, BLS12_G2MUL_GAS); }
uint64_t bls12_pairing_gas(uint64_t input_len) {
  uint64_t k = input_len / 384;
  return k == 0 ? 0 : k * BLS12_PAIRING_PAIR_GAS + BLS12_PAIRING_BASE_GAS;
}
uint64_t bls12_map_fp_to_g1_gas(void) { return BLS12_MAP_FP_TO_G1_GAS; }
uint64_t bls12_map_fp2_to_g2_gas(void) { return BLS12_MAP_FP2_TO_G2_GAS; }
