/*
 * eip2537.h -- drop-in replacement for /root/reference/src/eip2537.h (the reference's C ABI).
 *
 * Same 13 bls12_* entry points, same EIP2537_ERROR values, same gas symbols, same byte
 * encodings (128-byte padded G1, 256-byte G2, 32-byte big-endian scalars), so the reference's
 * Go (go/blst_eip2537.go:49-205, cgo) and Rust (rust/src/lib.rs:18-96, extern "C") bindings
 * link against libblst_eip2537 unchanged.  Unlike the reference header this one does not need
 * blst.h: it defines `byte` itself (reference: src/eip2537.h:12 gets it from blst.h).
 *
 * Each prototype cites the reference declaration it replaces.
 */
#ifndef __EIP2537_H__
#define __EIP2537_H__

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char byte;

/* src/eip2537.h:31-40 -- values are part of the ABI (Rust reads them as u32, lib.rs:8-16) */
typedef enum {
  EIP2537_SUCCESS = 0,
  EIP2537_POINT_NOT_ON_CURVE,
  EIP2537_POINT_NOT_IN_SUBGROUP,
  EIP2537_INVALID_ELEMENT,
  EIP2537_ENCODING_ERROR,
  EIP2537_INVALID_LENGTH,
  EIP2537_EMPTY_INPUT,
  EIP2537_MEMORY_ERROR,
} EIP2537_ERROR;

/* src/eip2537.h:42-46 */
EIP2537_ERROR bls12_g1add(byte out[128], const byte in[256], size_t in_len);
EIP2537_ERROR bls12_g1mul(byte out[128], const byte in[160], size_t in_len);
EIP2537_ERROR bls12_g1multiexp(byte out[128], byte* in, size_t in_len);
EIP2537_ERROR bls12_g1multiexp_naive(byte out[128], byte* in, size_t in_len);
EIP2537_ERROR bls12_g1multiexp_bc(byte out[128], byte* in, size_t in_len);

/* src/eip2537.h:48-52 */
EIP2537_ERROR bls12_g2add(byte out[256], const byte in[512], size_t in_len);
EIP2537_ERROR bls12_g2mul(byte out[256], const byte in[288], size_t in_len);
EIP2537_ERROR bls12_g2multiexp(byte out[256], byte* in, size_t in_len);
EIP2537_ERROR bls12_g2multiexp_naive(byte out[256], byte* in, size_t in_len);
EIP2537_ERROR bls12_g2multiexp_bc(byte out[256], byte* in, size_t in_len);

/* src/eip2537.h:54 */
EIP2537_ERROR bls12_pairing(byte out[32], byte* in, size_t in_len);

/* src/eip2537.h:56-59 (bodies src/eip2537.c:1093-1121, :1135-1163): MAP_FP_TO_G1 / MAP_FP2_TO_G2, GPU-backed
 * (csrc/map.cuh: RFC 9380 simplified SWU, isogeny, cofactor clearing -- what blst_map_to_g1/_g2 compute).
 * in_len != 64 / 128 -> INVALID_LENGTH; pad byte or value >= p -> INVALID_ELEMENT; `out` untouched on error. */
EIP2537_ERROR bls12_map_fp_to_g1(byte out[128], const byte in[64], size_t in_len);
EIP2537_ERROR bls12_map_fp2_to_g2(byte out[256], const byte in[128], size_t in_len);

/* src/eip2537.h:62-72 */
extern const uint64_t BLS12_G1ADD_GAS;
extern const uint64_t BLS12_G1MUL_GAS;
extern const uint64_t BLS12_G2ADD_GAS;
extern const uint64_t BLS12_G2MUL_GAS;
extern const uint64_t BLS12_PAIRING_BASE_GAS;
extern const uint64_t BLS12_PAIRING_PAIR_GAS;
extern const uint64_t BLS12_MAP_FP_TO_G1_GAS;
extern const uint64_t BLS12_MAP_FP2_TO_G2_GAS;
extern const uint64_t BLS12_MULTIEXP_MULTIPLIER_GAS;
extern const uint64_t BLS12_MULTIEXP_DISCOUNT_TABLE_LEN;
extern const uint64_t BLS12_MULTIEXP_DISCOUNT[128];

/* src/eip2537.h:74-82 */
uint64_t bls12_g1add_gas(void);
uint64_t bls12_g1mul_gas(void);
uint64_t bls12_g1multiexp_gas(uint64_t input_len);
uint64_t bls12_g2add_gas(void);
uint64_t bls12_g2mul_gas(void);
uint64_t bls12_g2multiexp_gas(uint64_t input_len);
uint64_t bls12_pairing_gas(uint64_t input_len);
uint64_t bls12_map_fp_to_g1_gas(void);
uint64_t bls12_map_fp2_to_g2_gas(void);

#ifdef __cplusplus
}
#endif
#endif /* __EIP2537_H__ */
