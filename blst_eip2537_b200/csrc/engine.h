/* engine.h -- internal C seam between the plain-C host layer and the CUDA engine.
 * Every function returns an EIP2537_ERROR value as int; CUDA failures map to 7 (MEMORY_ERROR). */
#ifndef B200_ENGINE_H
#define B200_ENGINE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* group: 1 = G1 (160-byte pairs, 128-byte result), 2 = G2 (288-byte pairs, 256-byte result) */
int b200_msm_host(int group, const unsigned char* in, size_t n_pairs, unsigned char* out);
int b200_add_host(int group, const unsigned char* in, unsigned char* out);
int b200_map_host(int group, const unsigned char* in, unsigned char* out);
int b200_pairing_host(const unsigned char* in, size_t k_pairs, unsigned char* out);
#ifdef __cplusplus
}
#endif
#endif
