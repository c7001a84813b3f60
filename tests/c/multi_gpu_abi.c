/*
 * multi_gpu_abi.c -- plain-C consumer of the reference ABI (what the Go / Rust bindings call,
 * /root/reference/go/blst_eip2537.go:70-81, rust/src/lib.rs:140-153) driving SEVERAL GPUs through ONE
 * bls12_g1multiexp / bls12_g2multiexp / bls12_pairing_batch call after bls12_b200_init_multi().
 *
 *   ./multi_gpu_abi [log2_pairs=19] [ngpu=0 (all)]
 *
 * Checks: the N-GPU result equals the 1-GPU result byte for byte (G1 and G2); a failing pair in the LAST shard
 * and an earlier one in the FIRST shard give the first one's code (eip2537.c:580-592 precedence) and leave
 * `out` untouched; every device launched kernels.  Prints one JSON line; exit status 0 = all checks passed.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "eip2537_b200.h"

static uint64_t rng_state = 0x2537ULL;
static uint64_t splitmix64(void) {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static void fill_random(byte* p, size_t n) {
  for (size_t i = 0; i < n; i += 8) { uint64_t v = splitmix64(); memcpy(p + i, &v, n - i < 8 ? n - i : 8); }
}
static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* n pairs: point_i = a_i * generator (product's generator kernel), scalar_i uniform 256-bit */
static byte* make_input(int group, size_t n) {
  const size_t plen = group == 1 ? 128 : 256, stride = plen + 32;
  byte* scal = malloc(32 * n);
  byte* pts = malloc(plen * n);
  byte* in = malloc(stride * n);
  if (!scal || !pts || !in) return NULL;
  fill_random(scal, 32 * n);
  for (size_t i = 0; i < n; i++) scal[32 * i] &= 0x3f;
  EIP2537_ERROR rc = group == 1 ? bls12_b200_g1_generator_mul(pts, scal, n) : bls12_b200_g2_generator_mul(pts, scal, n);
  if (rc) { fprintf(stderr, "generator_mul: %d %s\n", rc, bls12_b200_last_error()); return NULL; }
  fill_random(scal, 32 * n);
  for (size_t i = 0; i < n; i++) { memcpy(in + stride * i, pts + plen * i, plen); memcpy(in + stride * i + plen, scal + 32 * i, 32); }
  free(scal); free(pts);
  return in;
}

int main(int argc, char** argv) {
  const int logn = argc > 1 ? atoi(argv[1]) : 19;
  int ngpu = argc > 2 ? atoi(argv[2]) : 0;
  const size_t n = (size_t)1 << logn;
  int ok = 1;
  if (bls12_b200_init_multi(1)) { fprintf(stderr, "init: %s\n", bls12_b200_last_error()); return 2; }
  byte* in1 = make_input(1, n);
  byte* in2 = make_input(2, n / 4);
  if (!in1 || !in2) return 2;
  byte ref1[128], ref2[256], out1[128], out2[256];
  EIP2537_ERROR rc;
  if ((rc = bls12_g1multiexp(ref1, in1, 160 * n))) { fprintf(stderr, "1-GPU g1: %d %s\n", rc, bls12_b200_last_error()); return 2; }
  if ((rc = bls12_g2multiexp(ref2, in2, 288 * (n / 4)))) { fprintf(stderr, "1-GPU g2: %d %s\n", rc, bls12_b200_last_error()); return 2; }
  double t0 = now_ms();
  for (int r = 0; r < 3; r++) bls12_g1multiexp(out1, in1, 160 * n);
  const double ms_single = (now_ms() - t0) / 3;

  if (bls12_b200_init_multi(ngpu)) { fprintf(stderr, "init_multi: %s\n", bls12_b200_last_error()); return 2; }
  ngpu = bls12_b200_multi_gpus();
  uint64_t before[16];
  for (int d = 0; d < ngpu && d < 16; d++) before[d] = bls12_b200_device_launch_count(d);
  memset(out1, 0xA5, sizeof out1);
  memset(out2, 0xA5, sizeof out2);
  if ((rc = bls12_g1multiexp(out1, in1, 160 * n))) { fprintf(stderr, "N-GPU g1: %d %s\n", rc, bls12_b200_last_error()); ok = 0; }
  if ((rc = bls12_g2multiexp(out2, in2, 288 * (n / 4)))) { fprintf(stderr, "N-GPU g2: %d %s\n", rc, bls12_b200_last_error()); ok = 0; }
  const int same1 = !memcmp(out1, ref1, 128), same2 = !memcmp(out2, ref2, 256);
  t0 = now_ms();
  for (int r = 0; r < 3; r++) bls12_g1multiexp(out1, in1, 160 * n);
  const double ms_multi = (now_ms() - t0) / 3;
  int devices_used = 0;
  for (int d = 0; d < ngpu && d < 16; d++) devices_used += bls12_b200_device_launch_count(d) > before[d];

  /* first-error precedence across shards: pair n-5 (last shard) is off the curve (code 1), pair 3 (first shard) has
     a field element >= p (code 3): the earlier pair decides; `out` stays untouched */
  byte sentinel[128];
  memset(sentinel, 0xA5, 128);
  memset(out1, 0xA5, 128);
  in1[160 * (n - 5) + 127] ^= 1;
  const EIP2537_ERROR late_only = bls12_g1multiexp(out1, in1, 160 * n);
  memset(in1 + 160 * 3 + 16, 0xff, 48);
  const EIP2537_ERROR both = bls12_g1multiexp(out1, in1, 160 * n);
  const int untouched = !memcmp(out1, sentinel, 128);
  /* a call is sharded over at most n / 2^17 devices (each shard must be worth a GPU) */
  int expect_devices = (int)(n >> 17) < ngpu ? (int)(n >> 17) : ngpu;
  if (expect_devices < 1) expect_devices = 1;
  ok = ok && same1 && same2 && devices_used == expect_devices && late_only == EIP2537_POINT_NOT_ON_CURVE && both == EIP2537_INVALID_ELEMENT && untouched;
  printf("{\"test\": \"multi_gpu_abi\", \"pairs\": %zu, \"gpus\": %d, \"devices_used\": %d, \"devices_expected\": %d, \"g1_equal\": %s, \"g2_equal\": %s, "
         "\"late_error_code\": %d, \"first_error_code\": %d, \"out_untouched\": %s, \"ms_1gpu\": %.2f, \"ms_ngpu\": %.2f, \"ok\": %s}\n",
         n, ngpu, devices_used, expect_devices, same1 ? "true" : "false", same2 ? "true" : "false", (int)late_only, (int)both,
         untouched ? "true" : "false", ms_single, ms_multi, ok ? "true" : "false");
  bls12_b200_shutdown();
  return ok ? 0 : 1;
}
