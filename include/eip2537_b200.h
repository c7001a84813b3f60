/*
 * eip2537_b200.h -- ADDITIVE extensions to the reference ABI (nothing in the reference calls
 * these; SURVEY.md 8(b) "New (additive) exports for the batch configs").
 *
 * The legacy ABI (eip2537.h) is one call = one result with host buffers.  The data-parallel
 * configurations of BASELINE.json need (1) many independent calls in one submission,
 * (2) device-resident inputs for kernel-only timing and for multi-GPU sharding where each
 * rank holds a slice in HBM, and (3) explicit lifetime control.  Semantics per call are
 * identical to the single-call function of eip2537.h (same bytes, same codes).
 *
 * All pointers named d_* are CUDA device pointers, 16-byte aligned, on the current device.
 * `stream` is a cudaStream_t passed as void* and used as given (NULL = the CUDA default stream);
 * work submitted through these entry points is asynchronous with respect to the host.
 * Every function returns an EIP2537_ERROR; CUDA failures map to EIP2537_MEMORY_ERROR.
 *
 * Threading and streams.  Like the reference (src/eip2537.c has no mutable globals) every entry point may be
 * called from any number of host threads at once.  Each device owns a small pool of WORKSPACES (streams, scratch
 * buffers in HBM, a pinned staging ring; at most B200_WORKSPACES = 4 per device, created on demand); a call leases
 * one for its duration, so up to that many host calls per device run concurrently and further callers wait.
 * The asynchronous device-resident entry points return while their kernels still use the leased workspace: an
 * event recorded on the caller's stream marks the end of the submission, and whoever leases that workspace next
 * orders itself after the event (no wait at all for back-to-back submissions on the same stream).  Submissions
 * from different streams or threads therefore never share scratch memory in flight.
 * Host input buffers may be pinned or ordinary pageable memory (Go heap slices, Rust Vec); pageable input is
 * copied through the workspace's pinned ring by a few helper threads (B200_COPY_THREADS, default 4).
 */
#ifndef __EIP2537_B200_H__
#define __EIP2537_B200_H__

#include "eip2537.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lifetime: lazy init happens on first use; these make it explicit.  device < 0 = current. */
EIP2537_ERROR bls12_b200_init(int device);
void bls12_b200_shutdown(void);

/* ---- several GPUs behind the PLAIN ABI (one process).  After bls12_b200_init_multi(ngpu) (ngpu <= 0: all):
 *   - one bls12_g{1,2}multiexp call of >= 2^17 pairs per GPU is split into contiguous pair ranges, one host
 *     thread + stream per device, each device pulling its slice over its own PCIe link; the last kernel of every
 *     shard stores its partial sum and first-error key straight into device 0's gather buffer through NVLink peer
 *     memory (one exchange step, ngpu x 512 B); device 0 sums, inverts once, encodes.  Same bytes and same codes
 *     (first failing pair in input order) as the single-GPU call -- replaces the loops of src/eip2537.c:580-606 /
 *     :650-701 behind the unchanged prototype src/eip2537.h:44-46;
 *   - bls12_pairing_batch / bls12_g{1,2}multiexp_batch shard by call index, no collective.
 * bls12_b200_multi_gpus() = the current setting; bls12_b200_device_launch_count(d) = kernels launched on device d. */
EIP2537_ERROR bls12_b200_init_multi(int ngpu);
int bls12_b200_multi_gpus(void);
uint64_t bls12_b200_device_launch_count(int device);

/* ---- one process PER GPU (torchrun / MPI ranks): the C library owns an NCCL communicator (libnccl.so.2 is
 *      dlopen'ed on first use; single-GPU consumers carry no NCCL dependency).  Rank 0 calls
 *      bls12_b200_comm_unique_id, the 128 bytes travel to every rank by any means, every rank calls
 *      bls12_b200_comm_init(world, rank, id) with its device current.  bls12_b200_msm_sharded_* then run this rank's
 *      shard (pairs index_base .. index_base + n of the global input) and ONE ncclAllGather of world x 512 B
 *      {partial sum, first-error key}; every rank ends with the same encoded result / code.
 *      _device: device-resident shard, asynchronous on `stream`, d_status as for bls12_b200_msm_device;
 *      _host: host-resident shard, synchronous, returns the code, `out` written only on success. */
EIP2537_ERROR bls12_b200_comm_unique_id(byte* id128);
EIP2537_ERROR bls12_b200_comm_init(int world, int rank, const byte* id128);
void bls12_b200_comm_destroy(void);
EIP2537_ERROR bls12_b200_msm_sharded_device(int group, const void* d_in, size_t n, uint64_t index_base, void* d_out,
                                            uint64_t* d_status, void* stream);
EIP2537_ERROR bls12_b200_msm_sharded_host(int group, const byte* in, size_t n, uint64_t index_base, byte* out);
/* last CUDA error string seen by the engine on this thread ("" if none) */
const char* bls12_b200_last_error(void);
/* number of kernels the engine has launched since process start (bench.py's gpu_launches) */
uint64_t bls12_b200_launch_count(void);
/* force the Pippenger window width (2..16); 0 = automatic */
void bls12_b200_set_window(int c);

/* ---- batch of independent calls, host buffers (replaces n calls of bls12_pairing /
 *      bls12_g{1,2}multiexp; offsets[n+1] are byte offsets of each call's input in `in`).
 *      outs = n*32 / n*128 / n*256 bytes; errs[n] per-call codes; outs of failed calls are zeroed. */
EIP2537_ERROR bls12_pairing_batch(byte* outs, EIP2537_ERROR* errs, const byte* in,
                                  const uint64_t* offsets, size_t n);

/* Same for MULTIEXP: n independent calls (the EVM shape: many calls of a few to a few hundred pairs).
 * outs = n*128 (G1) / n*256 (G2) bytes.  Each pair is multiplied by its own thread (the reference's
 * naive strategy made data-parallel), so this is the right entry point for MANY SMALL calls; one large
 * call belongs to bls12_g{1,2}multiexp (Pippenger). */
EIP2537_ERROR bls12_g1multiexp_batch(byte* outs, EIP2537_ERROR* errs, const byte* in,
                                     const uint64_t* offsets, size_t n);
EIP2537_ERROR bls12_g2multiexp_batch(byte* outs, EIP2537_ERROR* errs, const byte* in,
                                     const uint64_t* offsets, size_t n);

/* n independent MAP_FP_TO_G1 / MAP_FP2_TO_G2 calls (src/eip2537.c:1093, :1135) in one submission.
 * in = n*64 (Fp) / n*128 (Fp2) bytes, outs = n*128 / n*256 bytes, errs[j] = 0 or INVALID_ELEMENT;
 * for n > 1, outs[j] is unspecified where errs[j] != 0.  Returns non-zero only for a CUDA failure. */
EIP2537_ERROR bls12_map_fp_to_g1_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, size_t n);
EIP2537_ERROR bls12_map_fp2_to_g2_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, size_t n);

/* ---- device-resident MSM (group = 1 for G1, 2 for G2).
 *  d_in: n pairs in wire format (160*n / 288*n bytes).  d_out: 128 / 256 encoded bytes.
 *  d_status: one uint64; ~0 on success else (first failing pair index << 8) | code.
 *  Asynchronous on `stream`; the caller synchronises and reads d_status / d_out. */
EIP2537_ERROR bls12_b200_msm_device(int group, const void* d_in, size_t n, void* d_out,
                                    uint64_t* d_status, void* stream);
/* size in bytes of one partial sum (XYZZ coordinates, Montgomery limbs) for group 1 / 2 */
size_t bls12_b200_partial_bytes(int group);
/* partial sum only (multi-GPU sharding): bls12_b200_partial_bytes(group) bytes at d_partial.
 * index_base is added to pair indices reported in d_status (global index of this shard).
 * PRECONDITION: *d_status must hold ~0 (all ones) before the call, set on the same stream (the kernels only
 * atomicMin into it so that several shards can share one word); bls12_b200_msm_device and the sharded / host
 * entry points initialise it themselves.  n must be < 2^31 and n * ceil(256 / window) < 2^32. */
EIP2537_ERROR bls12_b200_msm_partial_device(int group, const void* d_in, size_t n, uint64_t index_base,
                                            void* d_partial, uint64_t* d_status, void* stream);
/* same with a HOST-resident shard (pinned or pageable): the shard is streamed in chunks that accumulate
 * while the next chunk crosses PCIe; d_partial / d_status are device buffers; returns when they are ready */
EIP2537_ERROR bls12_b200_msm_partial_host(int group, const byte* in, size_t n, uint64_t index_base,
                                          void* d_partial, uint64_t* d_status);
/* sum `count` partials (e.g. after an NCCL all-gather), convert to affine and encode */
EIP2537_ERROR bls12_b200_msm_combine_device(int group, const void* d_partials, int count, void* d_out,
                                            void* stream);

/* ---- device-resident pairing batch: d_in = concatenated calls, d_offsets[n+1] byte offsets,
 *      d_outs n*32 bytes, d_errs n int32 codes.  Asynchronous on `stream`. */
EIP2537_ERROR bls12_b200_pairing_batch_device(const void* d_in, const uint64_t* d_offsets, size_t n,
                                              size_t total_pairs, void* d_outs, int32_t* d_errs, void* stream);

/* ---- batched point validation (K3).  points: n encoded points, `stride_bytes` apart (128 / 256 for bare
 *      point arrays, 160 / 288 to walk a MULTIEXP input, 384 for the G1 fields of a PAIRING input;
 *      a multiple of 16).  codes[i] = 0, EIP2537_INVALID_ELEMENT, EIP2537_POINT_NOT_ON_CURVE, or --
 *      when check_subgroup != 0 -- EIP2537_POINT_NOT_IN_SUBGROUP (the test the reference applies in
 *      PAIRING, src/eip2537.c:1041/:1051). */
EIP2537_ERROR bls12_b200_points_check(int group, const byte* points, size_t n, size_t stride_bytes,
                                      int check_subgroup, int32_t* codes);
EIP2537_ERROR bls12_b200_points_check_device(int group, const void* d_points, size_t n, size_t stride_bytes,
                                             int check_subgroup, int32_t* d_codes, void* stream);
/* opt-in "checked MSM": MULTIEXP additionally rejects points outside G1/G2 with
 * EIP2537_POINT_NOT_IN_SUBGROUP.  OFF by default because the reference does not check
 * (src/eip2537.c:340, :401 are TODOs) and the codes would differ. */
void bls12_b200_set_checked_msm(int on);

/* PAIRING batches of at most n_calls calls run on the warp-cooperative low-latency kernel (one warp per
 * call), larger ones on the dot-engine throughput kernels (three lanes per chunk, Fp12 in shared memory).  Default 128 (env B200_PAIRING_COOP_MAX);
 * n_calls < 0 restores the default.  Returns the previous threshold.  Results are identical either way. */
long bls12_b200_set_pairing_coop_max(long n_calls);

/* ---- workload generators (synthetic inputs): out[i] = encode(k_i * generator),
 *      k_i = 32-byte big-endian scalars.  Host buffers. */
EIP2537_ERROR bls12_b200_g1_generator_mul(byte* out, const byte* scalars, size_t n);
EIP2537_ERROR bls12_b200_g2_generator_mul(byte* out, const byte* scalars, size_t n);

/* ---- K1 microbenchmarks over n_threads threads x `iters` iterations; elapsed milliseconds in *ms
 *      (CUDA events).  mode 0 = dependent Fp multiplications (digest = thread 0's result),
 *      1 = IMAD.WIDE.U32 issue-rate probe (64 MAC32 per iteration), 2 = carry-chained
 *      IMAD.WIDE.U32.X probe (24 MAC32 per iteration), 3 = IMAD + IMAD.HI pairs (64 MAC32 per iteration) */
EIP2537_ERROR bls12_b200_fp_microbench(int mode, size_t n_threads, int iters, float* ms, byte* digest48);

/* ---- per-stage timing of the MSM pipeline (CUDA events on the launching stream), for the
 *      roofline report: enable, run one bls12_b200_msm_device, then read
 *      stage_ms4 = {decode+digits+sort, bucket accumulate, bucket reduce tree, window combine}
 *      and the number of non-zero signed digits (= point additions done by the accumulate kernel) */
void bls12_b200_set_profile(int on);
EIP2537_ERROR bls12_b200_last_msm_profile(float* stage_ms4, uint64_t* nonzero_digits);
/* work3 = { non-zero digits D of the last profiled MULTIEXP, bucket additions done pairwise in affine form with a shared
 * inversion (6 field multiplications each), bucket additions done by the XYZZ walk (10, G2: 28) }.  bench.py's roofline. */
EIP2537_ERROR bls12_b200_last_msm_work(uint64_t* work3);
/* same for the pairing batch: stage_ms4 = {decode + subgroup checks, line functions,
 * chunked multi-Miller accumulate, Fp12 product + final exponentiation} */
EIP2537_ERROR bls12_b200_last_pairing_profile(float* stage_ms4);
/* pairs per chunk the batch planner chose for the last pairing batch (0 if none) */
int bls12_b200_last_pairing_chunk(void);

/* ---- on-device self test of the PTX field arithmetic against portable C++ on n pseudo-random
 *      inputs; mismatches4 = {mul, add, sub, inv} mismatch counts (all must be 0) */
EIP2537_ERROR bls12_b200_selftest(uint64_t* mismatches4, size_t n);

#ifdef __cplusplus
}
#endif
#endif
