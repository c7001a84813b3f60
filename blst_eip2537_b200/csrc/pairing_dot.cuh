// pairing_dot.cuh -- K5 throughput kernels built on the dot-product Fp12 engine (dot12.cuh).
//
//   k_pairing_lines_slots     one thread per PAIR SLOT: the 68 line functions of a pair, evaluated at P, written
//                             word-transposed (lines_t[(step*72 + word) * stride + slot]) so that the accumulate
//                             kernel's warps read them fully coalesced.  Slots follow the task order of the batch
//                             plan (k-th pairs of consecutive tasks are consecutive slots).
//   k_pairing_accumulate_dot  32 chunks per block of three warps; the Fp12 accumulators live in shared memory
//                             (two buffers, word-transposed: conflict-free), role r = warp index computes the
//                             coefficients of w^r and w^(r+3) of every operation, one barrier per operation.
//
// Replaces blst_miller_loop + blst_fp12_mul of /root/reference/src/eip2537.c:1060-1065 for all pairs of a chunk at
// once (shared squarings; only the boolean of the call is observable).
#pragma once
#include "dot12.cuh"

namespace b200 {
#ifdef __CUDACC__

// blocks of 96 threads per SM: 5 (<= 136 registers) or 6 (<= 112 registers; 6 x 37 KB is all the shared memory there is)
static constexpr int DOT_BLOCKS_PER_SM = 6;
static constexpr int DOT_SMEM_BYTES = 2 * 144 * 32 * 4;
static constexpr int LINE_WORDS = 72;           // Line = 3 Fp2 = 6 Fp = 72 words

__global__ void __launch_bounds__(64, 6) k_pairing_lines_slots(const G1Affine* __restrict__ g1, const G2Affine* __restrict__ g2,
                                                            const PairingPlanState* __restrict__ st, const uint32_t* __restrict__ slot_pair,
                                                            size_t stride, uint32_t* __restrict__ lines_t, unsigned char* __restrict__ skip_slot) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= st->npair_slots) return;
  const size_t j = slot_pair[q];
  if (j == 0xFFFFFFFFu) { skip_slot[q] = 1; return; }           // padding slot (slot bases are multiples of 32)
  G1Affine p = g1[j];
  G2Affine qq = g2[j];
  if (is_inf(p) || is_inf(qq)) { skip_slot[q] = 1; return; }   // contributes 1
  skip_slot[q] = 0;
  G2Proj t;
  t.x = qq.x; t.y = qq.y; t.z = fp2_one();
  int s = 0;
  for (int i = 62; i >= 0; i--) {
    const int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
    for (int a = 0; a < nsteps; a++, s++) {
      Line ln;
      if (a == 0) ml_dbl_step(t, ln.l0, ln.l1, ln.l4); else ml_add_step(t, qq, ln.l0, ln.l1, ln.l4);
      ln.l1 = mulfpo(ln.l1, p.x); ln.l4 = mulfpo(ln.l4, p.y);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&ln);
      uint32_t* dst = lines_t + (size_t)s * LINE_WORDS * stride + q;
#pragma unroll 8
      for (int x = 0; x < LINE_WORDS; x++) dst[(size_t)x * stride] = w[x];
    }
  }
}

// f buffers (dynamic shared memory, dot::dot_smem): word (c, limb) of lane L at buf*4608 + (c*12 + limb)*32 + L,
// c = 2*k + {0 re, 1 im}, k = power of w
template <int BLOCKS, int PAIR64>
__global__ void __launch_bounds__(96, BLOCKS) k_pairing_accumulate_dot(
    const PairingTask* __restrict__ tasks, const PairingPlanState* __restrict__ st, const uint32_t* __restrict__ lines_t,
    const unsigned char* __restrict__ skip_slot, size_t stride, Fp12* __restrict__ fchunk) {
  using dot::dot_smem;
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const uint32_t ntasks = st->ntasks;
  if (blockIdx.x * 32u >= ntasks) return;
  const uint32_t t = blockIdx.x * 32u + lane;
  const bool live = t < ntasks;
  PairingTask task = PairingTask{0, 0, 0};
  if (live) task = tasks[t];
  const uint32_t npairs = live ? task.npairs : 0;
  uint32_t kmax = npairs;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { uint32_t v = __shfl_xor_sync(0xffffffffu, kmax, o); kmax = v > kmax ? v : kmax; }
  // f = 1
  for (int q = 0; q < 2; q++) {
    const int k = role + 3 * q;
    for (int comp = 0; comp < 2; comp++)
      for (int l = 0; l < 12; l++) dot_smem[((2 * k + comp) * 12 + l) * 32 + lane] = (k == 0 && comp == 0) ? C_ONE()[l] : 0u;
  }
  __syncthreads();
  uint32_t cur = 0;      // word offset of the current buffer: 0 or 4608
  auto run_op = [&](int op, const dot::GlobalWords& G, bool keep) {
    const dot::SmemWords F{cur + (uint32_t)lane};
    uint32_t* dst = dot_smem + (cur ^ 4608u) + lane;     // 4608 = 0x1200: xor toggles between 0 and 4608
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      const int c = 2 * (role + 3 * (q >> 1)) + (q & 1);
      Fp r;
      if (keep) r = dot::dot_eval_dev<BLOCKS + 10 * PAIR64>(dot::op_row(op, c), dot::op_count(op, c), F, G);
      else      r = F.load(c);
#pragma unroll
      for (int l = 0; l < 12; l++) dst[(c * 12 + l) * 32] = r.v[l];
    }
    __syncthreads();
    cur ^= 4608u;
  };
  int s = 0;
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    run_op(dot::OP_SQR, dot::GlobalWords{lines_t, 0}, true);
    const int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
    for (int a = 0; a < nsteps; a++, s++) {
      for (uint32_t k = 0; k < kmax; k++) {
        bool active = k < npairs;
        size_t slot = 0;
        if (active) { slot = (size_t)st->slot_base[k] + t; active = skip_slot[slot] == 0; }
        run_op(dot::OP_MUL014, dot::GlobalWords{lines_t + (size_t)s * LINE_WORDS * stride + slot, (uint32_t)stride}, active);
      }
    }
  }
  // conj (the loop ran over |z|, z < 0): negate the odd powers of w; back to the memory order of Fp12
  if (live) {
    const dot::SmemWords F{cur + (uint32_t)lane};
    Fp2* out = reinterpret_cast<Fp2*>(&fchunk[task.slot]);
    for (int q = 0; q < 2; q++) {
      const int k = role + 3 * q;
      Fp2 v;
      v.c0 = F.load(2 * k); v.c1 = F.load(2 * k + 1);
      if (k & 1) v = neg(v);
      out[dot::mem_of_wpow(k)] = v;
    }
  }
}

// ---- six roles per chunk, lines staged in shared memory by cp.async -------------------------------------------
// One warp per coefficient of w^r (r = 0..5): 32 chunks per block of 192 threads, four blocks per SM (24 warps:
// what the dot engine needs to saturate the multiply pipe, 8.9 TMAC32/s in isolation).  Shared memory per block:
// two f buffers (2 x 18 KB) + two line stages (2 x 9 KB); the line of the NEXT sparse product is in flight
// (cp.async, 16-byte pieces, issued by all threads) while the current one is multiplied in, so the loop body
// touches shared memory only.
static constexpr int DOT6_THREADS = 192, DOT6_BLOCKS_PER_SM = 3;   // 3 blocks: <= 112 registers, no spills; 4 blocks: 80 registers, spills
static constexpr int DOT6_F_WORDS = 144 * 32, DOT6_LINE_WORDS = LINE_WORDS * 32;
static constexpr int DOT6_SMEM_BYTES = (2 * DOT6_F_WORDS + 2 * DOT6_LINE_WORDS) * 4;     // 55,296 B

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

template <int BLOCKS>
__global__ void __maxnreg__(BLOCKS == 3 ? 112 : 80) k_pairing_accumulate_dot6(
    const PairingTask* __restrict__ tasks, const PairingPlanState* __restrict__ st, const uint32_t* __restrict__ lines_t,
    const unsigned char* __restrict__ skip_slot, size_t stride, Fp12* __restrict__ fchunk) {
  using dot::dot_smem;
  __shared__ uint32_t slot_base[PAIRING_MAX_CHUNK];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const uint32_t ntasks = st->ntasks;
  if (blockIdx.x * 32u >= ntasks) return;
  const uint32_t t = blockIdx.x * 32u + lane;
  const bool live = t < ntasks;
  PairingTask task = PairingTask{0, 0, 0};
  if (live) task = tasks[t];
  const uint32_t npairs = live ? task.npairs : 0;
  uint32_t kmax = npairs;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { uint32_t v = __shfl_xor_sync(0xffffffffu, kmax, o); kmax = v > kmax ? v : kmax; }
  if (threadIdx.x < PAIRING_MAX_CHUNK) slot_base[threadIdx.x] = st->slot_base[threadIdx.x];
  // pairs with an infinite member contribute 1: a per-lane mask, read once
  uint32_t skipmask = 0;
  for (uint32_t k = 0; k < npairs; k++) skipmask |= (uint32_t)(skip_slot[(size_t)st->slot_base[k] + t] != 0) << k;
  // f = 1
  for (int comp = 0; comp < 2; comp++)
    for (int l = 0; l < 12; l++) dot_smem[((2 * role + comp) * 12 + l) * 32 + lane] = (role == 0 && comp == 0) ? C_ONE()[l] : 0u;
  __syncthreads();
  const uint32_t stage_smem = (uint32_t)__cvta_generic_to_shared(dot_smem + 2 * DOT6_F_WORDS);
  // line of sparse product (step s, pair position k) of this block's 32 chunks -> stage `st_i`: 72 rows of 128 B
  auto prefetch = [&](int st_i, int s, uint32_t k) {
    const uint32_t* src = lines_t + (size_t)s * LINE_WORDS * stride + slot_base[k] + blockIdx.x * 32u;
    const uint32_t dst = stage_smem + (uint32_t)st_i * (DOT6_LINE_WORDS * 4);
#pragma unroll
    for (int p = threadIdx.x; p < LINE_WORDS * 8; p += DOT6_THREADS) {
      const int w = p >> 3, j = p & 7;
      cp_async16(dst + (uint32_t)(w * 32 + j * 4) * 4, src + (size_t)w * stride + j * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t cur = 0;      // word offset of the current f buffer: 0 or DOT6_F_WORDS
  const int M = ML_STEPS * (int)kmax;
  if (M > 0) prefetch(0, 0, 0);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  int s = 0, m = 0;
  // ONE operation site (the dot engine is inlined exactly once): per bit of |z| one squaring, then the sparse
  // products of this step's lines (two steps when the bit is set), pair position by pair position
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    const int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
    const int nops = 1 + nsteps * (int)kmax;
    uint32_t k = 0;
#pragma unroll 1
    for (int j = 0; j < nops; j++) {
      int op = dot::OP_SQR;
      uint32_t line_base = 0;
      bool active = true;
      if (j > 0) {
        op = dot::OP_MUL014;
        if (m + 1 < M) { if (k + 1 < kmax) prefetch((m + 1) & 1, s, k + 1); else prefetch((m + 1) & 1, s + 1, 0); }
        line_base = (uint32_t)(2 * DOT6_F_WORDS + (m & 1) * DOT6_LINE_WORDS);
        active = k < npairs && !((skipmask >> k) & 1u);
      }
      const dot::SmemWords F{cur + (uint32_t)lane}, G{line_base + (uint32_t)lane};
      uint32_t* dst = dot_smem + (cur ^ (uint32_t)DOT6_F_WORDS) + lane;
#pragma unroll 1
      for (int q = 0; q < 2; q++) {
        const int c = 2 * role + q;
        Fp r = dot::dot_eval_smem(dot::op_row(op, c), dot::op_count(op, c), F, G);
        if (!active) r = F.load(c);
#pragma unroll
        for (int l = 0; l < 12; l++) dst[(c * 12 + l) * 32] = r.v[l];
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      cur ^= (uint32_t)DOT6_F_WORDS;
      if (j > 0) { m++; if (++k == kmax) { k = 0; s++; } }
    }
  }
  // conj (the loop ran over |z|, z < 0): negate the odd powers of w; back to the memory order of Fp12
  if (live) {
    const dot::SmemWords F{cur + (uint32_t)lane};
    Fp2* out = reinterpret_cast<Fp2*>(&fchunk[task.slot]);
    Fp2 v;
    v.c0 = F.load(2 * role); v.c1 = F.load(2 * role + 1);
    if (role & 1) v = neg(v);
    out[dot::mem_of_wpow(role)] = v;
  }
}

#endif  // __CUDACC__
}  // namespace b200
