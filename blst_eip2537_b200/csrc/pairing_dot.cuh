// pairing_dot.cuh -- K5 throughput kernels built on the dot-product Fp12 engine (dot12.cuh).
//
//   k_pairing_lines_slots      one thread per PAIR SLOT: the 68 line functions of a pair, evaluated at P, written
//                              word-transposed (lines_t[(step*72 + word) * stride + slot]) so that the accumulate
//                              kernel reads them as 128-byte rows.  Slots follow the task order of the batch plan
//                              (k-th pairs of consecutive tasks are consecutive slots).  The walk T = [|z|]Q of the
//                              Miller loop doubles as the ladder of the G2 membership test (deferred from decode).
//   k_pairing_accumulate_dot6  32 chunks per block of SIX warps, warp r owns the coefficient of w^r of every chunk;
//                              Fp12 accumulators in shared memory (two buffers, word-transposed: conflict-free), the
//                              line of the next sparse product staged by cp.async, one barrier per operation.
//   k_pairing_final_dot6       same geometry, 32 CALLS per block: product of the chunk values, final exponentiation
//                              as an interpreted program, is-one.
//
// Replaces blst_miller_loop + blst_fp12_mul + blst_final_exp + blst_fp12_is_one of
// /root/reference/src/eip2537.c:1060-1078 for all pairs of a chunk / all calls of a batch at once (shared squarings;
// only the boolean of the call is observable).
#pragma once
#include "dot12.cuh"

namespace b200 {
#ifdef __CUDACC__

static constexpr int LINE_WORDS = 72;           // Line = 3 Fp2 = 6 Fp = 72 words

__global__ void __launch_bounds__(64, 6) k_pairing_lines_slots(const G1Affine* __restrict__ g1, const G2Affine* __restrict__ g2,
                                                            const PairingPlanState* __restrict__ st, const uint32_t* __restrict__ slot_pair,
                                                            size_t stride, uint32_t* __restrict__ lines_t, unsigned char* __restrict__ skip_slot,
                                                            int* __restrict__ status) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= st->npair_slots) return;
  const size_t j = slot_pair[q];
  if (j == 0xFFFFFFFFu) { skip_slot[q] = 1; return; }           // padding slot (slot bases are multiples of 32)
  G1Affine p = g1[j];
  G2Affine qq = g2[j];
  if (is_inf(qq)) { skip_slot[q] = 1; return; }                  // infinity is a member; the pair contributes 1 (Appendix D-2)
  if (is_inf(p)) {                                               // contributes 1, but Q must still be in G2 (eip2537.c:1051)
    skip_slot[q] = 1;
    if (!g2_in_subgroup(qq)) status[j] = E_NOT_IN_SUBGROUP;
    return;
  }
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
  // Deferred G2 membership test psi(Q) == [z]Q (blst_p2_affine_in_g2, eip2537.c:1051): the walk above ended at
  // T = [|z|]Q in homogeneous coordinates, and z < 0, so the test is psi(Q) == -T.  The step formulas are incomplete;
  // every exceptional event (T = infinity, T = +-Q at an addition step -- possible only for points of small order
  // outside G2) forces Z = 0 from then on, in which case the exact ladder decides.
  bool member;
  if (is_zero(t.z)) {
    member = g2_in_subgroup(qq);
  } else {
    const Fp2 px = mulo(conj(qq.x), fp2_load_const(C_PSI_CX()));
    const Fp2 py = mulo(conj(qq.y), fp2_load_const(C_PSI_CY()));
    member = eq(mulo(px, t.z), t.x) && eq(mulo(py, t.z), neg(t.y));
  }
  if (!member) status[j] = E_NOT_IN_SUBGROUP;
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
        const int c = dot::role_comp(op, role, q);
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

// ---- final exponentiation on the dot engine ---------------------------------------------------------------------
// Replaces blst_final_exp + blst_fp12_is_one (/root/reference/src/eip2537.c:1070-1078) for 32 calls per block of six
// role warps (same geometry as the accumulate kernel).  Round 1 ran this stage one THREAD per call: 16384 threads,
// a 12.9 ms chain of ~7,700 dependent multiplications.  Here the chain is walked operation by operation with the
// six coefficients of each Fp12 value computed side by side; three Fp12 buffers live in shared memory, values that
// are needed again much later (f, f^((z-1)^2 (z+p))) are parked in HBM.
// The exponent chain is a small PROGRAM (generated once on the host, engine.cu final_exp_program) interpreted by
// every block, so the dot engine is inlined at exactly one site.
enum FinalOp : uint32_t { FX_MUL = 0, FX_CYC = 1, FX_CONJ = 2, FX_FROB1 = 3, FX_FROB2 = 4, FX_COPY = 5, FX_PARK = 6, FX_UNPARK = 7, FX_INV = 8 };
static constexpr int FX_PARK_SLOTS = 2;
__host__ __device__ inline uint32_t fx_encode(uint32_t op, uint32_t d, uint32_t x, uint32_t y) { return op | (d << 8) | (x << 12) | (y << 16); }

__global__ void __maxnreg__(112) k_pairing_final_dot6(size_t n_calls, const unsigned long long* __restrict__ offsets,
                                                      const PairingPlanState* __restrict__ st, const uint32_t* __restrict__ call_first_task,
                                                      const Fp12* __restrict__ fchunk, const uint32_t* __restrict__ prog, int prog_len,
                                                      Fp12* __restrict__ park, uint32_t* __restrict__ outs, const int* __restrict__ errs) {
  using dot::dot_smem;
  __shared__ int not_one[32];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const size_t call = (size_t)blockIdx.x * 32 + lane;
  const bool in_range = call < n_calls;
  const bool live = in_range && errs[call] == E_SUCCESS;
  if (role == 0 && in_range) {
    uint32_t* out = outs + 8 * call;
    for (int k = 0; k < 8; k++) out[k] = 0;
  }
  if (role == 0) not_one[lane] = 0;
  const uint32_t chunk = st->chunk;
  uint32_t nch = 0, base = 0;
  if (live) {
    const uint32_t npairs = (uint32_t)(offsets[call + 1] / 384 - offsets[call] / 384);
    nch = (npairs + chunk - 1) / chunk;
    base = call_first_task[call];
  }
  uint32_t nchmax = nch;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { uint32_t v = __shfl_xor_sync(0xffffffffu, nchmax, o); nchmax = v > nchmax ? v : nchmax; }
  if (nchmax == 0) return;           // no valid call in this block (block-uniform)
  const int mem_idx = dot::mem_of_wpow(role);
  auto words = [&](uint32_t b) { return dot::SmemWords{b * (uint32_t)DOT6_F_WORDS + (uint32_t)lane}; };
  auto ld_coef = [&](uint32_t b) { Fp2 v; v.c0 = words(b).load(2 * role); v.c1 = words(b).load(2 * role + 1); return v; };
  auto st_coef = [&](uint32_t b, const Fp2& v) {
    uint32_t* dst = dot_smem + b * DOT6_F_WORDS + lane;
#pragma unroll
    for (int l = 0; l < 12; l++) { dst[((2 * role) * 12 + l) * 32] = v.c0.v[l]; dst[((2 * role + 1) * 12 + l) * 32] = v.c1.v[l]; }
  };
  auto load_chunk = [&](uint32_t b, uint32_t idx) {       // chunk value idx of this lane's call (1 beyond its last chunk)
    Fp2 v = role == 0 ? fp2_one() : fp2_zero();
    if (idx < nch) v = reinterpret_cast<const Fp2*>(&fchunk[base + idx])[mem_idx];
    st_coef(b, v);
  };
  load_chunk(0, 0);
  __syncthreads();
  // steps: (nchmax - 1) x {load chunk j -> buffer 1, product -> the other of buffers 0 / 2}, a copy back to buffer 0 if
  // the product ended in buffer 2, then the exponent program
  const int pre = 2 * ((int)nchmax - 1);
  const int fix = (nchmax > 1 && ((nchmax - 1) & 1)) ? 1 : 0;
  uint32_t acc = 0;
#pragma unroll 1
  for (int step = 0; step < pre + fix + prog_len; step++) {
    uint32_t ins;
    if (step < pre) {
      if ((step & 1) == 0) { load_chunk(1, (uint32_t)(step / 2 + 1)); __syncthreads(); continue; }
      ins = fx_encode(FX_MUL, acc ^ 2u, acc, 1);
      acc ^= 2u;
    } else if (step < pre + fix) {
      ins = fx_encode(FX_COPY, 0, 2, 0);
    } else {
      ins = prog[step - pre - fix];
    }
    const uint32_t opc = ins & 0xffu, d = (ins >> 8) & 15u, x = (ins >> 12) & 15u, y = (ins >> 16) & 15u;
    if (opc <= FX_CYC) {
      const dot::SmemWords F = words(x), G = words(y);
      const int op = opc == FX_MUL ? dot::OP_MUL : dot::OP_CYC;
      uint32_t* dst = dot_smem + d * DOT6_F_WORDS + lane;
#pragma unroll 1
      for (int q = 0; q < 2; q++) {
        const int c = dot::role_comp(op, role, q);
        Fp r = dot::dot_eval_smem(dot::op_row(op, c), dot::op_count(op, c), F, G);
        if (opc == FX_CYC) r = dot::cyc_epilogue(r, F.load(c), ((c >> 1) & 1) != 0);
#pragma unroll
        for (int l = 0; l < 12; l++) dst[(c * 12 + l) * 32] = r.v[l];
      }
    } else if (opc == FX_CONJ) {
      if (role & 1) st_coef(x, neg(ld_coef(x)));
    } else if (opc == FX_FROB1 || opc == FX_FROB2) {
      Fp2 v = ld_coef(x);
      if (opc == FX_FROB1) v = conj(v);
      st_coef(x, mulo(v, fp2_load_const((opc == FX_FROB1 ? C_FROB1() : C_FROB2()) + 24 * role)));
    } else if (opc == FX_COPY) {
      st_coef(d, ld_coef(x));
    } else if (opc == FX_PARK) {
      if (live) reinterpret_cast<Fp2*>(&park[(size_t)y * n_calls + call])[mem_idx] = ld_coef(x);
    } else if (opc == FX_UNPARK) {
      Fp2 v = role == 0 ? fp2_one() : fp2_zero();
      if (live) v = reinterpret_cast<const Fp2*>(&park[(size_t)y * n_calls + call])[mem_idx];
      st_coef(d, v);
    } else if (opc == FX_INV) {
      if (role == 0) {         // one lane per call runs the tower inversion (one Fp inversion inside)
        Fp12 v, vi;
        Fp2* m = reinterpret_cast<Fp2*>(&v);
        for (int k = 0; k < 6; k++) { m[dot::mem_of_wpow(k)].c0 = words(x).load(2 * k); m[dot::mem_of_wpow(k)].c1 = words(x).load(2 * k + 1); }
        fp12_inv(vi, v);
        const Fp2* mi = reinterpret_cast<const Fp2*>(&vi);
        uint32_t* dst = dot_smem + d * DOT6_F_WORDS + lane;
        for (int k = 0; k < 6; k++) {
          const Fp2& cf = mi[dot::mem_of_wpow(k)];
          for (int l = 0; l < 12; l++) { dst[((2 * k) * 12 + l) * 32] = cf.c0.v[l]; dst[((2 * k + 1) * 12 + l) * 32] = cf.c1.v[l]; }
        }
      }
    }
    __syncthreads();
  }
  // is-one: the program leaves the result in buffer 0
  {
    const Fp2 v = ld_coef(0);
    const bool ok = role == 0 ? eq(v, fp2_one()) : is_zero(v);
    if (!ok) atomicOr(&not_one[lane], 1);
  }
  __syncthreads();
  if (role == 0 && live && not_one[lane] == 0) outs[8 * call + 7] = 0x01000000u;   // byte 31 of the 32-byte big-endian word
}

#endif  // __CUDACC__
}  // namespace b200
