// msm.cuh -- K3 (batched decode) + K4 (Pippenger bucket MSM) kernels, field-generic (G1: Fp, G2: Fp2).
//
// Replaces the reference's MULTIEXP strategies -- k=1 blst_p*_mult, k<=4 naive per-pair
// scalar multiplications, k>4 Bos-Coster heap (/root/reference/src/eip2537.c:541-708 for G1,
// :829-998 for G2) -- with one data-parallel pipeline:
//
//   decode      one thread per pair: wire bytes -> Montgomery affine point + status; the first
//               failing pair (minimum index) decides the error code, as the reference's
//               sequential loops do (:580-592, :650-668)
//   digits      one thread per scalar: full 256-bit scalar (never reduced mod r, :417-420)
//               -> signed c-bit digits (top window unsigned, so no 257th-bit window) +
//               per-bucket histogram
//   scan        exclusive prefix sum of the histogram (bucket segment offsets)
//   scatter     counting-sort the (bucket <- point index | sign) entries into segments
//   accumulate  one thread per bucket: XYZZ += affine point (mixed add, exact exceptional cases)
//   reduce      sum_b b*B_b per window by a tree of (plain sum, weighted sum) nodes
//   final       Horner over windows, to-affine (one inversion), encode to wire bytes
//
// Result = encode(affine(sum k_i * P_i)) in the FULL curve group: no GLV, no mod-r reduction,
// because MULTIEXP inputs are not subgroup-checked.
#pragma once
#include "codec.cuh"
#include "coop.cuh"

namespace b200 {

// Window layout.  The scalar bits (256; 128 on the checked GLV path) are cut into nwin = ceil(bits/c) windows whose widths are c or
// c-1 bits (so every c is usable, not only the divisors of 256).  All windows but the top one use
// signed digits (|d| <= 2^(width-1)); the top window is unsigned and absorbs the last carry, so there
// is no 257th-bit window.  When 256 is not a multiple of c the top window is one of the narrow ones
// and needs exactly 2^(c-1) buckets like the others; when it is (c = 4, 8, 16) the top window has 2^c.
struct MsmPlan {
  int c;          // nominal window width in bits (2..16)
  int nwin;       // ceil(256 / c)
  int log_nb;     // log2(buckets per window), uniform layout
  uint32_t nb;    // buckets per window, power of two
  uint8_t width[128];    // bits of window w
  uint16_t bitpos[128];  // first bit of window w
};

static inline MsmPlan make_plan(int c, int nbits = 256) {
  MsmPlan p;
  p.c = c;
  p.nwin = (nbits + c - 1) / c;
  const int narrow = p.nwin * c - nbits;         // windows that are c-1 bits wide (top window first)
  int pos = 0;
  for (int w = 0; w < p.nwin; w++) {
    // narrow windows: the top one, then every other slot from the top down as needed
    bool is_narrow = (p.nwin - 1 - w) < narrow;
    p.width[w] = (uint8_t)(is_narrow ? c - 1 : c);
  }
  // (widths are assigned so that the narrow windows sit at the top; their sum is exactly nbits)
  for (int w = 0; w < p.nwin; w++) { p.bitpos[w] = (uint16_t)pos; pos += p.width[w]; }
  const int top_w = p.width[p.nwin - 1];
  p.log_nb = (c - 1 > top_w) ? c - 1 : top_w;
  p.nb = 1u << p.log_nb;
  return p;
}

// Upper levels of the bucket reduction (k_reduce_level / k_window_finish): level i folds 2^l_log[i] children that cover
// 2^cov[i] buckets each.  Four children per level (three dependent additions), two when an odd number of bits remains.
static constexpr int REDUCE_MAX_LEVELS = 16;
struct ReduceLevels {
  int n;
  uint8_t l_log[REDUCE_MAX_LEVELS], cov[REDUCE_MAX_LEVELS];
};
static inline ReduceLevels make_reduce_levels(int log_nb, int leaf_log, int inner_log = 2) {
  ReduceLevels lv;
  lv.n = 0;
  int cov = leaf_log;
  while (cov < log_nb) {
    const int rem = log_nb - cov, l = rem < inner_log ? rem : inner_log;
    lv.l_log[lv.n] = (uint8_t)l; lv.cov[lv.n] = (uint8_t)cov; lv.n++;
    cov += l;
  }
  return lv;
}
static inline int reduce_root_width(const ReduceLevels& lv) {      // accumulators per root node
  int k = 2;
  for (int i = 0; i < lv.n; i++) k += (1 << lv.l_log[i]) - 1;
  return k;
}

#ifdef __CUDACC__

static constexpr unsigned long long STATUS_OK = ~0ull;

// ------------------------------------------------------------------------------------------ decode
// status key = (pair index << 8) | code ; atomicMin keeps the first failing pair
// Checked MULTIEXP (opt-in): every point is PROVEN to lie in the r-torsion subgroup, so the fast path
// the reference left as a TODO (eip2537.c:340, :401) becomes legal: scalars are reduced mod r and split as
// k = q*z^2 + t (t, q < 2^128), and k*P = t*P + q*E(P) with the cheap endomorphism E = [z^2]:
//   G1: phi(P) = (beta x, y) = [-z^2]P  =>  E(P) = (beta x, -y);   G2: psi(Q) = [z]Q  =>  E(Q) = psi^2(Q) = (N(cx) x, N(cy) y)
// The pipeline then sees 2n points with 128-bit scalars: the same bucket additions, HALF the windows (half the tail).
__device__ __forceinline__ Affine<Fp> glv_image(const Affine<Fp>& p) {
  Affine<Fp> r;
  r.x = mul(p.x, fp_load_const(C_BETA()));
  r.y = neg(p.y);
  return r;
}
__device__ __forceinline__ Affine<Fp2> glv_image(const Affine<Fp2>& p) {
  const Fp2 cx = fp2_load_const(C_PSI_CX()), cy = fp2_load_const(C_PSI_CY());
  const Fp nx = add(sqr(cx.c0), sqr(cx.c1)), ny = add(sqr(cy.c0), sqr(cy.c1));      // psi^2 multiplies by the norms
  Affine<Fp2> r;
  r.x = mul_fp(p.x, nx);
  r.y = mul_fp(p.y, ny);
  return r;
}
// k (8 little-endian words, any 256-bit value) -> k mod r -> (t, q) with k mod r = q*z^2 + t, 4 words each
__device__ __forceinline__ void glv_split(const uint32_t* k_in, uint32_t* t4, uint32_t* q4) {
  const uint32_t R[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
  const uint32_t Z2[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};
  const uint32_t MU[5] = {0xf6cfee2eu, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u};   // floor(2^256 / z^2)
  uint32_t k[8];
  for (int i = 0; i < 8; i++) k[i] = k_in[i];
  for (int rep = 0; rep < 2; rep++) {          // 2^256 < 3r: at most two subtractions
    uint32_t d[8];
    uint64_t bw = 0;
    for (int i = 0; i < 8; i++) { uint64_t x = (uint64_t)k[i] - R[i] - bw; d[i] = (uint32_t)x; bw = (x >> 32) & 1; }
    if (!bw) for (int i = 0; i < 8; i++) k[i] = d[i];
  }
  // q = floor(k * MU / 2^256) (never above the true quotient, at most 1 below): words 8..11 of the 13-word product
  uint32_t prod[13];
  for (int i = 0; i < 13; i++) prod[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 5; j++) { uint64_t x = (uint64_t)k[i] * MU[j] + prod[i + j] + c; prod[i + j] = (uint32_t)x; c = x >> 32; }
    prod[i + 5] = (uint32_t)c;
  }
  uint32_t q[4] = {prod[8], prod[9], prod[10], prod[11]};
  // t = k - q*z^2 (fits in 5 words before the correction)
  uint32_t qz[8];
  for (int i = 0; i < 8; i++) qz[i] = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) { uint64_t x = (uint64_t)q[i] * Z2[j] + qz[i + j] + c; qz[i + j] = (uint32_t)x; c = x >> 32; }
    qz[i + 4] = (uint32_t)c;
  }
  uint32_t t[5];
  {
    uint64_t bw = 0;
    for (int i = 0; i < 5; i++) { uint64_t x = (uint64_t)k[i] - qz[i] - bw; t[i] = (uint32_t)x; bw = (x >> 32) & 1; }
  }
  for (int rep = 0; rep < 2; rep++) {          // t >= z^2: one more z^2 fits
    uint32_t d[5];
    uint64_t bw = 0;
    for (int i = 0; i < 5; i++) { uint64_t x = (uint64_t)t[i] - (i < 4 ? Z2[i] : 0u) - bw; d[i] = (uint32_t)x; bw = (x >> 32) & 1; }
    if (!bw) {
      for (int i = 0; i < 5; i++) t[i] = d[i];
      uint64_t c = 1;
      for (int i = 0; i < 4; i++) { uint64_t x = (uint64_t)q[i] + c; q[i] = (uint32_t)x; c = x >> 32; }
    }
  }
  for (int i = 0; i < 4; i++) { t4[i] = t[i]; q4[i] = q[i]; }
}

// glv != 0: also writes the endomorphism image E(P_i) at pts[n + i]
template <class F>
__global__ void __launch_bounds__(128) k_decode(const uint32_t* __restrict__ raw, size_t n, Affine<F>* __restrict__ pts,
                                                unsigned long long* status, size_t index_base, int glv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int PW = Wire<F>::POINT_WORDS, SW = Wire<F>::PAIR_WORDS;
  uint32_t w[PW];
  const uint4* src = reinterpret_cast<const uint4*>(raw + i * SW);
#pragma unroll
  for (int k = 0; k < PW / 4; k++) {
    uint4 q = __ldg(src + k);
    w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
  }
  Affine<F> pt;
  int code = decode_point(pt, w);
  if (code != E_SUCCESS) atomicMin(status, ((unsigned long long)(index_base + i) << 8) | (unsigned)code);
  pts[i] = pt;
  if (glv) pts[n + i] = is_inf(pt) ? pt : glv_image(pt);
}

// ------------------------------------------------------------------------------------------ digits
__device__ __forceinline__ uint32_t window_bits(const uint32_t* k, int bit, int c) {
  int word = bit >> 5, sh = bit & 31;
  uint32_t v = k[word] >> sh;
  if (sh + c > 32 && word + 1 < 8) v |= k[word + 1] << (32 - sh);
  return v & ((1u << c) - 1);
}

// digits[w*nv + v] = signed digit of virtual point v in window w (0 = skip); counts[w*nb + |d|-1]++.
// Plain path: nv = n, one 256-bit scalar per pair.  glv: nv = 2n, pair i feeds t to point i and q to point n + i.
template <class F>
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ raw, size_t n, const Affine<F>* __restrict__ pts,
                                                MsmPlan plan, int* __restrict__ digits, uint32_t* counts, int glv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int PW = Wire<F>::POINT_WORDS, SW = Wire<F>::PAIR_WORDS;
  const uint4* src = reinterpret_cast<const uint4*>(raw + i * SW + PW);
  uint4 q0 = __ldg(src), q1 = __ldg(src + 1);
  uint32_t sw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w}, k[8];
  scalar_from_slot(k, sw);
  // a point at infinity contributes nothing whatever its scalar
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(pts + i);
  uint32_t any = 0;
  for (int t = 0; t < (int)(sizeof(Affine<F>) / 4); t++) any |= pw[t];
  const bool inf = (any == 0);
  const size_t nv = glv ? 2 * n : n;
  uint32_t half[2][8];
  if (glv) {
    for (int t = 0; t < 8; t++) { half[0][t] = 0; half[1][t] = 0; }
    glv_split(k, half[0], half[1]);
  }
  for (int part = 0; part < (glv ? 2 : 1); part++) {
    const uint32_t* kk = glv ? half[part] : k;
    const size_t v = i + (size_t)part * n;
    uint32_t carry = 0;
    for (int w = 0; w < plan.nwin; w++) {
      int d;
      const int wd = plan.width[w];
      uint32_t raw_d = window_bits(kk, plan.bitpos[w], wd) + carry;
      if (w < plan.nwin - 1) {
        if (raw_d > (1u << (wd - 1))) { d = (int)raw_d - (int)(1u << wd); carry = 1; }
        else                          { d = (int)raw_d; carry = 0; }
      } else {
        d = (int)raw_d;   // top window: unsigned, <= 2^width
      }
      if (inf) d = 0;
      digits[(size_t)w * nv + v] = d;     // (|d| reaches 2^16 in an unsigned 16-bit top window: 17 bits with the sign)
      if (d != 0) {
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        atomicAdd(&counts[(size_t)w * plan.nb + (mag - 1)], 1u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ scan
// three-kernel exclusive scan over m counters (m <= ~2^21): block-local scan, scan of block sums, fix-up
// (pad_mask = 2^R - 1 rounds every count up to a multiple of 2^R: segments aligned for R rounds of pairwise additions)
__global__ void __launch_bounds__(1024) k_scan_blocks(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                      uint32_t* __restrict__ block_sums, uint32_t m, uint32_t pad_mask) {
  __shared__ uint32_t warp_sums[32];
  uint32_t i = blockIdx.x * 1024u + threadIdx.x;
  uint32_t v = i < m ? ((in[i] + pad_mask) & ~pad_mask) : 0, x = v;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = warp_sums[lane], t = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
    warp_sums[lane] = t - s;
    if (lane == 31) block_sums[blockIdx.x] = t;
  }
  __syncthreads();
  if (i < m) out[i] = x - v + warp_sums[wid];
}
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* block_sums, uint32_t nblocks) {
  // single block, sequential over tiles of 1024
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? block_sums[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t s = warp_sums[lane], t = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      warp_sums[lane] = t - s;
    }
    __syncthreads();
    uint32_t excl = x - v + warp_sums[wid] + carry_s;
    if (i < nblocks) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) k_scan_fix(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_sums, uint32_t m) {
  uint32_t i = blockIdx.x * 1024u + threadIdx.x;
  if (i < m) out[i] += block_sums[blockIdx.x];
}

// ------------------------------------------------------------------------------------------ scatter
// entries[offsets[b] + k] = point index | sign<<31 ; order inside a bucket is irrelevant
// (the group is commutative and the affine result is unique)
__global__ void __launch_bounds__(256) k_scatter(const int* __restrict__ digits, size_t n, MsmPlan plan,
                                                 const uint32_t* __restrict__ offsets, uint32_t* cursors,
                                                 uint32_t* __restrict__ entries) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)plan.nwin) return;
  size_t w = t / n, i = t - w * n;
  int d = digits[t];
  if (d == 0) return;
  uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
  size_t b = w * plan.nb + (mag - 1);
  uint32_t pos = atomicAdd(&cursors[b], 1u);
  entries[offsets[b] + pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
}

// ------------------------------------------------------------------------------------------ bucket order
// Buckets are processed in order of decreasing size so the 32 lanes of a warp walk segments of
// (nearly) equal length: a counting sort of bucket ids by min(count, ORDER_BINS-1), block-local
// shared-memory histograms, one global atomic per (block, bin).  The same pass plans the
// "overflow" work of oversized buckets (skewed / adversarial scalars, e.g. all scalars equal):
// a bucket thread only ever walks its first `cap` entries; the rest is cut into tasks of <= cap
// entries, accumulated by other threads and merged by one warp per oversized bucket.
static constexpr int ORDER_BINS = 1024;
struct BigBucket { uint32_t bucket, first_task, ntasks, pad; };
struct OverflowTask { uint32_t bucket, start, len, pad; };
struct OrderCounters { uint32_t ntasks, nbig; };

__global__ void __launch_bounds__(1024) k_order_hist(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                     uint32_t nbt, uint32_t cap, uint32_t* bin_total, OrderCounters* oc,
                                                     BigBucket* big, OverflowTask* tasks) {
  __shared__ uint32_t h[ORDER_BINS];
  for (int i = threadIdx.x; i < ORDER_BINS; i += 1024) h[i] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * 1024u + threadIdx.x;
  if (b < nbt) {
    uint32_t cnt = counts[b];
    atomicAdd(&h[cnt < ORDER_BINS - 1 ? cnt : ORDER_BINS - 1], 1u);
    if (cnt > cap) {
      uint32_t extra = cnt - cap, nt = (extra + cap - 1) / cap;
      uint32_t first = atomicAdd(&oc->ntasks, nt), slot = atomicAdd(&oc->nbig, 1u);
      big[slot] = BigBucket{b, first, nt, 0};
      uint32_t start = offsets[b] + cap;
      for (uint32_t t = 0; t < nt; t++) {
        uint32_t len = extra - t * cap < cap ? extra - t * cap : cap;
        tasks[first + t] = OverflowTask{b, start + t * cap, len, 0};
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ORDER_BINS; i += 1024)
    if (h[i]) atomicAdd(&bin_total[i], h[i]);
}
// bin_start[bin] = number of buckets in strictly larger bins (descending order); one block
__global__ void __launch_bounds__(1024) k_order_scan(const uint32_t* __restrict__ bin_total, uint32_t* __restrict__ bin_start) {
  __shared__ uint32_t warp_sums[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int bin = ORDER_BINS - 1 - threadIdx.x;
  uint32_t v = bin_total[bin], x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = warp_sums[lane], t = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
    warp_sums[lane] = t - s;
  }
  __syncthreads();
  bin_start[bin] = x - v + warp_sums[wid];
}
__global__ void __launch_bounds__(1024) k_order_scatter(const uint32_t* __restrict__ counts, uint32_t nbt, uint32_t* bin_cursor,
                                                        uint32_t* __restrict__ order) {
  __shared__ uint32_t h[ORDER_BINS], base[ORDER_BINS];
  for (int i = threadIdx.x; i < ORDER_BINS; i += 1024) h[i] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * 1024u + threadIdx.x, bin = 0, local = 0;
  if (b < nbt) {
    uint32_t cnt = counts[b];
    bin = cnt < ORDER_BINS - 1 ? cnt : ORDER_BINS - 1;
    local = atomicAdd(&h[bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ORDER_BINS; i += 1024)
    if (h[i]) base[i] = atomicAdd(&bin_cursor[i], h[i]);
  __syncthreads();
  if (b < nbt) order[base[bin] + local] = b;
}

// ------------------------------------------------------------------------------------------ accumulate
template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F>* p) {
  Affine<F> r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) d[k] = __ldg(s + k);
  return r;
}

// DIRECT: the segment is a run of affine partial sums (output of the pair rounds; (0,0) = infinity) instead of entries
template <class F, int ROWS = 2, bool DIRECT = false>
__device__ __forceinline__ void accumulate_segment(XYZZ<F>& acc, const Affine<F>* __restrict__ pts,
                                                   const uint32_t* __restrict__ entries, uint32_t start, uint32_t cnt) {
  for (uint32_t k = 0; k < cnt; k++) {
    Affine<F> pt;
    if constexpr (DIRECT) {
      pt = load_affine(pts + start + k);
      if (is_inf(pt)) continue;
    } else {
      uint32_t e = __ldg(entries + start + k);
      pt = load_affine(pts + (e & 0x7fffffffu));
      if (e >> 31) pt.y = neg(pt.y);
    }
    if constexpr (ROWS != 2 && sizeof(F) == sizeof(Fp)) xyzz_madd_unrolled<ROWS>(acc, pt);
    else xyzz_madd(acc, pt);
  }
}

// ---- batched-affine pair rounds (opt-in: B200_AFFINE_ROUNDS) --------------------------------------------------------------
// With segment offsets aligned to 2^R (k_scan_blocks pad_mask) and the padding slots holding ENTRY_NONE, the entries of a
// bucket can be summed PAIRWISE before the sequential XYZZ walk: round 1 adds entries (2q, 2q+1) -> affine partial sum q,
// round 2 adds partial sums (2q, 2q+1); a pair never straddles a bucket boundary, and bucket b's segment after R rounds is
// [offsets[b] >> R, +padded_count >> R).  Each thread owns PAIR_BATCH output slots and the 32 lanes of a warp share ONE
// inversion (ec.cuh pair_prepare / pair_finish: ~7 multiplications per addition, the trick included, instead of the 10 of the
// mixed XYZZ addition).  Measured (G1 2^20, profiles/r02_msm_tail.md): the round runs at 65 % of the multiply pipe
// (two random 96-byte gathers per addition in each of its two passes) and the stage ends at break-even, so the default
// pipeline does not use it.
static constexpr uint32_t ENTRY_NONE = 0xFFFFFFFFu;
static constexpr int PAIR_BATCH = 64;       // output slots per thread: 2048 additions share one inversion per warp
template <class F>
__device__ __forceinline__ Affine<F> affine_inf() { Affine<F> r; r.x = FieldOps<F>::zero(); r.y = FieldOps<F>::zero(); return r; }
template <class F, bool FIRST>
__device__ __forceinline__ void pair_load(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ entries, uint32_t q,
                                          Affine<F>& a, Affine<F>& b) {
  if constexpr (FIRST) {
    const uint2 e = __ldg(reinterpret_cast<const uint2*>(entries) + q);
    a = affine_inf<F>(); b = affine_inf<F>();
    if (e.x != ENTRY_NONE) { a = load_affine(src + (e.x & 0x7fffffffu)); if (e.x >> 31) a.y = neg(a.y); }
    if (e.y != ENTRY_NONE) { b = load_affine(src + (e.y & 0x7fffffffu)); if (e.y >> 31) b.y = neg(b.y); }
  } else {
    a = load_affine(src + 2 * (size_t)q);
    b = load_affine(src + 2 * (size_t)q + 1);
  }
}
// total_padded = sorted entries incl. alignment padding (device scalar); this round reads total_padded >> shift slots.
// ONE inversion per WARP: the 32 lanes' products are multiplied up a butterfly (every lane keeps the partner products it
// met: they are exactly the other lanes' contribution), all lanes invert the same total (no divergence), and each lane
// recovers its own inverse with five more multiplications.
template <class F>
__device__ __forceinline__ F shfl_xor_field(const F& v, int mask) {
  F r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask);
  return r;
}
template <class F, bool FIRST, int BATCH>
__global__ void __launch_bounds__(128) k_pair_round(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ entries,
                                                    const uint32_t* __restrict__ total_padded, int shift, Affine<F>* __restrict__ out) {
  const uint32_t n_out = (*total_padded >> shift) >> 1;
  const uint32_t base = blockIdx.x * (128u * BATCH) + threadIdx.x;
  if (base - (threadIdx.x & 31u) >= n_out) return;          // the whole warp is past the end
  F pre[BATCH];
  F run = FieldOps<F>::one();
#pragma unroll 1
  for (int i = 0; i < BATCH; i++) {
    const uint32_t q = base + (uint32_t)i * 128u;
    pre[i] = run;
    if (q < n_out) {
      Affine<F> a, b;
      pair_load<F, FIRST>(src, entries, q, a, b);
      F den;
      pair_prepare(a, b, den);
      run = mul(run, den);
    }
  }
  F inv_run;
  {
    F other[5];
    F cur = run;
#pragma unroll
    for (int k = 0; k < 5; k++) { other[k] = shfl_xor_field(cur, 1 << k); cur = mul(cur, other[k]); }
    inv_run = inv(cur);                                      // the same value on all 32 lanes
#pragma unroll
    for (int k = 0; k < 5; k++) inv_run = mul(inv_run, other[k]);
  }
#pragma unroll 1
  for (int i = BATCH - 1; i >= 0; i--) {
    const uint32_t q = base + (uint32_t)i * 128u;
    if (q >= n_out) continue;
    Affine<F> a, b;
    pair_load<F, FIRST>(src, entries, q, a, b);
    F den;
    const int kind = pair_prepare(a, b, den);
    const F inv_den = mul(inv_run, pre[i]);
    inv_run = mul(inv_run, den);
    const Affine<F> r = pair_finish(kind, a, b, inv_den);
    uint4* d = reinterpret_cast<uint4*>(out + q);
    const uint4* sv = reinterpret_cast<const uint4*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) d[k] = sv[k];
  }
}
// counts / offsets of the segments after R rounds, and the padded total (one thread per bucket)
__global__ void __launch_bounds__(256) k_pair_layout(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t nbt,
                                                     int R, uint32_t* __restrict__ counts_r, uint32_t* __restrict__ offsets_r) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbt) return;
  const uint32_t mask = (1u << R) - 1;
  counts_r[b] = (counts[b] + mask) >> R;
  offsets_r[b] = offsets[b] >> R;
}
__global__ void k_pair_total(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t nbt, int R,
                             uint32_t* __restrict__ total_padded) {
  const uint32_t mask = (1u << R) - 1;
  *total_padded = offsets[nbt - 1] + ((counts[nbt - 1] + mask) & ~mask);
}

// one thread per bucket, buckets taken in order of decreasing size; at most `cap` entries each
// MINB: resident blocks per SM the register budget is sized for (4 -> 128 registers, 3 -> 168, 2 -> 255)
template <class F, int ROWS = 2, int MINB = 4, bool DIRECT = false>
__global__ void __launch_bounds__(128, MINB) k_accumulate(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ entries,
                                                    const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                                                    const uint32_t* __restrict__ order, uint32_t nbuckets_total, uint32_t cap,
                                                    int add_to_existing, XYZZ<F>* __restrict__ buckets) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbuckets_total) return;
  uint32_t b = order[t];
  uint32_t cnt = counts[b];
  if (add_to_existing && cnt == 0) return;   // streamed chunks: the bucket keeps what earlier chunks left
  XYZZ<F> acc;
  if (add_to_existing) acc = buckets[b]; else acc = xyzz_inf<F>();
  accumulate_segment<F, ROWS, DIRECT>(acc, pts, entries, offsets[b], cnt < cap ? cnt : cap);
  buckets[b] = acc;
}
// one thread per overflow task
template <class F, bool DIRECT = false>
__global__ void __launch_bounds__(128) k_accumulate_overflow(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ entries,
                                                             const OverflowTask* __restrict__ tasks, const OrderCounters* __restrict__ oc,
                                                             XYZZ<F>* __restrict__ partials) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= oc->ntasks) return;
  OverflowTask task = tasks[t];
  XYZZ<F> acc = xyzz_inf<F>();
  accumulate_segment<F, 2, DIRECT>(acc, pts, entries, task.start, task.len);
  partials[t] = acc;
}
// one warp per oversized bucket: lanes sum strided subsets of its partials, then a shuffle tree
template <class F>
__global__ void __launch_bounds__(128) k_merge_overflow(const BigBucket* __restrict__ big, const OrderCounters* __restrict__ oc,
                                                        const XYZZ<F>* __restrict__ partials, XYZZ<F>* __restrict__ buckets) {
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= oc->nbig) return;
  BigBucket bb = big[w];
  XYZZ<F> acc = xyzz_inf<F>();
  for (uint32_t t = lane; t < bb.ntasks; t += 32) { XYZZ<F> p = partials[bb.first_task + t]; xyzz_add(acc, p); }
  for (int o = 16; o >= 1; o >>= 1) {
    XYZZ<F> other;
    uint32_t* dst = reinterpret_cast<uint32_t*>(&other);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&acc);
    for (int k = 0; k < (int)(sizeof(XYZZ<F>) / 4); k++) dst[k] = __shfl_down_sync(0xffffffffu, src[k], o);
    if (lane < o) xyzz_add(acc, other);
  }
  if (lane == 0) { XYZZ<F> cur = buckets[bb.bucket]; xyzz_add(cur, acc); buckets[bb.bucket] = cur; }
}

// ------------------------------------------------------------------------------------------ reduce
// A node covers a power-of-two range of consecutive buckets of one window and carries
//   s = sum of its buckets,  w = sum over its buckets of (1-based offset inside the node) * bucket
template <class F>
struct Node { XYZZ<F> s, w; };

// level 0: each thread folds `L` consecutive buckets (bucket magnitudes lo+1 .. lo+L) by a running sum
template <class F>
__global__ void __launch_bounds__(128, (sizeof(F) == sizeof(Fp) ? 4 : 2)) k_reduce_leaf(const XYZZ<F>* __restrict__ buckets, uint32_t nnodes_total, int L,
                                                     Node<F>* __restrict__ nodes) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnodes_total) return;
  const XYZZ<F>* b = buckets + (size_t)t * L;
  XYZZ<F> run = xyzz_inf<F>(), acc = xyzz_inf<F>();
  for (int j = L - 1; j >= 0; j--) {
    XYZZ<F> v = b[j];
    xyzz_add(run, v);
    xyzz_add(acc, run);
  }
  nodes[t].s = run;
  nodes[t].w = acc;
}
// Upper levels WITHOUT doublings.  Folding L children that cover 2^cov buckets each needs
//   w_parent = sum_t w_t + 2^cov * sum_t t*s_t ;
// round 1 evaluated this per node (a running sum plus `cov` doublings at every level: ~80 dependent additions and ~48
// doublings from the leaves to the root, 1.1 of the 1.8 ms of this stage).  The scaled term is linear, so it can be
// summed over ALL nodes of a level before it is scaled: a node carries a vector of plain sums
//   acc[0] = s,  acc[1] = sum of the leaf w's,  then for every level i below it and t = 1..L_i-1:
//   a_{i,t} = sum of s over the level-i children with index t
// a level only ADDS vectors component-wise (one lane group per component: 3 dependent additions per level) and appends
// its own a_{.,t} = s of child t; the root's  w + sum_i 2^cov_i * sum_t t*a_{i,t}  is one Horner walk per window
// (k_window_finish: cov_top doublings in all).
// COOP = false: one THREAD per component (the wide lower levels are throughput-bound, and a lane group spends 32 multiplication
// slots on an addition a thread does in 14); COOP = true: one lane group per component (narrow upper levels: latency).
template <class F, bool COOP>
__global__ void __launch_bounds__(128) k_reduce_level(const XYZZ<F>* __restrict__ in, int K_in, uint32_t nout_total, int l_log,
                                                      XYZZ<F>* __restrict__ out) {
  const int L = 1 << l_log, K_out = K_in + L - 1;
  const unsigned long long gid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) / (COOP ? Coop<F>::LANES : 1);
  if (gid >= (unsigned long long)nout_total * K_out) return;
  const uint32_t t = (uint32_t)(gid / K_out);
  const int k = (int)(gid % K_out);
  const XYZZ<F>* ch = in + (size_t)t * L * K_in;
  XYZZ<F> acc;
  if (k < K_in) {
    acc = ch[k];
    if (COOP) {
      const CoopGroup g = coop_group<F>();
      for (int j = 1; j < L; j++) { XYZZ<F> c = ch[(size_t)j * K_in + k]; coop_add(acc, c, g); }
    } else {
      for (int j = 1; j < L; j++) { XYZZ<F> c = ch[(size_t)j * K_in + k]; xyzz_add(acc, c); }
    }
  } else {
    acc = ch[(size_t)(k - K_in + 1) * K_in];      // a_{this level, t} = s of child t
  }
  if (COOP) {
    const CoopGroup g = coop_group<F>();
    if (g.lane != 0 || g.sub != 0) return;
  }
  out[(size_t)t * K_out + k] = acc;
}
// One block of 8 lane groups per window: the groups form U_i = sum_t t*a_{i,t} in parallel and convert them (and w) to
// homogeneous coordinates, then group 0 walks T = w + 2^cov_0 (U_0 + 2^(cov_1 - cov_0) (U_1 + ...)) with the complete
// two-level formulas of coop.cuh (coop_hadd / coop_hdbl).
template <class F>
__global__ void k_window_finish(const XYZZ<F>* __restrict__ roots, int K, ReduceLevels lv, Hom<F>* __restrict__ tw) {
  __shared__ Hom<F> U[REDUCE_MAX_LEVELS + 1];
  const CoopGroup g = coop_group<F>();
  const int grp = threadIdx.x / Coop<F>::LANES, ngrp = blockDim.x / Coop<F>::LANES;
  const XYZZ<F>* r = roots + (size_t)blockIdx.x * K;
  int base = 2;
  for (int i = 0; i <= lv.n; i++) {
    const int L = i < lv.n ? 1 << lv.l_log[i] : 0;
    if (i % ngrp == grp) {
      XYZZ<F> u;
      if (i == lv.n) {
        u = r[1];                                 // sum of the leaf w's
      } else if (L == 2) {
        u = r[base];
      } else if (L == 4) {                        // a1 + 2 a2 + 3 a3 = (a1 + a3) + 2 (a2 + a3)
        XYZZ<F> a3 = r[base + 2], y = r[base];
        u = r[base + 1];
        coop_add(u, a3, g); coop_dbl(u, g);
        coop_add(y, a3, g); coop_add(u, y, g);
      } else {                                    // running sum from the top child down
        XYZZ<F> run = xyzz_inf<F>();
        u = xyzz_inf<F>();
        for (int t = L - 1; t >= 1; t--) { XYZZ<F> a = r[base + t - 1]; coop_add(run, a, g); coop_add(u, run, g); }
      }
      const Hom<F> h = coop_xyzz_to_hom(u, g);
      if (g.lane == 0 && g.sub == 0) U[i] = h;
    }
    base += L - 1;
  }
  __syncthreads();
  if (grp != 0) return;
  Hom<F> acc = hom_inf<F>();
  for (int i = lv.n - 1; i >= 0; i--) {
    const Hom<F> u = U[i];
    coop_hadd(acc, u, g);
    const int nd = lv.cov[i] - (i > 0 ? lv.cov[i - 1] : 0);
    for (int k = 0; k < nd; k++) coop_hdbl(acc, g);
  }
  const Hom<F> w = U[lv.n];
  coop_hadd(acc, w, g);
  if (g.lane == 0 && g.sub == 0) tw[blockIdx.x] = acc;
}

// Horner over the windows from the top down: acc = 2^width[w] * acc + T_w, in homogeneous coordinates (two dependency
// levels per doubling instead of XYZZ's three).  One cooperative lane group walks the ~240 sequential doublings.
// The result (and, when asked, a copy of the shard's first-error key) may be written through a PEER pointer into
// another GPU's gather buffer: the multi-GPU exchange is fused into this last kernel of the shard's pipeline.
template <class F>
__global__ void k_window_combine(const Hom<F>* __restrict__ tw, MsmPlan plan, XYZZ<F>* __restrict__ acc_io,
                                 const unsigned long long* __restrict__ status_src, unsigned long long* __restrict__ status_dst) {
  if (blockIdx.x != 0 || threadIdx.x >= Coop<F>::LANES) return;
  if (threadIdx.x == 0 && status_dst) *status_dst = *status_src;
  const CoopGroup g = coop_group<F>();
  Hom<F> acc = tw[plan.nwin - 1];
  for (int w = plan.nwin - 2; w >= 0; w--) {
    for (int k = 0; k < plan.width[w]; k++) coop_hdbl(acc, g);
    const Hom<F> t = tw[w];
    coop_hadd(acc, t, g);
  }
  const XYZZ<F> out = coop_hom_to_xyzz(acc, g);
  if (g.lane == 0 && g.sub == 0) *acc_io = out;
}

// sum `count` partial results (multi-GPU gather, or count = 1), convert to affine, encode
template <class F>
__global__ void k_finalize(const XYZZ<F>* __restrict__ partials, int count, uint32_t* __restrict__ out_words) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  XYZZ<F> acc = partials[0];
  for (int i = 1; i < count; i++) { XYZZ<F> t = partials[i]; xyzz_add(acc, t); }
  Affine<F> a = xyzz_to_affine(acc);
  uint32_t w[Wire<F>::POINT_WORDS];
  encode_point(w, a);
  for (int i = 0; i < Wire<F>::POINT_WORDS; i++) out_words[i] = w[i];
}

// One shard's contribution to a multi-GPU MULTIEXP: XYZZ partial sum (192 / 384 B) + first-error key, 512 B
struct alignas(16) ShardRecord {
  unsigned char partial[448];
  unsigned long long status;
  unsigned long long pad[7];
};
// sum the shards' partials unless any shard reported an error (minimum key = first failing pair in input order,
// eip2537.c:580-592 / :650-668), convert to affine, encode
template <class F>
__global__ void k_finalize_records(const ShardRecord* __restrict__ recs, int count, uint32_t* __restrict__ out_words,
                                   unsigned long long* __restrict__ status_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long st = STATUS_OK;
  for (int i = 0; i < count; i++) st = recs[i].status < st ? recs[i].status : st;
  *status_out = st;
  if (st != STATUS_OK) return;
  XYZZ<F> acc = *reinterpret_cast<const XYZZ<F>*>(recs[0].partial);
  for (int i = 1; i < count; i++) { XYZZ<F> t = *reinterpret_cast<const XYZZ<F>*>(recs[i].partial); xyzz_add(acc, t); }
  Affine<F> a = xyzz_to_affine(acc);
  uint32_t w[Wire<F>::POINT_WORDS];
  encode_point(w, a);
  for (int i = 0; i < Wire<F>::POINT_WORDS; i++) out_words[i] = w[i];
}

// ------------------------------------------------------------------------------------------ batch of small calls
// Many independent MULTIEXP calls in one submission (the EVM shape: thousands of calls of a few to a
// few hundred pairs).  Pippenger's bucket machinery does not pay at this size and its serial tail costs
// ~2 ms per call; here every PAIR is one thread doing its own 256-bit double-and-add (the reference's
// naive strategy, eip2537.c:564-616, made data-parallel across all pairs of all calls), then one warp
// per call adds the per-pair results, converts to affine once and encodes.
template <class F>
__global__ void __launch_bounds__(128) k_batch_pair_mul(const uint32_t* __restrict__ raw, size_t total_pairs,
                                                        XYZZ<F>* __restrict__ partial, int* __restrict__ codes) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total_pairs) return;
  constexpr int PW = Wire<F>::POINT_WORDS, SW = Wire<F>::PAIR_WORDS;
  uint32_t w[SW];
  const uint4* src = reinterpret_cast<const uint4*>(raw + j * SW);
#pragma unroll
  for (int k = 0; k < SW / 4; k++) { uint4 q = __ldg(src + k); w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w; }
  Affine<F> pt;
  int code = decode_point(pt, w);
  codes[j] = code;
  if (code != E_SUCCESS) return;
  uint32_t k[8];
  scalar_from_slot(k, w + PW);
  partial[j] = xyzz_scalar_mul(pt, k, 256);
}
// offsets: byte offsets of the calls (multiples of the pair size, validated by the host layer)
template <class F>
__global__ void __launch_bounds__(128) k_batch_call_sum(const unsigned long long* __restrict__ offsets, size_t n_calls,
                                                        const XYZZ<F>* __restrict__ partial, const int* __restrict__ codes,
                                                        uint32_t* __restrict__ outs, int* __restrict__ errs) {
  const size_t call = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (call >= n_calls) return;
  constexpr int PW = Wire<F>::POINT_WORDS, PAIR_BYTES = Wire<F>::PAIR_WORDS * 4;
  const size_t first = (size_t)(offsets[call] / PAIR_BYTES), last = (size_t)(offsets[call + 1] / PAIR_BYTES);
  // first failing pair in input order (eip2537.c:580-592)
  unsigned first_bad = 0xffffffffu;
  for (size_t j = first + lane; j < last; j += 32)
    if (codes[j] != E_SUCCESS) { first_bad = (unsigned)(j - first); break; }
  for (int o = 16; o >= 1; o >>= 1) { unsigned other = __shfl_xor_sync(0xffffffffu, first_bad, o); first_bad = other < first_bad ? other : first_bad; }
  uint32_t* out = outs + call * PW;
  if (first == last || first_bad != 0xffffffffu) {
    if (lane == 0) errs[call] = first == last ? E_INVALID_LENGTH : codes[first + first_bad];
    for (int i = lane; i < PW; i += 32) out[i] = 0;
    return;
  }
  XYZZ<F> acc = xyzz_inf<F>();
  for (size_t j = first + lane; j < last; j += 32) { XYZZ<F> p = partial[j]; xyzz_add(acc, p); }
  for (int o = 16; o >= 1; o >>= 1) {
    XYZZ<F> other;
    uint32_t* dst = reinterpret_cast<uint32_t*>(&other);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&acc);
    for (int k = 0; k < (int)(sizeof(XYZZ<F>) / 4); k++) dst[k] = __shfl_down_sync(0xffffffffu, s[k], o);
    if (lane < o) xyzz_add(acc, other);
  }
  if (lane == 0) {
    Affine<F> a = xyzz_to_affine(acc);
    uint32_t w[PW];
    encode_point(w, a);
    for (int i = 0; i < PW; i++) out[i] = w[i];
    errs[call] = E_SUCCESS;
  }
}

#endif  // __CUDACC__
}  // namespace b200
