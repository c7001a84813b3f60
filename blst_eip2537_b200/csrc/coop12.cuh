// coop12.cuh -- warp-cooperative Fp12 arithmetic for LATENCY-bound pairing work.
//
// A single bls12_pairing call (or a small batch) has no data parallelism: one thread per call walks
// ~20,000 dependent Fp multiplications at 1.6 us each.  Here one WARP owns one Fp12 computation: the
// Fp12 values live in shared memory as six Fp2 coefficients, every Fp12-level operation is expressed as
// a list of independent Fp2 products ("slots": operand expressions over the coefficients) followed by
// a linear output stage, and the warp executes 8 product slots at a time, each slot on 4 lanes that
// split the three Fp multiplications of a Karatsuba Fp2 product (coop_product).  A sparse line
// multiplication takes 2 rounds of ~1 Fp-multiplication latency instead of 39 sequential ones.
//
// The operation tables (struct Op*) are plain B200_HD code, so tests/host_emul runs the very same
// tables sequentially on the CPU and checks them against the thread-level functions of pairing.cuh.
// Coefficient order = memory order of Fp12: f[0..2] = c0.{c0,c1,c2}, f[3..5] = c1.{c0,c1,c2}.
#pragma once
#include "pairing.cuh"
#include "coop.cuh"

namespace b200 {

// ---- Karatsuba helpers over coefficient triples -----------------------------------------------
// operand of product s (0..5) of X*Y, X = (x0,x1,x2):  x0, x1, x2, x1+x2, x0+x1, x0+x2
B200_HD Fp2 kara_operand(int s, const Fp2& x0, const Fp2& x1, const Fp2& x2) {
  switch (s) {
    case 0: return x0;
    case 1: return x1;
    case 2: return x2;
    case 3: return add(x1, x2);
    case 4: return add(x0, x1);
    default: return add(x0, x2);
  }
}
// coefficient k of X*Y from its six Karatsuba products t[0..5]
B200_HD Fp2 kara_output(int k, const Fp2* t) {
  switch (k) {
    case 0: return add(t[0], mul_xi(sub(sub(t[3], t[1]), t[2])));
    case 1: return add(sub(sub(t[4], t[0]), t[1]), mul_xi(t[2]));
    default: return add(sub(sub(t[5], t[0]), t[2]), t[1]);
  }
}
// coefficient k of v*X
B200_HD Fp2 mulv_coef(int k, const Fp2& x0, const Fp2& x1, const Fp2& x2) {
  return k == 0 ? mul_xi(x2) : (k == 1 ? x0 : x1);
}

// Every Op has: NP products, NO = 6 outputs;
//   operands(s, f, g, x, y): the two factors of product s   (f: in/out value, g: second operand)
//   output(j, P, f, g): new coefficient j from the products P[0..NP)
struct OpSqr {   // f <- f^2  (complex squaring: ab = A*B, st = (A+B)(A+vB); c = st - ab - v*ab, d = 2ab)
  static constexpr int NP = 12;
  static B200_HD void operands(int s, const Fp2* f, const Fp2*, Fp2& x, Fp2& y) {
    if (s < 6) { x = kara_operand(s, f[0], f[1], f[2]); y = kara_operand(s, f[3], f[4], f[5]); return; }
    Fp2 p0 = add(f[0], f[3]), p1 = add(f[1], f[4]), p2 = add(f[2], f[5]);
    Fp2 q0 = add(f[0], mul_xi(f[5])), q1 = add(f[1], f[3]), q2 = add(f[2], f[4]);
    x = kara_operand(s - 6, p0, p1, p2);
    y = kara_operand(s - 6, q0, q1, q2);
  }
  static B200_HD Fp2 output(int j, const Fp2* P, const Fp2*, const Fp2*) {
    if (j >= 3) return dbl(kara_output(j - 3, P));
    Fp2 ab0 = kara_output(0, P), ab1 = kara_output(1, P), ab2 = kara_output(2, P);
    Fp2 ab = j == 0 ? ab0 : (j == 1 ? ab1 : ab2);
    return sub(sub(kara_output(j, P + 6), ab), mulv_coef(j, ab0, ab1, ab2));
  }
};

struct OpMul {   // f <- f*g  (t0 = A*C, t1 = B*D, t2 = (A+B)(C+D); c = t0 + v*t1, d = t2 - t0 - t1)
  static constexpr int NP = 18;
  static B200_HD void operands(int s, const Fp2* f, const Fp2* g, Fp2& x, Fp2& y) {
    if (s < 6)       { x = kara_operand(s, f[0], f[1], f[2]); y = kara_operand(s, g[0], g[1], g[2]); }
    else if (s < 12) { x = kara_operand(s - 6, f[3], f[4], f[5]); y = kara_operand(s - 6, g[3], g[4], g[5]); }
    else {
      x = kara_operand(s - 12, add(f[0], f[3]), add(f[1], f[4]), add(f[2], f[5]));
      y = kara_operand(s - 12, add(g[0], g[3]), add(g[1], g[4]), add(g[2], g[5]));
    }
  }
  static B200_HD Fp2 output(int j, const Fp2* P, const Fp2*, const Fp2*) {
    if (j < 3) {
      Fp2 u0 = kara_output(0, P + 6), u1 = kara_output(1, P + 6), u2 = kara_output(2, P + 6);
      return add(kara_output(j, P), mulv_coef(j, u0, u1, u2));
    }
    return sub(sub(kara_output(j - 3, P + 12), kara_output(j - 3, P)), kara_output(j - 3, P + 6));
  }
};

struct OpMul014 {   // f <- f * (l0 + l1 v + l4 v w),  g = {l0, l1, l4}
  static constexpr int NP = 13;
  // products of X * (b0 + b1 v): x0*b0, x1*b1, (x1+x2)*b1, (x0+x1)*(b0+b1), (x0+x2)*b0
  static B200_HD void by01(int s, const Fp2& x0, const Fp2& x1, const Fp2& x2, const Fp2& b0, const Fp2& b1, Fp2& x, Fp2& y) {
    switch (s) {
      case 0: x = x0; y = b0; break;
      case 1: x = x1; y = b1; break;
      case 2: x = add(x1, x2); y = b1; break;
      case 3: x = add(x0, x1); y = add(b0, b1); break;
      default: x = add(x0, x2); y = b0; break;
    }
  }
  static B200_HD Fp2 by01_out(int k, const Fp2* p) {
    switch (k) {
      case 0: return add(mul_xi(sub(p[2], p[1])), p[0]);
      case 1: return sub(sub(p[3], p[0]), p[1]);
      default: return add(sub(p[4], p[0]), p[1]);
    }
  }
  static B200_HD void operands(int s, const Fp2* f, const Fp2* g, Fp2& x, Fp2& y) {
    if (s < 5)       by01(s, f[0], f[1], f[2], g[0], g[1], x, y);                       // aa = A*(l0,l1)
    else if (s < 8)  { x = f[3 + (s - 5)]; y = g[2]; }                                   // B_k * l4
    else             by01(s - 8, add(f[0], f[3]), add(f[1], f[4]), add(f[2], f[5]), g[0], add(g[1], g[2]), x, y);
  }
  static B200_HD Fp2 output(int j, const Fp2* P, const Fp2*, const Fp2*) {
    // bb = B*(l4 v) = (xi*b2l4, b0l4, b1l4) with P[5..7] = b0l4, b1l4, b2l4
    if (j < 3) {
      // c = aa + v*bb ;  v*bb = (xi*bb2, bb0, bb1) = (xi*b1l4, xi*b2l4, b0l4)
      Fp2 vbb = j == 0 ? mul_xi(P[6]) : (j == 1 ? mul_xi(P[7]) : P[5]);
      return add(by01_out(j, P), vbb);
    }
    int k = j - 3;
    Fp2 bb = k == 0 ? mul_xi(P[7]) : (k == 1 ? P[5] : P[6]);
    return sub(sub(by01_out(k, P + 8), by01_out(k, P)), bb);
  }
};

struct OpCycSqr {   // Granger-Scott squaring in the cyclotomic subgroup: 9 Fp2 squarings
  static constexpr int NP = 9;
  // (z0,z1) = (f0,f4), (z2,z3) = (f3,f2), (z4,z5) = (f1,f5); products: a^2, b^2, (a+b)^2 per pair
  static B200_HD void operands(int s, const Fp2* f, const Fp2*, Fp2& x, Fp2& y) {
    const int pa[3] = {0, 3, 1}, pb[3] = {4, 2, 5};
    int q = s / 3, r = s % 3;
    x = r == 0 ? f[pa[q]] : (r == 1 ? f[pb[q]] : add(f[pa[q]], f[pb[q]]));
    y = x;
  }
  static B200_HD Fp2 output(int j, const Fp2* P, const Fp2* f, const Fp2*) {
    // fp4_sqr(a,b) -> r0 = xi*b^2 + a^2, r1 = (a+b)^2 - a^2 - b^2
    auto r0 = [&](int q) { return add(mul_xi(P[3 * q + 1]), P[3 * q]); };
    auto r1 = [&](int q) { return sub(sub(P[3 * q + 2], P[3 * q]), P[3 * q + 1]); };
    switch (j) {
      case 0: { Fp2 t = r0(0); return add(dbl(sub(t, f[0])), t); }            // z0
      case 4: { Fp2 t = r1(0); return add(dbl(add(t, f[4])), t); }            // z1
      case 1: { Fp2 t = r0(1); return add(dbl(sub(t, f[1])), t); }            // z4 = 3*u0 - 2*z4
      case 5: { Fp2 t = r1(1); return add(dbl(add(t, f[5])), t); }            // z5
      case 3: { Fp2 t = mul_xi(r1(2)); return add(dbl(add(t, f[3])), t); }    // z2
      default: { Fp2 t = r0(2); return add(dbl(sub(t, f[2])), t); }           // z3
    }
  }
};

template <int K>
struct OpFrob {   // f <- f^(p^K), K = 1 or 2: coefficient-wise (conjugate for odd K) times gamma_K[w-power]
  static constexpr int NP = 6;
  static B200_HD void operands(int s, const Fp2* f, const Fp2*, Fp2& x, Fp2& y) {
    const int wpow[6] = {0, 2, 4, 1, 3, 5};
    x = (K & 1) ? conj(f[s]) : f[s];
    y = fp2_load_const((K == 1 ? C_FROB1() : C_FROB2()) + 24 * wpow[s]);
  }
  static B200_HD Fp2 output(int j, const Fp2* P, const Fp2*, const Fp2*) { return P[j]; }
};

// sequential executor (host emulation, and the reference for the device executor)
template <class OP>
B200_HD void coop12_exec_seq(Fp2* f, const Fp2* g) {
  Fp2 P[18], o[6];
  for (int s = 0; s < OP::NP; s++) { Fp2 x, y; OP::operands(s, f, g, x, y); P[s] = mul(x, y); }
  for (int j = 0; j < 6; j++) o[j] = OP::output(j, P, f, g);
  for (int j = 0; j < 6; j++) f[j] = o[j];
}

#ifdef __CUDACC__
// one warp; f, g, P (and the scratch behind P) in shared memory.  Every Fp2 product of the operation's table is three Fp
// products (Karatsuba: a0 b0, a1 b1, (a0 + a1)(b0 + b1)); the 3 * NP Fp products are dealt out FLAT over the 32 lanes, so
// a general product (54) takes two rounds and a cyclotomic squaring (27) one -- the slot geometry of coop.cuh (8 slots x 4
// lanes, one lane of four idle) needed three and two.
template <class OP>
__device__ __noinline__ void coop12_exec(Fp2* f, const Fp2* g, Fp2* P) {
  const int lane = threadIdx.x & 31;
  Fp* T = reinterpret_cast<Fp*>(P + 18);          // Coop12Smem::T
#pragma unroll 1
  for (int base = 0; base < 3 * OP::NP; base += 32) {
    const int j = base + lane;
    const bool active = j < 3 * OP::NP;
    const int s = active ? j / 3 : 0, part = j % 3;
    Fp2 x, y;
    OP::operands(s, f, g, x, y);
    Fp a = add(x.c0, x.c1), b = add(y.c0, y.c1);
    coop_pick(a, part == 0, x.c0); coop_pick(b, part == 0, y.c0);
    coop_pick(a, part == 1, x.c1); coop_pick(b, part == 1, y.c1);
    const Fp t = coop_mul(a, b);
    if (active) T[j] = t;
  }
  __syncwarp();
  if (lane < OP::NP) {
    const Fp t0 = T[3 * lane], t1 = T[3 * lane + 1], t2 = T[3 * lane + 2];
    Fp2 r;
    r.c0 = sub(t0, t1);
    r.c1 = sub(sub(t2, t0), t1);
    P[lane] = r;
  }
  __syncwarp();
  Fp2 o;
  if (lane < 6) o = OP::output(lane, P, f, g);
  __syncwarp();
  if (lane < 6) f[lane] = o;
  __syncwarp();
}

struct Coop12Smem {
  Fp2 f[6], a[6], t0[6], t1[6], t2[6], g[6], P[18];
  Fp T[54];      // the Fp products of one operation (must follow P: coop12_exec finds it there)
};

__device__ __noinline__ void c12_copy(Fp2* dst, const Fp2* src) {
  const int lane = threadIdx.x & 31;
  if (lane < 6) dst[lane] = src[lane];
  __syncwarp();
}
__device__ __noinline__ void c12_conj(Fp2* f) {
  const int lane = threadIdx.x & 31;
  if (lane >= 3 && lane < 6) f[lane] = neg(f[lane]);
  __syncwarp();
}
// r <- a^z (z < 0) in the cyclotomic subgroup; r and a distinct shared arrays
__device__ __noinline__ void c12_exp_z(Fp2* r, const Fp2* a, Fp2* P) {
  c12_copy(r, a);
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    coop12_exec<OpCycSqr>(r, nullptr, P);
    if ((B200_Z_ABS >> i) & 1) coop12_exec<OpMul>(r, a, P);
  }
  c12_conj(r);
}
// f <- f^(3 (p^12-1)/r), same chain as final_exp() in pairing.cuh
__device__ __noinline__ void c12_final_exp(Coop12Smem& S) {
  const int lane = threadIdx.x & 31;
  Fp2 *f = S.f, *t0 = S.t0, *t1 = S.t1, *t2 = S.t2, *u = S.g, *P = S.P;
  // easy part: f <- conj(f) * f^-1 ; the inversion runs on one lane (one Fp inversion inside)
  if (lane == 0) {
    Fp12 x, xi;
    Fp2* xa = reinterpret_cast<Fp2*>(&x);
    for (int j = 0; j < 6; j++) xa[j] = f[j];
    fp12_inv(xi, x);
    const Fp2* ia = reinterpret_cast<const Fp2*>(&xi);
    for (int j = 0; j < 6; j++) t0[j] = ia[j];
  }
  __syncwarp();
  c12_conj(f);
  coop12_exec<OpMul>(f, t0, P);
  c12_copy(t0, f);
  coop12_exec<OpFrob<2>>(t0, nullptr, P);
  coop12_exec<OpMul>(f, t0, P);                       // f^(p^2+1)
  // hard part
  c12_exp_z(t0, f, P); c12_copy(u, f); c12_conj(u); coop12_exec<OpMul>(t0, u, P);          // f^(z-1)
  c12_exp_z(t1, t0, P); c12_copy(u, t0); c12_conj(u); coop12_exec<OpMul>(t1, u, P); c12_copy(t0, t1);   // f^((z-1)^2)
  c12_exp_z(t1, t0, P); c12_copy(u, t0); coop12_exec<OpFrob<1>>(u, nullptr, P); coop12_exec<OpMul>(t1, u, P);   // ^(z+p)
  c12_exp_z(t2, t1, P); c12_copy(u, t2); c12_exp_z(t2, u, P);                               // t1^(z^2)
  c12_copy(u, t1); coop12_exec<OpFrob<2>>(u, nullptr, P); coop12_exec<OpMul>(t2, u, P);
  c12_copy(u, t1); c12_conj(u); coop12_exec<OpMul>(t2, u, P);                               // ^(z^2+p^2-1)
  c12_copy(u, f); coop12_exec<OpCycSqr>(u, nullptr, P); coop12_exec<OpMul>(u, f, P);        // f^3
  coop12_exec<OpMul>(t2, u, P);
  c12_copy(f, t2);
}

// One warp per call: shared-squaring multi-Miller accumulation over ALL pairs of the call from the
// precomputed line functions, then the final exponentiation and the is-one test.
__global__ void __launch_bounds__(32) k_pairing_call_coop(size_t n_calls, const unsigned long long* __restrict__ offsets,
                                                          const Line* __restrict__ lines, const unsigned char* __restrict__ skip,
                                                          size_t total_pairs, uint32_t* __restrict__ outs, const int* __restrict__ errs) {
  __shared__ Coop12Smem S;
  const size_t call = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (call >= n_calls) return;
  uint32_t* out = outs + 8 * call;
  if (lane < 8) out[lane] = 0;
  if (errs[call] != E_SUCCESS) return;
  const size_t first = (size_t)(offsets[call] / 384), last = (size_t)(offsets[call + 1] / 384);
  if (lane < 6) S.f[lane] = lane == 0 ? fp2_one() : fp2_zero();
  __syncwarp();
  bool started = false;
  int s = 0;
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    if (started) coop12_exec<OpSqr>(S.f, nullptr, S.P);
    const int nsteps = ((B200_Z_ABS >> i) & 1) ? 2 : 1;
    for (int a = 0; a < nsteps; a++, s++) {
      for (size_t j = first; j < last; j++) {
        if (skip[j]) continue;
        const Line& ln = lines[(size_t)s * total_pairs + j];
        if (lane < 3) S.g[lane] = lane == 0 ? ln.l0 : (lane == 1 ? ln.l1 : ln.l4);
        __syncwarp();
        coop12_exec<OpMul014>(S.f, S.g, S.P);
        started = true;
      }
    }
  }
  c12_conj(S.f);
  c12_final_exp(S);
  if (lane == 0) {
    bool one = eq(S.f[0], fp2_one());
    for (int j = 1; j < 6; j++) one = one && is_zero(S.f[j]);
    if (one) out[7] = 0x01000000u;
  }
}
// One warp per call, for calls with MANY pairs: the Miller values of the call's chunks were accumulated in
// parallel by k_pairing_accumulate (thread per chunk); the warp multiplies them together, then runs the final
// exponentiation and the is-one test.  (A lone warp walking hundreds of pairs itself would serialise them.)
__global__ void __launch_bounds__(32) k_pairing_call_coop_chunks(size_t n_calls, const unsigned long long* __restrict__ offsets,
                                                                 const PairingPlanState* __restrict__ st,
                                                                 const uint32_t* __restrict__ call_first_task, const Fp12* __restrict__ fchunk,
                                                                 uint32_t* __restrict__ outs, const int* __restrict__ errs) {
  __shared__ Coop12Smem S;
  const size_t call = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (call >= n_calls) return;
  uint32_t* out = outs + 8 * call;
  if (lane < 8) out[lane] = 0;
  if (errs[call] != E_SUCCESS) return;
  const uint32_t chunk = st->chunk;
  const uint32_t npairs = (uint32_t)(offsets[call + 1] / 384 - offsets[call] / 384), nch = (npairs + chunk - 1) / chunk;
  const uint32_t base = call_first_task[call];
  if (lane < 6) S.f[lane] = reinterpret_cast<const Fp2*>(&fchunk[base])[lane];
  __syncwarp();
#pragma unroll 1
  for (uint32_t c = 1; c < nch; c++) {
    if (lane < 6) S.a[lane] = reinterpret_cast<const Fp2*>(&fchunk[base + c])[lane];
    __syncwarp();
    coop12_exec<OpMul>(S.f, S.a, S.P);
  }
  c12_final_exp(S);
  if (lane == 0) {
    bool one = eq(S.f[0], fp2_one());
    for (int j = 1; j < 6; j++) one = one && is_zero(S.f[j]);
    if (one) out[7] = 0x01000000u;
  }
}
#endif  // __CUDACC__

}  // namespace b200
