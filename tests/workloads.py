"""Seeded synthetic inputs for tests (SURVEY.md 8(d) generator: splitmix64, seed 0x2537 + config).

Test infrastructure: uses the C oracle to build points.  bench.py has its own generator that
uses only the product library.
"""
from __future__ import annotations

import c_oracle
import py_oracle as po

MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def scalar256(self):
        return (self.next() << 192) | (self.next() << 128) | (self.next() << 64) | self.next()

    def below(self, n):
        return self.scalar256() % n


def g1_msm_input(n, seed, structured=True):
    """n pairs: P_i = (a0 + i*delta) G1, k_i uniform 256-bit.  Returns (bytes, expected scalar mod r)."""
    rng = SplitMix64(seed)
    a0, delta = rng.below(po.R), rng.below(po.R)
    pts = c_oracle.g1_progression(c_oracle.g1_gen_mul(a0), c_oracle.g1_gen_mul(delta), n)
    out = bytearray()
    acc = 0
    for i in range(n):
        k = rng.scalar256()
        out += pts[128 * i:128 * (i + 1)] + k.to_bytes(32, "big")
        acc = (acc + (a0 + i * delta) * k) % po.R
    return bytes(out), acc


def g2_msm_input(n, seed):
    rng = SplitMix64(seed)
    a0, delta = rng.below(po.R), rng.below(po.R)
    pts = c_oracle.g2_progression(c_oracle.g2_gen_mul(a0), c_oracle.g2_gen_mul(delta), n)
    out = bytearray()
    acc = 0
    for i in range(n):
        k = rng.scalar256()
        out += pts[256 * i:256 * (i + 1)] + k.to_bytes(32, "big")
        acc = (acc + (a0 + i * delta) * k) % po.R
    return bytes(out), acc


def pairing_call(k, rng, truth=True):
    """One PAIRING input of k pairs with sum a_i*b_i = 0 mod r (truth) or perturbed (false)."""
    out = bytearray()
    acc = 0
    for i in range(k - 1):
        a, b = rng.below(po.R - 1) + 1, rng.below(po.R - 1) + 1
        out += c_oracle.g1_gen_mul(a) + c_oracle.g2_gen_mul(b)
        acc = (acc + a * b) % po.R
    last = (-acc) % po.R
    if not truth:
        last = (last + 1) % po.R
    out += c_oracle.g1_gen_mul(last) + c_oracle.g2_gen_mul(1)
    return bytes(out)


def pairing_batch(n_calls, seed, kmin=2, kmax=16):
    """-> (data, offsets list, expected bools): call j has kmin + j % (kmax-kmin+1) pairs; every 4th false."""
    rng = SplitMix64(seed)
    data = bytearray()
    offs = [0]
    truth = []
    for j in range(n_calls):
        k = kmin + (j % (kmax - kmin + 1))
        t = (j % 4) != 3
        data += pairing_call(k, rng, t)
        offs.append(len(data))
        truth.append(t)
    return bytes(data), offs, truth
