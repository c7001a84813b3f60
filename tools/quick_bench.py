#!/usr/bin/env python3
"""Ad-hoc timing probe (developer tool, not the contract bench): K1 microbenchmarks and MSM/pairing timings."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import blst_eip2537_b200 as b

L = b._native.lib()
assert L.bls12_b200_init(-1) == 0, L.bls12_b200_last_error()
ms = ctypes.c_float()
what = sys.argv[1:] or ["micro", "msm", "pairing"]

if "micro" in what:
    nthr = 148 * 2048
    for mode, name, per in ((1, "imad.wide (operands vary)", 64), (3, "imad lo+hi pairs", 64), (2, "imad.wide carry chains", 24), (0, "fp_mul", 1),
                            (4, "dot6 engine, 32-bit accumulator words (MAC32)", 1020), (5, "dot6 engine, 64-bit accumulator pairs (MAC32)", 1020),
                            (6, "fp_mul, per-thread multiplier, 2 rows/iter (product)", 1), (7, "fp_mul, 4 rows/iter", 1), (8, "fp_mul, 6 rows/iter", 1),
                            (9, "fp_mul, fully unrolled", 1), (10, "fp_mul, multiplier limbs from shared memory", 1),
                            (11, "fp_mul, 2 rows/iter, ALU-pipe shifts", 1), (12, "fp_mul, 6 rows/iter, ALU-pipe shifts", 1)):
        iters = 2000 if mode == 1 else (300 if mode >= 4 else 1000)
        L.bls12_b200_fp_microbench(mode, nthr, iters, ctypes.byref(ms), None)
        ops = nthr * iters * per
        print("%s: %.3f ms, %.3e ops/s" % (name, ms.value, ops / (ms.value * 1e-3)), flush=True)

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
rng = np.random.default_rng(0x2537)


def gen_g1(n):
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 0] &= 0x3F   # < 2^254 < r
    pts = np.frombuffer(b.generator_mul(1, sc), dtype=np.uint8).reshape(n, 128)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    return np.concatenate([pts, ks], axis=1).reshape(-1)


if "latency" in what:
    # one warp per SM sub-partition at most: dependent Fp-mul chain latency, CIOS (rolled) vs product scanning
    for mode, name in ((0, "mul (rolled CIOS)"), (7, "CIOS 4 rows/iter"), (8, "CIOS 6 rows/iter"), (9, "CIOS fully unrolled"),
                       (13, "product + separated reduction, 64-bit pairs"), (14, "product + separated reduction, 32-bit words")):
        for nthr in (32, 148 * 128, 148 * 2048):
            iters = 2000
            L.bls12_b200_fp_microbench(mode, nthr, iters, ctypes.byref(ms), None)
            print("%s, %d threads: %.3f us per dependent mul, %.3e mul/s" % (name, nthr, ms.value * 1e3 / iters, nthr * iters / (ms.value * 1e-3)), flush=True)

if "msm" in what:
    for logn in (7, 10, 12, 14, 16, 18, 20):
        n = 1 << logn
        t0 = time.time(); data = gen_g1(n); tg = time.time() - t0
        h = torch.from_numpy(data).pin_memory()
        for rep in range(3):
            t0 = time.time(); out = b.G1Multiexp(h); dt = time.time() - t0
        d_in = h.cuda()
        d_out = torch.zeros(128, dtype=torch.uint8, device="cuda")
        d_st = torch.zeros(1, dtype=torch.int64, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        torch.cuda.synchronize()
        for c in (0,):
            b.set_window(c)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for rep in range(3):
                ev0.record()
                assert L.bls12_b200_msm_device(1, d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), s) == 0
                ev1.record(); torch.cuda.synchronize()
            assert bytes(d_out.cpu().numpy()) == out and d_st.item() == -1
            print("G1 MSM 2^%d c=%d: gen %.2fs  e2e %.2f ms  device %.3f ms  -> %.3e pts/s" % (logn, c, tg, dt * 1e3, ev0.elapsed_time(ev1), n / (ev0.elapsed_time(ev1) * 1e-3)), flush=True)
        b.set_window(0)

if "pairing" in what:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import workloads as wl
    data, offs, truth = wl.pairing_batch(64, 99, 2, 16)
    reps = 64
    blob = data * reps
    o = [0]
    for r in range(reps):
        for j in range(64):
            o.append(o[-1] + offs[j + 1] - offs[j])
    for rep in range(3):
        t0 = time.time(); outs, errs = b.PairingBatch(blob, o); dt = time.time() - t0
    assert [int(x[31]) for x in outs[:64]] == [1 if t else 0 for t in truth] and not errs.any()
    print("pairing batch %d calls (%d pairs): %.1f ms -> %.0f checks/s" % (len(o) - 1, len(blob) // 384, dt * 1e3, (len(o) - 1) / dt), flush=True)

if "check" in what:
    n = 1 << 20
    data = gen_g1(n)
    d_in = torch.from_numpy(data).cuda()
    d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for sub in (0, 1):
        for rep in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            assert L.bls12_b200_points_check_device(1, d_in.data_ptr(), n, 160, sub, d_codes.data_ptr(), s) == 0
            ev1.record(); torch.cuda.synchronize()
        assert int(d_codes.abs().sum().item()) == 0
        print("G1 batched decode%s over a 2^20-pair MULTIEXP input: %.3f ms -> %.3e points/s" % (" + subgroup check" if sub else "", ev0.elapsed_time(ev1), n / (ev0.elapsed_time(ev1) * 1e-3)), flush=True)

if "checked" in what:
    # opt-in checked MULTIEXP (subgroup check per point + GLV / mod-r fast path) against the plain path, host buffers
    n = 1 << 20
    data = gen_g1(n)
    h = torch.from_numpy(data).pin_memory()
    ref = b.G1Multiexp(h)
    for label, checked, noglv in (("plain (reference semantics)", False, "0"), ("checked, GLV fast path", True, "0")):
        b.set_checked_msm(checked)
        for rep in range(3):
            t0 = time.time(); out = b.G1Multiexp(h); dt = time.time() - t0
        assert out == ref
        print("G1 MULTIEXP 2^20 %s: %.2f ms end to end -> %.3e points/s" % (label, dt * 1e3, n / dt), flush=True)
    b.set_checked_msm(False)
    L.bls12_b200_set_profile(1)
    d_in = h.cuda()
    d_part = torch.zeros(512, dtype=torch.uint8, device="cuda")
    d_st = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    st = (ctypes.c_float * 4)()
    nz = ctypes.c_uint64()
    for checked in (False, True):
        b.set_checked_msm(checked)
        for rep in range(2):
            L.bls12_b200_msm_partial_device(1, d_in.data_ptr(), n, 0, d_part.data_ptr(), d_st.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        L.bls12_b200_last_msm_profile(st, ctypes.byref(nz))
        print("  stages (checked=%s): decode[+check]+sort %.2f, accumulate %.2f, reduce %.2f, combine %.2f ms; %d additions" % (checked, st[0], st[1], st[2], st[3], nz.value), flush=True)
    b.set_checked_msm(False)
    L.bls12_b200_set_profile(0)

if "batch" in what:
    for k in (4, 32, 128):
        ncalls = 16384 if k <= 32 else 4096
        data = gen_g1(ncalls * k)
        offs = np.arange(ncalls + 1, dtype=np.uint64) * (160 * k)
        for rep in range(3):
            t0 = time.time(); outs, errs = b.MultiexpBatch(1, data, offs); dt = time.time() - t0
        assert not errs.any() and bytes(outs[0]) == b.G1Multiexp(data[:160 * k])
        print("G1 MULTIEXP batch: %d calls x %d pairs: %.1f ms -> %.0f calls/s, %.3e pairs/s" % (ncalls, k, dt * 1e3, ncalls / dt, ncalls * k / dt), flush=True)

if "single" in what:
    # latency of ONE call through the legacy ABI (what a drop-in user of bls12_* sees per call)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import workloads as wl
    for k in (1, 4, 128):
        d = bytes(gen_g1(k))
        for rep in range(5):
            t0 = time.time(); b.G1Multiexp(d); dt = time.time() - t0
        print("single G1MULTIEXP k=%d: %.2f ms" % (k, dt * 1e3), flush=True)
    rng = wl.SplitMix64(3)
    for k in (2, 16):
        d = wl.pairing_call(k, rng, True)
        for rep in range(5):
            t0 = time.time(); out = b.Pairing(d); dt = time.time() - t0
        assert out[31] == 1
        print("single PAIRING k=%d: %.2f ms" % (k, dt * 1e3), flush=True)
if "map" in what:
    # MAP_FP_TO_G1 / MAP_FP2_TO_G2: single-call latency and batch throughput (host buffers, copies included)
    import numpy as np
    rng = np.random.default_rng(5)
    for group, width in ((1, 64), (2, 128)):
        for n in (1, 1 << 10, 1 << 14, 1 << 17):
            raw = rng.integers(0, 256, size=(n * (width // 64), 64), dtype=np.uint8)
            raw[:, :16] = 0
            raw[:, 16] &= 0x0F                      # < p
            blob = raw.tobytes()
            for rep in range(3):
                t0 = time.time(); outs, errs = b.MapBatch(group, blob); dt = time.time() - t0
            assert not errs.any()
            print("MAP group %d n=%d: %.2f ms  (%.0f maps/s)" % (group, n, dt * 1e3, n / dt), flush=True)
