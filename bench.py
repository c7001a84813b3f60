#!/usr/bin/env python3
"""bench.py -- headline benchmark of the EIP-2537 hot path (BASELINE.json metric / configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload g1msm|g2msm|pairing] [--logn L] [--calls C]

A "step" is one pass of the hot path over one batch of synthetic input:
  g1msm   (default, BASELINE configs[1]) one G1MULTIEXP call over 2^20 (point, scalar) pairs per GPU
  g2msm   (configs[2]) one G2MULTIEXP call over 2^18 pairs per GPU
  pairing (configs[3]) 16384 independent PAIRING calls of k = 2..16 pairs per GPU (calls sharded, no collective)
The plain invocation (no --workload / --logn) times the headline g1msm AND, for ~5 steps each, the other two
single-GPU configs; their value / e2e / roofline / clocks go under config.secondary.{g2msm,pairing}.
Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM, CUDA
events on the launching stream, max over ranks); `e2e` goes through the C ABI with HOST (pinned)
buffers, host<->device copies inside the timed region (pinned source; `e2e.pageable_value` repeats it with
ordinary heap memory, what a Go / Rust caller passes; `config.h2d_ceiling_*` is the bare copy).  Inputs are synthetic: points k_i*G made by the
product's own generator kernel, scalars uniform 256-bit (not reduced), as the reference's Rust bench
does (rust/benches/eip2537_benches.rs:72-80).  The 160 MiB / 72 MiB / 54 MiB inputs exceed nothing
the L2 could keep between steps for the MSMs (inputs > 126 MB L2 for g1msm; for the smaller
workloads a 256 MiB scratch buffer is written between steps to flush L2).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port": the
reference itself needs blst, which is not in the image) on a bounded sample with all host threads.
The oracle is used ONLY there and in the `cpu_baseline` leg; the GPU path never touches it.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FME_MAC32 = 300          # one 381-bit Montgomery multiplication = 2N^2+N = 300 32x32->64 MACs (N = 12)
DECODE_FME = 2256        # pairing: decode + G1/G2 subgroup checks per pair (Jacobian ladders); counted on the host-emulation build
G2_LADDER_FME = 1164     # ... of which the G2 membership ladder, which the dot-engine pipeline replaces by 12 Fp-mul in the line kernel
                         # (tests/test_host_emul.py::test_deferred_g2_membership_agrees_with_the_ladder prints both)
DOT_MAC = {"mul014": 12240, "sqr": 13536, "mul": 22608, "cyc": 6624}   # MAC32 per Fp12 operation of the dot engine (tables x 144 + 156 per output)
MADD_FME = 10            # XYZZ mixed addition 8M + 2S (G1); G2: 8*3 + 2*2 = 28
PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")


# ------------------------------------------------------------------------------------------------
# synthetic workloads (product library only -- no oracle)
# ------------------------------------------------------------------------------------------------
def _scalars_below_r(rng, n):
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    s[:, 0] &= 0x3F            # < 2^254 < r
    return s


def _be_to_int(rows):
    return [int.from_bytes(bytes(r), "big") for r in rows]


def make_msm_input(group, n, seed, with_expected=True):
    """-> (uint8[n*stride] wire bytes, sum a_i*k_i mod r or None)."""
    import blst_eip2537_b200 as b
    rng = np.random.default_rng(seed)
    a = _scalars_below_r(rng, n)
    plen = 128 if group == 1 else 256
    pts = np.frombuffer(b.generator_mul(group, a), dtype=np.uint8).reshape(n, plen)
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)      # uniform 256-bit, never reduced
    data = np.concatenate([pts, k], axis=1).reshape(-1)
    expected = None
    if with_expected:
        acc = 0
        for x, y in zip(_be_to_int(a), _be_to_int(k)):
            acc += x * y
        expected = acc % R_ORDER
    return data, expected


def make_pairing_batch(n_calls, seed, first_call=0, fixed_k=0):
    """-> (uint8 blob, uint64 offsets[n+1], expected bool[n]); call j has 2 + (j % 15) pairs (or fixed_k), every 4th false."""
    import blst_eip2537_b200 as b
    rng = np.random.default_rng(seed)
    ks = [fixed_k if fixed_k else 2 + ((first_call + j) % 15) for j in range(n_calls)]
    total = sum(ks)
    a = _scalars_below_r(rng, total)
    bb = _scalars_below_r(rng, total)
    ai, bi = _be_to_int(a), _be_to_int(bb)
    truth = np.ones(n_calls, dtype=bool)
    pos = 0
    for j, k in enumerate(ks):
        acc = 0
        for t in range(pos, pos + k - 1):
            acc += ai[t] * bi[t]
        last = (-acc) % R_ORDER
        if (first_call + j) % 4 == 3:
            last = (last + 1) % R_ORDER
            truth[j] = False
        ai[pos + k - 1], bi[pos + k - 1] = last, 1
        pos += k
    a = np.frombuffer(b"".join(x.to_bytes(32, "big") for x in ai), dtype=np.uint8)
    bb = np.frombuffer(b"".join(x.to_bytes(32, "big") for x in bi), dtype=np.uint8)
    g1 = np.frombuffer(b.generator_mul(1, a), dtype=np.uint8).reshape(total, 128)
    g2 = np.frombuffer(b.generator_mul(2, bb), dtype=np.uint8).reshape(total, 256)
    blob = np.concatenate([g1, g2], axis=1).reshape(-1)
    offs = np.zeros(n_calls + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(np.asarray(ks, dtype=np.uint64) * 384)
    return blob, offs, truth


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle) -- the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    c_oracle.lib()
    return c_oracle


def cpu_msm_points_per_s(group, data, n_pairs, threads):
    """Time the restated reference (dispatch -> Bos-Coster) on `threads` independent slices of `data`."""
    orc = _oracle()
    stride = 160 if group == 1 else 288
    per = n_pairs // threads
    fn = getattr(orc.lib(), "oracle_bls12_g%dmultiexp" % group)
    slices = [bytes(data[i * per * stride:(i + 1) * per * stride]) for i in range(threads)]
    outs = [ctypes.create_string_buffer(256) for _ in range(threads)]
    errs = [None] * threads

    def work(i):
        errs[i] = fn(outs[i], slices[i], len(slices[i]))   # ctypes releases the GIL

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    assert all(e == 0 for e in errs), errs
    return per * threads / dt, dt


def cpu_pairing_checks_per_s(blob, offs, threads):
    orc = _oracle()
    n = len(offs) - 1
    fn = orc.lib().oracle_bls12_pairing
    data = bytes(blob)
    errs = [0] * threads

    def work(i):
        out = ctypes.create_string_buffer(32)
        for j in range(i, n, threads):
            a, b = int(offs[j]), int(offs[j + 1])
            errs[i] |= fn(out, data[a:b], b - a)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    assert not any(errs)
    return n / dt, dt


# ------------------------------------------------------------------------------------------------
def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def metric_of(workload):
    return {"g1msm": ("G1 MULTIEXP points/s", "points/s"), "g2msm": ("G2 MULTIEXP points/s", "points/s"),
            "pairing": ("PAIRING checks/s", "checks/s")}[workload]


def workload_name(args):
    if args.workload == "g1msm":
        return "G1MULTIEXP single call, 2^%d random points+scalars per GPU (BASELINE configs[1])" % args.logn
    if args.workload == "g2msm":
        return "G2MULTIEXP single call, 2^%d points over Fp2 per GPU (BASELINE configs[2])" % args.logn
    if args.pairs:
        return "PAIRING batch: %d independent calls of k=%d pairs per GPU, sharded by call (fixed-k sweep of BASELINE configs[3])" % (args.calls, args.pairs)
    return "PAIRING batch: %d independent calls of k=2..16 pairs per GPU, sharded by call (BASELINE configs[3])" % args.calls


def run_reference(args, rank, world):
    """CPU arm: restated reference on the host cores, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    metric, unit = metric_of(args.workload)
    orc = _oracle()   # this arm runs entirely on the CPU: inputs are synthesised by the oracle as well
    rng = np.random.default_rng(0x2537 + 1)
    rand_r = lambda: int.from_bytes(bytes(rng.integers(0, 256, size=32, dtype=np.uint8)), "big") % R_ORDER
    if args.workload in ("g1msm", "g2msm"):
        group = 1 if args.workload == "g1msm" else 2
        per = (1 << 13) if group == 1 else (1 << 11)
        n = per * threads
        gen, prog, plen = (orc.g1_gen_mul, orc.g1_progression, 128) if group == 1 else (orc.g2_gen_mul, orc.g2_progression, 256)
        pts = np.frombuffer(prog(gen(rand_r()), gen(rand_r()), n), dtype=np.uint8).reshape(n, plen)
        data = np.concatenate([pts, rng.integers(0, 256, size=(n, 32), dtype=np.uint8)], axis=1).reshape(-1)
        step = lambda: cpu_msm_points_per_s(group, data, n, threads)
        units, sample = n, "%d threads x one %d-pair MULTIEXP call each (Bos-Coster path, eip2537.c:619-708) per step" % (threads, per)
    else:
        n = 4 * threads
        ks = [2 + (j % 15) for j in range(n)]
        chunks = []
        for k in ks:
            acc = 0
            for _ in range(k - 1):
                a_, b_ = rand_r(), rand_r()
                chunks.append(orc.g1_gen_mul(a_) + orc.g2_gen_mul(b_))
                acc += a_ * b_
            chunks.append(orc.g1_gen_mul((-acc) % R_ORDER) + orc.g2_gen_mul(1))
        blob = np.frombuffer(b"".join(chunks), dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(np.asarray(ks, dtype=np.uint64) * 384)
        step = lambda: cpu_pairing_checks_per_s(blob, offs, threads)
        units, sample = n, "%d PAIRING calls (k=2..16) spread over %d threads per step" % (n, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (6x64 Montgomery)", "data": "synthetic",
        "config": {"workload": workload_name(args), "arm": "CPU restatement of the reference algorithm (oracle/eip2537_oracle.c); "
                   "the reference binary needs blst, absent from the image"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Ctx:
    """Per-process state shared by the workloads of one bench.py run."""

    def __init__(self, rank, world, local_rank):
        import torch
        import blst_eip2537_b200 as b
        self.torch, self.b, self.L = torch, b, b._native.lib()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
        self.rank, self.world, self.local_rank = rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
        assert self.L.bls12_b200_init(-1) == 0, self.L.bls12_b200_last_error()
        self.stream = torch.cuda.Stream()
        self.flush = None
        self.peak_mac = None
        self.sharded_c = False
        if world > 1 and hasattr(self.L, "bls12_b200_comm_init"):
            # NCCL communicator owned by the C library (one rank per process): the unique id travels once over
            # torch.distributed; after that every step's exchange is issued from C (no eager torch ops in the step)
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                assert self.L.bls12_b200_comm_unique_id(uid.data_ptr()) == 0, self.L.bls12_b200_last_error()
            uid = uid.cuda()
            self.dist.broadcast(uid, 0)
            uid = uid.cpu()
            rc = self.L.bls12_b200_comm_init(world, rank, uid.data_ptr())
            assert rc == 0, self.L.bls12_b200_last_error()
            self.sharded_c = True

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def flush_l2(self):
        if self.flush is None:
            self.flush = self.torch.empty(256 * 1024 * 1024, dtype=self.torch.uint8, device="cuda")
        self.flush.fill_(1)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def measured_peak_mac(self):
        """int32 multiply-accumulate peak of THIS GPU, measured live (MEASURED_PEAKS.json has no integer peak)."""
        if self.peak_mac is None:
            ms = ctypes.c_float()
            nthr, iters = 148 * 2048, 2000
            self.L.bls12_b200_fp_microbench(1, nthr, iters, ctypes.byref(ms), None)   # IMAD.WIDE.U32, independent accumulators
            peak_plain = nthr * iters * 64 / (ms.value * 1e-3)
            self.L.bls12_b200_fp_microbench(2, nthr, iters, ctypes.byref(ms), None)   # carry-chained IMAD.WIDE.U32.X rows
            peak_chain = nthr * iters * 24 / (ms.value * 1e-3)
            self.peak_mac = max(peak_plain, peak_chain)
        return self.peak_mac


PEAK_SOURCE = ("measured live: best of two IMAD.WIDE.U32 issue-rate probes (independent accumulators / carry-chained rows), "
               "148x2048 threads; theoretical 32 lanes/clk/SM x 148 SM x 1.965 GHz = 9.3 TMAC32/s (MEASURED_PEAKS.json has no integer peak)")


def latest_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the NEWEST committed `ncu --set full` capture
    (profiles/rNN_traffic.json, written by tools/ncu_summary.py --traffic; one file per round)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_traffic.json")))
    for f in reversed(files):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if kernel_key in d:
            return d[kernel_key], os.path.basename(f)
    return None, None


def measure(ctx, workload, logn, calls, pairs, steps, warmup, window=0, cpu=True):
    """Time one workload: device-resident (`value`), end to end through the C ABI with host buffers (`e2e`, pinned
    and pageable), roofline of its dominant kernel, CPU baseline.  Returns the dict of result fields."""
    torch, b, L, dist = ctx.torch, ctx.b, ctx.L, ctx.dist
    rank, world = ctx.rank, ctx.world
    metric, unit = metric_of(workload)
    sampler = ClockSampler(ctx.local_rank)
    extra = {}
    if window:
        L.bls12_b200_set_window(window)
    with torch.cuda.stream(ctx.stream):
        s = ctx.stream.cuda_stream
        if workload in ("g1msm", "g2msm"):
            group = 1 if workload == "g1msm" else 2
            n = 1 << logn
            stride, plen = (160, 128) if group == 1 else (288, 256)
            xy = int(L.bls12_b200_partial_bytes(group))
            data, expected = make_msm_input(group, n, 0x2537 + group + 1000 * rank)
            h_in = torch.from_numpy(data).pin_memory()
            h_pageable = np.array(data, copy=True)          # ordinary heap memory: what a Go / Rust caller passes
            d_in = h_in.cuda()
            d_out = torch.zeros(plen, dtype=torch.uint8, device="cuda")
            d_status = torch.full((1,), -1, dtype=torch.int64, device="cuda")
            d_partial = torch.zeros(xy, dtype=torch.uint8, device="cuda")
            d_gather = torch.zeros(world * xy, dtype=torch.uint8, device="cuda")
            need_flush = d_in.numel() < 200 * 1024 * 1024

            def exchange_and_combine():
                # one exchange step: N x (192 | 384)-byte partial sums, plus the min of the first-error keys
                dist.all_gather_into_tensor(d_gather, d_partial)
                d_status.copy_(torch.where(d_status < 0, torch.iinfo(torch.int64).max, d_status))
                dist.all_reduce(d_status, op=dist.ReduceOp.MIN)
                d_status.copy_(torch.where(d_status == torch.iinfo(torch.int64).max, -1, d_status))
                assert L.bls12_b200_msm_combine_device(group, d_gather.data_ptr(), world, d_out.data_ptr(), s) == 0

            def step_device():
                if world == 1:
                    rc = L.bls12_b200_msm_device(group, d_in.data_ptr(), n, d_out.data_ptr(), d_status.data_ptr(), s)
                elif ctx.sharded_c:
                    rc = L.bls12_b200_msm_sharded_device(group, d_in.data_ptr(), n, rank * n, d_out.data_ptr(), d_status.data_ptr(), s)
                else:
                    d_status.fill_(-1)
                    rc = L.bls12_b200_msm_partial_device(group, d_in.data_ptr(), n, rank * n, d_partial.data_ptr(), d_status.data_ptr(), s)
                    exchange_and_combine()
                assert rc == 0, L.bls12_b200_last_error()

            def step_e2e(h):
                if world == 1:
                    return b.G1Multiexp(h) if group == 1 else b.G2Multiexp(h)
                ptr = h.data_ptr() if hasattr(h, "data_ptr") else h.ctypes.data
                if ctx.sharded_c:
                    out = np.zeros(plen, dtype=np.uint8)
                    code = L.bls12_b200_msm_sharded_host(group, ptr, n, rank * n, out.ctypes.data)
                    assert code == 0, (code, L.bls12_b200_last_error())
                    return out.tobytes()
                rc = L.bls12_b200_msm_partial_host(group, ptr, n, rank * n, d_partial.data_ptr(), d_status.data_ptr())
                assert rc == 0, L.bls12_b200_last_error()
                exchange_and_combine()
                return bytes(d_out.cpu().numpy())

            units_per_step = n * world
            h2d, d2h = n * stride, plen + 8

            def check():
                tot = expected
                if world > 1:
                    objs = [None] * world
                    dist.all_gather_object(objs, expected)
                    tot = sum(objs) % R_ORDER
                want = b.generator_mul(group, tot.to_bytes(32, "big"))
                got = bytes(d_out.cpu().numpy())
                return got == want and int(d_status.item()) == -1 and step_e2e(h_in) == want

        else:
            # PAIRING: `calls` independent calls PER GPU (no collective; each rank owns its calls and results)
            blob, offs, truth = make_pairing_batch(calls, 0x2537 + 4 + 1000 * rank, first_call=rank * calls, fixed_k=pairs)
            total_pairs = int(offs[-1]) // 384
            h_in = torch.from_numpy(blob).pin_memory()
            h_pageable = np.array(blob, copy=True)
            d_in = h_in.cuda()
            d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
            d_outs = torch.zeros(calls * 32, dtype=torch.uint8, device="cuda")
            d_errs = torch.zeros(calls, dtype=torch.int32, device="cuda")
            need_flush = True

            def step_device():
                rc = L.bls12_b200_pairing_batch_device(d_in.data_ptr(), d_offs.data_ptr(), calls, total_pairs,
                                                       d_outs.data_ptr(), d_errs.data_ptr(), s)
                assert rc == 0, L.bls12_b200_last_error()

            def step_e2e(h):
                return b.PairingBatch(h, offs)

            units_per_step = calls * world
            h2d, d2h = int(offs[-1]) + 8 * (calls + 1), calls * 36
            extra["pairs_per_step"] = total_pairs * world

            def check():
                got = d_outs.cpu().numpy().reshape(calls, 32)
                o2, e2 = step_e2e(h_in)
                return (bool((got[:, 31].astype(bool) == truth).all()) and not bool(d_errs.cpu().numpy().any())
                        and bool((o2[:, 31].astype(bool) == truth).all()) and not bool(e2.any()))

        def flush_l2():
            if need_flush:
                ctx.flush_l2()

        # ---------------- device-resident timing (value)
        sampler.start()
        for _ in range(max(warmup, 3)):
            flush_l2(); step_device()
        ctx.barrier()
        ok = check()
        launches0 = b.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ctx.barrier()
        for e0, e1 in evs:
            flush_l2()
            e0.record(); step_device(); e1.record()
        ctx.barrier()
        launches = b.launch_count() - launches0
        dev_ms = ctx.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs))
        value = units_per_step * steps / (dev_ms * 1e-3)

        # ---------------- end-to-end through the host-buffer API (e2e): pinned, then pageable caller memory
        def time_e2e(h):
            for _ in range(2):
                step_e2e(h)
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_e2e(h)
            torch.cuda.synchronize()
            return units_per_step * steps / ctx.max_over_ranks(time.perf_counter() - t0)

        e2e_value = time_e2e(h_in)
        e2e_pageable = time_e2e(h_pageable)

        # ---------------- bare H2D ceiling: the same host buffer copied to the device, all ranks at once
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        h2d_s = ctx.max_over_ranks(time.perf_counter() - t0) / 3
        extra["h2d_ceiling_gbs"] = h_in.numel() * world / h2d_s / 1e9      # aggregate over ranks, pinned source
        extra["h2d_ceiling_units_per_s"] = units_per_step / h2d_s          # what e2e could reach if only the copy counted
        clocks = sampler.stop()     # sampled from the first warm-up step to the last end-to-end step

        # ---------------- roofline of the dominant kernel (rank 0, live CUDA events inside the engine)
        roofline = None
        if workload in ("g1msm", "g2msm") and rank == 0:
            L.bls12_b200_set_profile(1)
            acc_ms, digits = [], 0
            st = (ctypes.c_float * 4)()
            nz = ctypes.c_uint64()
            for _ in range(3):
                flush_l2()
                L.bls12_b200_msm_partial_device(group, d_in.data_ptr(), n, 0, d_partial.data_ptr(), d_status.data_ptr(), s)
                torch.cuda.synchronize()
                if L.bls12_b200_last_msm_profile(st, ctypes.byref(nz)) == 0:
                    acc_ms.append(st[1]); digits = nz.value
                    extra["stage_ms"] = {"decode_digits_sort": st[0], "bucket_accumulate": st[1], "bucket_reduce": st[2], "window_combine": st[3]}
            L.bls12_b200_set_profile(0)
            peak_mac = ctx.measured_peak_mac()
            ms = ctypes.c_float()
            L.bls12_b200_fp_microbench(0, 148 * 2048, 1000, ctypes.byref(ms), None)    # dependent Fp-mul chains, full occupancy
            extra["fp_mul_per_s"] = 148 * 2048 * 1000 / (ms.value * 1e-3)
            if acc_ms:
                fme = MADD_FME if group == 1 else 28
                mac = digits * fme * FME_MAC32
                t_acc = float(np.mean(acc_ms)) * 1e-3
                achieved = mac / t_acc
                hbm = json.load(open(PEAKS_FILE))["hbm_gbs"] if os.path.exists(PEAKS_FILE) else 6650.0
                point_bytes = digits * (96 if group == 1 else 192)
                traffic, traffic_src = latest_traffic("k_accumulate_g%d_2p%d" % (group, logn))
                # whole-call fraction: 10 D + 28 B W + 9*256 Fp-mul for G1 (28 / 80 / 24 for G2)
                plan_c = window or 16
                nwin = (256 + plan_c - 1) // plan_c
                whole_fme = fme * digits + (28 if group == 1 else 80) * (1 << (plan_c - 1)) * nwin + (9 if group == 1 else 24) * 256
                roofline = {
                    "bound": "int32-mad", "kernel": "k_accumulate", "achieved": achieved / 1e12, "peak": peak_mac / 1e12,
                    "unit": "TMAC32/s", "frac": achieved / peak_mac, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": PEAK_SOURCE,
                    "algorithmic": "%d point additions x %d Fp-mul x 300 MAC32" % (digits, fme),
                    "whole_step_frac": whole_fme * FME_MAC32 / (dev_ms / steps * 1e-3) / peak_mac if logn >= 20 or group == 2 else None,
                    "hbm_secondary": {"achieved_gbs": point_bytes / t_acc / 1e9, "peak_gbs": hbm,
                                      "peak_source": "MEASURED_PEAKS.json" if os.path.exists(PEAKS_FILE) else "fallback"},
                }

        if workload == "pairing" and rank == 0:
            # Fp-multiplication counts of the reference formulation (Karatsuba tower), measured with the host-emulation
            # build (tests/test_host_emul.py::test_pairing_fme_constants pins them): per pair 2256 (decode + G1/G2
            # subgroup checks) + 1780 (68 line functions) + 2652 (68 sparse Fp12 products); per chunk 2232 (62 Fp12
            # squarings); per call 7688 (final exponentiation) + 54 per extra chunk.  The chunking the kernel actually
            # used is read back from the engine, so `algorithmic` counts the squarings really shared.
            ks = [pairs if pairs else 2 + ((rank * calls + j) % 15) for j in range(calls)]
            L.bls12_b200_set_profile(1)
            st = (ctypes.c_float * 4)()
            stage = []
            for _ in range(3):
                flush_l2(); step_device(); torch.cuda.synchronize()
                if L.bls12_b200_last_pairing_profile(st) == 0:
                    stage.append([st[0], st[1], st[2], st[3]])
            L.bls12_b200_set_profile(0)
            chunk = 3
            if hasattr(L, "bls12_b200_last_pairing_chunk"):
                chunk = max(1, int(L.bls12_b200_last_pairing_chunk()))
            nch = [(k + chunk - 1) // chunk for k in ks]
            # this batch size runs the dot-engine pipeline: G2 membership is decided in the line kernel
            fme = {"decode": (DECODE_FME - G2_LADDER_FME) * sum(ks), "lines": (1780 + 12) * sum(ks), "accumulate": 2652 * sum(ks) + 2232 * sum(nch),
                   "calls": sum(7688 + 54 * (c - 1) for c in nch)}
            # multiply-accumulates the dot-engine kernels really execute (more than the Karatsuba count: one reduction per output instead)
            actual_mac = {"accumulate": DOT_MAC["mul014"] * 68 * sum(ks) + DOT_MAC["sqr"] * 63 * sum(nch),
                          "calls": sum(DOT_MAC["cyc"] * 316 + DOT_MAC["mul"] * (37 + c - 1) for c in nch)}
            extra["pairs_per_chunk"] = chunk
            peak_mac = ctx.measured_peak_mac()
            if stage and calls <= 128:
                # small batches run the lane-cooperative latency kernels (pairing_coop.cuh, coop12.cuh): per-stage times only
                avg = np.mean(np.asarray(stage), axis=0)
                extra["stage_ms"] = {"decode_subgroup_lines": float(avg[0] + avg[1]), "miller_final_exp": float(avg[2] + avg[3])}
                roofline = {"bound": "latency", "kernel": "k_pairing_call_coop", "achieved": None, "peak": peak_mac / 1e12, "unit": "TMAC32/s",
                            "frac": None, "traffic": None, "peak_source": PEAK_SOURCE,
                            "algorithmic": "small batch: dependent-multiplication latency, not a throughput kernel (profiles/r02_msm_tail.md)"}
            elif stage:
                avg = np.mean(np.asarray(stage), axis=0)
                extra["stage_ms"] = {"decode_subgroup": float(avg[0]), "lines": float(avg[1]), "accumulate": float(avg[2]), "final_exp": float(avg[3])}
                extra["stage_frac_of_peak"] = {k: fme[k] * FME_MAC32 / (float(avg[i]) * 1e-3) / peak_mac for i, k in enumerate(("decode", "lines", "accumulate", "calls"))}
                extra["stage_actual_mac_frac"] = {"accumulate": actual_mac["accumulate"] / (float(avg[2]) * 1e-3) / peak_mac,
                                                  "final_exp": actual_mac["calls"] / (float(avg[3]) * 1e-3) / peak_mac}
                achieved = fme["accumulate"] * FME_MAC32 / (float(avg[2]) * 1e-3)
                traffic, traffic_src = latest_traffic("k_pairing_accumulate_%d" % calls)
                roofline = {"bound": "int32-mad", "kernel": "k_pairing_accumulate_dot6", "achieved": achieved / 1e12, "peak": peak_mac / 1e12,
                            "unit": "TMAC32/s", "frac": achieved / peak_mac, "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": PEAK_SOURCE,
                            "algorithmic": "%d Fp-mul x 300 MAC32: Karatsuba-tower count (2652 per pair + 2232 per chunk of %d pairs) of the products the kernel forms" % (fme["accumulate"], chunk),
                            "actual_mac_frac": actual_mac["accumulate"] / (float(avg[2]) * 1e-3) / peak_mac,
                            "algorithmic_line_bytes": 68 * 288 * sum(ks),
                            "whole_step_frac": sum(fme.values()) * FME_MAC32 / (dev_ms / steps * 1e-3) / peak_mac}

        # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
        cpu_baseline = None
        if rank == 0 and world == 1 and cpu:
            if workload in ("g1msm", "g2msm"):
                sample_n = min(n, (1 << 17) if group == 1 else (1 << 15))
                v, dt = cpu_msm_points_per_s(group, data, sample_n, 1)
                cpu_baseline = {"value": v, "unit": unit, "cores": 1, "kind": "port",
                                "sample": "first %d pairs of the same input, one MULTIEXP call (Bos-Coster path), %.1f s" % (sample_n, dt)}
            else:
                cn = min(calls, 256)
                v, dt = cpu_pairing_checks_per_s(blob[:int(offs[cn])], offs[:cn + 1], 1)
                cpu_baseline = {"value": v, "unit": unit, "cores": 1, "kind": "port",
                                "sample": "first %d calls of the same batch, serial, %.1f s" % (cn, dt)}
    if window:
        L.bls12_b200_set_window(0)
    return {
        "metric": metric, "value": value, "unit": unit, "ms_per_step": dev_ms / steps, "steps": steps,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_memory": "pinned", "pageable_value": e2e_pageable},
        "roofline": roofline, "clocks": clocks, "cpu_baseline": cpu_baseline, "gpu_launches": int(launches),
        "correct": bool(ok), "need_flush": need_flush, "extra": extra,
    }


def run_ours(args, rank, world, local_rank):
    ctx = Ctx(rank, world, local_rank)
    warm = max(args.warmup, 3)
    head = measure(ctx, args.workload, args.logn, args.calls, args.pairs, args.steps, warm, window=args.window, cpu=not args.no_cpu)
    secondary = {}
    if args.secondary:
        # the other two single-GPU BASELINE configs, ~5 steps each, so the driver's line carries their rooflines too
        k = min(5, args.steps)
        for name, wl_, logn in (("g2msm", "g2msm", 18), ("pairing", "pairing", 0)):
            if wl_ == args.workload:
                continue
            r = measure(ctx, wl_, logn, 16384, 0, k, 3, cpu=not args.no_cpu)
            a2 = argparse.Namespace(workload=wl_, logn=logn, calls=16384, pairs=0)
            secondary[name] = {
                "workload": workload_name(a2), "metric": r["metric"], "unit": r["unit"], "value": r["value"], "ms_per_step": r["ms_per_step"],
                "steps": r["steps"], "scaling": "weak", "e2e": r["e2e"], "roofline": r["roofline"], "clocks": r["clocks"],
                "cpu_baseline": r["cpu_baseline"], "gpu_launches": r["gpu_launches"], "correct": r["correct"],
                "l2": "L2 flushed between steps (256 MiB fill)", **r["extra"]}
    if rank == 0:
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (12x32 Montgomery, IMAD.WIDE)", "data": "synthetic",
            "config": dict({"workload": workload_name(args),
                            "l2": "L2 flushed between steps (256 MiB fill)" if head["need_flush"] else "inputs larger than L2",
                            "parallelism": ("points sharded, one NCCL all-gather of partial sums + status issued by the C library" if ctx.sharded_c else
                                            "points sharded, NCCL all-gather of partial sums") if args.workload != "pairing" else "calls sharded, no collective",
                            "correct": head["correct"]}, **head["extra"], **({"secondary": secondary} if secondary else {})),
            "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"] + sum(v["gpu_launches"] for v in secondary.values()),
            "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"],
        }
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="g1msm", choices=["g1msm", "g2msm", "pairing"])
    ap.add_argument("--logn", type=int, default=None)
    ap.add_argument("--calls", type=int, default=16384, help="pairing workload: independent calls PER GPU")
    ap.add_argument("--secondary", type=int, default=None, help="1: also time G2 2^18 and the 16384-call PAIRING batch (default for the plain invocation)")
    ap.add_argument("--pairs", type=int, default=0, help="pairing workload: fixed pairs per call (default: k = 2..16 mix)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--window", type=int, default=0, help="force the Pippenger window width (developer sweep)")
    args = ap.parse_args()
    if args.secondary is None:    # the driver's plain invocation: headline + the two other single-GPU configs
        args.secondary = 1 if (args.workload == "g1msm" and args.logn is None and not args.window and args.impl == "ours") else 0
    if args.logn is None:
        args.logn = 20 if args.workload == "g1msm" else 18
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
