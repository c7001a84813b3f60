#!/usr/bin/env python3
"""Developer tool: device-resident MSM time per window width at SMALL sizes (single-call latency floor)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blst_eip2537_b200 as b
L = b._native.lib()
rng = np.random.default_rng(1)
for group in (1, 2):
    plen = 128 if group == 1 else 256
    for logn in (1, 3, 5, 7, 9, 11, 13):
        n = 1 << logn
        sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); sc[:, 0] &= 0x3F
        pts = np.frombuffer(b.generator_mul(group, sc), dtype=np.uint8).reshape(n, plen)
        ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        d_in = torch.from_numpy(np.concatenate([pts, ks], axis=1).reshape(-1)).cuda()
        d_out = torch.zeros(plen, dtype=torch.uint8, device="cuda"); d_st = torch.zeros(1, dtype=torch.int64, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        res = []
        for c in range(3, 15):
            b.set_window(c)
            best = 1e9
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); L.bls12_b200_msm_device(group, d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), s); e1.record()
                torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            res.append((best, c))
        b.set_window(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.bls12_b200_msm_device(group, d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), s); e1.record(); torch.cuda.synchronize()
        print("G%d 2^%d:" % (group, logn), " ".join("c%d=%.2f" % (c, t) for t, c in res), " best c=%d (%.2f ms); rule: %.2f ms" % (min(res)[1], min(res)[0], e0.elapsed_time(e1)), flush=True)
