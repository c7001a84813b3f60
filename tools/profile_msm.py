#!/usr/bin/env python3
"""One device-resident G1/G2 MSM (or pairing batch) after a warm-up, for ncu (developer tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import blst_eip2537_b200 as b

group = int(sys.argv[1]) if len(sys.argv) > 1 else 1
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
L = b._native.lib()
n = 1 << logn
rng = np.random.default_rng(0x2537)
sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
sc[:, 0] &= 0x3F
plen = 128 if group == 1 else 256
pts = np.frombuffer(b.generator_mul(group, sc), dtype=np.uint8).reshape(n, plen)
ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
d_in = torch.from_numpy(np.concatenate([pts, ks], axis=1).reshape(-1)).cuda()
d_out = torch.zeros(plen, dtype=torch.uint8, device="cuda")
d_st = torch.zeros(1, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    assert L.bls12_b200_msm_device(group, d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), s) == 0
    ev1.record()
    torch.cuda.synchronize()
    print("group %d n=2^%d: %.3f ms" % (group, logn, ev0.elapsed_time(ev1)))
assert d_st.item() == -1
