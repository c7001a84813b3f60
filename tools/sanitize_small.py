#!/usr/bin/env python3
"""Small end-to-end calls for compute-sanitizer (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import blst_eip2537_b200 as b
import c_oracle, workloads as wl
for n in (1, 5, 200):
    d, _ = wl.g1_msm_input(n, n)
    assert b.G1Multiexp(d) == c_oracle.call("g1multiexp", d)[1]
d, _ = wl.g2_msm_input(9, 3)
assert b.G2Multiexp(d) == c_oracle.call("g2multiexp", d)[1]
data, offs, truth = wl.pairing_batch(6, 5, 2, 4)
outs, errs = b.PairingBatch(data, offs)
assert [bool(o[31]) for o in outs] == truth and not errs.any()
skew = (d[:256] + (12345).to_bytes(32, "big")) * 300
assert b.G2Multiexp(skew) == c_oracle.call("g2multiexp", skew)[1]
print("sanitize_small ok")
