#!/usr/bin/env python3
"""SASS opcode histogram of the hot kernels in libblst_eip2537.so (evidence for the IMAD.WIDE / LDS / cp.async claims):
   python tools/sass_hist.py [kernel-substring ...] > profiles/rNN_sass_hist.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "blst_eip2537_b200", "libblst_eip2537.so")
want = sys.argv[1:] or ["k_accumulateINS_2FpELi6", "k_accumulateINS_3Fp2", "k_pairing_accumulate_dot6ILi3", "k_pairing_final_dot6",
                        "k_pairing_lines_slots", "k_pairing_decodeILb0", "k_reduce_leafINS_2Fp", "k_dot_chainIN4b2003dot5Acc64", "k_fp_chainiP"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, hist = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        hist[fn][m.group(2)] += 1
print("# SASS opcode histograms (cuobjdump -sass, sm_100a cubin of libblst_eip2537.so)\n")
for key in want:
    for f in hist:
        if key in f:
            h = hist[f]
            total = sum(h.values())
            wide = sum(v for k, v in h.items() if k.startswith("IMAD.WIDE"))
            print("## `%s`\n\n%d instructions; IMAD.WIDE* %d (%.1f %%); LDL+STL %d; LDS %d; LDG/LD %d; LDGSTS (cp.async) %d; BAR %d\n" % (
                f[:110], total, wide, 100.0 * wide / total, h.get("LDL", 0) + h.get("STL", 0) + h.get("LDL.LU", 0) + h.get("LDL.64", 0) + h.get("STL.64", 0),
                sum(v for k, v in h.items() if k.startswith("LDS")), sum(v for k, v in h.items() if k.startswith("LDG") or k.startswith("LD.")),
                sum(v for k, v in h.items() if k.startswith("LDGSTS")), sum(v for k, v in h.items() if k.startswith("BAR"))))
            print("| opcode | count |\n|---|---|")
            for op, c in h.most_common(14):
                print("| %s | %d |" % (op, c))
            print()
