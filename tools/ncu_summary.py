#!/usr/bin/env python3
"""Print the interesting metrics of an .ncu-rep (developer tool): python tools/ncu_summary.py rep [kernel-substr]"""
import csv, subprocess, sys, io
# --traffic KEY OUT.json : also record dram bytes (read + write) of the first matching kernel under KEY in OUT.json
#                          (bench.py's roofline.traffic reads the newest profiles/rNN_traffic.json)
args = [a for a in sys.argv[1:]]
traffic_key = traffic_out = None
if "--traffic" in args:
    i = args.index("--traffic")
    traffic_key, traffic_out = args[i + 1], args[i + 2]
    del args[i:i + 3]
rep = args[0]
sub = args[1] if len(args) > 1 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if sub and sub not in r[ki]:
        continue
    print("==", r[ki][:100])
    if traffic_key:
        import json, os
        vals = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        def to_bytes(name):
            v, u = vals[name]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        total = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        d = json.load(open(traffic_out)) if os.path.exists(traffic_out) else {}
        d[traffic_key] = total
        json.dump(d, open(traffic_out, "w"), indent=1, sort_keys=True)
        traffic_key = None
    for h, u, v in zip(hdr, units, r):
        if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.2):
            print("  %-75s %12s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), v, u))
