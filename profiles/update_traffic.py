#!/usr/bin/env python
"""Regenerates profiles/roofline_traffic.json (the `traffic` field of the bench line) from an `ncu --set full`
capture of profiles/profile_target.py: DRAM bytes read + written by ms_fused_kernel per launch / pairs per launch.

usage: python profiles/update_traffic.py gpurun_out/prof_X.ncu-rep pairs_per_launch [tag]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, pairs = sys.argv[1], int(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ik = hdr.index("Kernel Name")
cols = {n: hdr.index(n) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
units = rows[1]
best = None
for r in rows[2:]:
    if "ms_fused_kernel" in r[ik]:
        best = r


def val(name):
    v = float(best[cols[name]].replace(",", ""))
    u = units[cols[name]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    return v * scale


rd, wr, ms = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
alg = 2 * 560 * 980 + 8 * 192 * 540 * 960 * 4
res = {"source": tag, "kernel": best[ik], "pairs_per_launch": pairs,
       "ms_fused_kernel_dram_bytes_per_pair": int((rd + wr) / pairs),
       "dram_read_bytes_per_pair": int(rd / pairs), "dram_write_bytes_per_pair": int(wr / pairs),
       "algorithmic_bytes_per_pair": alg, "traffic_over_algorithmic": round((rd + wr) / pairs / alg, 4),
       "kernel_ms_under_ncu": round(ms, 4)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roofline_traffic.json")
json.dump(res, open(path, "w"), indent=1)
print(json.dumps(res))
