#!/usr/bin/env python
"""Per-100-instruction stall histogram of a kernel from an .ncu-rep source page (SASS order).
usage: python profiles/ncu_blocks.py prof.ncu-rep kernel_regex [block]"""
import collections, csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
blk_n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    data.append(r)
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN]) for r in data)
tote = sum(int(r[iE]) for r in data)
print("total samples", tot, "static instrs", len(data), "warp-instr executed", tote)
for a in range(0, len(data), blk_n):
    blk = data[a:a + blk_n]
    s = sum(int(r[iN]) for r in blk)
    e = sum(int(r[iE]) for r in blk)
    if s == 0:
        continue
    agg = collections.Counter()
    for r in blk:
        for i, h in stall:
            if r[i].isdigit():
                agg[h] += int(r[i])
    print("%5d samples %5.1f%%  instr %5.1f%%  %s" % (a, 100.0 * s / tot, 100.0 * e / tote, ", ".join(
        "%s %.1f" % (h.replace("stall_", ""), 100.0 * v / tot) for h, v in agg.most_common(6))))
print("hottest:")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][iN]))[:20]:
    best = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall), reverse=True)[:2]
    print("  %5.2f%% idx %4d exec %9s %-60s %s" % (100.0 * int(r[iN]) / tot, k, r[iE], r[iS].strip()[:60], best))
