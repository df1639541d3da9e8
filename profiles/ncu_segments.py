#!/usr/bin/env python
"""Splits a kernel's stall samples and executed instructions into the segments between its
BAR.SYNC instructions (= the phases of ms_fused_kernel) from an .ncu-rep source page.

usage: python profiles/ncu_segments.py prof.ncu-rep kernel_regex
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break          # second kernel instance: stop
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        data.append(r)
    iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iN]) for r in data)
    tote = sum(int(r[iE]) for r in data)
    bars = [k for k, r in enumerate(data) if "BAR.SYNC" in r[iS]]
    print("instructions %d, samples %d, BAR.SYNC at %s" % (len(data), tot, bars))
    bounds = [0] + bars + [len(data)]
    for a, b in zip(bounds, bounds[1:]):
        s = sum(int(r[iN]) for r in data[a:b])
        e = sum(int(r[iE]) for r in data[a:b])
        agg = collections.Counter()
        for r in data[a:b]:
            for i, h in stall:
                if r[i].isdigit():
                    agg[h] += int(r[i])
        top = ", ".join("%s %.1f%%" % (h.replace("stall_", ""), 100.0 * v / tot) for h, v in agg.most_common(5))
        print("segment %4d..%4d  samples %5.1f%%  warp-instr %5.1f%%   %s" % (a, b, 100.0 * s / tot, 100.0 * e / tote, top))
    print("hottest instructions:")
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][iN]))[:12]:
        best = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall), reverse=True)[:2]
        print("  %5.2f%%  idx %4d  exec %9s  %-50s %s" % (100.0 * int(r[iN]) / tot, k, r[iE], r[iS].strip()[:50], best))


if __name__ == "__main__":
    main()
