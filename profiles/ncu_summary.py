#!/usr/bin/env python
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw/source --csv`) into the
numbers DESIGN.md / bench.py cite: duration, DRAM bytes, issue utilisation, stall
reasons and the per-opcode instruction mix of selected kernels.

usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep kernel_regex [voxels_per_launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep, pat = sys.argv[1], re.compile(sys.argv[2])
    vox = float(sys.argv[3]) if len(sys.argv) > 3 else None
    hdr, units, rows = raw_rows(rep)
    kn = hdr.index("Kernel Name")
    seen = set()
    for r in rows:
        name = r[kn]
        short = re.sub(r"\(.*", "", name)
        if not pat.search(name) or short in seen:
            continue
        seen.add(short)
        print("=" * 100)
        print(short)
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
            if m:
                try:
                    stalls.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        print("  stalls (warps per issue-active cycle): " +
              ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:7]))
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                              "regex:" + re.escape(short.split("::")[-1].split("<")[0])],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        h2 = None
        cnt, samp = collections.Counter(), collections.Counter()
        for sr in srows:
            if sr and sr[0] == "Address":
                if h2 is not None:
                    break
                h2 = sr
                continue
            if h2 is None or len(sr) < len(h2):
                continue
            txt = sr[h2.index("Source")].strip().split()
            op = txt[1] if txt[0].startswith("@") else txt[0]
            op = op.rstrip(";").split(".")[0]
            cnt[op] += int(sr[h2.index("Instructions Executed")])
            samp[op] += int(sr[h2.index("# Samples")])
        tot, tots = sum(cnt.values()), max(sum(samp.values()), 1)
        print("  warp instructions %d%s" % (tot, ("  = %.1f per voxel" % (tot * 32 / vox)) if vox else ""))
        for op, v in cnt.most_common(18):
            print("    %-10s %5.1f%%  %s stall-samples %4.1f%%" % (
                op, 100.0 * v / tot, ("%6.2f/voxel" % (v * 32 / vox)) if vox else "", 100.0 * samp[op] / tots))


if __name__ == "__main__":
    main()
