#!/usr/bin/env python
"""L1 / shared-memory data-pipe view of a kernel from an .ncu-rep: wavefront counts per voxel
and utilisation (the pipe that bounds ms_fused_kernel).

usage: python profiles/ncu_pipe.py prof.ncu-rep kernel_regex [voxels_per_launch]
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "derived__memory_l1_wavefronts_shared_excessive",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, pat = sys.argv[1], re.compile(sys.argv[2])
    vox = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        if not pat.search(r[kn]):
            continue
        print(re.sub(r"\(.*", "", r[kn]))
        for k in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]
                extra = ""
                if vox and ("wavefront" in k or "conflict" in k or "inst_executed" in k) and "pct" not in k:
                    try:
                        f = float(v.replace(",", ""))
                        if "lgds.avg" in k:
                            f *= 148
                        if "inst_executed" in k:
                            f *= 32
                        extra = "   = %.3f per voxel" % (f / vox)
                    except ValueError:
                        pass
                print("  %-75s %s %s%s" % (k, v, units[hdr.index(k)], extra))


if __name__ == "__main__":
    main()
