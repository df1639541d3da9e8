#!/usr/bin/env python
"""Static look at a kernel's SASS: instruction mix per loop body (backward branch targets) and the
TMA / packed-fp32 / popcount mnemonic counts that the DESIGN claims rest on.

usage: python profiles/sass_loops.py <lib.so> <kernel-name-substring> [--all]
       (no GPU needed: cuobjdump -sass on the built library)
"""
import collections
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def opcode(txt):
    t = txt.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    show_all = "--all" in sys.argv
    for name, ins in kernels(lib).items():
        if pat not in name:
            continue
        print("=" * 100)
        print(name, "instructions:", len(ins))
        tot = collections.Counter(opcode(t) for _, t in ins)
        keys = ["UTMALDG", "UBLKCP", "FADD2", "FFMA2", "FMUL2", "POPC", "MUFU", "DMUL", "LDS", "STS", "STG", "LDG", "BAR",
                "SYNCS", "MOV", "STL", "LDL"]
        print("  static counts: " + "  ".join("%s %d" % (k, tot[k]) for k in keys if tot[k]))
        addr = [a for a, _ in ins]
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L_x_\d+\)?|BRA\S*\s+.*0x([0-9a-f]+)", t)
            if "BRA" in t:
                m2 = re.search(r"0x([0-9a-f]+)", t)
                if m2:
                    tgt = int(m2.group(1), 16)
                    if tgt <= a and tgt in addr:
                        loops.append((addr.index(tgt), i))
        for (b, e) in sorted(set(loops)):
            body = ins[b:e + 1]
            if len(body) < 8 and not show_all:
                continue
            c = collections.Counter(opcode(t) for _, t in body)
            top = ", ".join("%s %d" % kv for kv in c.most_common(14))
            print("  loop [%04x..%04x] %4d instr: %s" % (ins[b][0], ins[e][0], len(body), top))


if __name__ == "__main__":
    main()
