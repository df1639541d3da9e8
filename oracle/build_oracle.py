#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- compiles oracle/ms_oracle.c -> oracle/libms_oracle.so.

Flags keep IEEE semantics (no fast-math, no FMA contraction) because the oracle
replays the reference's fp32 operation order.  -march is left at the x86-64
baseline so the .so runs on whatever host the GPU box has.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ms_oracle.c")
OUT = os.path.join(HERE, "libms_oracle.so")


def build(force=False, verbose=True):
    if (not force and os.path.isfile(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)
            and os.path.getmtime(OUT) >= os.path.getmtime(__file__)):
        return OUT
    cmd = ["gcc", "-std=c11", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-fPIC", "-shared", "-Wall", SRC, "-o", OUT, "-lm"]
    if verbose:
        print("[oracle] " + " ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
