#!/usr/bin/env python
"""Builds ms-nets_b200/libmsnets_b200.so (sm_100a only) with nvcc, in-tree.

nvcc cross-compiles without a GPU.  No -use_fast_math: several kernels must be
IEEE-exact (true division, ordered fp32 adds, fp64 sqrt) to stay bit-identical
to the reference.  -lineinfo keeps ncu's source page usable.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libmsnets_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["capi.cu", "prep.cu", "matchers.cu", "sadsob.cu", "fte.cu", "features.cu", "slab.cu",
           "regress.cu", "ms_fused.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-fmad=true"]


def _deps():
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(PKG), "include", "msnets_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return hdrs


def _stamp(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True, ptxas_info=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_stamp = _stamp(_deps())
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(HERE, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        tag = obj + ".stamp"
        want = _stamp([sp]) + hdr_stamp
        objs.append(obj)
        have = open(tag).read() if os.path.isfile(tag) else ""
        if force or have != want or not os.path.isfile(obj):
            jobs.append((sp, obj, tag, want))

    def compile_one(job):
        sp, obj, tag, want = job
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", sp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (sp, r.stderr[-6000:]))
        if ptxas_info or verbose:
            msg = r.stderr.strip()
            if msg:
                print(msg)
        with open(tag, "w") as f:
            f.write(want)
        return obj

    if jobs:
        if verbose:
            print("[msnets_b200] nvcc %s" % " ".join(os.path.basename(j[0]) for j in jobs))
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.isfile(OUT):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
        if verbose:
            print("[msnets_b200] linked " + OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, ptxas_info="--ptxas" in sys.argv)
