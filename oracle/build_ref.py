#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- builds oracle/_ref from the UNMODIFIED reference.

Compiles /root/reference/src/cpp/matchers/matchers.cpp and
/root/reference/src/cpp/featextract/featextract.cpp where they lie (zero edits,
nothing copied) against the Boost.Python->pybind11 stand-in in oracle/shim, and
writes the two extension modules into oracle/_ref/<variant>/:

    libmatchers.so      (census, nccNister, zsad, sobel, sadsob, initthreads)
    libfeatextract.so   (swap_axes, get_right_cost, extract_likelihood, ...)

Variants: ``sse41`` (-msse4.1, runs on any x86-64 box), ``avx2`` (-march=core-avx2,
the flag the reference's own CMakeLists.txt:11 uses) -- both with the shipped
``THREADS_NUM_USED 8`` (paramSetting.hpp:11) -- and ``avx2_nproc``: the same
sources with THREADS_NUM_USED re-defined, from the command line (``-include``;
still zero edits), to a run-time value = the host's core count (or
MSN_REF_THREADS), which is the second build BASELINE.md section 3 asks for.

It also copies the reference's NumPy glue ``src/dataloader/cbmv_generator.py``
VERBATIM into ``oracle/_ref/pyref/`` (git-ignored like the .so files; it travels
to the GPU box with them) so that the CPU baseline and the drop-in tests drive
the UNMODIFIED ``get_costs`` / ``extract_features_left`` there as well
(oracle/ref_glue.py imports it).
The reference's cmake build itself is unbuildable here (needs Boost 1.72
python37/numpy37, OpenCV 3, PythonLibs 3.7) -- see DESIGN.md.

oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot;
/root/reference does not exist there, so this script is a no-op without it.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
legs may load what this builds.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MSNETS_REFERENCE_ROOT", "/root/reference")
OUT_ROOT = os.path.join(HERE, "_ref")

SOURCES = {
    "libmatchers": "src/cpp/matchers/matchers.cpp",
    "libfeatextract": "src/cpp/featextract/featextract.cpp",
}
VARIANTS = {
    # reference flags: -std=c++14 -msse4.1 -march=core-avx2 -O3 -funroll-loops (+OpenMP)
    "sse41": ["-msse4.1", "-mssse3"],
    "avx2": ["-msse4.1", "-march=core-avx2"],
    # paramSetting.hpp is `#pragma once`: including it first and re-defining the macro afterwards
    # overrides THREADS_NUM_USED for the translation unit without touching any reference file
    "avx2_nproc": ["-msse4.1", "-march=core-avx2", "-include",
                   os.path.join(REF_ROOT, "src/cpp/paramSetting.hpp"), "-include",
                   os.path.join(HERE, "shim", "threads_override.h")],
}
# copied verbatim into _ref/pyref/: the generator plus the package files its relative imports need, so
# that `from src.dataloader import cbmv_generator` works on the GPU box with oracle/_ref/pyref on sys.path
GLUE = ["src/dataloader/cbmv_generator.py", "src/__init__.py", "src/cpp/__init__.py", "src/dataloader/__init__.py",
        "src/pfmutil.py", "src/funcs_utili.py"]


def _includes():
    import numpy
    import pybind11
    return [
        "-I" + os.path.join(HERE, "shim"),
        "-I" + sysconfig.get_paths()["include"],
        "-I" + numpy.get_include(),
        "-I" + pybind11.get_include(),
    ]


def build(force=False, verbose=True):
    """Returns the list of variant directories that now hold both modules."""
    if not os.path.isdir(REF_ROOT):
        if verbose:
            print("[oracle/_ref] %s absent: using prebuilt files only" % REF_ROOT)
        return [d for d in (os.path.join(OUT_ROOT, v) for v in VARIANTS)
                if all(os.path.isfile(os.path.join(d, m + ".so")) for m in SOURCES)]
    done = []
    import shutil
    for rel in GLUE:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT_ROOT, "pyref", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
    for variant, arch_flags in VARIANTS.items():
        out_dir = os.path.join(OUT_ROOT, variant)
        os.makedirs(out_dir, exist_ok=True)
        ok = True
        for mod, rel in SOURCES.items():
            src = os.path.join(REF_ROOT, rel)
            dst = os.path.join(out_dir, mod + ".so")
            if (not force and os.path.isfile(dst)
                    and os.path.getmtime(dst) >= os.path.getmtime(src)
                    and os.path.getmtime(dst) >= os.path.getmtime(__file__)):
                continue
            cmd = (["g++", "-std=c++14", "-O3", "-funroll-loops", "-fopenmp", "-fPIC",
                    "-shared", "-w", "-fvisibility=hidden"] + arch_flags + _includes()
                   + [src, "-o", dst])
            if verbose:
                print("[oracle/_ref] " + " ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                ok = False
                sys.stderr.write(r.stderr[-4000:])
        if ok:
            done.append(out_dir)
    return done


if __name__ == "__main__":
    dirs = build(force="--force" in sys.argv)
    print("built:", dirs)
    sys.exit(0 if dirs else 1)
