#!/usr/bin/env python
"""A/B harness for kernel experiments: builds extra copies of the library with preprocessor
switches (python profiles/ab_variants.py build NAME -DFLAG=1 ...  -> profiles/variants/libNAME.so,
git-ignored, travels to the GPU box) and times/checks them there
(python profiles/ab_variants.py time NAME [N]).  `time` prints the per-kernel CUDA-event split of the
fused path at config B and compares channels 0-3 / 4-7 of one pair with the default library."""
import os, subprocess, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CS = os.path.join(ROOT, "ms-nets_b200", "csrc")
VD = os.path.join(ROOT, "profiles", "variants")


def build(name, flags):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(CS, "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    b.build(verbose=False)
    os.makedirs(VD, exist_ok=True)
    objs = []
    for src in b.SOURCES:
        o = os.path.join(b.OBJ_DIR, src.replace(".cu", ".o"))
        if src in ("ms_fused.cu", "sadsob.cu") and flags:
            o = os.path.join(VD, name + "_" + src.replace(".cu", ".o"))
            subprocess.check_call([b.NVCC] + b.FLAGS + flags + ["-c", os.path.join(CS, src), "-o", o])
        objs.append(o)
    out = os.path.join(VD, "lib%s.so" % name)
    subprocess.check_call([b.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart"])
    print("built", out)


def time_variant(name, N):
    import torch
    from msnets_b200 import _lib
    if name != "default":
        _lib.LIB_PATH = os.path.join(VD, "lib%s.so" % name)
    from msnets_b200 import cbmv
    from tests._synth import bordered_pair
    H, W, D, B = 540, 960, 192, 10
    pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
    l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
    r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
    ex = cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
    out = ex.empty_output()
    for _ in range(3):
        ex(l, r, out=out)
    torch.cuda.synchronize()
    _lib.lib().msn_profile_enable(1)
    for _ in range(10):
        ex(l, r, out=out)
    a, b, c, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    _lib.lib().msn_profile_read(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(n))
    import hashlib
    o = out[0].cpu().numpy()
    print("%-12s N=%d per pair: prep %.4f  sadsob %.4f  fused %.4f ms   sha ch0-3 %s ch4-7 %s" % (
        name, N, a.value / n.value / N, b.value / n.value / N, c.value / n.value / N,
        hashlib.sha256(o[:4].tobytes()).hexdigest()[:12], hashlib.sha256(o[4:].tobytes()).hexdigest()[:12]))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3:])
    else:
        time_variant(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 8)
