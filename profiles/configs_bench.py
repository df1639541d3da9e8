#!/usr/bin/env python
"""One-GPU timings of the BASELINE.json configurations other than the bench.py headline (configs[1]):
A 256x512 D192 (one pair), C KITTI-shaped 375x1242 padded to 384x1248 D192 (one pair, features +
soft-argmin), P PSMNet-shaped 544x960 D192 batch 16 (concat 4D volume + soft-argmin).  CUDA events,
inputs resident in HBM.  One JSON object.  (Config M: profiles/config_m_single.py, slab_bench.py.)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from msnets_b200 import cbmv, regression, volume  # noqa: E402
from tests._synth import bordered_pair  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ms_config(H, W, D, N):
    B = 10
    pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
    l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
    r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
    ex = cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
    out = ex.empty_output()
    logits = torch.randn((N, D, H, W), device="cuda")
    disp = torch.empty((N, H, W), device="cuda")

    def step():
        ex(l, r, out=out)
        regression.soft_argmin(logits, out=disp)
    ms = timed(step)
    nbytes = N * (8 * D * H * W * 4 + 2 * (H + 2 * B) * (W + 2 * B) + 4 * D * H * W + 4 * H * W)
    return {"ms_per_step": round(ms, 4), "pairs_per_s": round(N * 1e3 / ms, 1),
            "algorithmic_GBps": round(nbytes / ms / 1e6, 1)}


res = {}
res["A: 256x512 D192, 1 pair (features + soft-argmin)"] = ms_config(256, 512, 192, 1)
res["C: 384x1248 (375x1242 padded) D192, 1 pair (features + soft-argmin)"] = ms_config(384, 1248, 192, 1)
N, C, h, w, D4 = 16, 32, 136, 240, 48
fl, fr = torch.randn((N, C, h, w), device="cuda"), torch.randn((N, C, h, w), device="cuda")
vol = torch.empty((N, 2 * C, D4, h, w), device="cuda")
logits = torch.randn((N, 192, 544, 960), device="cuda")
disp = torch.empty((N, 544, 960), device="cuda")


def step_p():
    volume.concat_volume(fl, fr, D4, out=vol)
    regression.soft_argmin(logits, out=disp)


ms = timed(step_p)
nbytes = 2 * fl.numel() * 4 + vol.numel() * 4 + logits.numel() * 4 + disp.numel() * 4
res["P: 544x960 D192 batch 16 (concat volume [16,64,48,136,240] + soft-argmin)"] = {
    "ms_per_step": round(ms, 4), "samples_per_s": round(N * 1e3 / ms, 1), "algorithmic_GBps": round(nbytes / ms / 1e6, 1)}
print(json.dumps(res))
