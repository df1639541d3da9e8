#!/usr/bin/env python
"""Achieved HBM bandwidth of the kernels around the fused path (confidence pass, 4D volume
builders, soft-argmin at the PSMNet shape), CUDA events on the launch stream, algorithmic
bytes / time against MEASURED_PEAKS.json.  Prints one JSON object.  Not the headline metric."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from msnets_b200 import confidence, regression, volume  # noqa: E402

PEAK = 6551.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


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


def main():
    res = {}
    g = torch.Generator(device="cuda").manual_seed(1234)

    def add(name, ms, nbytes, note):
        gbs = nbytes / ms / 1e6
        res[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(gbs, 1),
                     "frac_of_measured_peak": round(gbs / PEAK, 3), "note": note}

    # config P: PSMNet-shaped 4D volume, batch 16, unary features [16,32,136,240], D/4 = 48
    N, C, h, w, D4 = 16, 32, 136, 240, 48
    fl = torch.randn((N, C, h, w), generator=g, device="cuda")
    fr = torch.randn((N, C, h, w), generator=g, device="cuda")
    out = torch.empty((N, 2 * C, D4, h, w), device="cuda")
    add("concat_volume_P", timed(lambda: volume.concat_volume(fl, fr, D4, out=out)),
        2 * fl.numel() * 4 + out.numel() * 4, "[16,64,48,136,240] fp32 written")
    outd = torch.empty((N, C, D4, h, w), device="cuda")
    add("diff_volume_P", timed(lambda: volume.diff_volume(fl, fr, D4, out=outd)),
        2 * fl.numel() * 4 + outd.numel() * 4, "[16,32,48,136,240] fp32 written")
    del out, outd
    logits = torch.randn((4, 192, 544, 960), generator=g, device="cuda")
    disp = torch.empty((4, 544, 960), device="cuda")
    add("soft_argmin_P", timed(lambda: regression.soft_argmin(logits, out=disp)),
        logits.numel() * 4 + disp.numel() * 4, "[4,192,544,960] logits read")
    # confidence pass on one config-B cost volume, both layouts
    H, W, D = 540, 960, 192
    c_dhw = torch.rand((D, H, W), generator=g, device="cuda")
    add("wta_dhw_B", timed(lambda: confidence.wta(c_dhw, layout="dhw")), c_dhw.numel() * 4 + 3 * H * W * 4,
        "argmin + min + second min over D, feature-plane layout")
    c_hwd = c_dhw.permute(1, 2, 0).contiguous()
    add("wta_hwd_B", timed(lambda: confidence.wta(c_hwd, layout="hwd")), c_hwd.numel() * 4 + 3 * H * W * 4,
        "same, reference [H,W,D] layout (warp-shuffle (min, argmin, second-min) merges)")
    add("lr_consistency_B", timed(lambda: confidence.lr_consistency(c_hwd, 1)), 2 * c_hwd.numel() * 4 + 9 * H * W,
        "left and right-view argmin (volume read twice) + mask")
    am, m1, m2 = confidence.wta(c_hwd, layout="hwd")
    add("pkrn_confidence_B", timed(lambda: confidence.pkrn_confidence(m1, m2, 0.01)), 3 * H * W * 4, "elementwise")
    print(json.dumps({"peak_GBps": PEAK, "kernels": res}))


if __name__ == "__main__":
    main()
