#!/usr/bin/env python
"""BASELINE config M (1984x2880, D=640, one pair) on ONE GPU: the volume (117 GB) is above the
fused kernel's parking limit, so it is cut into slabs of 192 disparities (capi.cu
use_fused_slabs).  CUDA events; one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from msnets_b200 import cbmv  # noqa: E402
from tests._synth import bordered_pair  # noqa: E402

H, W, D, B = 1984, 2880, 640, 10
L, R = bordered_pair(H, W, 99, border=B, shift=13)
l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
ex = cbmv.MSFeatureExtractor(1, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
out = ex.empty_output()
ex(l, r, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ex(l, r, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
vox = D * H * W
s = out[0, 4:8, :, 1000, 1000:1032].sum(1)          # AML columns sum to 1 where the pixel has a cost
print(json.dumps({"workload": "config M: %dx%d D=%d, 1 pair, 1 GPU, %d slabs of 192" % (H, W, D, (D + 191) // 192),
                  "ms_per_pair": round(ms, 2), "output_GB": round(32.0 * vox / 1e9, 1),
                  "output_GBps": round(32.0 * vox / ms / 1e6, 1), "aml_col_sum_err": float((s - 1).abs().max())}))
