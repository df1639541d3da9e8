"""Experiment: per-tile phase timeline of ms_fused_kernel on SM 0 (MSNETS_TRACE=1 makes thread 0 of the
CTAs resident there stamp clock64 at the phase boundaries into the head of the workspace).
usage: MSNETS_TRACE=1 [MSNETS_FUSED_CTAS=n] python profiles/fused_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msnets_b200 import cbmv
from tests._synth import bordered_pair
N, H, W, D, B = 8, 540, 960, 192, 10
pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
ex = cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
out = ex.empty_output()
for _ in range(3):
    ex(l, r, out=out)
torch.cuda.synchronize()
t = ex.workspace[:65536].cpu().numpy().view(np.int64)
n = int(t[0] & 0xffffffff)
print("CTAs seen on SM 0:", n)
rec = t[8:].reshape(-1, 8)
if os.environ.get("MSNETS_FUSED_CTAS", "") in ("", "0"):
    for slot in range(2):
        rr = rec[slot * 500:(slot + 1) * 500]
        rr = rr[rr[:, 0] != 0][5:105]
        if len(rr) == 0: continue
        t0 = rr[:, 0]
        print("slot %d tiles %d: period %.0f  rows-wait %.0f  phase1(w0) %.0f  to-barrierA %.0f  back-half %.0f  barrierD %.0f cycles (medians)" % (
            slot, len(rr), np.median(np.diff(t0)), np.median(rr[:, 1] - rr[:, 0]), np.median(rr[:, 2] - rr[:, 1]),
            np.median(rr[:, 3] - rr[:, 2]), np.median(rr[:, 5] - rr[:, 3]), np.median(rr[:, 6] - rr[:, 5])))
else:
    rr = rec[:1000]
    rr = rr[rr[:, 0] != 0]
    rr = rr[np.argsort(rr[:, 0])][10:900]
    print("one tile per CTA, %d CTAs: start-to-start (2 slots interleaved) %.0f  rows-wait %.0f  phase1(w0) %.0f  to-barrierA %.0f  back-half %.0f cycles; CTA lifetime %.0f" % (
        len(rr), np.median(np.diff(rr[:, 0])), np.median(rr[:, 1] - rr[:, 0]), np.median(rr[:, 2] - rr[:, 1]),
        np.median(rr[:, 3] - rr[:, 2]), np.median(rr[:, 5] - rr[:, 3]), np.median(rr[:, 5] - rr[:, 0])))
