import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__
__graft_entry__.build()
from msnets_b200 import cbmv
from tests._synth import bordered_pair
N, H, W, D, B = 2, 540, 960, 192, 10
pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
ex = cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, left_only=False, board_h=B, board_w_left=B, board_w_right=B)
out = ex.empty_output()
for _ in range(2):
    ex(l, r, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ex(l, r, out=out)
e1.record(); torch.cuda.synchronize()
print("16-channel volume: %.3f ms/pair (%s)" % (e0.elapsed_time(e1) / 5 / N, os.environ.get("MSNETS_FORCE_GENERIC", "")))
