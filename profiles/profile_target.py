"""Small driver for ncu captures: config-B-shaped pairs through the fused MS-volume path and
soft-argmin, `reps` times (default 2: one warm-up + one to capture).  Never a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from msnets_b200 import cbmv, regression  # noqa: E402
from tests._synth import bordered_pair  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W, D, B = 540, 960, 192, 10
pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
ex = cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
out = ex.empty_output()
logits = torch.randn((N, D, H, W), device="cuda")
for _ in range(reps):
    ex(l, r, out=out)
    regression.soft_argmin(logits)
torch.cuda.synchronize()
print("done", tuple(out.shape))
