"""Times the fused MS-volume path alone (CUDA events) for quick A/B experiments."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__
__graft_entry__.build()
from msnets_b200 import cbmv, _lib
from tests._synth import bordered_pair
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
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
print("N=%d  per pair: prep %.4f ms  sadsob %.4f ms  fused %.4f ms   (%s)" % (
    N, a.value / n.value / N, b.value / n.value / N, c.value / n.value / N,
    " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("MSNETS_"))))
