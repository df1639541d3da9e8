"""Fused kernel in the exact AML mode vs the fast mode (config B, batch 4): ms per pair."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msnets_b200 as ms
from msnets_b200 import cbmv, _lib
from tests._synth import bordered_pair
N, H, W, D, B = 4, 540, 960, 192, 10
pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
ex = cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
out = ex.empty_output()
for mode in (False, True):
    ms.set_aml_exact(mode)
    for _ in range(3): ex(l, r, out=out)
    torch.cuda.synchronize()
    _lib.lib().msn_profile_enable(1)
    for _ in range(10): ex(l, r, out=out)
    a, b, c, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    _lib.lib().msn_profile_read(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(n))
    _lib.lib().msn_profile_enable(0)
    print("aml_exact=%s: fused %.4f ms/pair" % (mode, c.value / n.value / N))
ms.set_aml_exact(False)
