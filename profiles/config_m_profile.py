import os, sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
from msnets_b200 import cbmv, _lib
from tests._synth import bordered_pair
H, W, D, B = 1984, 2880, int(sys.argv[1]) if len(sys.argv) > 1 else 640, 10
L, R = bordered_pair(H, W, 99, border=B, shift=13)
l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
ex = cbmv.MSFeatureExtractor(1, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
out = ex.empty_output()
ex(l, r, out=out); torch.cuda.synchronize()
_lib.lib().msn_profile_enable(1)
for _ in range(3): ex(l, r, out=out)
a, b, c, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
_lib.lib().msn_profile_read(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(n))
print("D=%d calls %d: prep %.2f sadsob %.2f fused %.2f ms per call" % (D, n.value, a.value/n.value, b.value/n.value, c.value/n.value))
