"""Experiment: cost of the in-kernel exchange protocol by itself -- the Middlebury-sized frame with D disparities on ONE
GPU through (a) the plain fused kernel (D <= 192), (b) the exchange form with world = 1 and `subs` virtual ranks."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msnets_b200 import cbmv, sharding, _lib
from tests._synth import bordered_pair
H, W, B = 1984, 2880, 10
L, R = bordered_pair(H, W, 99, border=B, shift=13)
l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    _lib.lib().msn_profile_enable(1)
    for _ in range(n): fn()
    a, b, c, k = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    _lib.lib().msn_profile_read(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(k))
    _lib.lib().msn_profile_enable(0)
    return c.value / k.value


for D in (int(v) for v in sys.argv[1:] or ["192", "384"]):
    pe = D * H * W / (192 * 540 * 960.0)
    if D <= 192:
        ex = cbmv.MSFeatureExtractor(1, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B)
        out = ex.empty_output()
        ms = timed(lambda: ex(l, r, out=out))
        print("D=%d plain fused kernel: %.2f ms = %.3f ms per config-B pair equivalent" % (D, ms, ms / pe))
        del ex, out
    xs = sharding.ExchangeSlabMSFeatures(1, H + 2 * B, W + 2 * B, maxdisp=D, rank=0, world=1, connect=False, board_h=B,
                                         board_w_left=B, board_w_right=B)
    out = torch.empty(xs.shape, dtype=torch.float32, device="cuda")
    ms = timed(lambda: xs(l, r, out=out))
    print("D=%d exchange form, 1 rank x %d sub-slab(s): %.2f ms = %.3f ms per config-B pair equivalent" % (D, xs.subs, ms, ms / pe))
    xs.close(); del out
