"""Experiment: exchange-form cost vs sub-slab width and count on one GPU (virtual ranks = sub-slabs)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msnets_b200 import cbmv, sharding, _lib
from tests._synth import bordered_pair
H, W, B = 992, 2880, 10
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


for ds in (tuple(int(v) for v in sys.argv[1:]) or (64, 80, 96, 128, 160, 192)):
    row = []
    for S in (1, 2, 3, 4):
        D = ds * S
        sharding.ExchangeSlabMSFeatures.TILE_D_CHOICES = (ds,)
        xs = sharding.ExchangeSlabMSFeatures(1, H + 2 * B, W + 2 * B, maxdisp=D, rank=0, world=1, connect=False, board_h=B,
                                             board_w_left=B, board_w_right=B)
        assert xs.subs == S
        out = torch.empty(xs.shape, dtype=torch.float32, device="cuda")
        ms = timed(lambda: xs(l, r, out=out))
        row.append(ms / (D * H * W / (192 * 540 * 960.0)))
        xs.close(); del out
    print("sub-slab width %3d: ms per config-B pair equivalent at 1/2/3/4 sub-slabs: %s" % (ds, "  ".join("%.3f" % v for v in row)))
