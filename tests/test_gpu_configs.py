"""GPU: the five BASELINE.json configurations at (or close to) their stated sizes.

A  256x512  D=192 one pair           full comparison against the reference / oracle
C  375x1242 D=192 KITTI-shaped       padded to 384x1248 as generate_test_cbmv does, full comparison
P  544x960  D=192 PSMNet-shaped      concat volume at 1/4 resolution + soft-argmin vs the reference's torch ops
M  1984x2880 D=640 Middlebury-shaped one disparity slab (80 of 640) as rank 2 of 8 would compute it;
                                     the local matchers are checked bit-exactly on a window against the oracle
(B, 540x960 D=192 batch 8, is bench.py's workload and test_gpu_parity.test_full_size_properties.)
"""
import ctypes

import numpy as np
import pytest

from tests._synth import bordered_pair, synth_pair

pytestmark = pytest.mark.gpu
AML_ATOL = 2e-6


def _reference_features(oracle, L, R, D, border):
    """The unmodified reference C++ when oracle/_ref travelled to this box, else the C oracle."""
    ref = oracle.load_ref()
    if ref is not None:
        return oracle.ms_features(L, R, D, board_h=border, board_w_left=border, board_w_right=border,
                                  mtc=ref[0], fte=ref[1]), "oracle/_ref"
    return oracle.ms_features(L, R, D, board_h=border, board_w_left=border, board_w_right=border), "C oracle"


# AML on rows that repeat one cost many times (zero padding, flat regions): the reference's
# sequential fp32 denominator then rounds the SAME addend the SAME way up to D times, a drift of
# up to D * 2^-24 = 1.1e-5 whose direction depends on the last bit of expf.  The SFU exponential
# differs from glibc's in that bit, so such rows can differ by the drift itself; everywhere else
# the 2e-6 bound holds.  Both bounds are asserted: a hard cap and a rarity bound.
AML_DEGENERATE_ATOL = 1.2e-5
AML_DEGENERATE_FRACTION = 1e-6


def _compare(got, want):
    assert got.shape == want.shape
    for c in range(4):
        assert np.array_equal(got[c], want[c]), "channel %d must be bit-exact" % c
    for c in range(4, 8):
        err = np.abs(got[c] - want[c])
        assert float(err.max()) <= AML_DEGENERATE_ATOL, "AML channel %d" % c
        assert float((err > AML_ATOL).mean()) <= AML_DEGENERATE_FRACTION, "AML channel %d" % c


def test_config_a_256x512_d192(oracle):
    import msnets_b200 as ms
    L, R = bordered_pair(256, 512, 1234, border=10)
    want, who = _reference_features(oracle, L, R, 192, 10)
    got = ms.cbmv.ms_features(L, R, 192, board_h=10, board_w_left=10, board_w_right=10)
    assert got.shape == (8, 192, 256, 512)
    _compare(got, want)


def test_config_a_256x512_d192_both_views(oracle):
    """configs[0] with is_left_only=False: all 16 channels (extract_features_lr, cbmv_generator.py:84-254) from the
    two one-pass launches, against the unmodified reference C++ when it travelled, else the C oracle."""
    import msnets_b200 as ms
    L, R = bordered_pair(256, 512, 4321, border=10)
    ref = oracle.load_ref()
    kw = dict(board_h=10, board_w_left=10, board_w_right=10, left_only=False)
    want = oracle.ms_features(L, R, 192, mtc=ref[0], fte=ref[1], **kw) if ref is not None else oracle.ms_features(L, R, 192, **kw)
    got = ms.cbmv.ms_features(L, R, 192, **kw)
    assert got.shape == (16, 192, 256, 512)
    _compare(got[:8], want[:8])
    _compare(got[8:], want[8:])


def test_config_b_sceneflow_540x960_d192_full_compare(oracle):
    """bench.py's own workload at full size, value by value: one 540x960 (560x980 bordered), D=192 pair
    computed as member 1 of a batch of two through the batched device API (the bench path), against the
    unmodified reference C++ (oracle/_ref) when it travelled, else the C oracle."""
    import torch
    import msnets_b200 as ms
    D, border = 192, 10
    pairs = [bordered_pair(540, 960, 1234 + i, border=border) for i in range(2)]
    want, who = _reference_features(oracle, pairs[1][0], pairs[1][1], D, border)
    l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
    r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
    ex = ms.cbmv.MSFeatureExtractor(2, 560, 980, maxdisp=D, board_h=border, board_w_left=border,
                                    board_w_right=border)
    got = ex(l, r)
    assert tuple(got.shape) == (2, 8, 192, 540, 960)
    _compare(got[1].cpu().numpy(), want)


def test_config_c_kitti_375x1242_d192(oracle):
    import msnets_b200 as ms
    h, w = 375, 1242
    L0, R0 = synth_pair(h, w, 4321, shift=9, patches=True)
    ph, pw = (32 - h % 32) % 32, (32 - w % 32) % 32                       # cbmv_generator.py:780-788
    L = np.pad(L0, ((ph, 0), (0, pw)), "constant")
    R = np.pad(R0, ((ph, 0), (0, pw)), "constant")
    assert L.shape == (384, 1248)
    Lb = np.ascontiguousarray(np.pad(L, ((10, 10), (10, 10)), "constant"))  # :819-823
    Rb = np.ascontiguousarray(np.pad(R, ((10, 10), (10, 10)), "constant"))
    want, who = _reference_features(oracle, Lb, Rb, 192, 10)
    got = ms.cbmv.ms_features(Lb, Rb, 192, board_h=10, board_w_left=10, board_w_right=10)
    assert got.shape == (8, 192, 384, 1248)
    _compare(got, want)


def test_config_p_psmnet_volume_and_softargmin():
    import torch
    import torch.nn.functional as F
    import msnets_b200 as ms
    g = torch.Generator("cuda").manual_seed(1234)
    N, C, h4, w4, D4 = 4, 32, 136, 240, 48                                  # 544x960 at 1/4, D=192/4
    fl = torch.randn((N, C, h4, w4), generator=g, device="cuda")
    fr = torch.randn((N, C, h4, w4), generator=g, device="cuda")
    vol = ms.volume.concat_volume(fl, fr, D4)
    assert tuple(vol.shape) == (N, 2 * C, D4, h4, w4)                       # dres0's 64 channels, psmnet_3dcnn.py:96
    for d in (0, 1, 17, 47):
        assert torch.equal(vol[:, :C, d, :, d:], fl[:, :, :, d:])
        assert torch.equal(vol[:, C:, d, :, d:], fr[:, :, :, :w4 - d])
        assert float(vol[:, :, d, :, :d].abs().sum()) == 0.0
    logits = torch.randn((2, 192, 544, 960), generator=g, device="cuda")
    disp = ms.regression.soft_argmin(logits)
    prob = F.softmax(logits, 1)                                             # psmnet_3dcnn.py:170-174
    ref = torch.sum(prob * torch.arange(192, device="cuda", dtype=torch.float32).view(1, 192, 1, 1), 1)
    assert float((disp - ref).abs().max()) <= 1e-3


def test_config_m_middlebury_slab(oracle):
    """Rank 2 of 8 of the 1984x2880, D=640 frame: disparities [160, 240)."""
    import torch
    import msnets_b200 as ms
    from msnets_b200 import _lib, cbmv, sharding
    H, W, D, border = 1984 + 20, 2880 + 20, 640, 10
    L, R = bordered_pair(1984, 2880, 99, border=border, shift=13)
    d0, dn = sharding.shard_range(D, 2, 8)
    assert (d0, dn) == (160, 80)
    p = cbmv.make_params(D, board_h=border, board_w_left=border, board_w_right=border, d_begin=d0, d_count=dn)
    shape = cbmv.output_shape(1, H, W, p)
    assert shape == (1, 8, 80, 1984, 2880)
    lib = _lib.lib()
    l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
    ws = torch.empty(lib.msn_ms_slab_workspace_bytes(1, H, W, ctypes.byref(p)), dtype=torch.uint8, device="cuda")
    out = torch.empty(shape, dtype=torch.float32, device="cuda")
    mins = torch.empty((1, 4, shape[3], shape[4]), dtype=torch.float32, device="cuda")
    den = torch.empty_like(mins)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.msn_ms_slab_phase_a_dev(l.data_ptr(), r.data_ptr(), 1, H, W, ctypes.byref(p), None,
                                           out.data_ptr(), mins.data_ptr(), ws.data_ptr(), ws.numel(), st))
    _lib.check(lib.msn_ms_slab_phase_b_dev(out.data_ptr(), mins.data_ptr(), 1, shape[3], shape[4], ctypes.byref(p),
                                           den.data_ptr(), st))
    _lib.check(lib.msn_ms_slab_phase_c_dev(out.data_ptr(), mins.data_ptr(), den.data_ptr(), 1, shape[3], shape[4],
                                           ctypes.byref(p), st))
    torch.cuda.synchronize()
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0 and bool(torch.isfinite(out).all())
    # single-rank "merge": the slab's AML columns sum to 1 over its own disparities where any cost is valid
    s = out[0, 4:8].sum(1)
    valid = mins[0] != 2147483648.0
    assert float((s[valid] - 1).abs().max()) <= 2e-5
    # the three local matchers, bit-exact on a window: oracle on a crop that holds every tap
    y0, y1, x0, x1 = 700, 740, 1500, 1560                                   # window in bordered coordinates
    cx0 = x0 - (d0 + dn - 1) - 6
    Lc = np.ascontiguousarray(L[y0 - 6:y1 + 6, cx0:x1 + 6])
    Rc = np.ascontiguousarray(R[y0 - 6:y1 + 6, cx0:x1 + 6])
    Dc = d0 + dn
    cen = oracle.census(Lc, Rc, Dc, 11)[6:-6, x0 - cx0:x1 - cx0, d0:]        # [y,x,d]
    ncc = oracle.nccNister(Lc, Rc, Dc, 3)[d0:, 6:-6, x0 - cx0:x1 - cx0]      # [d,y,x]
    zs = oracle.zsad(Lc, Rc, Dc, 5)[d0:, 6:-6, x0 - cx0:x1 - cx0]
    win = out[0, :, :, y0 - border:y1 - border, x0 - border:x1 - border].cpu().numpy()
    assert np.array_equal(win[0], (np.clip(cen, 0., 120.) / 120.).transpose(2, 0, 1))
    assert np.array_equal(win[1], (1 + np.clip(ncc, -1., 1.)) / 2)
    assert np.array_equal(win[3], np.clip(zs, 0., 2 ** 13) / float(2 ** 13))


def _virtual_slab_ranks(L, R, D, border, world):
    """The slab-sharded pipeline with `world` virtual ranks on one GPU: phase A per slab, the two
    all-reduces done with torch ops, phases B and C -> the assembled [8, D, h, w] volume."""
    import torch
    from msnets_b200 import _lib, cbmv, sharding
    lib = _lib.lib()
    N, (H, W) = 1, L.shape
    l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
    st = torch.cuda.current_stream().cuda_stream
    slabs = []
    for rank in range(world):
        d0, dn = sharding.shard_range(D, rank, world)
        p = cbmv.make_params(D, board_h=border, board_w_left=border, board_w_right=border, d_begin=d0, d_count=dn)
        shape = cbmv.output_shape(N, H, W, p)
        ws = torch.empty(lib.msn_ms_slab_workspace_bytes(N, H, W, ctypes.byref(p)), dtype=torch.uint8, device="cuda")
        out = torch.empty(shape, dtype=torch.float32, device="cuda")
        mins = torch.empty((N, 4, shape[3], shape[4]), dtype=torch.float32, device="cuda")
        _lib.check(lib.msn_ms_slab_phase_a_dev(l.data_ptr(), r.data_ptr(), N, H, W, ctypes.byref(p), None,
                                               out.data_ptr(), mins.data_ptr(), ws.data_ptr(), ws.numel(), st))
        slabs.append((p, out, mins, shape))
        del ws
    gmin = slabs[0][2].clone()
    for s_ in slabs[1:]:
        gmin = torch.minimum(gmin, s_[2])                       # all-reduce(min)
    gden = torch.zeros_like(gmin)
    for p, out, _, shape in slabs:
        den = torch.empty_like(gmin)
        _lib.check(lib.msn_ms_slab_phase_b_dev(out.data_ptr(), gmin.data_ptr(), N, shape[3], shape[4],
                                               ctypes.byref(p), den.data_ptr(), st))
        gden += den                                              # all-reduce(sum), rank order
    for p, out, _, shape in slabs:
        _lib.check(lib.msn_ms_slab_phase_c_dev(out.data_ptr(), gmin.data_ptr(), gden.data_ptr(), N, shape[3],
                                               shape[4], ctypes.byref(p), st))
    torch.cuda.synchronize()
    return torch.cat([s_[1] for s_ in slabs], dim=2)[0].cpu().numpy()


def test_config_m_full_width_strip_all_channels_after_merge(oracle):
    """Config M's width and disparity range on a row strip: 2880 (+20) columns, D = 640 in eight slabs
    of 80, 76 (+20) rows.  Every channel of the merged volume is compared with the reference run on the
    same strip: the SAD-of-Sobel channel (whole-row fp32 scans at W = 2900, sums far above 2^24) and
    the AML channels after the cross-slab min / sum merges included."""
    D, border = 640, 10
    Lf, Rf = bordered_pair(1984, 2880, 99, border=0, shift=13)
    L0, R0 = Lf[900:976], Rf[900:976]
    L = np.ascontiguousarray(np.pad(L0, ((border, border), (border, border)), "constant"))
    R = np.ascontiguousarray(np.pad(R0, ((border, border), (border, border)), "constant"))
    want, who = _reference_features(oracle, L, R, D, border)
    got = _virtual_slab_ranks(L, R, D, border, 8)
    assert got.shape == want.shape == (8, 640, 76, 2880)
    for c in range(4):
        assert np.array_equal(got[c], want[c]), "channel %d must be bit-exact" % c
    for c in range(4, 8):                                        # cross-slab sum order differs by construction
        assert float(np.abs(got[c] - want[c]).max()) <= AML_DEGENERATE_ATOL
        assert float((np.abs(got[c] - want[c]) > AML_ATOL).mean()) <= 1e-5


def test_config_m_sadsob_deep_rows_full_width(oracle):
    """SAD-of-Sobel at config M's width 330 rows down the frame (13 row bands of the scan): the fp32 table
    of a row depends on every row above it, so the oracle runs on the top 336 rows at full width; one
    slab of disparities (rank 2 of 8: [160, 240)) is compared bit for bit on rows 300..320."""
    import msnets_b200 as ms
    from msnets_b200 import sharding
    border = 10
    L, R = bordered_pair(1984, 2880, 99, border=border, shift=13)
    Ls, Rs = np.ascontiguousarray(L[:336]), np.ascontiguousarray(R[:336])
    d0, dn = sharding.shard_range(640, 2, 8)
    sl, sr = oracle.sobel(Ls), oracle.sobel(Rs)
    want = oracle.sadsob(sl, sr, d0 + dn, 5)[d0:, 300:320]        # [d, y, x], raw costs
    mtc = ms.libmatchers
    gl, gr = mtc.sobel(Ls), mtc.sobel(Rs)
    assert np.array_equal(gl, sl) and np.array_equal(gr, sr)
    got = mtc.sadsob(gl, gr, d0 + dn, 5)[d0:, 300:320]
    assert np.array_equal(got, want)
    assert float(want[want < 1e9].max()) > 2 ** 13 / 4
    # the same rows through the slab path's own scan (sadsob_scan5_kernel, pitch 4096) and phase A: channel 2
    import torch
    from msnets_b200 import _lib, cbmv
    lib = _lib.lib()
    H, W = Ls.shape
    p = cbmv.make_params(640, board_h=border, board_w_left=border, board_w_right=border, d_begin=d0, d_count=dn)
    shape = cbmv.output_shape(1, H, W, p)
    l, r = torch.from_numpy(Ls[None]).cuda(), torch.from_numpy(Rs[None]).cuda()
    ws = torch.empty(lib.msn_ms_slab_workspace_bytes(1, H, W, ctypes.byref(p)), dtype=torch.uint8, device="cuda")
    out = torch.empty(shape, dtype=torch.float32, device="cuda")
    mins = torch.empty((1, 4, shape[3], shape[4]), dtype=torch.float32, device="cuda")
    _lib.check(lib.msn_ms_slab_phase_a_dev(l.data_ptr(), r.data_ptr(), 1, H, W, ctypes.byref(p), None, out.data_ptr(),
                                           mins.data_ptr(), ws.data_ptr(), ws.numel(),
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ch2 = out[0, 2, :, 300 - border:320 - border].cpu().numpy()                 # [d, y, x] cropped
    want2 = (np.clip(want[:, :, border:W - border], 0., 2 ** 13) / float(2 ** 13)).astype(np.float32)
    assert np.array_equal(ch2, want2)


def test_config_a_exact_aml_mode_all_channels_bit_exact(oracle):
    """Config A (256x512, D=192) with msn_set_aml_exact(1): all eight channels array_equal to the unmodified
    reference C++ -- no tolerance class at all."""
    import msnets_b200 as ms
    L, R = bordered_pair(256, 512, 1234, border=10)
    want, who = _reference_features(oracle, L, R, 192, 10)
    prev = ms.set_aml_exact(True)
    try:
        got = ms.cbmv.ms_features(L, R, 192, board_h=10, board_w_left=10, board_w_right=10)
    finally:
        ms.set_aml_exact(prev)
    assert np.array_equal(got, want)
