"""GPU parity tests proper: every function on the hot path, called THROUGH THE C ABI
(ctypes -> libmsnets_b200.so), against the CPU oracle on the same seeded inputs and
against the committed reference golden vectors.

Parity classes (stated here as the bar):
  bit-exact (np.array_equal): census, sobel, zsad, sadsob, nccNister, swap_axes,
      get_right_cost / get_left_cost, extract_ratio, feature channels 0-3, WTA argmin,
      LR mask, concat / diff volume
  tolerance: AML (extract_likelihood, channels 4-7)  <= 2e-6 abs on [0,1] outputs
             soft-argmin                             <= 1e-3 px
"""
import os

import numpy as np
import pytest

from tests._synth import bordered_pair, synth_pair

pytestmark = pytest.mark.gpu

AML_ATOL = 2e-6
SOFTARGMIN_ATOL = 1e-3


@pytest.fixture(scope="module")
def ms():
    import msnets_b200
    assert msnets_b200.device_count() >= 1
    return msnets_b200


SHAPES = [(40, 56, 12, 11), (37, 53, 64, 12), (64, 97, 33, 13), (30, 130, 128, 14)]


@pytest.mark.parametrize("H,W,D,seed", SHAPES)
def test_matchers_default_windows_bit_exact(ms, oracle, H, W, D, seed):
    L, R = synth_pair(H, W, seed, shift=5)
    mtc = ms.libmatchers
    assert np.array_equal(mtc.census(L, R, D, 11), oracle.census(L, R, D, 11))
    assert np.array_equal(mtc.nccNister(L, R, D, 3), oracle.nccNister(L, R, D, 3))
    assert np.array_equal(mtc.zsad(L, R, D, 5), oracle.zsad(L, R, D, 5))
    sl, sr = mtc.sobel(L), mtc.sobel(R)
    assert np.array_equal(sl, oracle.sobel(L)) and np.array_equal(sr, oracle.sobel(R))
    assert np.array_equal(mtc.sadsob(sl, sr, D, 5), oracle.sadsob(sl, sr, D, 5))


@pytest.mark.parametrize("censw,nccw,sadw,sobelw", [(7, 5, 3, 3), (9, 7, 7, 7), (5, 1, 1, 9), (16, 4, 6, 16)])
def test_matchers_other_windows_bit_exact(ms, oracle, censw, nccw, sadw, sobelw):
    H, W, D = 48, 75, 20
    L, R = synth_pair(H, W, 99 + censw, shift=4)
    mtc = ms.libmatchers
    assert np.array_equal(mtc.census(L, R, D, censw), oracle.census(L, R, D, censw))
    assert np.array_equal(mtc.nccNister(L, R, D, nccw), oracle.nccNister(L, R, D, nccw))
    assert np.array_equal(mtc.zsad(L, R, D, sadw), oracle.zsad(L, R, D, sadw))
    sl, sr = oracle.sobel(L), oracle.sobel(R)
    assert np.array_equal(mtc.sadsob(sl, sr, D, sobelw), oracle.sadsob(sl, sr, D, sobelw))


def test_sadsob_generic_float_inputs_and_big_sums(ms, oracle):
    """sadsob takes any float32 images; large values push the running sums far past 2^24,
    where only the replayed add order stays bit-exact."""
    rng = np.random.default_rng(5)
    H, W, D = 70, 150, 40
    a = (rng.standard_normal((H, W)) * 3000).astype(np.float32)
    b = (rng.standard_normal((H, W)) * 3000).astype(np.float32)
    assert np.array_equal(ms.libmatchers.sadsob(a, b, D, 5), oracle.sadsob(a, b, D, 5))


def test_edge_images_smaller_than_window(ms, oracle):
    L, R = synth_pair(9, 14, 3, shift=2, patches=False)
    mtc = ms.libmatchers
    c = mtc.census(L, R, 6, 11)
    assert np.all(c == oracle.FILL) and np.array_equal(c, oracle.census(L, R, 6, 11))
    assert np.array_equal(mtc.zsad(L, R, 6, 5), oracle.zsad(L, R, 6, 5))
    assert np.array_equal(mtc.nccNister(L, R, 1, 3), oracle.nccNister(L, R, 1, 3))
    sl, sr = mtc.sobel(L), mtc.sobel(R)
    assert np.array_equal(mtc.sadsob(sl, sr, 6, 5), oracle.sadsob(sl, sr, 6, 5))
    tiny = np.zeros((2, 3), np.uint8)
    assert np.array_equal(mtc.sobel(tiny), oracle.sobel(tiny))


def test_saturated_and_flat_images(ms, oracle):
    """all-255 / all-0 pairs: every NCC window is degenerate (cost +1), census all zero."""
    for val in (0, 255):
        L = np.full((30, 44), val, np.uint8)
        R = L.copy()
        mtc = ms.libmatchers
        assert np.array_equal(mtc.nccNister(L, R, 8, 3), oracle.nccNister(L, R, 8, 3))
        assert np.array_equal(mtc.census(L, R, 8, 11), oracle.census(L, R, 8, 11))
        assert np.array_equal(mtc.zsad(L, R, 8, 5), oracle.zsad(L, R, 8, 5))


def test_featextract_functions(ms, oracle):
    H, W, D = 33, 47, 24
    L, R = synth_pair(H, W, 21, shift=3)
    fte = ms.libfeatextract
    dhw = oracle.zsad(L, R, D, 5)
    hwd = oracle.census(L, R, D, 11)
    assert np.array_equal(fte.swap_axes(dhw), oracle.swap_axes(dhw))
    assert np.array_equal(fte.swap_axes_back(hwd), oracle.swap_axes_back(hwd))
    assert np.array_equal(fte.get_right_cost(hwd), oracle.get_right_cost(hwd))
    assert np.array_equal(fte.get_left_cost(hwd), oracle.get_left_cost(hwd))
    rows = hwd.reshape(-1, D)
    assert np.array_equal(fte.extract_ratio(rows, 0.01), oracle.extract_ratio(rows, 0.01))
    for sigma, vol in ((128.0, hwd), (0.02, oracle.swap_axes(oracle.nccNister(L, R, D, 3))),
                       (20000.0, oracle.swap_axes(dhw))):
        rows = np.ascontiguousarray(vol.reshape(-1, D))
        got, want = fte.extract_likelihood(rows, sigma), oracle.extract_likelihood(rows, sigma)
        assert np.abs(got - want).max() <= AML_ATOL
        assert np.array_equal(got == 0, want == 0) or np.abs(got - want).max() <= AML_ATOL
    with pytest.raises(NotImplementedError):
        fte.extract_likelihood(rows, rows, 1.0)
    with pytest.raises(NotImplementedError):
        fte.get_samples(rows, rows)


def _check_features(got, want, lr=False):
    assert got.shape == want.shape and got.dtype == np.float32
    nviews = 2 if lr else 1
    for v in range(nviews):
        b = 8 * v
        assert np.array_equal(got[b:b + 4], want[b:b + 4]), "normalised channels must be bit-exact"
        assert np.abs(got[b + 4:b + 8] - want[b + 4:b + 8]).max() <= AML_ATOL


@pytest.mark.parametrize("H,W,D,border", [(44, 60, 16, 10), (52, 71, 40, 12), (36, 90, 100, 10)])
def test_get_costs_and_extract_features_mirror(ms, oracle, H, W, D, border):
    L, R = synth_pair(H, W, 300 + D, shift=6)
    costs = ms.cbmv.get_costs(L, R, D, 11, 3, 5, 5, border, border, border)
    want = oracle.get_costs(L, R, D, 11, 3, 5, 5, border, border, border)
    for a, b in zip(costs, want):
        assert a.flags["C_CONTIGUOUS"] and np.array_equal(a, b)
    _check_features(ms.cbmv.extract_features_left(*costs), oracle.extract_features_left(*want))
    _check_features(ms.cbmv.extract_features_lr(*costs), oracle.extract_features_lr(*want), lr=True)


@pytest.mark.parametrize("left_only", [True, False])
@pytest.mark.parametrize("H,W,D,bh,bl,br", [(44, 60, 16, 10, 10, 10), (50, 83, 48, 12, 20, 0), (40, 70, 96, 10, 10, 10)])
def test_ms_features_one_call(ms, oracle, H, W, D, bh, bl, br, left_only):
    L, R = synth_pair(H, W, 500 + W, shift=5)
    got = ms.cbmv.ms_features(L, R, D, left_only=left_only, board_h=bh, board_w_left=bl, board_w_right=br)
    want = oracle.ms_features(L, R, D, board_h=bh, board_w_left=bl, board_w_right=br, left_only=left_only)
    _check_features(got, want, lr=not left_only)


@pytest.mark.parametrize("H,W,D", [(34, 300, 200), (32, 352, 300), (32, 470, 448)])
def test_ms_features_fused_large_d(ms, oracle, H, W, D):
    """D above 192 selects the 256 / 384 / 448 instantiations of the fused kernel (TMA box up to
    256 disparities, LDGSTS staging above); interior tiles exist (W > D + 37 + border)."""
    L, R = synth_pair(H, W, 900 + D, shift=9)
    got = ms.cbmv.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10)
    want = oracle.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10)
    _check_features(got, want)


@pytest.mark.parametrize("D", [500, 640])
def test_ms_features_above_fused_limit_uses_slabs(ms, oracle, D):
    """More than 448 disparities: the volume is cut into slabs of 192 through the fused kernel's
    phase-A form, minima folded across slabs, phases B/C over all D (one GPU)."""
    L, R = synth_pair(31, 700, 1500 + D, shift=9)
    got = ms.cbmv.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10)
    want = oracle.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10)
    _check_features(got, want)


def test_ms_features_fused_interior_tiles(ms, oracle):
    """D = 8 * 6: the d-groups cover D exactly, so tiles right of column D + 5 take the
    instantiation without validity selects; the image is wide enough to have several."""
    L, R = synth_pair(40, 260, 77, shift=11)
    got = ms.cbmv.ms_features(L, R, 48, board_h=10, board_w_left=10, board_w_right=10)
    want = oracle.ms_features(L, R, 48, board_h=10, board_w_left=10, board_w_right=10)
    _check_features(got, want)


@pytest.mark.parametrize("H,W,D,bl,br", [(34, 300, 48, 10, 10), (33, 420, 192, 10, 0), (32, 352, 300, 12, 10),
                                         (31, 275, 64, 0, 3)])
def test_ms_features_both_views_one_pass(ms, oracle, H, W, D, bl, br):
    """extract_features_lr through the two one-pass launches: the right view (channels 8-15) is recomputed
    by tiles that fix the right pixel and slide the left window (ms_fused.cu: phase1_tile_right).  Images wide
    enough for interior (guard-free) blocks, every kernel size, ragged widths, right border 0 (the c.flat[0] fill
    of get_right_cost, featextract.cpp:151, then starts inside the valid costs) and a left border of 0."""
    L, R = synth_pair(H, W, 4100 + D, shift=7)
    got = ms.cbmv.ms_features(L, R, D, left_only=False, board_h=10, board_w_left=bl, board_w_right=br)
    want = oracle.ms_features(L, R, D, left_only=False, board_h=10, board_w_left=bl, board_w_right=br)
    _check_features(got, want, lr=True)


def test_ms_features_both_views_batched_matches_single(ms):
    """A batch through the device API: every pair's right view uses its OWN first-element fill."""
    import torch
    H, W, D, B = 40, 200, 64, 10
    pairs = [synth_pair(H, W, 9000 + i, shift=3 + i) for i in range(3)]
    ex = ms.cbmv.MSFeatureExtractor(3, H, W, maxdisp=D, left_only=False, board_h=B, board_w_left=B, board_w_right=B)
    l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
    r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
    got = ex(l, r).cpu().numpy()
    for i, (L, R) in enumerate(pairs):
        one = ms.cbmv.ms_features(L, R, D, left_only=False, board_h=B, board_w_left=B, board_w_right=B)
        assert np.array_equal(got[i], one)


@pytest.mark.parametrize("H,W,D", [(150, 140, 32), (75, 96, 16), (100, 200, 48), (64, 70, 16)])
def test_sadsob_scan_64_row_bands(ms, oracle, monkeypatch, H, W, D):
    """The SAD-of-Sobel scan of the fused path on bands of 64 table rows (sadsob_scan5x2_kernel; batches that fill
    the machine pick it by themselves, MSNETS_SCAN64=1 forces it here): images with several bands, a last band whose
    second half is empty (processed as a 32-row band) and one whose second half is partial; bit-exact channel 2."""
    monkeypatch.setenv("MSNETS_SCAN64", "1")
    L, R = synth_pair(H, W, 7300 + H, shift=5)
    got = ms.cbmv.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10)
    want = oracle.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10)
    assert np.array_equal(got[:4], want[:4]), "channels 0-3 (2 = SAD-of-Sobel) must be bit-exact"
    from tests._synth import assert_aml_close          # (the one stated AML class, degenerate rows included)
    assert_aml_close(got[4:], want[4:])
    monkeypatch.setenv("MSNETS_SCAN32", "1")          # and the 32-row form gives the same bits
    assert np.array_equal(ms.cbmv.ms_features(L, R, D, board_h=10, board_w_left=10, board_w_right=10), got)


def test_ms_features_generic_path_matches(ms, oracle, monkeypatch):
    """the three-phase global-memory path (used for slabs and non-default windows)."""
    monkeypatch.setenv("MSNETS_FORCE_GENERIC", "1")
    L, R = synth_pair(48, 66, 77, shift=4)
    got = ms.cbmv.ms_features(L, R, 24, board_h=10, board_w_left=10, board_w_right=10)
    want = oracle.ms_features(L, R, 24)
    _check_features(got, want)
    got = ms.cbmv.ms_features(L, R, 24, censw=9, nccw=5, sadw=3, sobelw=7, board_h=8, board_w_left=8,
                              board_w_right=8)
    costs = oracle.get_costs(L, R, 24, 9, 5, 3, 7, 8, 8, 8)
    _check_features(got, oracle.extract_features_left(*costs))


def test_ms_features_batch_is_independent(ms):
    pairs = [synth_pair(44, 64, s, shift=3) for s in (1, 2, 3)]
    L = np.stack([p[0] for p in pairs])
    R = np.stack([p[1] for p in pairs])
    both = ms.cbmv.ms_features(L, R, 16)
    for i in range(3):
        assert np.array_equal(both[i], ms.cbmv.ms_features(L[i], R[i], 16))


def test_against_reference_golden(ms, golden_dir):
    """the committed outputs of the unmodified reference (tests/golden/make_golden.py)."""
    for case in ("small_a", "small_b", "small_c"):
        g = np.load(os.path.join(golden_dir, case + ".npz"))
        H, W, D, seed, shift, border = (int(v) for v in g["meta"])
        L, R = g["L"], g["R"]
        mtc = ms.libmatchers
        assert np.array_equal(mtc.census(L, R, D, 11), g["census"])
        assert np.array_equal(mtc.nccNister(L, R, D, 3), g["ncc"])
        assert np.array_equal(mtc.zsad(L, R, D, 5), g["zsad"])
        assert np.array_equal(mtc.sobel(L), g["sobel_l"])
        assert np.array_equal(mtc.sadsob(mtc.sobel(L), mtc.sobel(R), D, 5), g["sadsob"])
        f = ms.cbmv.ms_features(L, R, D, board_h=border, board_w_left=border, board_w_right=border)
        _check_features(f, g["features_left"])


def test_soft_argmin(ms, oracle, golden_dir):
    import torch
    g = np.load(os.path.join(golden_dir, "soft_argmin.npz"))
    for name in ("sa_small", "sa_peaky", "sa_flat"):
        x = g[name + "_x"]
        assert np.abs(ms.regression.soft_argmin(x) - g[name + "_y"]).max() <= SOFTARGMIN_ATOL
    rng = np.random.default_rng(8)
    for shape in ((2, 192, 20, 32), (1, 48, 7, 9), (3, 5, 4, 4)):
        x = (rng.standard_normal(shape) * 4).astype(np.float32)
        want = oracle.soft_argmin(x)
        got = ms.regression.soft_argmin(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.abs(got - want).max() <= SOFTARGMIN_ATOL
        prob = torch.softmax(torch.from_numpy(x).cuda(), 1)
        mod = ms.regression.disparityregression(shape[1])
        assert np.abs(mod(prob).cpu().numpy() - want).max() <= SOFTARGMIN_ATOL
    onehot = np.full((1, 32, 4, 8), -1e4, np.float32)
    onehot[0, 17] = 50.0
    assert np.allclose(ms.regression.soft_argmin(onehot), 17.0, atol=1e-4)


def test_wta_confidence_lrc(ms, oracle):
    import torch
    L, R = synth_pair(40, 70, 17, shift=6)
    D = 32
    hwd = oracle.census(L, R, D, 11)
    am, m1, m2 = ms.confidence.wta(hwd, "hwd")
    oam, om1, om2 = oracle.wta(hwd)
    assert np.array_equal(am, oam) and np.array_equal(m1, om1) and np.array_equal(m2, om2)
    ncc = oracle.nccNister(L, R, D, 3)  # negative costs, [D,H,W] plane layout
    am, m1, m2 = ms.confidence.wta(torch.from_numpy(ncc).cuda(), "dhw")
    oam, om1, om2 = oracle.wta(np.ascontiguousarray(ncc.transpose(1, 2, 0)))
    assert np.array_equal(am.cpu().numpy(), oam) and np.array_equal(m1.cpu().numpy(), om1)
    assert np.array_equal(m2.cpu().numpy(), om2)
    conf = ms.confidence.pkrn_confidence(m1, m2, 0.01).cpu().numpy()
    assert np.array_equal(conf, oracle.pkrn_confidence(om1, om2, 0.01))
    dl, dr, mask = ms.confidence.lr_consistency(hwd, 1)
    odl, odr, omask = oracle.lr_consistency(hwd, 1)
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr) and np.array_equal(mask, omask)


def test_volume_builders(ms, oracle):
    import torch
    rng = np.random.default_rng(1234)
    for (N, C, H, W, D) in ((2, 4, 6, 20, 8), (1, 3, 5, 13, 16)):
        fl = rng.standard_normal((N, C, H, W)).astype(np.float32)
        fr = rng.standard_normal((N, C, H, W)).astype(np.float32)
        tl, tr = torch.from_numpy(fl).cuda(), torch.from_numpy(fr).cuda()
        assert np.array_equal(ms.volume.concat_volume(tl, tr, D).cpu().numpy(), oracle.concat_volume(fl, fr, D))
        assert np.array_equal(ms.volume.diff_volume(tl, tr, D).cpu().numpy(), oracle.diff_volume(fl, fr, D))


def test_device_resident_extractor_matches_host_api(ms):
    import torch
    L, R = bordered_pair(32, 48, 9, border=10)
    H, W = L.shape
    ex = ms.cbmv.MSFeatureExtractor(2, H, W, maxdisp=24, board_h=10, board_w_left=10, board_w_right=10)
    l = torch.from_numpy(np.stack([L, L])).cuda()
    r = torch.from_numpy(np.stack([R, R])).cuda()
    out = ex(l, r)
    torch.cuda.synchronize()
    want = ms.cbmv.ms_features(L, R, 24, board_h=10, board_w_left=10, board_w_right=10)
    assert np.array_equal(out[0].cpu().numpy(), want) and np.array_equal(out[1].cpu().numpy(), want)


def test_full_size_properties(ms):
    """Config B of BASELINE.json (540x960, D=192, 10 px border) through size-independent
    properties: channels in [0,1]; valid AML columns sum to 1; channel 0 holds k/120;
    the run is deterministic; WTA over channel 0 equals WTA over its AML channel's argmax."""
    import torch
    L, R = bordered_pair(540, 960, 1234, border=10)
    H, W = L.shape
    ex = ms.cbmv.MSFeatureExtractor(1, H, W, maxdisp=192, board_h=10, board_w_left=10, board_w_right=10)
    l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
    f = ex(l, r)
    assert tuple(f.shape) == (1, 8, 192, 540, 960)
    assert float(f.min()) >= 0.0 and float(f.max()) <= 1.0
    s = f[0, 4:8].sum(1)
    assert float((s - 1).abs().max()) <= 2e-5
    k = f[0, 0] * 120.0
    assert float((k - k.round()).abs().max()) <= 1e-4
    f2 = ex(l, r)
    assert torch.equal(f, f2)
    am0 = ms.confidence.wta(f[0, 0].contiguous(), "dhw")[0]
    assert torch.equal(am0.long(), f[0, 0].argmin(0))
    # the true shift (7 px) wins the census WTA on the bulk of the image
    assert float((am0[:, 200:] == 7).float().mean()) > 0.99


@pytest.mark.parametrize("left_only", [True, False])
def test_generate_test_cbmv_device_resident(ms, oracle, golden_dir, left_only):
    """cbmv.generate_test_cbmv mirrors cbmv_generator.py:727-861 but hands back a CUDA tensor."""
    import torch
    g = np.load(os.path.join(golden_dir, "test_cbmv.npz"))
    tag = "left" if left_only else "lr"
    f, h, w, ch, cw = ms.cbmv.generate_test_cbmv(g["L"], g["R"], encoder_ds=16, maxdisp=24, is_left_only=left_only,
                                                 args_dict={"ds_scale": 1})
    assert isinstance(f, torch.Tensor) and f.is_cuda and f.dtype == torch.float32
    want, h2, w2, ch2, cw2 = oracle.generate_test_cbmv(g["L"], g["R"], encoder_ds=16, maxdisp=24,
                                                       is_left_only=left_only)
    assert (h, w, ch, cw) == (h2, w2, ch2, cw2)
    assert [h, w, ch, cw] + list(f.shape) == g["meta_" + tag].tolist()
    got = f.cpu().numpy()
    _check_features(got, want, lr=not left_only)
    # and against the reference's own output (every 5th value is stored)
    ref = g["sub_" + tag]
    sub = got.reshape(-1)[::5]
    C = got.shape[0]
    chan = (np.arange(got.size)[::5] // (got.size // C)) % 8
    assert np.array_equal(sub[chan < 4], ref[chan < 4])
    assert np.abs(sub - ref).max() <= AML_ATOL
    # the reference's default (args_dict=None -> ds_scale = 2): half-size volume (tests/test_gpu_prematch.py)
    f2 = ms.cbmv.generate_test_cbmv(g["L"], g["R"], encoder_ds=16, maxdisp=24, is_left_only=left_only)[0]
    assert tuple(f2.shape) == (f.shape[0], 12, ch // 2, cw // 2)


@pytest.mark.parametrize("H,W,D,seed", [(40, 70, 24, 5), (37, 101, 64, 6), (48, 60, 130, 7)])
def test_fused_wta_byproduct_matches_argmin_of_the_volume(ms, oracle, H, W, D, seed):
    """msn_ms_features_wta_dev: argmin / min / second min of channels 0-3 straight from the fused kernel equal
    what the reference's consumer computes from the volume (np.argmin over D, main_msnet.py:444-448) and the
    oracle's second minimum -- on the kernel's own volume AND on the oracle's."""
    import torch
    L, R = bordered_pair(H, W, seed, border=10, patches=True)
    Hb, Wb = L.shape
    ex = ms.cbmv.MSFeatureExtractor(2, Hb, Wb, maxdisp=D, board_h=10, board_w_left=10, board_w_right=10)
    l = torch.from_numpy(np.stack([L, R])).cuda()          # second pair: swapped images, different content
    r = torch.from_numpy(np.stack([R, L])).cuda()
    wta = ex.empty_wta()
    vol = ex(l, r, wta=wta)
    torch.cuda.synchronize()
    vol = vol.cpu().numpy()
    am, m1, m2 = [t.cpu().numpy() for t in wta]
    want = oracle.ms_features(L, R, D)
    for n in range(2):
        for c in range(4):
            hwd = np.ascontiguousarray(vol[n, c].transpose(1, 2, 0))
            wi, w1, w2 = oracle.wta(hwd)
            assert np.array_equal(am[n, c], wi) and np.array_equal(m1[n, c], w1) and np.array_equal(m2[n, c], w2)
            assert np.array_equal(am[n, c], np.argmin(vol[n, c], axis=0))
    for c in range(4):
        assert np.array_equal(am[0, c], np.argmin(want[c], axis=0))
    conf = ms.confidence.pkrn_confidence(wta[1], wta[2], 0.01).cpu().numpy()
    assert np.array_equal(conf, oracle.pkrn_confidence(m1, m2, 0.01))


def test_soft_argmin_autograd(ms):
    """The reference trains THROUGH F.softmax + disparityregression (gcnet_3dcnn.py:127-141): the kernels carry
    gradients (msn_soft_argmin_backward_dev; expectation backward = grad x arange(D))."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator("cuda").manual_seed(3)
    x = (torch.randn((2, 48, 9, 20), generator=g, device="cuda") * 2).requires_grad_(True)
    w = torch.randn((2, 9, 20), generator=g, device="cuda")
    disp = ms.regression.soft_argmin(x)
    assert disp.grad_fn is not None
    (disp * w).sum().backward()
    got = x.grad.clone()
    x.grad = None
    dvec = torch.arange(48, device="cuda", dtype=torch.float32).view(1, 48, 1, 1)
    ref = torch.sum(F.softmax(x, 1) * dvec, 1)
    (ref * w).sum().backward()
    assert float((disp - ref).detach().abs().max()) <= SOFTARGMIN_ATOL
    assert float((got - x.grad).abs().max()) <= 1e-5 * max(1.0, float(x.grad.abs().max()))
    # the regression half alone, as the patched model uses it
    p = F.softmax(x.detach(), 1).requires_grad_(True)
    e = ms.regression.expected_disparity(p)
    assert e.grad_fn is not None
    (e * w).sum().backward()
    assert torch.allclose(p.grad, w.unsqueeze(1) * dvec)
    with torch.no_grad():
        assert ms.regression.soft_argmin(x).grad_fn is None


@pytest.mark.parametrize("H,W,D", [(40, 70, 24), (37, 101, 64)])
def test_bf16_volume_is_the_rounded_fp32_volume(ms, H, W, D):
    """msn_ms_features_bf16_dev: round-to-nearest-even of the fp32 volume, bit for bit."""
    import torch
    L, R = bordered_pair(H, W, 3, border=10, patches=True)
    l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
    kw = dict(maxdisp=D, board_h=10, board_w_left=10, board_w_right=10)
    f32 = ms.cbmv.MSFeatureExtractor(1, L.shape[0], L.shape[1], **kw)(l, r)
    b16 = ms.cbmv.MSFeatureExtractor(1, L.shape[0], L.shape[1], out_dtype=torch.bfloat16, **kw)(l, r)
    assert b16.dtype == torch.bfloat16 and b16.shape == f32.shape
    assert torch.equal(b16, f32.to(torch.bfloat16))


@pytest.fixture
def exact_aml(ms):
    prev = ms.set_aml_exact(True)
    yield ms
    ms.set_aml_exact(prev)


def test_exact_aml_mode_is_bit_exact_everywhere(exact_aml, oracle, golden_dir):
    """msn_set_aml_exact(1): the reference's fp32 operations with glibc's expf replayed (feature_math.cuh) --
    extract_likelihood, extract_features_left / _lr and the fused volume equal the reference BIT FOR BIT, AML
    channels included: against the golden vectors the unmodified reference produced and against the oracle."""
    import torch
    ms = exact_aml
    assert ms.aml_exact()
    for case in ("small_a", "small_b", "small_c"):
        g = np.load(os.path.join(golden_dir, case + ".npz"))
        H, W, D, seed, shift, border = [int(v) for v in g["meta"]]
        costs = ms.cbmv.get_costs(g["L"], g["R"], D, 11, 3, 5, 5, border, border, border)
        aml = ms.libfeatextract.extract_likelihood(costs[3].reshape(-1, D), 20000.0)
        assert np.array_equal(aml, g["aml_sad"])                                  # featextract.cpp:415-462
        f8 = ms.cbmv.extract_features_left(*costs)
        assert np.array_equal(f8, g["features_left"])                             # all 8 channels
        from tests._synth import digest
        assert digest(ms.cbmv.extract_features_lr(*costs)) == str(g["sha_features_lr"])   # all 16 channels
        fused = ms.cbmv.ms_features(g["L"], g["R"], D, board_h=border, board_w_left=border, board_w_right=border)
        assert np.array_equal(fused, g["features_left"])                          # the one-pass kernel
    # larger, D = 192, degenerate rows included (zero border, flat patch): the fused kernel vs the oracle
    L, R = bordered_pair(70, 260, 17, border=10, patches=True)
    got = ms.cbmv.ms_features(L, R, 192, board_h=10, board_w_left=10, board_w_right=10)
    assert np.array_equal(got, oracle.ms_features(L, R, 192))
    # D > 448: slabs of 192 through phase A, then phases B / C (generic AML kernels in exact mode)
    L, R = bordered_pair(30, 60, 18, border=10, patches=True)
    got = ms.cbmv.ms_features(L, R, 500, board_h=10, board_w_left=10, board_w_right=10)
    assert np.array_equal(got, oracle.ms_features(L, R, 500))
    with pytest.raises(ms.MsnetsError):
        ms.cbmv.MSFeatureExtractor(1, L.shape[0], L.shape[1], maxdisp=64, out_dtype=torch.bfloat16, board_h=10,
                                   board_w_left=10, board_w_right=10)(torch.from_numpy(L[None]).cuda(),
                                                                      torch.from_numpy(R[None]).cuda())
