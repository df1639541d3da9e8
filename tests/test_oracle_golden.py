"""CPU: pins oracle/ (the C restatement + NumPy glue) to the golden vectors that
tests/golden/make_golden.py produced by running the unmodified reference
(compiled C++ + imported cbmv_generator.py) in the build container."""
import glob
import os

import numpy as np
import pytest

from tests._synth import digest, synth_pair

CASES = sorted(os.path.basename(p)[:-4] for p in
               glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
               if not p.endswith("soft_argmin.npz") and not p.endswith("test_cbmv.npz"))


def _oracle_outputs(O, L, R, D, border):
    cen = O.census(L, R, D, 11)
    ncc = O.nccNister(L, R, D, 3)
    zs = O.zsad(L, R, D, 5)
    sl, sr = O.sobel(L), O.sobel(R)
    ss = O.sadsob(sl, sr, D, 5)
    costs = O.get_costs(L, R, D, 11, 3, 5, 5, border, border, border)
    return {"census": cen, "ncc": ncc, "zsad": zs, "sobel_l": sl, "sobel_r": sr, "sadsob": ss,
            "cost_census": costs[0], "cost_ncc": costs[1], "cost_sobel": costs[2],
            "cost_sad": costs[3],
            "features_left": O.extract_features_left(*costs),
            "features_lr": O.extract_features_lr(*costs),
            "right_census": O.get_right_cost(costs[0]),
            "aml_sad": O.extract_likelihood(costs[3].reshape(-1, D), 20000.0),
            "pkrn_census": O.extract_ratio(costs[0].reshape(-1, D), 0.01)}


def test_golden_files_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(oracle, golden_dir, case):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    H, W, D, seed, shift, border = (int(v) for v in g["meta"])
    L, R = synth_pair(H, W, seed, shift)
    assert np.array_equal(L, g["L"]) and np.array_equal(R, g["R"]), "synthetic generator drifted"
    got = _oracle_outputs(oracle, L, R, D, border)
    for key, val in got.items():
        # every function on the path reproduces the reference bit for bit on the CPU
        assert digest(val) == str(g["sha_" + key]), "%s: %s differs from the reference" % (case, key)
        if key in g.files:
            assert np.array_equal(val, g[key])


def test_soft_argmin_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "soft_argmin.npz"))
    for name in ("sa_small", "sa_peaky", "sa_flat"):
        y = oracle.soft_argmin(g[name + "_x"])
        # north-star tolerance for soft-argmin: <= 1e-3 px
        assert np.abs(y - g[name + "_y"]).max() <= 1e-3, name


def test_documented_support_regions(oracle):
    """where each matcher writes; everything else is fill / zero."""
    H, W, D = 40, 50, 8
    L, R = synth_pair(H, W, 7, 3, patches=False)
    fill = oracle.FILL
    cen = oracle.census(L, R, D, 11)
    valid = cen != fill
    assert valid[5:H - 6, 5 + D:W - 6, :].all() and not valid[:5].any() and not valid[H - 6:].any()
    assert not valid[:, W - 6:].any()
    for d in range(D):
        assert not valid[:, :5 + d, d].any() and valid[5:H - 6, 5 + d:W - 6, d].all()
    for fn, w in ((oracle.nccNister, 3), (oracle.zsad, 5)):
        v = fn(L, R, D, w) != fill
        wc = w // 2
        for d in range(D):
            assert v[d, wc:H - w + wc, wc + d:W - w + wc].all()
            v[d, wc:H - w + wc, wc + d:W - w + wc] = False
        assert not v.any()
    s = oracle.sobel(L)
    assert not s[0].any() and not s[H - 2:].any() and not s[:, 0].any() and not s[:, W - 2:].any()


def test_aml_properties(oracle):
    rng = np.random.default_rng(3)
    c = rng.uniform(0, 120, (64, 48)).astype(np.float32)
    c[5] = oracle.FILL
    c[6, 10:] = oracle.FILL
    a = oracle.extract_likelihood(c, 128.0)
    assert np.all(a[5] == 0)
    s = a.sum(1)
    assert np.allclose(np.delete(s, 5), 1.0, atol=1e-5)
    assert np.all(a[6, 10:] == 0)


@pytest.mark.parametrize("tag,left_only", [("left", True), ("lr", False)])
def test_oracle_generate_test_cbmv_matches_reference(oracle, golden_dir, tag, left_only):
    """generate_test_cbmv (cbmv_generator.py:727-861): pad-to-multiple policy, 10 px border and
    feature assembly, against the reference function run on PNG files (ds_scale = 1)."""
    g = np.load(os.path.join(golden_dir, "test_cbmv.npz"))
    f, h, w, ch, cw = oracle.generate_test_cbmv(g["L"], g["R"], encoder_ds=16, maxdisp=24, is_left_only=left_only)
    assert [h, w, ch, cw] + list(f.shape) == g["meta_" + tag].tolist()
    assert digest(f) == str(g["sha_" + tag])
    assert np.array_equal(f.reshape(-1)[::5], g["sub_" + tag])


@pytest.mark.parametrize("shape,scale", [((64, 96), 0.5), ((37, 53), 0.5), ((40, 60), 0.25), ((30, 30), 0.5)])
def test_rescale_replay_equals_scipy_composition(oracle, shape, scale):
    """rescale_antialiased = scipy.ndimage.gaussian_filter + zoom + clip as skimage >= 0.19 composes them;
    rescale_antialiased_replay = the same arithmetic written out (the form the CUDA kernel follows)."""
    rng = np.random.default_rng(7)
    im = rng.integers(0, 256, shape, dtype=np.uint8)
    if shape == (30, 30):
        im = np.maximum(im, 17).astype(np.uint8)      # range clip with a positive minimum
    f = im.astype(np.float32) / 255.0
    a, b = oracle.rescale_antialiased(f, scale), oracle.rescale_antialiased_replay(f, scale)
    assert a.dtype == b.dtype == np.float32 and np.array_equal(a, b)


def test_oracle_generate_test_cbmv_ds2_equals_reference_function(oracle):
    """The oracle's generate_test_cbmv at the reference's default ds_scale = 2 against the unmodified
    reference function (its skimage.transform.rescale call served by the oracle's restatement)."""
    import tempfile
    import cv2
    from oracle import ref_glue
    oracle.lib()
    gen = ref_glue.load_generator(oracle.MTC, oracle.FTE, rescale=oracle.rescale_antialiased)
    if gen is None:
        pytest.skip("cbmv_generator.py not available")
    L, R = synth_pair(70, 150, 5, 6)
    with tempfile.TemporaryDirectory() as td:
        fl, fr = os.path.join(td, "l.png"), os.path.join(td, "r.png")
        cv2.imwrite(fl, L)
        cv2.imwrite(fr, R)
        want, h, w, ch, cw = gen.generate_test_cbmv(fl, fr, encoder_ds=16, maxdisp=32)
    got = oracle.generate_test_cbmv(L, R, encoder_ds=16, maxdisp=32, args_dict={"ds_scale": 2})
    assert (h, w, ch, cw) == got[1:]
    assert np.array_equal(want.numpy(), got[0])
