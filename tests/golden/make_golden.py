#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE ITSELF in this container.

What runs: the unmodified reference C++ (oracle/_ref, built by
oracle/build_ref.py from /root/reference/src/cpp/{matchers,featextract}) driven
by the unmodified reference Python glue imported from /root/reference
(src/dataloader/cbmv_generator.py: get_costs :27, extract_features_left :258,
extract_features_lr :84).  The reference has no tests, golden files or known-
answer vectors of its own, so these outputs are what pins
the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

The soft-argmin vectors come from the reference's own three torch lines
(src/models/gcnet_3dcnn.py:127,136-139) evaluated on CPU; the method itself
hard-codes .cuda() (:137) so it cannot be called here.

Cannot run on the GPU box (/root/reference is absent there); the .npz files are
committed.  Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


from tests._synth import synth_pair, digest  # noqa: E402


def import_reference_glue():
    from oracle import ms_oracle as O
    ref = O.load_ref("sse41")
    assert ref is not None, "run oracle/build_ref.py first"
    mtc, fte, _ = ref
    for name in ("skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.image"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
    sys.path.insert(0, REF)
    import src.cpp  # noqa: F401  (reference package)
    lib_pkg = types.ModuleType("src.cpp.lib")
    lib_pkg.__path__ = []
    lib_pkg.libmatchers, lib_pkg.libfeatextract = mtc, fte
    sys.modules["src.cpp.lib"] = lib_pkg
    sys.modules["src.cpp.lib.libmatchers"] = mtc
    sys.modules["src.cpp.lib.libfeatextract"] = fte
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import src.dataloader.cbmv_generator as gen
    return mtc, fte, gen


# name: (H, W, D, seed, shift, border)  -- sizes are of the bordered pair
SMALL = {
    "small_a": (36, 48, 10, 101, 5, 10),     # generic
    "small_b": (33, 41, 48, 102, 3, 10),     # D > W (every x has a truncated d range)
    "small_c": (40, 51, 14, 103, 7, 12),     # W not a multiple of 16, board_h 12 as in training
}
# arrays stored in full for the small cases (everything else: sha256 only)
FULL_KEYS = ("census", "ncc", "zsad", "sobel_l", "sadsob", "features_left", "aml_sad")
MEDIUM = {
    "medium_a": (120, 168, 48, 201, 7, 10),
    "medium_b": (96, 250, 96, 202, 11, 10),
}


def run_case(mtc, fte, gen, H, W, D, seed, shift, border, full):
    L, R = synth_pair(H, W, seed, shift)
    out = {"L": L, "R": R, "meta": np.array([H, W, D, seed, shift, border], np.int64)}
    cen = mtc.census(L, R, D, 11)
    ncc = mtc.nccNister(L, R, D, 3)
    zs = mtc.zsad(L, R, D, 5)
    sl, sr = mtc.sobel(L), mtc.sobel(R)
    ss = mtc.sadsob(sl, sr, D, 5)
    costs = gen.get_costs(L, R, D, 11, 3, 5, 5, border, border, border)
    f8 = gen.extract_features_left(*costs)
    f16 = gen.extract_features_lr(*costs)
    rc = fte.get_right_cost(costs[0])
    aml = fte.extract_likelihood(costs[3].reshape(-1, D), 20000.0)
    pk = fte.extract_ratio(costs[0].reshape(-1, D), 0.01)
    named = {"census": cen, "ncc": ncc, "zsad": zs, "sobel_l": sl, "sobel_r": sr, "sadsob": ss,
             "cost_census": costs[0], "cost_ncc": costs[1], "cost_sobel": costs[2],
             "cost_sad": costs[3], "features_left": f8, "features_lr": f16,
             "right_census": rc, "aml_sad": aml, "pkrn_census": pk}
    for k, v in named.items():
        assert v.dtype == np.float32, (k, v.dtype)
        out["sha_" + k] = np.array(digest(v))
        if full and k in FULL_KEYS:
            out[k] = v
    return out


def soft_argmin_vectors():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(1234)
    out = {}
    for name, (N, D, H, W), scale in (("sa_small", (2, 24, 9, 13), 1.0),
                                      ("sa_peaky", (1, 192, 6, 10), 12.0),
                                      ("sa_flat", (1, 48, 5, 7), 0.0)):
        x = (rng.standard_normal((N, D, H, W)) * scale).astype(np.float32)
        t = torch.from_numpy(x)
        prob = F.softmax(t, 1)                                          # gcnet_3dcnn.py:127
        disp = torch.tensor(np.array(range(D)), dtype=torch.float32).view(1, D, 1, 1)
        disp = disp.repeat(N, 1, H, W)                                   # :136-138
        res = torch.sum(prob * disp, 1)                                  # :139
        out[name + "_x"] = x
        out[name + "_y"] = res.numpy().astype(np.float32)
    return out


def test_cbmv_vectors(gen):
    """generate_test_cbmv (cbmv_generator.py:727-861) run on two PNG files written here, with
    ds_scale = 1 (the reference's ds_scale = 2 branch needs skimage, absent in this container).
    Stores the inputs, the returned sizes, the SHA-256 of the feature tensors and every 5th
    value of them."""
    import tempfile
    import cv2
    L, R = synth_pair(27, 45, 777, 4)
    out = {"L": L, "R": R}
    with tempfile.TemporaryDirectory() as td:
        fl, fr = os.path.join(td, "l.png"), os.path.join(td, "r.png")
        cv2.imwrite(fl, L)
        cv2.imwrite(fr, R)
        for tag, left_only in (("left", True), ("lr", False)):
            ad = gen.get_default_args_dict()
            ad["ds_scale"] = 1
            f, h, w, ch, cw = gen.generate_test_cbmv(fl, fr, encoder_ds=16, maxdisp=24, args_dict=ad,
                                                     is_left_only=left_only)
            f = f.numpy()
            assert f.dtype == np.float32
            out["meta_" + tag] = np.array([h, w, ch, cw] + list(f.shape), np.int64)
            out["sha_" + tag] = np.array(digest(f))
            out["sub_" + tag] = f.reshape(-1)[::5].copy()
    return out


def main():
    mtc, fte, gen = import_reference_glue()
    np.savez_compressed(os.path.join(HERE, "test_cbmv.npz"), **test_cbmv_vectors(gen))
    for name, cfg in SMALL.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(mtc, fte, gen, *cfg, full=True))
    for name, cfg in MEDIUM.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(mtc, fte, gen, *cfg, full=False))
    np.savez_compressed(os.path.join(HERE, "soft_argmin.npz"), **soft_argmin_vectors())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
