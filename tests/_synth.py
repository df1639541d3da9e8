"""Deterministic synthetic inputs shared by the golden generator, the tests and bench.py."""
import hashlib

import numpy as np


def synth_pair(H, W, seed, shift=7, patches=True):
    """SURVEY.md 8d synthetic pair: L ~ U{0..255}, R = L rolled left by `shift`
    with the wrapped strip re-randomised; parity sets add a flat patch
    (degenerate NCC window), a saturated block and a smooth ramp."""
    rng = np.random.default_rng(seed)
    L = rng.integers(0, 256, (H, W), dtype=np.uint8)
    R = np.roll(L, -shift, axis=1).copy()
    R[:, W - shift:] = rng.integers(0, 256, (H, shift), dtype=np.uint8)
    if patches:
        h4, w4 = H // 4, W // 4
        L[h4:h4 + 10, w4:w4 + 20] = 77
        R[h4 + 1:h4 + 11, max(w4 - shift, 0):max(w4 - shift, 0) + 20] = 77
        L[2 * h4:2 * h4 + 6, 2 * w4:2 * w4 + 9] = 255
        R[2 * h4:2 * h4 + 6, 2 * w4:2 * w4 + 9] = 255
        ramp = (np.arange(W)[None, :] * 2 + np.arange(H)[:, None]).astype(np.uint8)
        L[3 * h4:, :w4] = ramp[3 * h4:, :w4]
        R[3 * h4:, :w4] = ramp[3 * h4:, shift:w4 + shift]
    return np.ascontiguousarray(L), np.ascontiguousarray(R)


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + ":" + "x".join(map(str, a.shape)) + ":" + str(a.dtype)


def bordered_pair(h, w, seed, border=10, shift=7, patches=False):
    """Test-mode policy of generate_test_cbmv (cbmv_generator.py:819-823): an
    h x w pair zero-padded by `border` px on all four sides."""
    L, R = synth_pair(h, w, seed, shift, patches)
    pad = ((border, border), (border, border))
    return (np.ascontiguousarray(np.pad(L, pad, "constant")),
            np.ascontiguousarray(np.pad(R, pad, "constant")))
