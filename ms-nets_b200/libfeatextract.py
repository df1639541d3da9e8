"""Drop-in for the reference's `src.cpp.lib.libfeatextract` (featextract.cpp:529-564).

Live exports (the ones cbmv_generator.py calls) run as CUDA kernels:
swap_axes, get_right_cost, extract_likelihood(vol, sigma), extract_ratio(vol, e),
plus swap_axes_back / get_left_cost.  The CBMV random-forest training helpers the
reference still exports but never calls (get_cost, generate_d_indices,
get_samples, generate_labels and the 3-argument overloads; SURVEY.md 8a "dead
exports") raise NotImplementedError naming the reference lines.
"""
import numpy as np

from . import _lib


def _vol(a, ndim, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != ndim:
        raise ValueError("%s: expected a %d-D float32 numpy array" % (name, ndim))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("%s: array must be C-contiguous" % name)
    return a


def swap_axes(cost):
    """featextract.cpp:49-76: [D,H,W] -> [H,W,D]."""
    c = _vol(cost, 3, "cost")
    D, H, W = c.shape
    out = np.empty((H, W, D), np.float32)
    _lib.check(_lib.lib().msn_swap_axes_host(c.ctypes.data, D, H, W, out.ctypes.data))
    return out


def swap_axes_back(cost):
    """featextract.cpp:78-105: [H,W,D] -> [D,H,W]."""
    c = _vol(cost, 3, "cost")
    H, W, D = c.shape
    out = np.empty((D, H, W), np.float32)
    _lib.check(_lib.lib().msn_swap_axes_back_host(c.ctypes.data, H, W, D, out.ctypes.data))
    return out


def get_right_cost(cost):
    """featextract.cpp:136-172: res[y,x,d] = c[y,x+d,d] for x < W-d, else c.flat[0]."""
    c = _vol(cost, 3, "cost")
    H, W, D = c.shape
    out = np.empty_like(c)
    _lib.check(_lib.lib().msn_right_cost_host(c.ctypes.data, H, W, D, out.ctypes.data))
    return out


def get_left_cost(cost):
    """featextract.cpp:464-499: res[y,x,d] = c[y,x-d,d] for x >= d, else c.flat[0]."""
    c = _vol(cost, 3, "cost")
    H, W, D = c.shape
    out = np.empty_like(c)
    _lib.check(_lib.lib().msn_left_cost_host(c.ctypes.data, H, W, D, out.ctypes.data))
    return out


def extract_likelihood(vol, *args):
    """featextract.cpp:415-462 (2-arg overload `extract_aml_testing`): AML per row of [n,D]."""
    if len(args) != 1:
        raise NotImplementedError("extract_likelihood(vol, r_samp, sigma) (featextract.cpp:359-412) is a CBMV "
                                  "training helper with no caller in MS-Nets; only (vol, sigma) is provided")
    c = _vol(vol, 2, "vol")
    n, D = c.shape
    out = np.empty_like(c)
    _lib.check(_lib.lib().msn_aml_host(c.ctypes.data, n, D, float(args[0]), out.ctypes.data))
    return out


def extract_ratio(vol, *args):
    """featextract.cpp:320-356 (2-arg overload `extract_pkrn_test`): (min+e)/(c+e) per row."""
    if len(args) != 1:
        raise NotImplementedError("extract_ratio(vol, r_samp, e) (featextract.cpp:272-317) is a CBMV training "
                                  "helper with no caller in MS-Nets; only (vol, e) is provided")
    c = _vol(vol, 2, "vol")
    n, D = c.shape
    out = np.empty_like(c)
    _lib.check(_lib.lib().msn_pkrn_host(c.ctypes.data, n, D, float(args[0]), out.ctypes.data))
    return out


def _dead(name, where):
    def fn(*a, **k):
        raise NotImplementedError("%s (%s) is never called by MS-Nets (SURVEY.md 8a, dead exports) and is "
                                  "outside the accelerated hot path" % (name, where))
    fn.__name__ = name
    return fn


get_cost = _dead("get_cost", "featextract.cpp:107-134")
generate_d_indices = _dead("generate_d_indices", "featextract.cpp:174-234")
get_samples = _dead("get_samples", "featextract.cpp:236-270")
generate_labels = _dead("generate_labels", "featextract.cpp:501-526")
