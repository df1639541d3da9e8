"""Drop-in for the reference's `src.cpp.lib.libmatchers` (matchers.cpp:565-589).

Same names, argument order, dtypes and OUTPUT LAYOUTS as the Boost.Python module:
census -> float32 [H,W,D]; nccNister / zsad / sadsob -> float32 [D,H,W];
sobel -> float32 [H,W].  Inputs are borrowed, a new ndarray is returned.
The reference reinterpret_casts whatever it is handed; this module
checks dtype/shape/contiguity and raises ValueError instead of misreading memory.
Everything runs in hand-written CUDA kernels through the C ABI; no CPU path.
"""
import numpy as np

from . import _lib


def _img(a, dtype, name):
    if not isinstance(a, np.ndarray):
        raise ValueError("%s: expected a numpy array" % name)
    if a.dtype != dtype or a.ndim != 2:
        raise ValueError("%s: expected a 2-D %s array, got %s %s" % (name, np.dtype(dtype), a.dtype, a.shape))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("%s: array must be C-contiguous" % name)
    return a


def _pair(left, right, dtype):
    l, r = _img(left, dtype, "left"), _img(right, dtype, "right")
    if l.shape != r.shape:
        raise ValueError("left/right shapes differ: %s vs %s" % (l.shape, r.shape))
    return l, r


def census(left, right, ndisp, wsize):
    """matchers.cpp:232-353: Hamming distance of wsize x wsize census codes."""
    l, r = _pair(left, right, np.uint8)
    H, W = l.shape
    out = np.empty((H, W, int(ndisp)), np.float32)
    _lib.check(_lib.lib().msn_census_host(l.ctypes.data, r.ctypes.data, H, W, int(ndisp), int(wsize),
                                          out.ctypes.data))
    return out


def nccNister(left, right, ndisp, wsize):
    """matchers.cpp:47-228: negated window NCC, fp64 inside, +1 on flat windows."""
    l, r = _pair(left, right, np.uint8)
    H, W = l.shape
    out = np.empty((int(ndisp), H, W), np.float32)
    _lib.check(_lib.lib().msn_ncc_host(l.ctypes.data, r.ctypes.data, H, W, int(ndisp), int(wsize),
                                       out.ctypes.data))
    return out


def zsad(left, right, ndisp, wsize):
    """matchers.cpp:442-512: zero-mean SAD, ordered fp32."""
    l, r = _pair(left, right, np.uint8)
    H, W = l.shape
    out = np.empty((int(ndisp), H, W), np.float32)
    _lib.check(_lib.lib().msn_zsad_host(l.ctypes.data, r.ctypes.data, H, W, int(ndisp), int(wsize),
                                        out.ctypes.data))
    return out


def sobel(img):
    """matchers.cpp:515-554: horizontal 3x3 Sobel."""
    a = _img(img, np.uint8, "img")
    H, W = a.shape
    out = np.empty((H, W), np.float32)
    _lib.check(_lib.lib().msn_sobel_host(a.ctypes.data, H, W, out.ctypes.data))
    return out


def sadsob(left, right, ndisp, wsize):
    """matchers.cpp:356-438: SAD of two float32 (Sobel) images via an fp32 summed-area table."""
    l, r = _pair(left, right, np.float32)
    H, W = l.shape
    out = np.empty((int(ndisp), H, W), np.float32)
    _lib.check(_lib.lib().msn_sadsob_host(l.ctypes.data, r.ctypes.data, H, W, int(ndisp), int(wsize),
                                          out.ctypes.data))
    return out


def initthreads():
    """matchers.cpp:556-563 returns the OpenMP team size (8); here: the SM count."""
    import ctypes
    n = ctypes.c_int(0)
    _lib.check(_lib.lib().msn_initthreads(ctypes.byref(n)))
    return n.value
