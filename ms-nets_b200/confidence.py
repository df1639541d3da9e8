"""Confidence pass over a cost volume: winner-take-all, second minimum, peak-ratio and
the left-right consistency mask.

The only reference consumer is the TensorBoard WTA picture
(`np.argmin(dsi[:,k],axis=1)`, main_msnet.py:444-448): first minimal index wins, an
all-fill pixel yields 0.  Second-min / peak-ratio / LR-check have no reference code
(SURVEY.md 0.3); their definitions are stated in oracle/ms_oracle.py and mirrored here.
NumPy arrays go through the host C ABI, CUDA tensors through the device one.
"""
import numpy as np

from . import _lib


def wta(cost, layout="hwd"):
    """cost: [...,D] ("hwd": D innermost, the reference's [H,W,D]) or [D,...] ("dhw": the
    feature-plane layout).  Returns (argmin int32, min float32, second_min float32)."""
    if layout not in ("hwd", "dhw"):
        raise ValueError("layout must be 'hwd' or 'dhw'")
    lay = 0 if layout == "hwd" else 1
    if isinstance(cost, np.ndarray):
        c = np.ascontiguousarray(cost, dtype=np.float32)
        D = c.shape[-1] if lay == 0 else c.shape[0]
        shp = c.shape[:-1] if lay == 0 else c.shape[1:]
        n = int(np.prod(shp))
        am = np.empty(shp, np.int32)
        m1 = np.empty(shp, np.float32)
        m2 = np.empty(shp, np.float32)
        _lib.check(_lib.lib().msn_wta_host(c.ctypes.data, n, D, lay, am.ctypes.data, m1.ctypes.data,
                                           m2.ctypes.data))
        return am, m1, m2
    import torch
    if not cost.is_cuda or cost.dtype != torch.float32:
        raise _lib.MsnetsError("wta: expected a float32 CUDA tensor (no CPU fallback)")
    c = cost.contiguous()
    D = c.shape[-1] if lay == 0 else c.shape[0]
    shp = tuple(c.shape[:-1]) if lay == 0 else tuple(c.shape[1:])
    n = int(np.prod(shp))
    am = torch.empty(shp, dtype=torch.int32, device=c.device)
    m1 = torch.empty(shp, dtype=torch.float32, device=c.device)
    m2 = torch.empty(shp, dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _lib.check(_lib.lib().msn_wta_dev(c.data_ptr(), n, D, lay, am.data_ptr(), m1.data_ptr(), m2.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    return am, m1, m2


def pkrn_confidence(min1, min2, e=0.01):
    """(min1+e)/(min2+e), 0 where min1 is fill; CUDA tensors in, CUDA tensor out."""
    import torch
    if not (min1.is_cuda and min2.is_cuda):
        raise _lib.MsnetsError("pkrn_confidence: expected CUDA tensors (no CPU fallback)")
    a, b = min1.contiguous(), min2.contiguous()
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().msn_pkrn_conf_dev(a.data_ptr(), b.data_ptr(), a.numel(), float(e), out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    return out


def lr_consistency(cost_hwd, thresh=1):
    """cost [H,W,D] -> (dL int32 [H,W], dR int32 [H,W], mask uint8 [H,W]);
    mask = x-dL >= 0 and |dL(x) - dR(x-dL)| <= thresh, dR from the right-view volume
    get_right_cost would build (featextract.cpp:136-172), never materialised here."""
    if isinstance(cost_hwd, np.ndarray):
        c = np.ascontiguousarray(cost_hwd, dtype=np.float32)
        H, W, D = c.shape
        dl = np.empty((H, W), np.int32)
        dr = np.empty((H, W), np.int32)
        mask = np.empty((H, W), np.uint8)
        _lib.check(_lib.lib().msn_lrc_host(c.ctypes.data, H, W, D, int(thresh), dl.ctypes.data, dr.ctypes.data,
                                           mask.ctypes.data))
        return dl, dr, mask
    import torch
    if not cost_hwd.is_cuda or cost_hwd.dtype != torch.float32:
        raise _lib.MsnetsError("lr_consistency: expected a float32 CUDA tensor (no CPU fallback)")
    c = cost_hwd.contiguous()
    H, W, D = c.shape
    dl = torch.empty((H, W), dtype=torch.int32, device=c.device)
    dr = torch.empty((H, W), dtype=torch.int32, device=c.device)
    mask = torch.empty((H, W), dtype=torch.uint8, device=c.device)
    with torch.cuda.device(c.device):
        _lib.check(_lib.lib().msn_lrc_dev(c.data_ptr(), H, W, D, int(thresh), dl.data_ptr(), dr.data_ptr(),
                                          mask.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return dl, dr, mask
