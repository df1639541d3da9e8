"""4D cost-volume builders for learned unary features (GC-Net / PSMNet style).

The reference has no such builder: its PSMNet head expects the
classic 2x32-channel concat volume (`dres0`, psmnet_3dcnn.py:96) but nothing
produces it.  Definitions (oracle/ms_oracle.py:concat_volume / diff_volume):

    concat[n, :C, d, y, x] = fl[n, :, y, x]      for x >= d, else 0
    concat[n, C:, d, y, x] = fr[n, :, y, x - d]  for x >= d, else 0
    diff  [n, c, d, y, x]  = fl[n,c,y,x] - fr[n,c,y,x-d]  for x >= d, else 0
"""
from . import _lib


def _build(fl, fr, ndisp, diff, out):
    import torch
    if not (fl.is_cuda and fr.is_cuda):
        raise _lib.MsnetsError("volume builders need CUDA tensors (no CPU fallback)")
    if fl.dtype != torch.float32 or fr.dtype != torch.float32 or fl.shape != fr.shape or fl.dim() != 4:
        raise ValueError("expected two float32 [N,C,H,W] tensors of equal shape")
    a, b = fl.contiguous(), fr.contiguous()
    N, C, H, W = a.shape
    shape = (N, C if diff else 2 * C, int(ndisp), H, W)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=a.device)
    elif tuple(out.shape) != shape or not out.is_contiguous():
        raise ValueError("out must be contiguous with shape %s" % (shape,))
    fn = _lib.lib().msn_diff_volume_dev if diff else _lib.lib().msn_concat_volume_dev
    with torch.cuda.device(a.device):
        _lib.check(fn(a.data_ptr(), b.data_ptr(), N, C, H, W, int(ndisp), out.data_ptr(),
                      torch.cuda.current_stream().cuda_stream))
    return out


def concat_volume(fl, fr, ndisp, out=None):
    return _build(fl, fr, ndisp, False, out)


def diff_volume(fl, fr, ndisp, out=None):
    return _build(fl, fr, ndisp, True, out)
