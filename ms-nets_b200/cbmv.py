"""Host-side mirror of the reference's MS feature generator
(src/dataloader/cbmv_generator.py), same names / arguments / return layouts:

    get_costs              cbmv_generator.py:27-79
    extract_features_left  cbmv_generator.py:258-308
    extract_features_lr    cbmv_generator.py:84-254
    get_default_args_dict  cbmv_generator.py:434-462
    generate_test_cbmv     cbmv_generator.py:727-861 (device-resident: returns a CUDA tensor)

plus the B200-native entry points that skip the reference's intermediate layouts:

    ms_features            NumPy in -> NumPy out, one fused device pass
    MSFeatureExtractor     CUDA tensors in -> CUDA tensor out ([N,C,D,h,w]), stream
                           ordered, buffers resident in HBM (the throughput path)

All arithmetic runs in the CUDA kernels behind the C ABI (include/msnets_b200.h).
"""
import ctypes

import numpy as np

from . import _lib
from . import libfeatextract as fte
from . import libmatchers as mtc


def get_default_args_dict():
    """cbmv_generator.py:434-462 (values the MS hyper-parameters are hard-coded to)."""
    return {
        "censw": 11, "nccw": 3, "sadw": 5, "sobelw": 5,
        "cens_sigma": 128.0, "ncc_sigma": 0.02, "sad_sigma": 20000.0, "sobel_sigma": 20000.0,
        "cbmv_F": 8, "w_padding": 1248, "h_padding": 384, "board_h": 12, "seed": 1234,
        "batch_h": 256, "overlap_board": 20, "batch_in_image": 0, "ds_scale": 2,
        "sf_frames_type": "frames_finalpass",
    }


def get_costs(iml, imr, maxdisp=192, censw=11, nccw=3, sadw=5, sobelw=5, board_h=10,
              board_w_left=10, board_w_right=0):
    """cbmv_generator.py:27-79 -> (census, ncc, sobel, sad), float32 [h,w,D] C-contiguous."""
    costcensus = mtc.census(iml, imr, maxdisp, censw)
    costncc = fte.swap_axes(mtc.nccNister(iml, imr, maxdisp, nccw))
    costsad = fte.swap_axes(mtc.zsad(iml, imr, maxdisp, sadw))
    costsob = fte.swap_axes(mtc.sadsob(mtc.sobel(iml), mtc.sobel(imr), maxdisp, sobelw))
    vld_w_end = -board_w_right if board_w_right > 0 else None
    vld_h_end = -board_h if board_h > 0 else None
    crop = lambda a: a[board_h:vld_h_end, board_w_left:vld_w_end, :].copy(order="C")
    return crop(costcensus), crop(costncc), crop(costsob), crop(costsad)


def _features_from_costs(census, ncc, sobel, sad, cens_sigma, ncc_sigma, sad_sigma, lr):
    vols = []
    for name, a in (("census", census), ("ncc", ncc), ("sobel", sobel), ("sad", sad)):
        if not isinstance(a, np.ndarray) or a.ndim != 3:
            raise ValueError("%s: expected a 3-D [h,w,D] numpy array" % name)
        vols.append(np.ascontiguousarray(a, dtype=np.float32))
    h, w, D = vols[0].shape
    if any(v.shape != (h, w, D) for v in vols):
        raise ValueError("cost volumes must share one [h,w,D] shape")
    out = np.empty((16 if lr else 8, D, h, w), np.float32)
    _lib.check(_lib.lib().msn_features_from_costs_host(
        vols[0].ctypes.data, vols[1].ctypes.data, vols[2].ctypes.data, vols[3].ctypes.data, h, w, D,
        float(cens_sigma), float(ncc_sigma), float(sad_sigma), 1 if lr else 0, out.ctypes.data))
    return out


def extract_features_left(census, ncc, sobel, sad, cens_sigma=128.0, ncc_sigma=0.02,
                          sad_sigma=20000.0, sobel_sigma=20000.0, disp_image=None):
    """cbmv_generator.py:258-308 -> float32 [8,D,h,w].  `sobel_sigma` is accepted and
    ignored exactly as in the reference (:298 passes sad_sigma for the sobel channel)."""
    return _features_from_costs(census, ncc, sobel, sad, cens_sigma, ncc_sigma, sad_sigma, False)


def extract_features_lr(census, ncc, sobel, sad, cens_sigma=128.0, ncc_sigma=0.02,
                        sad_sigma=20000.0, sobel_sigma=20000.0, disp_image=None):
    """cbmv_generator.py:84-254 -> float32 [16,D,h,w] (right-view channels 8-15)."""
    return _features_from_costs(census, ncc, sobel, sad, cens_sigma, ncc_sigma, sad_sigma, True)


def make_params(maxdisp=192, censw=11, nccw=3, sadw=5, sobelw=5, board_h=10, board_w_left=10,
                board_w_right=0, cens_sigma=128.0, ncc_sigma=0.02, sad_sigma=20000.0, left_only=True,
                d_begin=0, d_count=0):
    return _lib.default_params(ndisp=int(maxdisp), censw=int(censw), nccw=int(nccw), sadw=int(sadw),
                               sobelw=int(sobelw), board_h=int(board_h), board_w_left=int(board_w_left),
                               board_w_right=int(board_w_right), cens_sigma=float(cens_sigma),
                               ncc_sigma=float(ncc_sigma), sad_sigma=float(sad_sigma),
                               lr=0 if left_only else 1, d_begin=int(d_begin), d_count=int(d_count))


def output_shape(N, H, W, p):
    C = 16 if p.lr else 8
    Dn = p.d_count if p.d_count > 0 else p.ndisp
    return (N, C, Dn, H - 2 * p.board_h, W - p.board_w_left - p.board_w_right)


def ms_features(iml, imr, maxdisp=192, left_only=True, **kw):
    """get_costs + extract_features_* in one device pass (what generate_test_cbmv
    chains, cbmv_generator.py:826-843).  iml/imr: uint8 [H,W] or [N,H,W] (already
    bordered); returns float32 [C,D,h,w] or [N,C,D,h,w].  Keyword arguments are the
    get_costs ones (censw, nccw, sadw, sobelw, board_h, board_w_left, board_w_right)
    and the sigmas."""
    l = np.ascontiguousarray(iml)
    r = np.ascontiguousarray(imr)
    if l.dtype != np.uint8 or r.dtype != np.uint8 or l.shape != r.shape or l.ndim not in (2, 3):
        raise ValueError("ms_features: expected two uint8 arrays of equal shape [H,W] or [N,H,W]")
    single = l.ndim == 2
    if single:
        l, r = l[None], r[None]
    N, H, W = l.shape
    p = make_params(maxdisp, left_only=left_only, **kw)
    shape = output_shape(N, H, W, p)
    if min(shape) < 1:
        raise ValueError("ms_features: borders leave an empty image")
    out = np.empty(shape, np.float32)
    _lib.check(_lib.lib().msn_ms_features_host(l.ctypes.data, r.ctypes.data, N, H, W, ctypes.byref(p),
                                               out.ctypes.data))
    return out[0] if single else out


class MSFeatureExtractor(object):
    """Device-resident MS volume builder: uint8 CUDA tensors [N,H,W] in, float32 CUDA
    tensor [N,C,D,h,w] out, enqueued on the current torch stream.  Replaces the
    DataLoader-worker CPU extraction + `cost.cuda()` copy (main_msnet.py:375-377,
    571-572): the 3.2 GB/pair volume is produced where the 3D CNN consumes it."""

    def __init__(self, N, H, W, maxdisp=192, left_only=True, device=None, **kw):
        import torch
        if not torch.cuda.is_available():
            raise _lib.MsnetsError("MSFeatureExtractor needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.N, self.H, self.W = int(N), int(H), int(W)
        self.params = make_params(maxdisp, left_only=left_only, **kw)
        self.shape = output_shape(self.N, self.H, self.W, self.params)
        with torch.cuda.device(self.device):
            nbytes = _lib.lib().msn_ms_features_workspace_bytes(self.N, self.H, self.W, ctypes.byref(self.params))
            if nbytes == 0:
                raise _lib.MsnetsError(_lib.lib().msn_last_error().decode())
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def empty_output(self):
        return self.torch.empty(self.shape, dtype=self.torch.float32, device=self.device)

    def empty_wta(self):
        """(argmin int32, min1 float32, min2 float32), each [N,4,h,w]: buffers for the `wta=` by-product."""
        torch, shp = self.torch, (self.N, 4, self.shape[3], self.shape[4])
        return (torch.empty(shp, dtype=torch.int32, device=self.device),
                torch.empty(shp, dtype=torch.float32, device=self.device),
                torch.empty(shp, dtype=torch.float32, device=self.device))

    def __call__(self, left, right, out=None, wta=None):
        """wta: None, or the three tensors of empty_wta(): filled with, for each of channels 0-3, the
        winner-take-all disparity (np.argmin's rule: what main_msnet.py:443-448 computes from a host copy of
        the volume), the smallest and the second smallest channel value -- taken inside the kernel from the
        costs in shared memory, no pass over the volume."""
        torch = self.torch
        for t in (left, right):
            if t.dtype != torch.uint8 or tuple(t.shape) != (self.N, self.H, self.W) or not t.is_cuda \
                    or not t.is_contiguous():
                raise ValueError("expected contiguous uint8 CUDA tensors of shape %s" % ((self.N, self.H, self.W),))
        if out is None:
            out = self.empty_output()
        elif tuple(out.shape) != self.shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 tensor of shape %s" % (self.shape,))
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            if wta is None:
                _lib.check(_lib.lib().msn_ms_features_dev(left.data_ptr(), right.data_ptr(), self.N, self.H, self.W,
                                                          ctypes.byref(self.params), out.data_ptr(),
                                                          self.workspace.data_ptr(), self.workspace.numel(), stream))
            else:
                am, m1, m2 = wta
                shp = (self.N, 4, self.shape[3], self.shape[4])
                for t, dt in ((am, torch.int32), (m1, torch.float32), (m2, torch.float32)):
                    if tuple(t.shape) != shp or t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
                        raise ValueError("wta: expected the three contiguous CUDA tensors of empty_wta()")
                _lib.check(_lib.lib().msn_ms_features_wta_dev(
                    left.data_ptr(), right.data_ptr(), self.N, self.H, self.W, ctypes.byref(self.params),
                    out.data_ptr(), am.data_ptr(), m1.data_ptr(), m2.data_ptr(), self.workspace.data_ptr(),
                    self.workspace.numel(), stream))
        return out


_extractors = {}


def generate_test_cbmv(limg_name, rimg_name, crop_height=384, crop_width=1248, encoder_ds=64, maxdisp=192,
                       args_dict=None, is_left_only=True, device=None):
    """cbmv_generator.py:727-861, device-resident: same arguments and return
    tuple `(features, h, w, crop_height, crop_width)`, but `features` is a float32 CUDA tensor
    [C, D, crop_height, crop_width] produced where the 3D CNN consumes it -- the caller's
    `features.cuda()` (main_msnet.py:571-572) becomes a no-op.

    limg_name / rimg_name: file names (read with cv2.imread(name, 0) as the reference does) or
    uint8 gray images.  As in the reference, crop_height / crop_width are recomputed from the
    image size: zero padding on top and on the right up to a multiple of encoder_ds (:780-788),
    then a 10-pixel zero border on all sides so the matchers' border fill lands outside the crop
    (:819-834).  ds_scale (args_dict) must be 1: the reference's default 2 goes through
    skimage.transform.rescale, which is not ported yet."""
    import torch
    ad = get_default_args_dict()
    ad["ds_scale"] = 1
    if args_dict is not None:
        ad.update(args_dict)
    if int(ad["ds_scale"]) != 1:
        raise NotImplementedError("generate_test_cbmv: ds_scale=%r needs the anti-aliased rescale "
                                  "(skimage.transform.rescale), not ported yet" % (ad["ds_scale"],))
    imgs = []
    for src in (limg_name, rimg_name):
        if isinstance(src, str):
            import cv2
            im = cv2.imread(src, 0)
            if im is None:
                raise ValueError("generate_test_cbmv: cannot read %r" % (src,))
            src = im
        a = np.ascontiguousarray(src)
        if a.dtype != np.uint8 or a.ndim != 2:
            raise ValueError("generate_test_cbmv: expected uint8 gray images [H,W]")
        imgs.append(a)
    if imgs[0].shape != imgs[1].shape:
        raise ValueError("generate_test_cbmv: left and right image sizes differ")
    h, w = imgs[0].shape
    crop_width = w + (encoder_ds - w % encoder_ds) % encoder_ds
    crop_height = h + (encoder_ds - h % encoder_ds) % encoder_ds
    if not torch.cuda.is_available():
        raise _lib.MsnetsError("generate_test_cbmv needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    board = 10
    Hb, Wb = crop_height + 2 * board, crop_width + 2 * board
    pair = torch.zeros((2, 1, Hb, Wb), dtype=torch.uint8, device=dev)
    for i, a in enumerate(imgs):   # top/right padding and the border are the zeros already there
        pair[i, 0, board + crop_height - h:board + crop_height, board:board + w] = torch.from_numpy(a).to(dev)
    key = (dev.index, Hb, Wb, int(maxdisp), bool(is_left_only), ad["censw"], ad["nccw"], ad["sadw"], ad["sobelw"],
           float(ad["cens_sigma"]), float(ad["ncc_sigma"]), float(ad["sad_sigma"]))
    ex = _extractors.get(key)
    if ex is None:
        _extractors.clear()        # one cached workspace: test images usually share a size
        ex = MSFeatureExtractor(1, Hb, Wb, maxdisp=int(maxdisp) // int(ad["ds_scale"]), left_only=is_left_only,
                                device=dev, censw=ad["censw"], nccw=ad["nccw"], sadw=ad["sadw"],
                                sobelw=ad["sobelw"], board_h=board, board_w_left=board, board_w_right=board,
                                cens_sigma=ad["cens_sigma"], ncc_sigma=ad["ncc_sigma"], sad_sigma=ad["sad_sigma"])
        _extractors[key] = ex
    with torch.cuda.device(dev):
        features = ex(pair[0], pair[1])[0]
    return features, h, w, crop_height, crop_width
