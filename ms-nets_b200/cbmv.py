"""Host-side mirror of the reference's MS feature generator
(src/dataloader/cbmv_generator.py), same names / arguments / return layouts:

    get_costs              cbmv_generator.py:27-79
    extract_features_left  cbmv_generator.py:258-308
    extract_features_lr    cbmv_generator.py:84-254
    get_default_args_dict  cbmv_generator.py:434-462
    generate_test_cbmv     cbmv_generator.py:727-861 (device-resident: returns a CUDA tensor)
    generate_crop_train_cbmv  cbmv_generator.py:549-725 (device-resident features)
    down_sampling_input    cbmv_generator.py:465-482 (anti-aliased rescale on the device)
    get_crop_position      cbmv_generator.py:398-432

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
                d_begin=0, d_count=0, row_begin=0, row_count=0):
    return _lib.default_params(ndisp=int(maxdisp), censw=int(censw), nccw=int(nccw), sadw=int(sadw),
                               sobelw=int(sobelw), board_h=int(board_h), board_w_left=int(board_w_left),
                               board_w_right=int(board_w_right), cens_sigma=float(cens_sigma),
                               ncc_sigma=float(ncc_sigma), sad_sigma=float(sad_sigma),
                               lr=0 if left_only else 1, d_begin=int(d_begin), d_count=int(d_count),
                               row_begin=int(row_begin), row_count=int(row_count))


def output_shape(N, H, W, p):
    C = 16 if p.lr else 8
    Dn = p.d_count if p.d_count > 0 else p.ndisp
    h = p.row_count if p.row_count > 0 else H - 2 * p.board_h
    return (N, C, Dn, h, W - p.board_w_left - p.board_w_right)


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

    def __init__(self, N, H, W, maxdisp=192, left_only=True, device=None, out_dtype=None, **kw):
        """out_dtype: torch.float32 (default) or torch.bfloat16 -- the volume rounded to bf16 inside the kernel
        (msn_ms_features_bf16_dev), half the bytes for a consumer under bf16 autocast."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.MsnetsError("MSFeatureExtractor needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.N, self.H, self.W = int(N), int(H), int(W)
        self.params = make_params(maxdisp, left_only=left_only, **kw)
        self.shape = output_shape(self.N, self.H, self.W, self.params)
        self.out_dtype = torch.float32 if out_dtype is None else out_dtype
        if self.out_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
        with torch.cuda.device(self.device):
            nbytes = _lib.lib().msn_ms_features_workspace_bytes(self.N, self.H, self.W, ctypes.byref(self.params))
            if nbytes == 0:
                raise _lib.MsnetsError(_lib.lib().msn_last_error().decode())
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def empty_output(self):
        return self.torch.empty(self.shape, dtype=self.out_dtype, device=self.device)

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
        elif tuple(out.shape) != self.shape or out.dtype != self.out_dtype or not out.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor of shape %s" % (self.out_dtype, self.shape))
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            if self.out_dtype == torch.bfloat16:
                if wta is not None:
                    raise ValueError("wta= is not combined with the bf16 volume")
                _lib.check(_lib.lib().msn_ms_features_bf16_dev(left.data_ptr(), right.data_ptr(), self.N, self.H,
                                                               self.W, ctypes.byref(self.params), out.data_ptr(),
                                                               self.workspace.data_ptr(), self.workspace.numel(),
                                                               stream))
            elif wta is None:
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


def _cached_extractor(dev, Hb, Wb, maxdisp, is_left_only, ad, board_h, board_w_left, board_w_right):
    key = (dev.index, Hb, Wb, int(maxdisp), bool(is_left_only), ad["censw"], ad["nccw"], ad["sadw"], ad["sobelw"],
           float(ad["cens_sigma"]), float(ad["ncc_sigma"]), float(ad["sad_sigma"]), board_h, board_w_left, board_w_right)
    ex = _extractors.get(key)
    if ex is None:
        _extractors.clear()        # one cached workspace: the images of a run usually share a size
        ex = MSFeatureExtractor(1, Hb, Wb, maxdisp=int(maxdisp), left_only=is_left_only, device=dev,
                                censw=ad["censw"], nccw=ad["nccw"], sadw=ad["sadw"], sobelw=ad["sobelw"],
                                board_h=board_h, board_w_left=board_w_left, board_w_right=board_w_right,
                                cens_sigma=ad["cens_sigma"], ncc_sigma=ad["ncc_sigma"], sad_sigma=ad["sad_sigma"])
        _extractors[key] = ex
    return ex


# ------------------------------------------------------- pre-matching image ops --
def _rescale_plan(H, W, scale):
    """What skimage.transform.rescale (>= 0.19: _warps.py rescale -> resize) derives from (shape, scale) for a
    2-D image with anti_aliasing=True, order 1: output shape = round(scale * shape), per-axis zoom factor =
    input / output extent, Gaussian sigma = max(0, (factor - 1) / 2), and scipy.ndimage's kernel for it
    (_gaussian_kernel1d: radius int(4 sigma + 0.5), exp(-0.5 / sigma^2 * x^2) normalised) -- evaluated with
    NumPy exactly as scipy evaluates it, because the device replay needs the same fp64 weights."""
    oh, ow = (int(v) for v in np.round(np.asarray((H, W)) * scale))
    if oh < 1 or ow < 1:
        raise ValueError("rescale: scale %r leaves an empty image" % (scale,))
    plan = []
    for n_in, n_out in ((H, oh), (W, ow)):
        factor = n_in / n_out
        sigma = max(0.0, (factor - 1) / 2)
        if sigma > 0:
            r = int(4.0 * float(sigma) + 0.5)
            x = np.arange(-r, r + 1)
            w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
            w = np.ascontiguousarray(w / w.sum(), dtype=np.float64)
        else:
            r, w = -1, None
        plan.append((r, w, float(factor)))
    return oh, ow, plan


def rescale_u8(img, scale):
    """One half of down_sampling_input (cbmv_generator.py:465-482) for uint8 gray images [H,W] or [N,H,W]:
    (img / 255 -> skimage.transform.rescale(scale, anti_aliasing=True, mode='constant') -> * 255).astype(uint8),
    computed on the device (msn_rescale_*).  NumPy array in -> NumPy array out; CUDA tensor in -> CUDA tensor
    out (stream ordered)."""
    is_np = isinstance(img, np.ndarray)
    a = np.ascontiguousarray(img) if is_np else img.contiguous()
    if (a.dtype != np.uint8) if is_np else (str(a.dtype) != "torch.uint8"):
        raise ValueError("rescale_u8: expected uint8 images")
    if a.ndim not in (2, 3):
        raise ValueError("rescale_u8: expected [H,W] or [N,H,W]")
    single = a.ndim == 2
    N = 1 if single else a.shape[0]
    H, W = a.shape[-2], a.shape[-1]
    oh, ow, ((rr, wr, zr), (rc, wc, zc)) = _rescale_plan(H, W, scale)
    pr = wr.ctypes.data if wr is not None else None
    pc = wc.ctypes.data if wc is not None else None
    L = _lib.lib()
    if is_np:
        out = np.empty((oh, ow) if single else (N, oh, ow), np.uint8)
        _lib.check(L.msn_rescale_host(a.ctypes.data, N, H, W, oh, ow, pr, rr, pc, rc, zr, zc, out.ctypes.data))
        return out
    import torch
    if not a.is_cuda:
        raise _lib.MsnetsError("rescale_u8: tensor must live on a CUDA device (no CPU fallback)")
    out = torch.empty((oh, ow) if single else (N, oh, ow), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        ws = torch.empty(L.msn_rescale_workspace_bytes(N, H, W), dtype=torch.uint8, device=a.device)
        _lib.check(L.msn_rescale_dev(a.data_ptr(), N, H, W, oh, ow, pr, rr, pc, rc, zr, zc, out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    return out


def down_sampling_input(ds_scale, imgl, imgr, anti_aliasing=True, multichannel=False, preserve_range=True):
    """cbmv_generator.py:465-482, same arguments: -> (imgl, imgr) uint8, rescaled by `ds_scale` (e.g. 0.5)."""
    if not anti_aliasing or multichannel or not preserve_range:
        raise NotImplementedError("down_sampling_input: only the reference's own call form is provided "
                                  "(anti_aliasing=True, multichannel=False, preserve_range=True)")
    return rescale_u8(imgl, ds_scale), rescale_u8(imgr, ds_scale)


def get_crop_position(w, h, crop_width=512, crop_height=256, board_w_left=256, board_w_right=0, board_h=10,
                      is_fixed_center_around_crop=False):
    """cbmv_generator.py:398-432: random (or centred) crop origin, the side borders halved once when the image is
    too narrow; draws from Python's `random` exactly as the reference does (same stream for the same seed)."""
    import random
    tmp_diff = w - crop_width - board_w_left - board_w_right
    if tmp_diff >= 0:
        new_left, new_right = board_w_left, board_w_right
    else:
        while tmp_diff < 0:
            new_left, new_right = board_w_left // 2, board_w_right // 2
            tmp_diff = w - crop_width - new_left - new_right
            if new_left == board_w_left // 2 and tmp_diff < 0:   # the reference would spin forever here
                raise ValueError("get_crop_position: image width %d too small for a %d px crop" % (w, crop_width))
    start_w = random.randint(0, w - crop_width - new_left - new_right)
    start_h = random.randint(0, h - crop_height - 2 * board_h)
    if is_fixed_center_around_crop:
        start_w = (w - crop_width - new_left - new_right) // 2 - 1
        start_h = (h - crop_height - 2 * board_h) // 2 - 1
    finish_h = start_h + (crop_height + 2 * board_h)
    finish_w = start_w + (crop_width + new_left + new_right)
    return start_w, start_h, finish_w, finish_h, new_left, new_right


def read_pfm(name):
    """PFM reader (the reference's src/pfmutil.py readPFM): float32 [H,W], rows flipped to top-down."""
    import re
    with open(name, "rb") as f:
        kind = f.readline().decode("latin-1")
        if "PF" in kind:
            channels = 3
        elif "Pf" in kind:
            channels = 1
        else:
            raise ValueError("%s: not a PFM file" % name)
        width, height = (int(v) for v in re.findall(r"\d+", f.readline().decode("latin-1"))[:2])
        big = "-" not in f.readline().decode("latin-1")
        data = np.frombuffer(f.read(width * height * channels * 4), dtype=(">f4" if big else "<f4"))
    return np.flipud(data.reshape(height, width)).astype(np.float32)


def _read_gray(src, who):
    if isinstance(src, str):
        import cv2
        im = cv2.imread(src, 0)
        if im is None:
            raise ValueError("%s: cannot read %r" % (who, src))
        src = im
    a = np.ascontiguousarray(src)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("%s: expected uint8 gray images [H,W]" % who)
    return a


def generate_test_cbmv(limg_name, rimg_name, crop_height=384, crop_width=1248, encoder_ds=64, maxdisp=192,
                       args_dict=None, is_left_only=True, device=None):
    """cbmv_generator.py:727-861, device-resident: same arguments, same defaults
    (args_dict=None means get_default_args_dict(), ds_scale = 2) and the same return tuple
    `(features, h, w, crop_height, crop_width)`, but `features` is a float32 CUDA tensor
    [C, maxdisp/ds, crop_height/ds, crop_width/ds] produced where the 3D CNN consumes it -- the caller's
    `features.cuda()` (main_msnet.py:571-572) becomes a no-op.

    limg_name / rimg_name: file names (read with cv2.imread(name, 0) as the reference does) or
    uint8 gray images.  As in the reference, crop_height / crop_width are recomputed from the
    image size: zero padding on top and on the right up to a multiple of encoder_ds (:780-788),
    the anti-aliased down-sampling by ds_scale (:802-803, rescale_u8 on the device), then a 10-pixel
    zero border on all sides so the matchers' border fill lands outside the crop (:819-834)."""
    import torch
    ad = get_default_args_dict() if args_dict is None else dict(get_default_args_dict(), **args_dict)
    ds = int(ad["ds_scale"])
    if ds < 1:
        raise ValueError("generate_test_cbmv: ds_scale must be >= 1")
    imgs = [_read_gray(src, "generate_test_cbmv") for src in (limg_name, rimg_name)]
    if imgs[0].shape != imgs[1].shape:
        raise ValueError("generate_test_cbmv: left and right image sizes differ")
    h, w = imgs[0].shape
    crop_width = w + (encoder_ds - w % encoder_ds) % encoder_ds
    crop_height = h + (encoder_ds - h % encoder_ds) % encoder_ds
    if not torch.cuda.is_available():
        raise _lib.MsnetsError("generate_test_cbmv needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    board = 10
    with torch.cuda.device(dev):
        padded = torch.zeros((2, crop_height, crop_width), dtype=torch.uint8, device=dev)
        for i, a in enumerate(imgs):   # top/right padding = the zeros already there
            padded[i, crop_height - h:, :w] = torch.from_numpy(a).to(dev)
        if ds > 1:
            padded = rescale_u8(padded, 1.0 / float(ds))
        ph, pw = padded.shape[1], padded.shape[2]
        Hb, Wb = ph + 2 * board, pw + 2 * board
        pair = torch.zeros((2, 1, Hb, Wb), dtype=torch.uint8, device=dev)
        pair[:, 0, board:board + ph, board:board + pw] = padded
        ex = _cached_extractor(dev, Hb, Wb, int(maxdisp) // ds, is_left_only, ad, board, board, board)
        features = ex(pair[0], pair[1])[0]
    return features, h, w, crop_height, crop_width


def generate_crop_train_cbmv(limg_name, rimg_name, disp_name, lseg_name, crop_height=256, crop_width=512,
                             maxdisp=192, is_fixed_center_around_crop=False, args_dict=None, is_left_only=True,
                             device=None):
    """cbmv_generator.py:549-725, device-resident: same arguments and return tuple
    `(features, disp_image, imgl_rgb, imgr_rgb, semantic_label)`; `features` is a float32 CUDA tensor
    [C, maxdisp/ds, crop_height/ds, crop_width/ds], the other four are the reference's CPU float tensors.

    The random crop (get_crop_position, `random` module), the border policy (board_h rows above and below,
    maxdisp columns on the left -- and on the right for the two-view volume -- halved when the image is too
    narrow), the anti-aliased down-sampling of the cropped pair and the integer division of maxdisp and of the
    borders by ds_scale follow the reference line by line; the pair is cropped on the host (two small uint8
    arrays) and everything after it runs on the device.  This is the path that starves the reference's trainer
    (do_main_msnet.sh:138-192): here a 256x512 crop costs a fraction of a millisecond."""
    import cv2
    import torch
    ad = get_default_args_dict() if args_dict is None else dict(get_default_args_dict(), **args_dict)
    ds = int(ad["ds_scale"])
    board_h = int(ad["board_h"])
    board_w_left = int(maxdisp)
    board_w_right = 0 if is_left_only else int(maxdisp)
    imgl = _read_gray(limg_name, "generate_crop_train_cbmv")
    imgr = _read_gray(rimg_name, "generate_crop_train_cbmv")
    imgl_rgb = cv2.imread(limg_name, 1).astype(np.uint8)[:, :, ::-1]
    imgr_rgb = cv2.imread(rimg_name, 1).astype(np.uint8)[:, :, ::-1]
    h, w = imgl.shape[:2]
    start_w, start_h, finish_w, finish_h, board_w_left, board_w_right = get_crop_position(
        w, h, crop_width, crop_height, board_w_left, board_w_right, board_h, is_fixed_center_around_crop)
    vld_w_end = -board_w_right if board_w_right > 0 else None
    vld_h_end = -board_h if board_h > 0 else None

    def inner(a):                                                       # remove_border, :484-503
        a = a[start_h:finish_h, start_w:finish_w]
        return np.ascontiguousarray(a[board_h:vld_h_end][:, board_w_left:vld_w_end])
    disp_image = read_pfm(disp_name)[start_h:finish_h, start_w:finish_w].copy()
    disp_image[disp_image == np.inf] = .0
    disp_image = np.ascontiguousarray(disp_image[board_h:vld_h_end][:, board_w_left:vld_w_end])
    imgl_rgb, imgr_rgb = inner(imgl_rgb), inner(imgr_rgb)
    if lseg_name is not None:
        from PIL import Image
        semantic_label = inner(np.asarray(Image.open(lseg_name), dtype=np.float32, order="C"))
    else:
        semantic_label = None
    cl = np.ascontiguousarray(imgl[start_h:finish_h, start_w:finish_w])
    cr = np.ascontiguousarray(imgr[start_h:finish_h, start_w:finish_w])
    if not torch.cuda.is_available():
        raise _lib.MsnetsError("generate_crop_train_cbmv needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        pair = torch.from_numpy(np.stack([cl, cr])).to(dev)
        if ds > 1:
            pair = rescale_u8(pair, 1.0 / float(ds))
        Hb, Wb = pair.shape[1], pair.shape[2]
        ex = _cached_extractor(dev, Hb, Wb, int(maxdisp) // ds, is_left_only, ad, board_h // ds, board_w_left // ds,
                               board_w_right // ds)
        features = ex(pair[0:1].contiguous(), pair[1:2].contiguous())[0]
    imgl_rgb = torch.from_numpy(imgl_rgb.transpose((2, 0, 1)).astype(np.float32) / 255.0).float()
    imgr_rgb = torch.from_numpy(imgr_rgb.transpose((2, 0, 1)).astype(np.float32) / 255.0).float()
    disp_image = torch.from_numpy(disp_image).float()
    if semantic_label is not None:
        semantic_label = torch.from_numpy(semantic_label[None, ...]).float()
    else:
        semantic_label = torch.zeros([1, disp_image.size()[0], disp_image.size(1)], dtype=torch.float32)
    return features, disp_image, imgl_rgb, imgr_rgb, semantic_label
