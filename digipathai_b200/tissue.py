"""Tissue mask and tile grid: the prologue of the hot path (one-off per slide).

Two evaluations of the same arithmetic: the host one (numpy + OpenCV, this module's original) and -- when a CUDA
device is named -- the device one (``tissue_mask_device`` / ``morpho_process_device``: histogram, mask and
rectangular min/max kernels of csrc/tissue.cuh through the C ABI), bit-identical to each other
(tests/test_gpu_tissue.py) and, for the morphology, to the reference's own cv2 calls (tests/golden/morph_golden.npz).
On a sharded run every rank needs the grid before it can start; on the device it costs a few milliseconds instead of
~0.2 s of host time per rank.

Mirrors ``TissueMaskGenerationOS`` (DigiPathAI/helpers/utils.py:336-354), ``BinMorphoProcessMaskOS``
(utils.py:200-219) and ``WSIStridedPatchDataset._preprocess`` / ``__getitem__`` coordinate logic
(DigiPathAI/loaders/dataloader.py:238-312,340-353).  scikit-image is not in this image, so the two skimage
calls the reference makes are restated from skimage's published algorithm: ``rgb2hsv`` (saturation =
(max-min)/max on the [0,1]-scaled image, 0 where max == min) and ``threshold_otsu`` (256-bin histogram over the
value range -- one bin per integer for integer images --, threshold = centre of the bin maximising the
between-class variance, mask = ``> threshold``).  Morphology goes through OpenCV exactly as in the reference.
"""
from __future__ import annotations

import numpy as np


def _hist_u8(*planes: np.ndarray) -> np.ndarray:
    """Exact histogram of one uint8 plane (256 bins) or joint histogram of two (256 x 256, first plane major).
    OpenCV's calcHist counts in float32, exact below 2**24 per bin: rows are fed in chunks of at most 2**24 pixels
    and the chunk histograms summed in int64."""
    import cv2
    rows = max(1, (1 << 24) // max(1, planes[0].shape[1]))
    n = len(planes)
    total = np.zeros((256,) * n, np.int64)
    for r0 in range(0, planes[0].shape[0], rows):
        h = cv2.calcHist([p[r0:r0 + rows] for p in planes], list(range(n)), None, [256] * n, [0, 256] * n)
        total += h.reshape(total.shape).astype(np.int64)
    return total.ravel()


def threshold_otsu(image: np.ndarray, nbins: int = 256) -> float:
    a = np.asarray(image)
    if a.dtype == np.uint8:                      # the path's case: one bincount pass, no int64 copy of the image
        full = _hist_u8(np.ascontiguousarray(a)) if a.ndim == 2 else np.bincount(a.ravel(), minlength=256)
        present = np.flatnonzero(full)
        lo, hi = int(present[0]), int(present[-1])
        if lo == hi:
            raise ValueError("threshold_otsu is expected to work with images having more than one color")
        return _otsu_from_hist(full[lo:hi + 1], np.arange(lo, hi + 1, dtype=np.float64))
    if a.min() == a.max():
        raise ValueError("threshold_otsu is expected to work with images having more than one color")
    flat = a.ravel()
    if np.issubdtype(a.dtype, np.integer):
        lo = int(flat.min())
        hist = np.bincount(flat.astype(np.int64) - lo).astype(np.float64)
        centers = np.arange(lo, lo + len(hist), dtype=np.float64)
    else:
        hist, edges = np.histogram(flat, bins=nbins)
        hist = hist.astype(np.float64)
        centers = (edges[:-1] + edges[1:]) / 2.0
    w1 = np.cumsum(hist)
    w2 = np.cumsum(hist[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        m1 = np.cumsum(hist * centers) / w1
        m2 = (np.cumsum((hist * centers)[::-1]) / w2[::-1])[::-1]
    var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return float(centers[:-1][int(np.argmax(var12))])


def saturation(img_rgb_u8: np.ndarray) -> np.ndarray:
    """HSV saturation channel of ``skimage.color.rgb2hsv`` for a uint8 RGB image (float64 in [0, 1])."""
    arr = img_rgb_u8.astype(np.float64) / 255.0
    v = arr.max(-1)
    delta = arr.max(-1) - arr.min(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = delta / v
    s[delta == 0.0] = 0.0
    s[np.isnan(s)] = 0.0
    return s


def _otsu_from_hist(hist: np.ndarray, centers: np.ndarray) -> float:
    hist = hist.astype(np.float64)
    w1 = np.cumsum(hist)
    w2 = np.cumsum(hist[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        m1 = np.cumsum(hist * centers) / w1
        m2 = (np.cumsum((hist * centers)[::-1]) / w2[::-1])[::-1]
    var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return float(centers[:-1][int(np.argmax(var12))])


def _saturation_lut():
    """float64 saturation of every (max, max - min) pair of a uint8 RGB pixel: ``(max/255 - min/255) / (max/255)``,
    the literal expression of skimage's rgb2hsv on the [0, 1]-scaled image (utils.py:339), 0 where max == min."""
    vv, dd = np.divmod(np.arange(65536), 256)
    mn = vv - dd                                  # negative for impossible pairs (never present in a histogram)
    with np.errstate(divide="ignore", invalid="ignore"):
        lut = (vv / 255.0 - mn / 255.0) / (vv / 255.0)
    lut[dd == 0] = 0.0
    lut[np.isnan(lut)] = 0.0
    return lut


def _saturation_threshold_table(joint: np.ndarray) -> np.ndarray:
    """bool [65536] table ``S > otsu(S)`` from the joint (max, max - min) histogram (int64 [65536])."""
    lut = _saturation_lut()
    present = joint > 0
    lo, hi = lut[present].min(), lut[present].max()
    if lo == hi:
        raise ValueError("threshold_otsu is expected to work with images having more than one color")
    hist, edges = np.histogram(lut[present], bins=256, range=(lo, hi), weights=joint[present])
    thr = _otsu_from_hist(hist, (edges[:-1] + edges[1:]) / 2.0)
    return lut > thr


def _otsu_u8_from_full_hist(full: np.ndarray) -> float:
    present = np.flatnonzero(full)
    lo, hi = int(present[0]), int(present[-1])
    if lo == hi:
        raise ValueError("threshold_otsu is expected to work with images having more than one color")
    return _otsu_from_hist(full[lo:hi + 1], np.arange(lo, hi + 1, dtype=np.float64))


def tissue_mask_device(rgb_dev, rgb_min: int = 50):
    """``tissue_mask`` for a level image already in HBM: ``rgb_dev`` cuda uint8 [a, b, 3] (any orientation) ->
    cuda uint8 [a, b] in {0, 1}.  Histograms and the mask run on the device (csrc/tissue.cuh); the four Otsu
    thresholds are derived on the host from 66 k counters with the same float64 code as the host path."""
    import ctypes as C
    import torch
    from . import _lib
    assert rgb_dev.is_cuda and rgb_dev.dtype == torch.uint8 and rgb_dev.dim() == 3 and rgb_dev.shape[2] == 3
    rgb_dev = rgb_dev.contiguous()
    n_pix = rgb_dev.shape[0] * rgb_dev.shape[1]
    st = C.c_void_p(torch.cuda.current_stream(rgb_dev.device).cuda_stream)
    with torch.cuda.device(rgb_dev.device):
        hist = torch.empty(768 + 65536, dtype=torch.int32, device=rgb_dev.device)
        _lib.check(_lib.lib.dp_tissue_hist(C.c_void_p(rgb_dev.data_ptr()), n_pix, C.c_void_p(hist.data_ptr()), st),
                   "dp_tissue_hist")
        h = hist.cpu().numpy().astype(np.int64)
        thr = [int(_otsu_u8_from_full_hist(h[256 * c:256 * (c + 1)])) for c in range(3)]   # integer bin centres
        table = torch.from_numpy(_saturation_threshold_table(h[768:]).astype(np.uint8)).to(rgb_dev.device)
        mask = torch.empty(rgb_dev.shape[:2], dtype=torch.uint8, device=rgb_dev.device)
        _lib.check(_lib.lib.dp_tissue_mask(C.c_void_p(rgb_dev.data_ptr()), n_pix, thr[0], thr[1], thr[2], int(rgb_min),
                                           C.c_void_p(table.data_ptr()), C.c_void_p(mask.data_ptr()), st),
                   "dp_tissue_mask")
    return mask


def morpho_process_device(mask_dev, level: int):
    """``morpho_process`` on a cuda uint8 [a, b] mask: close 20, open 5, dilate 60 / 35 / 10 (utils.py:200-219)."""
    import ctypes as C
    import torch
    from . import _lib
    if level <= 2:
        k_last = 60
    elif level == 3:
        k_last = 35
    elif level == 4:
        k_last = 10
    else:
        print(level)
        raise ValueError("Kernel for this level not fixed")
    assert mask_dev.is_cuda and mask_dev.dtype == torch.uint8 and mask_dev.dim() == 2
    m = mask_dev.contiguous().clone()
    tmp = torch.empty_like(m)
    st = C.c_void_p(torch.cuda.current_stream(m.device).cuda_stream)
    n0, n1 = m.shape
    with torch.cuda.device(m.device):
        for k, dil in ((20, 1), (20, 0), (5, 0), (5, 1), (k_last, 1)):      # close = dilate, erode; open = erode, dilate
            _lib.check(_lib.lib.dp_morph_rect(C.c_void_p(m.data_ptr()), C.c_void_p(m.data_ptr()),
                                              C.c_void_p(tmp.data_ptr()), n0, n1, k, dil, st), "dp_morph_rect")
    return m


def _level_rgb_device(slide, level: int, device):
    """Lowest-level RGB image as a cuda uint8 [x, y, 3] tensor, or None when the slide cannot provide one cheaply."""
    import torch
    from .slide import DeviceSlide
    if isinstance(slide, DeviceSlide):
        s = 2 ** int(level)
        w, h = slide.level_dimensions[level]
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        return slide.raster_xy[::s, ::s][:w, :h].contiguous().to(dev)
    return None


def tissue_mask(slide, level: int, rgb_min: int = 50) -> np.ndarray:
    """bool [x, y] mask at pyramid ``level`` (utils.py:336-354).

    Same result as the literal float64 evaluation (``saturation`` + ``threshold_otsu`` above, asserted in
    tests/test_host_pipeline.py), computed without float64 images: for uint8 RGB the saturation
    ``(max-min)/max`` takes one of 65 536 values indexed by ``(max, max-min)``, so its 256-bin histogram, the
    Otsu threshold and the ``S > t`` mask all follow from one joint histogram and a lookup table.
    """
    region = slide.read_region((0, 0), level, slide.level_dimensions[level])
    rgb = np.asarray(region.convert("RGB") if hasattr(region, "convert") else region)
    r, g, b = (np.ascontiguousarray(rgb[:, :, c]) for c in range(3))
    bg = (r > threshold_otsu(r)) & (g > threshold_otsu(g)) & (b > threshold_otsu(b))
    v = np.maximum(np.maximum(r, g), b)          # channel planes: a reduction over a length-3 axis is ~10x slower
    delta = v - np.minimum(np.minimum(r, g), b)
    idx = v.astype(np.uint16) * 256 + delta
    joint = _hist_u8(v, delta)
    tissue_s = _saturation_threshold_table(joint)[idx]
    above = (r > rgb_min) & (g > rgb_min) & (b > rgb_min)
    return np.ascontiguousarray((tissue_s & ~bg & above).T)


def mask_from_slide(mask_slide, level: int) -> np.ndarray:
    """uint8 [x, y] raw mask in {0, 255} from a mask image opened like a slide (dataloader.py:256-263): the whole
    ``level`` is read, reduced to PIL's 'L' luma, transposed, and every non-zero pixel set to 255."""
    region = mask_slide.read_region((0, 0), level, mask_slide.level_dimensions[level])
    if hasattr(region, "convert"):
        g = np.asarray(region.convert("L"))
    else:                                        # PIL's RGB -> L: ITU-R 601 luma in 16.16 fixed point, rounded
        rgb = np.asarray(region).astype(np.uint32)
        g = (rgb[..., 0] * 19595 + rgb[..., 1] * 38470 + rgb[..., 2] * 7471 + 0x8000) >> 16
    m = np.ascontiguousarray(g.T).astype(np.uint8)
    m[m > 0] = 255
    return m


def morpho_process(mask_u8: np.ndarray, level: int) -> np.ndarray:
    """close 20x20, open 5x5, dilate 60/35/10 by level; level > 4 raises like utils.py:200-219."""
    import cv2
    m = np.array(mask_u8, dtype=np.uint8)
    m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, np.ones((20, 20), np.uint8))
    m = cv2.morphologyEx(m, cv2.MORPH_OPEN, np.ones((5, 5), np.uint8))
    if level <= 2:
        k = np.ones((60, 60), np.uint8)
    elif level == 3:
        k = np.ones((35, 35), np.uint8)
    elif level == 4:
        k = np.ones((10, 10), np.uint8)
    else:
        print(level)
        raise ValueError("Kernel for this level not fixed")
    return cv2.dilate(m, k, iterations=1)


class TileGrid:
    """Tile origins of one slide in the reference's order, after clamping and ``drop_last``.

    ``coords`` is int32 [n, 2] (x, y) for the tiles that WILL be predicted: ``DataLoader(drop_last=True)``
    (Segmentation.py:92) silently drops the final ``n mod batch`` tiles.
    ``mask`` is a RAW tissue mask (the dataset applies its morphology to whatever mask it is given,
    dataloader.py:256-270): feeding ``self.mask`` back in dilates twice -- pass ``self.raw_mask``.
    """

    def __init__(self, slide, patch_size: int = 256, stride_size: int = 128, batch_size: int = 32,
                 roi_masking: bool = True, mask: np.ndarray | None = None, device=None):
        self.slide = slide
        self.patch_size = int(patch_size)
        self.stride_size = int(stride_size)
        self.level = len(slide.level_dimensions) - 1                       # dataloader.py:241 (mask_level ignored)
        self.factor = int(stride_size) // int(slide.level_downsamples[self.level])  # dataloader.py:242
        if self.factor < 1:
            raise ValueError("stride_size smaller than the mask level's downsample")
        X_slide, Y_slide = slide.level_dimensions[0]
        strided_dev = None
        rgb_dev = _level_rgb_device(slide, self.level, device) if (mask is None and device is not None) else None
        if rgb_dev is not None:
            # raster resident in HBM: histogram, mask, morphology and the stride sub-sampling on the device
            raw_dev = tissue_mask_device(rgb_dev)
            mask_dev = morpho_process_device(raw_dev, self.level)
            self.raw_mask = raw_dev.cpu().numpy().astype(bool)
            self.mask = mask_dev.cpu().numpy()
            strided_dev = mask_dev
        else:
            if mask is None:
                mask = tissue_mask(slide, self.level)
            self.raw_mask = mask                                           # before morphology: what ``mask=`` takes
            self.mask = morpho_process(np.uint8(mask), self.level)
        X_mask, Y_mask = self.mask.shape
        if X_slide // X_mask != Y_slide // Y_mask:
            raise Exception('Slide/Mask dimension does not match ,'
                            ' X_slide / X_mask : {} / {},'
                            ' Y_slide / Y_mask : {} / {}'.format(X_slide, X_mask, Y_slide, Y_mask))
        self.resolution = np.round(X_slide * 1.0 / X_mask)
        if not np.log2(self.resolution).is_integer():
            raise Exception('Resolution (X_slide / X_mask) is not power of 2 :'
                            ' {}'.format(self.resolution))
        if strided_dev is not None and roi_masking:
            import torch
            nz = torch.nonzero(strided_dev[::self.factor, ::self.factor]).cpu().numpy()   # row-major = x-major order
            self.X_idcs, self.Y_idcs = nz[:, 0] * self.factor, nz[:, 1] * self.factor
        else:
            ones = np.zeros_like(self.mask)
            ones[::self.factor, ::self.factor] = 1
            strided = ones * self.mask if roi_masking else ones
            self.X_idcs, self.Y_idcs = np.where(strided)                   # x-major order (dataloader.py:311)
        P = self.patch_size
        x = np.trunc(self.X_idcs * self.resolution - P // 2).astype(np.int64)   # int(...) of dataloader.py:348-349
        y = np.trunc(self.Y_idcs * self.resolution - P // 2).astype(np.int64)
        x = np.maximum(0, np.minimum(x, X_slide - P))                      # dataloader.py:351-353
        y = np.maximum(0, np.minimum(y, Y_slide - P))
        self.all_coords = np.stack([x, y], axis=1).astype(np.int32)
        self.batch_size = int(batch_size)
        self.n_batches = len(self.all_coords) // self.batch_size           # drop_last=True
        self.coords = self.all_coords[: self.n_batches * self.batch_size]

    def __len__(self):
        return self.n_batches
