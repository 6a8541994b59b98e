"""Tissue mask and tile grid: the host-side prologue of the hot path (one-off per slide, CPU).

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
    vv, dd = np.divmod(np.arange(65536), 256)
    with np.errstate(divide="ignore", invalid="ignore"):
        lut = (dd / 255.0) / (vv / 255.0)        # identical float64 arithmetic to saturation()
    lut[dd == 0] = 0.0
    lut[np.isnan(lut)] = 0.0
    present = joint > 0
    lo, hi = lut[present].min(), lut[present].max()
    if lo == hi:
        raise ValueError("threshold_otsu is expected to work with images having more than one color")
    hist, edges = np.histogram(lut[present], bins=256, range=(lo, hi), weights=joint[present])
    thr = _otsu_from_hist(hist, (edges[:-1] + edges[1:]) / 2.0)
    tissue_s = (lut > thr)[idx]
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
                 roi_masking: bool = True, mask: np.ndarray | None = None):
        self.slide = slide
        self.patch_size = int(patch_size)
        self.stride_size = int(stride_size)
        self.level = len(slide.level_dimensions) - 1                       # dataloader.py:241 (mask_level ignored)
        self.factor = int(stride_size) // int(slide.level_downsamples[self.level])  # dataloader.py:242
        if self.factor < 1:
            raise ValueError("stride_size smaller than the mask level's downsample")
        X_slide, Y_slide = slide.level_dimensions[0]
        if mask is None:
            mask = tissue_mask(slide, self.level)
        self.raw_mask = mask                                               # before morphology: what ``mask=`` takes
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
        ones = np.zeros_like(self.mask)
        ones[::self.factor, ::self.factor] = 1
        strided = ones * self.mask if roi_masking else ones
        self.X_idcs, self.Y_idcs = np.where(strided)                       # x-major order (dataloader.py:311)
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
