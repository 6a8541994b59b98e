"""Slide sources: what stands where ``openslide.OpenSlide`` stood (DigiPathAI/loaders/dataloader.py:239).

The hot path only uses four things of an OpenSlide handle -- ``level_dimensions``, ``level_downsamples``,
``level_count`` and ``read_region`` (dataloader.py:241-246,357; utils.py:337) -- so any object with those works.

``ArraySlide`` holds a level-0 RGB raster in host memory (``uint8 [H, W, 3]``, ordinary image layout) with a
*virtual* pyramid: level ``k`` is the level-0 raster sub-sampled with stride ``2**k``.  OpenSlide is not
available in this image; when it is importable, ``open_slide`` hands real WSI files to it and decodes level 0
into an ``ArraySlide`` so that tiles can be cropped on the GPU (real-slide ingest proper is SURVEY.md N2).
"""
from __future__ import annotations

import os

import numpy as np


class ArraySlide:
    def __init__(self, raster_hwc: np.ndarray, n_levels: int = 1):
        a = np.asarray(raster_hwc)
        if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
            raise ValueError("ArraySlide wants a uint8 [H, W, 3] raster")
        self.raster = a
        H, W = a.shape[:2]
        self.level_count = int(n_levels)
        self.level_downsamples = tuple(float(2 ** k) for k in range(n_levels))
        self.level_dimensions = tuple((W // 2 ** k, H // 2 ** k) for k in range(n_levels))
        self.dimensions = self.level_dimensions[0]

    def read_region(self, location, level, size) -> np.ndarray:
        """RGB ``uint8 [h, w, 3]`` of the region whose level-0 top-left corner is ``location = (x, y)``.

        Like OpenSlide, pixels outside the slide are background; the reference converts OpenSlide's transparent
        RGBA to RGB via PIL (dataloader.py:357), which yields black there.
        """
        x, y = int(location[0]), int(location[1])
        w, h = int(size[0]), int(size[1])
        s = 2 ** int(level)
        lvl = self.raster[::s, ::s]
        lx, ly = x // s, y // s
        out = np.zeros((h, w, 3), dtype=np.uint8)
        H, W = lvl.shape[:2]
        x0, y0, x1, y1 = max(lx, 0), max(ly, 0), min(lx + w, W), min(ly + h, H)
        if x1 > x0 and y1 > y0:
            out[y0 - ly:y1 - ly, x0 - lx:x1 - lx] = lvl[y0:y1, x0:x1]
        return out

    def close(self):
        pass


def open_slide(path_or_obj, n_levels: int | None = None):
    """Path (``.npy`` raster, any PIL-readable image, a WSI if openslide is installed, or -- with
    ``DIGIPATH_DEVICE_INGEST=1`` -- a JPEG-tiled TIFF / SVS kept compressed for GPU decode) or slide-like object."""
    if hasattr(path_or_obj, "read_region") and hasattr(path_or_obj, "level_dimensions"):
        return path_or_obj
    if isinstance(path_or_obj, np.ndarray):
        return ArraySlide(path_or_obj, n_levels or 1)
    path = os.fspath(path_or_obj)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.endswith(".npy"):
        return ArraySlide(np.load(path, mmap_mode="r"), n_levels or 1)
    try:
        import openslide  # noqa: F401  (absent in the build image; present on a DigiPathAI deployment)
    except ImportError:
        openslide = None
    if openslide is not None:
        try:
            return openslide.OpenSlide(path)
        except Exception:  # noqa: BLE001 -- fall through to PIL for plain images
            pass
    if n_levels is None and os.environ.get("DIGIPATH_DEVICE_INGEST", "0") not in ("", "0") \
            and path.lower().endswith((".tif", ".tiff", ".svs")):
        # JPEG-tiled TIFF / SVS: keep the tiles compressed and decode level 0 on the GPU (ingest.py).
        # Opt-in until the nvJPEG path has been verified on hardware; a TiffSlide object can also be passed directly.
        from .wsi_tiff import TiffSlide
        try:
            ts = TiffSlide(path)
            if ts.device_decodable(0):
                return ts
            ts.close()
        except ValueError:
            pass
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        if n_levels is None:
            # a pyramidal TIFF (pages that halve in size, e.g. ImageMagick `ptif:` or tiffio.save_pyramidal) keeps its
            # level count -- the tissue mask is computed on the lowest level (dataloader.py:241-246); the levels
            # themselves are re-derived by sub-sampling level 0 like every ArraySlide
            n_levels, (w, h) = 1, im.size
            for k in range(1, getattr(im, "n_frames", 1)):
                im.seek(k)
                if abs(im.size[0] - (w >> k)) > 1 or abs(im.size[1] - (h >> k)) > 1:
                    break
                n_levels = k + 1
            im.seek(0)
        return ArraySlide(np.asarray(im.convert("RGB")), n_levels)


def level0_xy_raster(slide) -> np.ndarray:
    """Level-0 raster in the reference's tile orientation ``[x, y, c]`` (dataloader.py:357-358), host uint8."""
    if isinstance(slide, ArraySlide):
        return np.ascontiguousarray(np.transpose(slide.raster, (1, 0, 2)))
    W, H = slide.level_dimensions[0]
    img = slide.read_region((0, 0), 0, (W, H))
    arr = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img)
    return np.ascontiguousarray(np.transpose(arr, (1, 0, 2)))


class DeviceSlide:
    """Slide whose level-0 raster already lives in HBM as ``uint8 [W, H, 3]`` ([x, y, c], the layout the stem
    gather reads).  OpenSlide-shaped like ``ArraySlide`` (virtual stride-2**k pyramid); ``read_region`` copies
    only the requested (sub-sampled) window back to the host -- the tissue mask needs the lowest level only.
    """

    def __init__(self, raster_xy, n_levels: int = 1):
        assert raster_xy.is_cuda and raster_xy.dim() == 3 and raster_xy.shape[2] == 3
        self.raster_xy = raster_xy
        W, H = int(raster_xy.shape[0]), int(raster_xy.shape[1])
        self.level_count = int(n_levels)
        self.level_downsamples = tuple(float(2 ** k) for k in range(n_levels))
        self.level_dimensions = tuple((W // 2 ** k, H // 2 ** k) for k in range(n_levels))
        self.dimensions = self.level_dimensions[0]

    def read_region(self, location, level, size) -> np.ndarray:
        x, y = int(location[0]), int(location[1])
        w, h = int(size[0]), int(size[1])
        s = 2 ** int(level)
        lvl = self.raster_xy[::s, ::s]
        lx, ly = x // s, y // s
        out = np.zeros((h, w, 3), dtype=np.uint8)
        W, H = lvl.shape[0], lvl.shape[1]
        x0, y0, x1, y1 = max(lx, 0), max(ly, 0), min(lx + w, W), min(ly + h, H)
        if x1 > x0 and y1 > y0:
            win = lvl[x0:x1, y0:y1].permute(1, 0, 2).contiguous().cpu().numpy()
            out[y0 - ly:y1 - ly, x0 - lx:x1 - lx] = win
        return out

    def close(self):
        pass


def synthetic_slide_device(width: int, height: int, device, seed: int = 0, n_levels: int = 1, n_blobs: int = 4):
    """``synthetic_slide`` generated directly in HBM (same recipe, torch RNG instead of numpy's, so the pixels
    differ from the host version): white background 240+-3, elliptical tissue (170, 90, 160)+-20."""
    import torch
    rng = np.random.default_rng(seed)
    cy = rng.uniform(0.25, 0.75, n_blobs) * height
    cx = rng.uniform(0.25, 0.75, n_blobs) * width
    ry = rng.uniform(0.18, 0.30, n_blobs) * height
    rx = rng.uniform(0.18, 0.30, n_blobs) * width
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((width, height, 3), dtype=torch.uint8, device=device)
    ys = torch.arange(height, dtype=torch.float32, device=device)[None, :]
    tissue = torch.tensor([170.0, 90.0, 160.0], device=device)
    step = 2048
    for x0 in range(0, width, step):
        x1 = min(width, x0 + step)
        xs = torch.arange(x0, x1, dtype=torch.float32, device=device)[:, None]
        inside = torch.zeros((x1 - x0, height), dtype=torch.bool, device=device)
        for k in range(n_blobs):
            inside |= ((ys - cy[k]) / ry[k]) ** 2 + ((xs - cx[k]) / rx[k]) ** 2 <= 1.0
        noise = torch.randn((x1 - x0, height, 3), generator=g, device=device)
        blk = torch.where(inside[..., None], tissue + 20.0 * noise, 240.0 + 3.0 * noise)
        out[x0:x1] = blk.round_().clamp_(0, 255).to(torch.uint8)
    return DeviceSlide(out, n_levels)


def upload_xy_raster(slide, x_lo: int, x_hi: int, device, rows_per_chunk: int = 4096):
    """Level-0 stripe ``[x_lo, x_hi)`` as a CUDA ``uint8 [x, y, c]`` tensor (the layout the stem gather reads).

    The ``[y, x, c] -> [x, y, c]`` transposition the reference does per tile on the host
    (dataloader.py:357-358) happens once, on the device, while the raster is uploaded in row chunks.
    """
    import torch
    if isinstance(slide, DeviceSlide):
        if slide.raster_xy.device != torch.device(device):
            return slide.raster_xy[x_lo:x_hi].to(device)
        return slide.raster_xy[x_lo:x_hi]      # contiguous row slice of the resident raster: no copy
    from .wsi_tiff import TiffSlide
    if isinstance(slide, TiffSlide) and slide.device_decodable(0):
        from .ingest import upload_tiff_raster      # compressed tiles -> nvJPEG -> scatter, no host pixels
        return upload_tiff_raster(slide, x_lo, x_hi, device)
    W, H = slide.level_dimensions[0]
    out = torch.empty((x_hi - x_lo, H, 3), dtype=torch.uint8, device=device)
    for y0 in range(0, H, rows_per_chunk):
        y1 = min(H, y0 + rows_per_chunk)
        if isinstance(slide, ArraySlide):
            blk = slide.raster[y0:y1, x_lo:x_hi]
        else:
            img = slide.read_region((x_lo, y0), 0, (x_hi - x_lo, y1 - y0))
            blk = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img)
        t = torch.from_numpy(np.ascontiguousarray(blk)).to(device, non_blocking=False)
        out[:, y0:y1] = t.permute(1, 0, 2)
    return out


def synthetic_slide(width: int, height: int, seed: int = 0, n_levels: int = 1, n_blobs: int = 4) -> ArraySlide:
    """Synthetic H&E-like slide: white background 240+-3, elliptical 'tissue' (170, 90, 160) +- 20 covering ~50 %.

    Recipe from SURVEY.md 8(d) config 1; built blockwise so that 40k x 40k needs no float64 temporaries.
    """
    rng = np.random.default_rng(seed)
    img = np.empty((height, width, 3), dtype=np.uint8)
    cy = rng.uniform(0.25, 0.75, n_blobs) * height
    cx = rng.uniform(0.25, 0.75, n_blobs) * width
    ry = rng.uniform(0.18, 0.30, n_blobs) * height
    rx = rng.uniform(0.18, 0.30, n_blobs) * width
    tissue_rgb = np.array([170.0, 90.0, 160.0], dtype=np.float32)
    step = 1024
    xs_all = np.arange(width, dtype=np.float32)
    for y0 in range(0, height, step):
        y1 = min(height, y0 + step)
        ys = np.arange(y0, y1, dtype=np.float32)[:, None]
        inside = np.zeros((y1 - y0, width), dtype=bool)
        for k in range(n_blobs):
            inside |= ((ys - cy[k]) / ry[k]) ** 2 + ((xs_all[None, :] - cx[k]) / rx[k]) ** 2 <= 1.0
        blk = 240.0 + 3.0 * rng.standard_normal((y1 - y0, width, 3), dtype=np.float32)
        tis = tissue_rgb + 20.0 * rng.standard_normal((y1 - y0, width, 3), dtype=np.float32)
        blk[inside] = tis[inside]
        img[y0:y1] = np.clip(np.rint(blk), 0, 255).astype(np.uint8)
    return ArraySlide(img, n_levels)
