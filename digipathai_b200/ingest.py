"""Device ingest of JPEG-tiled slides: ctypes binding of ``libdigipath_ingest.so`` (C ABI in
``include/digipath_ingest.h``) and the loop that fills the HBM raster from a ``wsi_tiff.TiffSlide``.

Stands where the reference's DataLoader workers stood (openslide ``read_region`` -> PIL -> numpy -> transpose per
patch on 8 CPU processes, DigiPathAI/loaders/dataloader.py:239,357-358; Segmentation.py:92): level-0 tiles go to
nvJPEG in batches as compressed streams and are scattered on the device into the ``[x, y, c]`` raster that
``dp_forward_tiles`` crops from.  No fallback: without the library this module raises on import.

Status: container-side logic tested on CPU (tests/test_wsi_tiff.py); decode + scatter parity on the B200 in
tests/test_gpu_wsi_ingest.py.  ``slide.open_slide`` returns a ``TiffSlide`` when asked to
(``DIGIPATH_DEVICE_INGEST=1``) or when handed one.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdigipath_ingest.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: run __graft_entry__.build(). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_SIGS = {
    "dp_ingest_abi_version": (C.c_int, []),
    "dp_ingest_last_error": (C.c_char_p, []),
    "dp_jpeg_decoder_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "dp_jpeg_decoder_destroy": (C.c_int, [C.c_void_p]),
    "dp_jpeg_decode_tiles": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]),
    "dp_jpeg_decode_tiles_ex": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dp_scatter_tiles_xy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class IngestError(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise IngestError(f"{what}: {lib.dp_ingest_last_error().decode('utf-8', 'replace')}")


class JpegTileDecoder:
    """nvJPEG batch decoder bound to one CUDA device."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self.device = int(device)
        _check(lib.dp_jpeg_decoder_create(self.device, C.byref(self._h)), "dp_jpeg_decoder_create")

    def close(self):
        if self._h:
            lib.dp_jpeg_decoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    def decode(self, streams, tile_w: int, tile_h: int, out=None, components_are_rgb: bool = False):
        """``streams``: list of self-contained JPEG byte strings -> cuda uint8 ``[n, tile_h, tile_w, 3]`` (RGB).
        Slots of streams smaller than the tile keep whatever ``out`` held (zeros when allocated here).
        ``components_are_rgb``: the stored components are R, G, B (TIFF photometric 2), not YCbCr."""
        import torch
        n = len(streams)
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.zeros((n, tile_h, tile_w, 3), dtype=torch.uint8, device=dev)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.shape == (n, tile_h, tile_w, 3)
        ptrs = (C.c_char_p * n)(*streams)
        lens = (C.c_size_t * n)(*[len(s) for s in streams])
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            _check(lib.dp_jpeg_decode_tiles_ex(self._h, ptrs, lens, n, int(tile_w), int(tile_h),
                                               C.c_void_p(out.data_ptr()), int(bool(components_are_rgb)),
                                               C.c_void_p(st)), "dp_jpeg_decode_tiles_ex")
        return out


def scatter_tiles_xy(tiles, origins, raster, x_lo: int):
    """tiles cuda uint8 [n, th, tw, 3]; origins cuda int32 [n, 2] (level-0 x, y); raster cuda uint8 [x, y, 3] that
    covers slide columns [x_lo, x_lo + raster.shape[0])."""
    import torch
    assert tiles.is_cuda and tiles.dtype == torch.uint8 and tiles.is_contiguous() and tiles.dim() == 4
    assert origins.is_cuda and origins.dtype == torch.int32 and origins.is_contiguous() and origins.shape == (tiles.shape[0], 2)
    assert raster.is_cuda and raster.dtype == torch.uint8 and raster.is_contiguous() and raster.shape[2] == 3
    n, th, tw, _ = tiles.shape
    with torch.cuda.device(raster.device):
        st = torch.cuda.current_stream().cuda_stream
        _check(lib.dp_scatter_tiles_xy(C.c_void_p(tiles.data_ptr()), n, tw, th, C.c_void_p(origins.data_ptr()),
                                       C.c_void_p(raster.data_ptr()), int(x_lo), int(x_lo) + raster.shape[0],
                                       raster.shape[1], C.c_void_p(st)), "dp_scatter_tiles_xy")


def upload_tiff_raster(slide, x_lo: int, x_hi: int, device, batch: int = 256):
    """Level-0 columns ``[x_lo, x_hi)`` of a ``TiffSlide`` as a cuda ``uint8 [x, y, c]`` raster: only the tiles that
    intersect the stripe are read; each batch of compressed streams is decoded by nvJPEG and scattered in place."""
    import torch
    dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
    W, H = slide.level_dimensions[0]
    tiles_x, tiles_y, tw, th = slide.tile_grid(0)
    out = torch.zeros((x_hi - x_lo, H, 3), dtype=torch.uint8, device=dev)
    cols = range(max(0, x_lo // tw), min(tiles_x, -(-x_hi // tw)))
    todo = [ty * tiles_x + tx for ty in range(tiles_y) for tx in cols]
    dec = JpegTileDecoder(dev.index if dev.index is not None else torch.cuda.current_device())
    try:
        buf = torch.zeros((batch, th, tw, 3), dtype=torch.uint8, device=dev)
        for s in range(0, len(todo), batch):
            idx = todo[s:s + batch]
            streams = [slide.jpeg_stream(0, k) for k in idx]
            org = torch.tensor([slide.tile_origin(0, k) for k in idx], dtype=torch.int32).to(dev)
            tiles = dec.decode(streams, tw, th, out=buf[:len(idx)], components_are_rgb=slide.components_are_rgb(0))
            scatter_tiles_xy(tiles, org, out, x_lo)
        torch.cuda.synchronize(dev)
    finally:
        dec.close()
    return out
