"""Plane export: plain TIFF and tiled pyramidal JPEG TIFF.

The reference writes each result plane with ``tifffile.imsave(compress=9)`` and then re-encodes it in place with
ImageMagick -- ``convert <p> -compress jpeg -quality 90 -define tiff:tile-geometry=256x256 ptif:<p>``
(DigiPathAI/Segmentation.py:333-334,345-346,351-352; DigiPathAI/helpers/convert_to_pyramidal.py:32-37) -- because the
viewer's DeepZoom layer reads the masks back through OpenSlide (DigiPathAI/main_server.py:54-55,198-203).
Neither tool exists here, and a deflate-9 pass plus an ImageMagick pass over a 40 000^2 plane costs far more than
the whole B200 segmentation.  ``save_pyramidal`` writes the final file directly:

  * reduction pyramid: 2x mean-pool levels, on the GPU (``dp_pyramid_down2``) when the plane is a CUDA tensor;
  * every level is a TIFF page of 256x256 JPEG-compressed tiles (Compression = 7, quality 90, one self-contained
    JPEG stream per tile, edge tiles padded by border replication), reduced levels flagged NewSubfileType = 1 -- the layout ImageMagick's
    ``ptif:`` coder produces and OpenSlide's generic-TIFF backend reads;
  * tiles are encoded by Pillow's libjpeg on a thread pool (the encoder releases the GIL); constant tiles -- most
    of a result plane: background 0, saturated mask 255 -- are detected and share ONE JPEG stream per value
    (TileOffsets may alias; libtiff / Pillow / OpenSlide address tiles purely by offset and byte count); classic TIFF offsets
    (the compressed planes of this path stay far below 4 GiB; larger files raise).

``save_plane`` keeps the lossless single-level export (float32 'F' / uint8 'L').
"""
from __future__ import annotations

import io
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

TILE = 256


def save_plane(path: str, plane) -> None:
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    a = plane.detach().cpu().numpy() if hasattr(plane, "detach") else np.asarray(plane)
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint8:
        im = Image.fromarray(a, mode="L")
    else:
        im = Image.fromarray(a.astype(np.float32), mode="F")
    im.save(path, format="TIFF", compression="tiff_adobe_deflate")


def _to_uint8(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.uint8:
        return a
    return np.clip(np.rint(a.astype(np.float32)), 0, 255).astype(np.uint8)


def pyramid_levels(plane, min_side: int = TILE):
    """[plane, plane/2, plane/4, ...] as uint8 numpy arrays ([rows, cols]); stops once both sides fit one tile.

    CUDA tensors are reduced on the device in fp32 (exact 2x2 means, rounded once per level on export); a value
    range of [0, 255] is expected (masks are {0, 255}; probabilities / uncertainties are scaled by the caller).
    """
    levels = []
    if hasattr(plane, "is_cuda") and plane.is_cuda:
        import torch
        from . import engine
        cur = plane.to(torch.float32).contiguous()
        while True:
            # 8-bit conversion on the device (round-half-even like np.rint): a quarter of the bytes to download
            levels.append(cur.round().clamp_(0, 255).to(torch.uint8).cpu().numpy())
            if max(cur.shape) <= min_side or min(cur.shape) < 2:
                break
            with torch.cuda.device(cur.device):
                cur = engine.pyramid_down2(cur)
        return levels
    cur = (plane.detach().cpu().numpy() if hasattr(plane, "detach") else np.asarray(plane)).astype(np.float32)
    while True:
        levels.append(_to_uint8(cur))
        if max(cur.shape) <= min_side or min(cur.shape) < 2:
            break
        h, w = cur.shape[0] // 2 * 2, cur.shape[1] // 2 * 2
        c = cur[:h, :w]
        cur = 0.25 * ((c[0::2, 0::2] + c[0::2, 1::2]) + (c[1::2, 0::2] + c[1::2, 1::2]))
    return levels


def _jpeg(tile: np.ndarray, quality: int) -> bytes:
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(tile, mode="L").save(buf, format="JPEG", quality=quality)
    return buf.getvalue()


def _encode_tile(args):
    """One tile of level ``a`` -> JPEG bytes, or the tile's value (int) when it is constant: result planes are mostly
    flat (background 0, saturated mask 255), and the writer stores one shared JPEG stream per constant value."""
    a, j, i, tile, quality = args
    t = a[j * tile:(j + 1) * tile, i * tile:(i + 1) * tile]
    lo = t.min()
    if lo == t.max():
        return int(lo)
    if t.shape != (tile, tile):       # edge tile: replicate the border (zero padding would ring into the image)
        t = np.pad(t, ((0, tile - t.shape[0]), (0, tile - t.shape[1])), mode="edge")
    return _jpeg(np.ascontiguousarray(t), quality)


def save_pyramidal(path: str, plane, quality: int = 90, tile: int = TILE, threads: int | None = None) -> int:
    """Writes ``plane`` ([rows, cols], uint8 or float in [0, 255]) as a tiled pyramidal JPEG TIFF; returns the
    number of levels.  ``threads`` = JPEG encoder threads (default: every host core, at most 32)."""
    if threads is None:
        import os
        threads = max(1, min(32, os.cpu_count() or 1))
    levels = pyramid_levels(plane, tile)
    out = bytearray(b"II*\0\0\0\0\0")            # little-endian classic TIFF, first-IFD offset patched below
    ifd_offset_pos = 4
    const_at = {}                                # value -> (offset, byte count) of the shared constant-tile stream
    with ThreadPoolExecutor(max_workers=threads) as pool:
        for li, a in enumerate(levels):
            rows, cols = a.shape
            ty, tx = -(-rows // tile), -(-cols // tile)
            jobs = ((a, j, i, tile, quality) for j in range(ty) for i in range(tx))
            offsets, counts = [], []
            for b in pool.map(_encode_tile, jobs, chunksize=16):
                if isinstance(b, int):            # constant tile: every such tile points at one stream per value
                    if b not in const_at:
                        blob = _jpeg(np.full((tile, tile), b, np.uint8), quality)
                        if len(out) & 1:
                            out.append(0)
                        const_at[b] = (len(out), len(blob))
                        out.extend(blob)
                    off, cnt = const_at[b]
                    offsets.append(off)
                    counts.append(cnt)
                    continue
                if len(out) & 1:
                    out.append(0)
                offsets.append(len(out))
                counts.append(len(b))
                out.extend(b)
            n = len(offsets)

            def arr(vals):
                """Offset of an out-of-line LONG array (or the value itself when it fits the entry)."""
                if len(vals) == 1:
                    return vals[0]
                if len(out) & 1:
                    out.append(0)
                off = len(out)
                out.extend(struct.pack(f"<{len(vals)}I", *vals))
                return off

            off_tab, cnt_tab = arr(offsets), arr(counts)
            entries = [  # (tag, type, count, value)   types: 3 = SHORT, 4 = LONG
                (254, 4, 1, 0 if li == 0 else 1),     # NewSubfileType: reduced-resolution image
                (256, 4, 1, cols), (257, 4, 1, rows),
                (258, 3, 1, 8), (259, 3, 1, 7),       # 8 bits, JPEG
                (262, 3, 1, 1), (277, 3, 1, 1),       # BlackIsZero, 1 sample per pixel
                (322, 4, 1, tile), (323, 4, 1, tile),
                (324, 4, n, off_tab), (325, 4, n, cnt_tab),
            ]
            if len(out) & 1:
                out.append(0)
            ifd = len(out)
            if ifd >= 2 ** 32 - 4096:
                raise ValueError("pyramidal TIFF exceeds the 4 GiB classic-TIFF limit")
            out[ifd_offset_pos:ifd_offset_pos + 4] = struct.pack("<I", ifd)
            out.extend(struct.pack("<H", len(entries)))
            for tag, typ, cnt, val in sorted(entries):
                out.extend(struct.pack("<HHI", tag, typ, cnt))
                out.extend(struct.pack("<HH", val, 0) if (typ == 3 and cnt == 1) else struct.pack("<I", val))
            ifd_offset_pos = len(out)
            out.extend(b"\0\0\0\0")                  # next-IFD offset (0 terminates the chain)
    with open(path, "wb") as f:
        f.write(out)
    return len(levels)
