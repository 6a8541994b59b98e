"""Tiled / stripped JPEG TIFF slides (generic pyramidal TIFF, Aperio SVS, this repo's own result planes) read without
OpenSlide: SURVEY.md 8(f) row N2, the step in front of the hot path.

The reference opens a slide with ``openslide.OpenSlide`` and decodes every 256 x 256 patch on CPU worker processes
(``read_region``, DigiPathAI/loaders/dataloader.py:239,357-358).  A whole-slide TIFF already stores level 0 as
independent JPEG streams, one per tile; ``TiffSlide`` exposes them so that ``slide.upload_xy_raster`` can hand
batches of streams to nvJPEG (``ingest.JpegTileDecoder``) and scatter the decoded tiles straight into the
HBM-resident ``[x, y, c]`` raster -- no host-side pixel ever exists for level 0.

Only container parsing happens here (classic TIFF and BigTIFF, either byte order): IFD chain, tile or strip layout,
``JPEGTables`` (tag 347) merged back into each abbreviated stream.  ``read_region`` -- used by the path for the low
resolution mask level only (utils.py:337) -- decodes the covered tiles on the host with Pillow's libjpeg.
"""
from __future__ import annotations

import io
import struct

import numpy as np

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             13: "I", 16: "Q", 17: "q", 18: "Q"}
# Adobe APP14 segment with transform = 0: "components are RGB (or CMYK), no YCbCr transform"
_ADOBE_RGB = b"\xff\xee\x00\x0eAdobe\x00\x64\x00\x00\x00\x00\x00"


class TiffPage:
    """One IFD: geometry + where its compressed tiles (or strips, treated as full-width tiles) live."""

    def __init__(self, tags: dict):
        self.tags = tags
        self.width = int(tags[256][0])
        self.height = int(tags[257][0])
        self.samples = int(tags.get(277, (1,))[0])
        self.compression = int(tags.get(259, (1,))[0])
        self.photometric = int(tags.get(262, (1,))[0])
        self.subfile = int(tags.get(254, (0,))[0])
        self.tables = bytes(tags[347]) if 347 in tags else None
        self.bits = tuple(int(b) for b in tags.get(258, (1,)))
        if 322 in tags:
            self.tiled = True
            self.tile_w, self.tile_h = int(tags[322][0]), int(tags[323][0])
            self.offsets, self.counts = tags[324], tags[325]
        else:
            self.tiled = False
            self.tile_w = self.width
            self.tile_h = min(int(tags.get(278, (self.height,))[0]), self.height)
            self.offsets, self.counts = tags[273], tags[279]
        self.tiles_x = -(-self.width // self.tile_w)
        self.tiles_y = -(-self.height // self.tile_h)
        if len(self.offsets) < self.tiles_x * self.tiles_y:
            raise ValueError("TIFF page lists fewer tiles than its geometry needs (planar-separate data?)")


def _read_ifds(buf) -> list:
    """All IFDs of the main chain as {tag: tuple of values} (bytes for BYTE/ASCII/UNDEFINED arrays)."""
    bo = {b"II": "<", b"MM": ">"}.get(bytes(buf[:2]))
    if bo is None:
        raise ValueError("not a TIFF file")
    magic = struct.unpack(bo + "H", buf[2:4])[0]
    if magic == 42:
        big, off = False, struct.unpack(bo + "I", buf[4:8])[0]
    elif magic == 43:
        big, off = True, struct.unpack(bo + "Q", buf[8:16])[0]
    else:
        raise ValueError("not a TIFF file")
    n_fmt, n_size, e_size, c_fmt, v_size = ("Q", 8, 20, "Q", 8) if big else ("H", 2, 12, "I", 4)
    pages, seen = [], set()
    while off and off not in seen:
        seen.add(off)
        n = struct.unpack(bo + n_fmt, buf[off:off + n_size])[0]
        tags = {}
        for i in range(n):
            e = off + n_size + i * e_size
            tag, typ = struct.unpack(bo + "HH", buf[e:e + 4])
            cnt = struct.unpack(bo + c_fmt, buf[e + 4:e + 4 + v_size])[0]
            fmt = _TYPE_FMT.get(typ)
            if fmt is None:
                continue
            size = struct.calcsize("=" + fmt) * cnt
            vpos = e + 4 + v_size
            if size > v_size:
                vpos = struct.unpack(bo + c_fmt, buf[vpos:vpos + v_size])[0]
            raw = bytes(buf[vpos:vpos + size])
            if typ in (1, 2, 7):
                tags[tag] = raw if typ != 1 or cnt > 4 else tuple(raw)
            elif typ in (5, 10):
                v = struct.unpack(bo + fmt[0] * (2 * cnt), raw)
                tags[tag] = tuple((v[2 * k], v[2 * k + 1]) for k in range(cnt))
            else:
                tags[tag] = struct.unpack(bo + fmt * cnt, raw)
        pages.append(tags)
        nxt = off + n_size + n * e_size
        off = struct.unpack(bo + c_fmt, buf[nxt:nxt + v_size])[0]
    return pages


class TiffSlide:
    """OpenSlide-shaped handle (``level_dimensions``, ``level_downsamples``, ``level_count``, ``read_region``) on a
    JPEG-compressed tiled or stripped TIFF, plus ``jpeg_stream`` / ``tile_origin`` for the device ingest."""

    def __init__(self, path: str):
        import mmap
        self.path = path
        self._f = open(path, "rb")
        self._buf = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        pages = [TiffPage(t) for t in _read_ifds(self._buf) if 256 in t and 257 in t]
        if not pages:
            raise ValueError("TIFF without image pages")
        base = pages[0]
        # pyramid = the first page plus every later page of the same kind (tiled like the base, same channels)
        # whose size is the base size divided by one common factor; SVS thumbnails / label / macro pages
        # (stripped, other aspect ratio) drop out
        levels = [base]
        for p in pages[1:]:
            if p.tiled != base.tiled or p.samples != base.samples or p.width >= levels[-1].width:
                continue
            fx, fy = base.width / p.width, base.height / p.height
            if abs(fx - fy) / fx < 0.02:
                levels.append(p)
        self.pages = levels
        self.level_count = len(levels)
        self.level_dimensions = tuple((p.width, p.height) for p in levels)
        self.level_downsamples = tuple(float(round(base.width / p.width)) if k else 1.0
                                       for k, p in enumerate(levels))
        self.dimensions = self.level_dimensions[0]

    # ---- compressed access (device ingest) -------------------------------------------------------------------
    def device_decodable(self, level: int = 0) -> bool:
        p = self.pages[level]
        return p.compression == 7 and p.samples in (1, 3) and all(b == 8 for b in p.bits)

    def components_are_rgb(self, level: int = 0) -> bool:
        """True when the page stores R, G, B components directly (photometric 2) rather than YCbCr (6)."""
        p = self.pages[level]
        return p.photometric == 2 and p.samples == 3

    def tile_grid(self, level: int = 0):
        p = self.pages[level]
        return p.tiles_x, p.tiles_y, p.tile_w, p.tile_h

    def tile_origin(self, level: int, index: int):
        p = self.pages[level]
        return (index % p.tiles_x) * p.tile_w, (index // p.tiles_x) * p.tile_h

    def jpeg_stream(self, level: int, index: int) -> bytes:
        """Self-contained JPEG stream of tile ``index`` (row-major).  Abbreviated streams get the page's
        ``JPEGTables`` spliced in after SOI; pages whose photometric interpretation is RGB get an Adobe APP14
        segment saying so, because a bare 3-component stream is YCbCr by convention."""
        p = self.pages[level]
        if p.compression != 7:
            raise ValueError(f"tile data is not JPEG (TIFF compression {p.compression})")
        o, c = int(p.offsets[index]), int(p.counts[index])
        data = bytes(self._buf[o:o + c])
        if data[:2] != b"\xff\xd8":
            raise ValueError("tile does not start with a JPEG SOI marker")
        head = b"\xff\xd8"
        if p.photometric == 2 and p.samples == 3 and b"Adobe" not in data[:64]:
            head += _ADOBE_RGB
        if p.tables:
            t = p.tables
            if t[:2] != b"\xff\xd8" or t[-2:] != b"\xff\xd9":
                raise ValueError("malformed JPEGTables")
            head += t[2:-2]
        return head + data[2:] if len(head) > 2 else data

    # ---- host access (mask level) ----------------------------------------------------------------------------
    def _decode_tile_host(self, level: int, index: int) -> np.ndarray:
        from PIL import Image
        p = self.pages[level]
        if p.compression == 7:
            a = np.asarray(Image.open(io.BytesIO(self.jpeg_stream(level, index))).convert("RGB"))
        elif p.compression == 1:
            o, c = int(p.offsets[index]), int(p.counts[index])
            a = np.frombuffer(self._buf[o:o + c], np.uint8)
            a = a[:(len(a) // (p.tile_w * p.samples)) * p.tile_w * p.samples].reshape(-1, p.tile_w, p.samples)
            a = np.repeat(a, 3, axis=2) if p.samples == 1 else a[..., :3]
        else:
            raise NotImplementedError(f"TIFF compression {p.compression}: open the file with OpenSlide or Pillow")
        return a

    def read_region(self, location, level, size) -> np.ndarray:
        """RGB ``uint8 [h, w, 3]``; ``location`` is the level-0 position of the top-left corner (OpenSlide's
        convention); pixels outside the level are black."""
        p = self.pages[level]
        s = self.level_downsamples[level]
        lx, ly = int(location[0] // s), int(location[1] // s)
        w, h = int(size[0]), int(size[1])
        out = np.zeros((h, w, 3), np.uint8)
        x0, y0, x1, y1 = max(lx, 0), max(ly, 0), min(lx + w, p.width), min(ly + h, p.height)
        if x1 <= x0 or y1 <= y0:
            return out
        for ty in range(y0 // p.tile_h, -(-y1 // p.tile_h)):
            for tx in range(x0 // p.tile_w, -(-x1 // p.tile_w)):
                t = self._decode_tile_host(level, ty * p.tiles_x + tx)
                gx, gy = tx * p.tile_w, ty * p.tile_h
                ax0, ay0 = max(x0, gx), max(y0, gy)
                ax1, ay1 = min(x1, gx + t.shape[1]), min(y1, gy + t.shape[0])
                if ax1 > ax0 and ay1 > ay0:
                    out[ay0 - ly:ay1 - ly, ax0 - lx:ax1 - lx] = t[ay0 - gy:ay1 - gy, ax0 - gx:ax1 - gx]
        return out

    def close(self):
        try:
            self._buf.close()
        finally:
            self._f.close()
