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
  * a CUDA plane is encoded ON THE DEVICE (``dp_jpeg_encode_gray_tiles``, csrc/jpeg_enc.cuh: one CTA per tile -- DCT,
    quantisation, Huffman coding at prefix-summed bit offsets, byte stuffing); only the compressed streams cross PCIe and
    the host prepends the JPEG header whose tables the kernel used.  Host arrays are encoded by Pillow's libjpeg on a
    thread pool (the encoder releases the GIL).  Either way constant tiles -- most of a result plane: background 0,
    saturated mask 255 -- are detected and share ONE JPEG stream per value (TileOffsets may alias; libtiff / Pillow /
    OpenSlide address tiles purely by offset and byte count); classic TIFF offsets (the compressed planes of this path
    stay far below 4 GiB; larger files raise).

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


_ZIGZAG = (0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47,
           55, 62, 63)
_TABLES = {}


def _jpeg_tables(quality: int, tile: int = TILE):
    """(header, packed tables) for the device encoder: the bytes SOI .. end of the SOS header of a grayscale baseline JPEG
    of a ``tile`` x ``tile`` image as the host's libjpeg writes it at this quality, and the quantiser / Huffman tables
    parsed out of those very bytes, packed as csrc/jpeg_enc.cuh:JpegTables.  Scan data emitted with these tables after
    this header (+ EOI) is a valid stream by construction."""
    key = (quality, tile)
    if key in _TABLES:
        return _TABLES[key]
    ref = _jpeg(np.arange(tile * tile, dtype=np.uint32).reshape(tile, tile).astype(np.uint8), quality)
    assert ref[:2] == b"\xff\xd8"
    pos, q, huff, header = 2, None, {}, None
    while pos < len(ref):
        assert ref[pos] == 0xFF, "JPEG marker expected"
        marker = ref[pos + 1]
        n = struct.unpack(">H", ref[pos + 2:pos + 4])[0]
        body = ref[pos + 4:pos + 2 + n]
        if marker == 0xDB:                                   # DQT: (Pq << 4 | Tq), 64 values in zig-zag order
            i = 0
            while i < len(body):
                pq, tq = body[i] >> 4, body[i] & 15
                vals = (np.frombuffer(body[i + 1:i + 65], np.uint8) if pq == 0 else
                        np.frombuffer(body[i + 1:i + 129], ">u2")).astype(np.float64)
                if tq == 0:
                    q = vals
                i += 65 if pq == 0 else 129
        elif marker == 0xC4:                                 # DHT: (Tc << 4 | Th), 16 counts, values
            i = 0
            while i < len(body):
                tc, th = body[i] >> 4, body[i] & 15
                counts = list(body[i + 1:i + 17])
                vals = list(body[i + 17:i + 17 + sum(counts)])
                huff[(tc, th)] = (counts, vals)
                i += 17 + sum(counts)
        elif marker == 0xC0:
            assert body[0] == 8 and struct.unpack(">HH", body[1:5]) == (tile, tile) and body[5] == 1, "baseline gray expected"
        elif marker in (0xC1, 0xC2):
            raise ValueError("the host JPEG library did not write a baseline stream")
        elif marker == 0xDA:                                 # SOS: scan data follows
            header = ref[:pos + 2 + n]
            break
        pos += 2 + n
    assert q is not None and header is not None and (0, 0) in huff and (1, 0) in huff
    inv_q = np.zeros(64, np.float32)
    for k, nat in enumerate(_ZIGZAG):
        inv_q[nat] = 1.0 / q[k]

    def canonical(counts, vals, size):
        code_of, len_of = np.zeros(size, np.uint16), np.zeros(size, np.uint8)
        code, it = 0, iter(vals)
        for length, cnt in enumerate(counts, start=1):
            for _ in range(cnt):
                v = next(it)
                code_of[v], len_of[v] = code, length
                code += 1
            code <<= 1
        return code_of, len_of

    dc_code, dc_len = canonical(*huff[(0, 0)], 16)
    ac_code, ac_len = canonical(*huff[(1, 0)], 256)
    packed = inv_q.tobytes() + dc_code.tobytes() + dc_len.tobytes() + ac_code.tobytes() + ac_len.tobytes()
    assert len(packed) == 1072
    _TABLES[key] = (header, packed)
    return _TABLES[key]


def pyramid_levels_device(plane, min_side: int = TILE):
    """Like ``pyramid_levels`` for a CUDA plane, but the uint8 levels stay on the device."""
    import torch
    from . import engine
    levels = []
    cur = plane.to(torch.float32).contiguous()
    while True:
        levels.append(cur.round().clamp_(0, 255).to(torch.uint8))
        if max(cur.shape) <= min_side or min(cur.shape) < 2:
            break
        with torch.cuda.device(cur.device):
            cur = engine.pyramid_down2(cur)
    return levels


def encode_level_device(level, quality: int = 90, chunk_tiles: int = 2048, scratch_per_tile: int = 64 << 10,
                        out_cap: int = 96 << 10):
    """One pyramid level (CUDA uint8 [rows, cols]) -> list over its tiles (row-major) of either ``bytes`` (a complete
    JPEG stream) or ``int`` (the value of a constant tile).  Tiles whose stream exceeds the device capacities (never seen
    on result planes; needs > 2 bits per pixel) are encoded on the host."""
    import ctypes as C

    import torch
    from . import _lib
    assert level.is_cuda and level.dtype == torch.uint8 and level.dim() == 2 and level.is_contiguous()
    header, packed = _jpeg_tables(quality)
    tab = (C.c_char * len(packed)).from_buffer_copy(packed)
    rows, cols = level.shape
    ty, tx = -(-rows // TILE), -(-cols // TILE)
    n_all = ty * tx
    out_list = []
    dev = level.device
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for t0 in range(0, n_all, chunk_tiles):
            n = min(chunk_tiles, n_all - t0)
            ws_bytes = int(_lib.lib.dp_jpeg_encode_workspace_bytes(n, scratch_per_tile))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            fixed = torch.empty(n * out_cap, dtype=torch.uint8, device=dev)
            sizes = torch.empty(n, dtype=torch.int32, device=dev)
            flags = torch.empty(n, dtype=torch.int32, device=dev)
            _lib.check(_lib.lib.dp_jpeg_encode_gray_tiles(C.c_void_p(level.data_ptr()), rows, cols, t0, n, tab, len(packed),
                                                          C.c_void_p(ws.data_ptr()), ws_bytes, scratch_per_tile,
                                                          C.c_void_p(fixed.data_ptr()), out_cap, C.c_void_p(sizes.data_ptr()),
                                                          C.c_void_p(flags.data_ptr()), st), "dp_jpeg_encode_gray_tiles")
            offs = torch.cumsum(sizes.to(torch.int64), 0) - sizes
            total = int((offs[-1] + sizes[-1]).item())
            compact = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib.dp_jpeg_compact(C.c_void_p(fixed.data_ptr()), out_cap, C.c_void_p(sizes.data_ptr()),
                                                C.c_void_p(offs.data_ptr()), C.c_void_p(compact.data_ptr()), n, st),
                       "dp_jpeg_compact")
            blob = compact.cpu().numpy().tobytes()
            sz, fl, of = sizes.cpu().numpy(), flags.cpu().numpy(), offs.cpu().numpy()
            for i in range(n):
                if fl[i] & 1:
                    out_list.append(int(fl[i] >> 8) & 0xFF)
                elif fl[i] & 2:                                # capacity exceeded: this tile on the host
                    j, k = divmod(t0 + i, tx)
                    t = level[j * TILE:(j + 1) * TILE, k * TILE:(k + 1) * TILE].cpu().numpy()
                    if t.shape != (TILE, TILE):
                        t = np.pad(t, ((0, TILE - t.shape[0]), (0, TILE - t.shape[1])), mode="edge")
                    out_list.append(_jpeg(np.ascontiguousarray(t), quality))
                else:
                    out_list.append(header + blob[of[i]:of[i] + sz[i]] + b"\xff\xd9")
    return out_list


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
    on_device = hasattr(plane, "is_cuda") and plane.is_cuda and tile == TILE
    levels = pyramid_levels_device(plane, tile) if on_device else pyramid_levels(plane, tile)
    out = bytearray(b"II*\0\0\0\0\0")            # little-endian classic TIFF, first-IFD offset patched below
    ifd_offset_pos = 4
    const_at = {}                                # value -> (offset, byte count) of the shared constant-tile stream
    with ThreadPoolExecutor(max_workers=threads) as pool:
        for li, a in enumerate(levels):
            rows, cols = a.shape
            ty, tx = -(-rows // tile), -(-cols // tile)
            jobs = ((a, j, i, tile, quality) for j in range(ty) for i in range(tx))
            offsets, counts = [], []
            for b in (encode_level_device(a, quality) if on_device else pool.map(_encode_tile, jobs, chunksize=16)):
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
