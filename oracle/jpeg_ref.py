"""CPU restatement of the baseline-JPEG tile encoder of the pyramidal writer -- TEST INFRASTRUCTURE (row N1).

What it restates: the entropy-coded scan of one 256 x 256 grayscale tile as ITU-T T.81 defines it (level shift, 8 x 8
DCT-II, quantisation, zig-zag, DC differences, run/size Huffman symbols with ZRL and EOB, 1-bit padding, 0xFF byte
stuffing), i.e. what ImageMagick's `-compress jpeg` produces for the reference's result files
(DigiPathAI/Segmentation.py:333-334) up to libjpeg's choice of DCT arithmetic.  How it is pinned: the quantiser and Huffman
tables are parsed out of a stream written by libjpeg (Pillow) and the streams built here are DECODED by libjpeg in
tests/test_jpeg_oracle.py -- libjpeg is the golden decoder; the decoded tiles must equal what libjpeg's own encoder at
the same quality decodes to, within the float-vs-integer DCT difference.
"""
from __future__ import annotations

import struct

import numpy as np

ZIGZAG = (0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
          62, 63)


def parse_tables(jpeg: bytes):
    """(header = SOI .. end of the SOS header, quantiser[64] in zig-zag order, {(class, id): (counts, values)})."""
    assert jpeg[:2] == b"\xff\xd8"
    pos, q, huff = 2, None, {}
    while pos < len(jpeg):
        assert jpeg[pos] == 0xFF
        marker = jpeg[pos + 1]
        n = struct.unpack(">H", jpeg[pos + 2:pos + 4])[0]
        body = jpeg[pos + 4:pos + 2 + n]
        if marker == 0xDB:
            i = 0
            while i < len(body):
                assert body[i] >> 4 == 0, "8-bit quantiser expected"
                if body[i] & 15 == 0:
                    q = np.frombuffer(body[i + 1:i + 65], np.uint8).astype(np.int64)
                i += 65
        elif marker == 0xC4:
            i = 0
            while i < len(body):
                counts = list(body[i + 1:i + 17])
                huff[(body[i] >> 4, body[i] & 15)] = (counts, list(body[i + 17:i + 17 + sum(counts)]))
                i += 17 + sum(counts)
        elif marker == 0xDA:
            return jpeg[:pos + 2 + n], q, huff
        pos += 2 + n
    raise ValueError("no SOS marker")


def canonical_codes(counts, values):
    """Huffman table -> {symbol: (code, length)} (T.81 Annex C)."""
    out, code, it = {}, 0, iter(values)
    for length, cnt in enumerate(counts, start=1):
        for _ in range(cnt):
            out[next(it)] = (code, length)
            code += 1
        code <<= 1
    return out


def _dct_matrix():
    k = np.arange(8)
    m = np.cos((2 * k[None, :] + 1) * k[:, None] * np.pi / 16) * 0.5
    m[0] *= np.sqrt(0.5)
    return m                                           # F = M f M^T


def encode_scan(tile: np.ndarray, q_zigzag, huff) -> bytes:
    """Entropy-coded scan (byte-stuffed, padded, no markers) of a uint8 tile whose sides are multiples of 8."""
    h, w = tile.shape
    assert h % 8 == 0 and w % 8 == 0
    M = _dct_matrix()
    dc, ac = canonical_codes(*huff[(0, 0)]), canonical_codes(*huff[(1, 0)])
    qnat = np.zeros(64)
    for k, n in enumerate(ZIGZAG):
        qnat[n] = q_zigzag[k]
    bits = []                                          # list of (code, length)

    def put_value(v, cat):
        if cat:
            bits.append(((v - 1 if v < 0 else v) & ((1 << cat) - 1), cat))

    prev = 0
    f = tile.astype(np.float64) - 128.0
    for by in range(h // 8):
        for bx in range(w // 8):
            blk = M @ f[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] @ M.T
            c = np.rint(blk.reshape(64) / qnat).astype(np.int64)[list(ZIGZAG)]
            d = int(c[0]) - prev
            prev = int(c[0])
            cat = 0 if d == 0 else int(abs(d)).bit_length()
            bits.append(dc[cat])
            put_value(d, cat)
            run = 0
            for k in range(1, 64):
                v = int(c[k])
                if v == 0:
                    run += 1
                    continue
                while run >= 16:
                    bits.append(ac[0xF0])
                    run -= 16
                cat = int(abs(v)).bit_length()
                bits.append(ac[(run << 4) | cat])
                put_value(v, cat)
                run = 0
            if run:
                bits.append(ac[0x00])
    acc, n = 0, 0
    for code, length in bits:
        acc = (acc << length) | code
        n += length
    pad = -n % 8
    acc = (acc << pad) | ((1 << pad) - 1)
    raw = acc.to_bytes((n + pad) // 8, "big")
    return raw.replace(b"\xff", b"\xff\x00")


def encode_tile(tile: np.ndarray, reference_jpeg: bytes) -> bytes:
    """A complete JPEG stream for ``tile`` with the header and tables of ``reference_jpeg`` (same dimensions)."""
    header, q, huff = parse_tables(reference_jpeg)
    return header + encode_scan(tile, q, huff) + b"\xff\xd9"
