"""Pyramidal tiled JPEG TIFF writer (reference: Segmentation.py:333-352 + ImageMagick `ptif:`)."""
import struct

import numpy as np

from digipathai_b200 import tiffio


def _smooth(rows, cols):
    y, x = np.mgrid[0:rows, 0:cols]
    return (127.5 + 127.5 * np.sin(x / 37.0) * np.cos(y / 53.0)).astype(np.float32)


def _ifds(path):
    b = open(path, "rb").read()
    assert b[:4] == b"II*\0"
    off = struct.unpack("<I", b[4:8])[0]
    out = []
    while off:
        n = struct.unpack("<H", b[off:off + 2])[0]
        tags = {}
        for i in range(n):
            tag, typ, cnt, val = struct.unpack("<HHII", b[off + 2 + 12 * i: off + 14 + 12 * i])
            tags[tag] = (typ, cnt, val & 0xFFFF if (typ == 3 and cnt == 1) else val)
        out.append(tags)
        off = struct.unpack("<I", b[off + 2 + 12 * n: off + 6 + 12 * n])[0]
    return out, b


def test_pyramid_structure_and_content(tmp_path):
    from PIL import Image
    a = _smooth(700, 1000)
    path = str(tmp_path / "p.tiff")
    n = tiffio.save_pyramidal(path, a)
    assert n == 3                                             # 1000x700 -> 500x350 -> 250x175
    ifds, raw = _ifds(path)
    assert [(d[256][2], d[257][2]) for d in ifds] == [(1000, 700), (500, 350), (250, 175)]
    assert [d[254][2] for d in ifds] == [0, 1, 1]             # reduced-resolution flag on the sub-levels
    for d in ifds:
        assert d[259][2] == 7 and d[322][2] == 256 and d[323][2] == 256 and d[258][2] == 8 and d[277][2] == 1
        tiles = -(-d[256][2] // 256) * -(-d[257][2] // 256)
        assert d[324][1] == tiles and d[325][1] == tiles
    # every tile is a self-contained JPEG stream
    d0 = ifds[0]
    offs = struct.unpack(f"<{d0[324][1]}I", raw[d0[324][2]: d0[324][2] + 4 * d0[324][1]])
    assert all(raw[o:o + 2] == b"\xff\xd8" for o in offs)
    im = Image.open(path)
    assert im.n_frames == 3
    lvl0 = np.asarray(im).astype(np.float32)
    assert lvl0.shape == (700, 1000) and np.abs(lvl0 - a).mean() < 1.5
    im.seek(1)
    lvl1 = np.asarray(im).astype(np.float32)
    want1 = 0.25 * (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2])
    assert lvl1.shape == (350, 500) and np.abs(lvl1 - want1).mean() < 1.5


def test_small_and_binary_planes(tmp_path):
    from PIL import Image
    m = np.zeros((300, 200), np.uint8)
    m[50:200, 40:160] = 255
    path = str(tmp_path / "m.tiff")
    assert tiffio.save_pyramidal(path, m) == 2
    back = np.asarray(Image.open(path))
    assert ((back > 127) == (m > 127)).mean() > 0.998
    one = str(tmp_path / "one.tiff")
    assert tiffio.save_pyramidal(one, np.full((100, 90), 7, np.uint8)) == 1     # fits one tile: single level
    assert np.abs(np.asarray(Image.open(one)).astype(int) - 7).max() <= 2


def test_pyramidal_tiff_opens_as_a_multi_level_slide(tmp_path):
    from digipathai_b200.slide import open_slide
    a = _smooth(600, 900)
    path = str(tmp_path / "s.tiff")
    assert tiffio.save_pyramidal(path, a) == 3
    s = open_slide(path)
    assert s.level_count == 3 and s.level_dimensions[0] == (900, 600) and s.level_dimensions[2] == (225, 150)
    region = s.read_region((0, 0), 0, (64, 64))
    assert region.shape == (64, 64, 3) and np.abs(region[..., 0].astype(np.float32) - a[:64, :64]).mean() < 2.0


def test_constant_tiles_share_one_stream_per_value(tmp_path):
    """A mask-like plane (flat 0 / flat 255 with one ragged blob edge): every flat tile points at one shared JPEG
    stream per value, the file stays small, and the image reads back exactly on the flat tiles."""
    from PIL import Image
    m = np.zeros((2100, 2600), np.uint8)                      # 9 x 11 tiles on level 0, ragged right / bottom edge
    m[300:1700, 500:2600] = 255
    m[1000, 1000] = 0                                          # one non-flat tile inside the blob
    path = str(tmp_path / "flat.tiff")
    assert tiffio.save_pyramidal(path, m) == 5
    ifds, raw = _ifds(path)
    d0 = ifds[0]
    n = d0[324][1]
    offs = struct.unpack(f"<{n}I", raw[d0[324][2]: d0[324][2] + 4 * n])
    cnts = struct.unpack(f"<{n}I", raw[d0[325][2]: d0[325][2] + 4 * n])
    assert n == 99 and len(set(offs)) < 45                    # 99 tiles, the flat ones aliased to two streams
    assert all(raw[o:o + 2] == b"\xff\xd8" and raw[o + c - 2:o + c] == b"\xff\xd9" for o, c in zip(offs, cnts))
    flat = [(j, i) for j in range(9) for i in range(11)
            if m[j * 256:(j + 1) * 256, i * 256:(i + 1) * 256].min() == m[j * 256:(j + 1) * 256, i * 256:(i + 1) * 256].max()]
    assert len({offs[j * 11 + i] for j, i in flat}) == 2
    im = Image.open(path)
    back = np.asarray(im)
    for j, i in flat:
        assert np.array_equal(back[j * 256:(j + 1) * 256, i * 256:(i + 1) * 256],
                              m[j * 256:(j + 1) * 256, i * 256:(i + 1) * 256])
    assert ((back > 127) == (m > 127)).mean() > 0.9999
    im.seek(4)
    assert np.asarray(im).shape == (131, 162)
    assert len(raw) < 200_000
