"""Container parsing for the device ingest (digipathai_b200/wsi_tiff.py): tiles / strips, JPEGTables
splicing, pyramid levels, read_region -- checked against Pillow's own reading of the same files."""
import io

import numpy as np
import pytest
from PIL import Image

from digipathai_b200 import tiffio
from digipathai_b200.wsi_tiff import TiffSlide


def _rgb(h, w):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([(x * 0.5) % 256, (y * 0.8) % 256, ((x + y) * 0.3) % 256], -1).astype(np.uint8)


def test_stripped_rgb_jpeg_tiff_with_jpegtables(tmp_path):
    img = _rgb(300, 500)
    path = str(tmp_path / "s.tiff")
    Image.fromarray(img).save(path, compression="jpeg", quality=90)
    want = np.asarray(Image.open(path).convert("RGB"))
    s = TiffSlide(path)
    p = s.pages[0]
    assert s.level_count == 1 and s.dimensions == (500, 300) and not p.tiled and p.tables and p.compression == 7
    assert s.device_decodable(0)
    tx, ty, tw, th = s.tile_grid(0)
    assert (tx, tw) == (1, 500) and ty * th >= 300
    # every strip, made self-contained, decodes to exactly what Pillow's TIFF reader shows for those rows
    for k in range(ty):
        stream = s.jpeg_stream(0, k)
        assert stream[:2] == b"\xff\xd8" and stream[-2:] == b"\xff\xd9"
        a = np.asarray(Image.open(io.BytesIO(stream)).convert("RGB"))
        x0, y0 = s.tile_origin(0, k)
        assert x0 == 0 and np.array_equal(a, want[y0:y0 + a.shape[0]])
    assert np.array_equal(s.read_region((0, 0), 0, (500, 300)), want)
    reg = s.read_region((-7, 250), 0, (40, 80))                     # hangs over the left and bottom edges
    assert not reg[:, :7].any() and not reg[50:].any() and np.array_equal(reg[:50, 7:], want[250:300, :33])
    s.close()


def test_own_pyramidal_result_plane_round_trips(tmp_path):
    y, x = np.mgrid[0:900, 0:1300]
    a = (127.5 + 127.5 * np.sin(x / 37.0) * np.cos(y / 53.0)).astype(np.float32)
    a[:300, :600] = 0                                                # flat tiles share one stream
    path = str(tmp_path / "p.tiff")
    assert tiffio.save_pyramidal(path, a) == 4
    s = TiffSlide(path)
    assert s.level_count == 4 and s.level_dimensions[0] == (1300, 900) and s.level_dimensions[2] == (325, 225)
    assert s.level_downsamples == (1.0, 2.0, 4.0, 8.0)
    assert s.tile_grid(0) == (6, 4, 256, 256) and s.pages[0].tiled and s.pages[0].tables is None
    im = Image.open(path)
    for lvl in (0, 2):
        im.seek(lvl)
        want = np.asarray(im.convert("RGB"))
        w, h = s.level_dimensions[lvl]
        assert np.array_equal(s.read_region((0, 0), lvl, (w, h)), want)
    im.seek(1)
    want1 = np.asarray(im.convert("RGB"))
    assert np.array_equal(s.read_region((512, 256), 1, (100, 60)), want1[128:188, 256:356])   # level-0 coordinates
    s.close()


def test_uncompressed_and_bigtiff(tmp_path):
    img = _rgb(70, 90)
    path = str(tmp_path / "raw.tiff")
    Image.fromarray(img).save(path)
    s = TiffSlide(path)
    assert not s.device_decodable(0) and np.array_equal(s.read_region((0, 0), 0, (90, 70)), img)
    with pytest.raises(ValueError):
        s.jpeg_stream(0, 0)
    s.close()
    big = str(tmp_path / "big.tiff")
    try:
        Image.fromarray(img).save(big, big_tiff=True)
    except Exception:                                                 # noqa: BLE001 -- older Pillow: no BigTIFF writer
        pytest.skip("this Pillow cannot write BigTIFF")
    assert open(big, "rb").read(4) in (b"II+\0", b"MM\0+")
    s = TiffSlide(big)
    assert np.array_equal(s.read_region((0, 0), 0, (90, 70)), img)
    s.close()


def test_not_a_tiff(tmp_path):
    p = tmp_path / "x.tiff"
    p.write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(ValueError):
        TiffSlide(str(p))


def test_open_slide_opt_in_and_tile_grid_on_a_tiff_slide(tmp_path, monkeypatch):
    from digipathai_b200.slide import ArraySlide, open_slide
    from digipathai_b200.tissue import TileGrid
    img = np.full((600, 800, 3), 240, np.uint8)
    img[150:450, 200:650] = (170, 90, 160)
    img = np.clip(img.astype(np.int16) + np.random.default_rng(0).integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    path = str(tmp_path / "slide.tiff")
    Image.fromarray(img).save(path, compression="jpeg", quality=90)
    monkeypatch.delenv("DIGIPATH_DEVICE_INGEST", raising=False)
    host = open_slide(path)
    assert isinstance(host, ArraySlide)                       # default: decoded on the host through Pillow
    monkeypatch.setenv("DIGIPATH_DEVICE_INGEST", "1")
    dev = open_slide(path)
    assert isinstance(dev, TiffSlide) and dev.device_decodable(0)
    # the host-side prologue (tissue mask + tile grid) sees the same pixels either way
    g_host, g_dev = TileGrid(host, 128, 64, 4), TileGrid(dev, 128, 64, 4)
    assert len(g_host.coords) > 8 and np.array_equal(g_host.coords, g_dev.coords)
    raw = str(tmp_path / "raw.tiff")
    Image.fromarray(img).save(raw)                            # not JPEG: stays on the Pillow path even when opted in
    assert isinstance(open_slide(raw), ArraySlide)
    dev.close()
