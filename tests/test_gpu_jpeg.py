"""The device JPEG tile encoder of the pyramidal writer (csrc/jpeg_enc.cuh, row N1): every stream it emits must decode with
libjpeg (Pillow) to what libjpeg's own encoder at the same quality decodes to, within the difference between a float and
an integer DCT; constant tiles, ragged plane edges (border replication), the capacity fallback and the whole
``save_pyramidal`` file against the host path."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _decode(b):
    from PIL import Image
    im = Image.open(io.BytesIO(b))
    assert im.mode == "L" and im.size == (256, 256)
    return np.asarray(im).astype(np.int32)


def _planes():
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:700, 0:1000].astype(np.float32)
    smooth = 127 + 100 * np.sin(xx / 37.0) * np.cos(yy / 53.0) + rng.normal(0, 2, xx.shape)
    blobs = 255.0 * (((xx - 400) / 260) ** 2 + ((yy - 300) / 170) ** 2 < 1)          # mask-like: flat regions + an edge
    noise = rng.integers(0, 256, (300, 520)).astype(np.float32)                       # worst case for the entropy coder
    faint = np.clip(rng.normal(3, 1.5, (512, 512)), 0, 255)                            # uncertainty-plane-like
    return {"smooth": smooth, "blobs": blobs, "noise": noise, "faint": faint}


@pytest.mark.parametrize("name", ["smooth", "blobs", "noise", "faint"])
def test_device_streams_decode_like_libjpegs_own(name):
    import torch
    from digipathai_b200 import tiffio
    a = np.clip(np.rint(_planes()[name]), 0, 255).astype(np.uint8)
    got = tiffio.encode_level_device(torch.from_numpy(a).cuda(), 90)
    rows, cols = a.shape
    ty, tx = -(-rows // 256), -(-cols // 256)
    assert len(got) == ty * tx
    n_const = 0
    for idx, g in enumerate(got):
        j, i = divmod(idx, tx)
        t = a[j * 256:(j + 1) * 256, i * 256:(i + 1) * 256]
        t = np.pad(t, ((0, 256 - t.shape[0]), (0, 256 - t.shape[1])), mode="edge")
        if isinstance(g, int):
            assert t.min() == t.max() == g
            n_const += 1
            continue
        assert t.min() != t.max()
        mine, ref = _decode(g), _decode(tiffio._jpeg(np.ascontiguousarray(t), 90))
        d = np.abs(mine - ref)
        # same tables, same quantiser; the DCTs differ (float here, libjpeg's scaled-integer "islow"): a coefficient now
        # and then rounds to the neighbouring level
        assert d.mean() <= (0.6 if name == "noise" else 0.35) and d.max() <= 12, (name, idx, d.mean(), d.max())
        e_mine, e_ref = np.abs(mine - t).mean(), np.abs(ref - t).mean()
        assert e_mine <= e_ref * 1.03 + 0.02, (name, idx, e_mine, e_ref)
        assert len(g) <= len(tiffio._jpeg(np.ascontiguousarray(t), 90)) * 1.03 + 16      # and it is as compact
    if name == "blobs":
        assert n_const >= 3


def test_capacity_fallback_and_pyramidal_file(tmp_path):
    import torch
    from PIL import Image
    from digipathai_b200 import tiffio
    a = np.clip(np.rint(_planes()["noise"]), 0, 255).astype(np.uint8)
    lvl = torch.from_numpy(a).cuda()
    small = tiffio.encode_level_device(lvl, 90, scratch_per_tile=4096, out_cap=4096)      # every tile overflows -> host
    full = tiffio.encode_level_device(lvl, 90)
    for s, f in zip(small, full):
        assert np.abs(_decode(s) - _decode(f)).mean() <= 0.6
    # whole file: device-encoded vs host-encoded pyramid, page by page
    plane = np.clip(_planes()["smooth"], 0, 255).astype(np.float32)
    p_dev, p_host = str(tmp_path / "dev.tiff"), str(tmp_path / "host.tiff")
    n_dev = tiffio.save_pyramidal(p_dev, torch.from_numpy(plane).cuda())
    n_host = tiffio.save_pyramidal(p_host, plane)
    assert n_dev == n_host >= 3
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(p_dev) as d, Image.open(p_host) as h:
        assert d.n_frames == h.n_frames == n_dev
        for k in range(n_dev):
            d.seek(k); h.seek(k)
            x, y = np.asarray(d).astype(np.int32), np.asarray(h).astype(np.int32)
            assert x.shape == y.shape
            assert np.abs(x - y).mean() <= 0.5 and np.abs(x - y).max() <= 14, (k, np.abs(x - y).mean(), np.abs(x - y).max())
