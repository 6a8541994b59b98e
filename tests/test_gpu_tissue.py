"""Device tissue mask + morphology (csrc/tissue.cuh through dp_tissue_hist / dp_tissue_mask / dp_morph_rect) --
bit-exact against the reference's own cv2 morphology calls (tests/golden/morph_golden.npz, produced by executing
BinMorphoProcessMaskOS, DigiPathAI/helpers/utils.py:200-219), against the host evaluation of
TissueMaskGenerationOS (utils.py:336-354) in tissue.py, and -- end to end -- the tile grid built on the device equals
the tile grid built on the host for the same slide."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


def test_device_morphology_matches_the_reference_cv2_calls():
    import torch
    from digipathai_b200.tissue import morpho_process_device
    z = np.load(os.path.join(G, "morph_golden.npz"))
    shape = tuple(z["shape"])
    m = np.unpackbits(z["mask"])[: shape[0] * shape[1]].reshape(shape)
    for lvl in (0, 3, 4):
        want = np.unpackbits(z[f"out_{lvl}"])[: shape[0] * shape[1]].reshape(shape)
        got = morpho_process_device(torch.from_numpy(m.copy()).cuda(), lvl).cpu().numpy()
        assert np.array_equal(got, want), lvl
    with pytest.raises(ValueError):
        morpho_process_device(torch.from_numpy(m.copy()).cuda(), 5)


@pytest.mark.parametrize("k", [1, 2, 5, 10, 20, 35, 60])
def test_device_rect_dilate_erode_match_cv2_on_ragged_masks(k):
    """Odd and even kernels, maps smaller than the kernel, non-square maps, values other than {0, 1}."""
    import ctypes as C
    import cv2
    import torch
    from digipathai_b200 import _lib
    rng = np.random.default_rng(k)
    for shape in ((7, 13), (64, 33), (200, 257)):
        m = (rng.random(shape) < 0.08).astype(np.uint8) * rng.integers(1, 255, shape).astype(np.uint8)
        d = torch.from_numpy(m).cuda()
        for dil, fn in ((1, cv2.dilate), (0, cv2.erode)):
            out, tmp = torch.empty_like(d), torch.empty_like(d)
            _lib.check(_lib.lib.dp_morph_rect(C.c_void_p(d.data_ptr()), C.c_void_p(out.data_ptr()),
                                              C.c_void_p(tmp.data_ptr()), shape[0], shape[1], k, dil, C.c_void_p(0)))
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), fn(m, np.ones((k, k), np.uint8))), (shape, k, dil)


def test_device_tissue_mask_equals_host_evaluation():
    import torch
    from digipathai_b200.slide import synthetic_slide
    from digipathai_b200.tissue import tissue_mask, tissue_mask_device
    for seed, (w, h) in ((0, (700, 500)), (3, (1250, 1250))):
        s = synthetic_slide(w, h, seed=seed, n_levels=1)
        want = tissue_mask(s, 0)                                           # bool [x, y]
        rgb_xy = torch.from_numpy(np.ascontiguousarray(np.transpose(s.raster, (1, 0, 2)))).cuda()
        got = tissue_mask_device(rgb_xy).cpu().numpy().astype(bool)
        assert got.shape == want.shape and np.array_equal(got, want)
    # uniform-noise image: every (max, max-min) pair occurs, thresholds land mid-range
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (300, 400, 3)).astype(np.uint8)

    class _S:
        level_dimensions = ((400, 300),)
        def read_region(self, loc, lvl, size):
            return img
    want = tissue_mask(_S(), 0)
    got = tissue_mask_device(torch.from_numpy(np.ascontiguousarray(np.transpose(img, (1, 0, 2)))).cuda()).cpu().numpy()
    assert np.array_equal(got.astype(bool), want)


def test_tile_grid_built_on_the_device_equals_the_host_grid():
    import torch
    from digipathai_b200.slide import ArraySlide, DeviceSlide, synthetic_slide
    from digipathai_b200.tissue import TileGrid
    host = synthetic_slide(4096, 3072, seed=2, n_levels=3)
    dev = DeviceSlide(torch.from_numpy(np.ascontiguousarray(np.transpose(host.raster, (1, 0, 2)))).cuda(), 3)
    a = TileGrid(host, 256, 128, 8)
    b = TileGrid(dev, 256, 128, 8, device=0)
    assert np.array_equal(a.coords, b.coords) and a.n_batches == b.n_batches and len(a.coords) > 0
    assert np.array_equal(np.uint8(a.mask), np.uint8(b.mask)) and np.array_equal(a.raw_mask, b.raw_mask)
    assert np.array_equal(a.X_idcs, b.X_idcs) and np.array_equal(a.Y_idcs, b.Y_idcs)
