"""Host-side logic and oracle known-answer tests (CPU): grid / clamp / drop_last quirks, stitch arithmetic."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from digipathai_b200.slide import ArraySlide, synthetic_slide
from digipathai_b200.tissue import TileGrid, saturation, threshold_otsu, tissue_mask
from oracle import pipeline_ref
from standin import StandInModel


class ConstModel:
    def __init__(self, p):
        self.p = np.float32(p)

    def predict(self, x, **kw):
        out = np.empty(x.shape[:3] + (2,), np.float32)
        out[..., 1] = self.p
        out[..., 0] = np.float32(1) - self.p
        return out


def full_mask(slide):
    W, H = slide.level_dimensions[-1]
    return np.ones((W, H), np.uint8)


def test_clamped_origins_duplicate_the_first_tile():
    # Q5: grid points 0 and 1 on each axis both map to origin 0 -> that tile is predicted (and counted) twice
    s = ArraySlide(np.full((512, 512, 3), 200, np.uint8))
    g = TileGrid(s, 256, 128, 1, mask=full_mask(s))
    xs = sorted(set(g.all_coords[:, 0].tolist()))
    assert xs == [0, 128, 256]
    assert len(g.all_coords) == 16 and (g.all_coords[:, 0] == 0).sum() == 8
    _, pm = pipeline_ref.get_prediction(s, batch_size=1, models={"m": ConstModel(0.5)}, patch_size=256,
                                        stride_size=128, mask=full_mask(s), normalise=False)
    cnt = pm["count"]
    assert cnt[0, 0] == 4 and cnt[200, 200] == 9 and cnt[300, 300] == 4 and cnt[511, 511] == 1 and cnt[0, 511] == 2
    assert np.allclose(pm["mean"], 0.5 * cnt)


def test_drop_last_and_count_wrap():
    s = ArraySlide(np.full((512, 512, 3), 200, np.uint8))
    g = TileGrid(s, 256, 128, 5, mask=full_mask(s))
    assert g.n_batches == 3 and len(g.coords) == 15                      # Q3: the 16th tile is never predicted
    # Q6: uint8 count wraps at 256 overlaps, zero count divides by one
    s2 = ArraySlide(np.full((64, 64, 3), 200, np.uint8))
    m = np.ones((64, 64), np.uint8)
    _, pm = pipeline_ref.get_prediction(s2, batch_size=1, models={"m": ConstModel(0.25)}, patch_size=64,
                                        stride_size=1, mask=m, normalise=False)
    assert int(pm["count"][0, 0]) == (64 * 64) % 256


def test_variance_is_population_variance_over_passes():
    s = ArraySlide(np.full((512, 512, 3), 200, np.uint8))
    models = {"a": ConstModel(0.2), "b": ConstModel(0.6)}
    _, pm = pipeline_ref.get_prediction(s, batch_size=1, models=models, patch_size=256, stride_size=256,
                                        mask=full_mask(s))
    # grid points {0, 256} -> origins {0, 128} per axis: pixel (0,0) is covered once, (200,200) four times
    assert np.allclose(pm["mean"][:384, :384], 0.4)
    assert np.isclose(pm["var"][0, 0], 0.04) and np.isclose(pm["var"][200, 200], 0.04 * 4 / 16)   # var / count**2
    assert pm["mean"][500, 500] == 0 and pm["count"][500, 500] == 1                              # zero count -> 1


def test_level_above_four_raises_like_the_reference():
    s = synthetic_slide(2048, 2048, 0, n_levels=6)
    with pytest.raises(ValueError):
        TileGrid(s, 256, 128, 4)


def test_otsu_and_saturation_restatements_agree_with_oracle():
    s = synthetic_slide(512, 384, 1)
    rgb = np.transpose(s.raster, (1, 0, 2))
    for c in range(3):
        assert threshold_otsu(rgb[..., c]) == pipeline_ref._otsu(rgb[..., c])
    sat = saturation(rgb)
    assert np.array_equal(sat, pipeline_ref._rgb2hsv_s(rgb))
    assert threshold_otsu(sat) == pipeline_ref._otsu(sat)
    assert np.array_equal(tissue_mask(s, 0), pipeline_ref.TissueMaskGenerationOS(s, 0))
    # two-level known answer: bimodal uint8 image splits between the modes
    img = np.array([[10] * 50 + [200] * 50], np.uint8)
    assert 10 <= threshold_otsu(img) < 200


@settings(max_examples=25, deadline=None)
@given(w=st.integers(2, 6), h=st.integers(2, 6), stride_pow=st.integers(5, 7), batch=st.integers(1, 5))
def test_grid_properties(w, h, stride_pow, batch):
    P, stride = 128, 2 ** stride_pow
    W, H = w * 128, h * 128
    s = ArraySlide(np.full((H, W, 3), 180, np.uint8))
    m = np.ones((W, H), np.uint8)
    g = TileGrid(s, P, stride, batch, mask=m)
    # every origin is in range, ordered x-major, and the analytic overlap count equals a brute-force count
    c = g.all_coords
    assert c[:, 0].min() >= 0 and c[:, 0].max() <= W - P and c[:, 1].max() <= H - P
    assert np.all(np.diff(g.X_idcs) >= 0)
    assert len(g.coords) == (len(c) // batch) * batch
    cnt = np.zeros((W, H), np.int64)
    for x, y in g.coords:
        cnt[x:x + P, y:y + P] += 1
    ds = pipeline_ref.WSIStridedPatchDataset(s, P, True, stride, True, mask=m)
    assert len(ds) == len(c)
    if stride <= P and len(g.coords) == len(c):
        # the reference covers [0, last_grid_point + P/2) per axis -- the far margin beyond that stays
        # unpredicted (count 0 -> divided by 1 -> probability 0), a property the drop-in keeps
        gx, gy = (W - 1) // stride * stride, (H - 1) // stride * stride
        assert cnt[:min(W, gx + P // 2), :min(H, gy + P // 2)].min() >= 1


def test_array_slide_read_region_and_levels():
    r = np.arange(64 * 48 * 3, dtype=np.uint8).reshape(48, 64, 3)
    s = ArraySlide(r, n_levels=3)
    assert s.level_dimensions == ((64, 48), (32, 24), (16, 12))
    assert np.array_equal(s.read_region((8, 4), 0, (16, 8)), r[4:12, 8:24])
    assert np.array_equal(s.read_region((8, 4), 1, (4, 4)), r[::2, ::2][2:6, 4:8])
    edge = s.read_region((60, 44), 0, (8, 8))
    assert np.array_equal(edge[:4, :4], r[44:, 60:]) and edge[4:, :].sum() == 0


def test_grid_raw_mask_round_trips_and_the_processed_mask_does_not():
    """``mask=`` takes a RAW mask (the dataset applies its morphology to any mask it is given,
    dataloader.py:256-270): rebuilding from ``raw_mask`` reproduces the grid, from ``mask`` it dilates twice.
    The multi-GPU path (dist.local_part) therefore hands the grid object over, never the processed mask."""
    from digipathai_b200.slide import synthetic_slide
    s = synthetic_slide(4096, 3072, seed=0, n_levels=3)
    g = TileGrid(s, 256, 128, 8)
    again = TileGrid(s, 256, 128, 8, mask=g.raw_mask)
    assert np.array_equal(again.coords, g.coords) and np.array_equal(again.mask, g.mask)
    twice = TileGrid(s, 256, 128, 8, mask=g.mask)
    assert twice.mask.sum() > g.mask.sum()
    assert (g.stride_size, g.patch_size, g.batch_size) == (128, 256, 8)


def test_exact_histograms_across_float32_chunks():
    """_hist_u8 feeds OpenCV's float32 histogram at most 2**24 pixels at a time: a constant 4200 x 4200 plane
    (17.6 M pixels in one bin, beyond float32's exact integers) must still count exactly."""
    from digipathai_b200.tissue import _hist_u8
    a = np.full((4200, 4200), 7, np.uint8)
    a[0, :3] = (0, 255, 255)
    h = _hist_u8(a)
    assert h.dtype == np.int64 and h[7] == 4200 * 4200 - 3 and h[0] == 1 and h[255] == 2 and h.sum() == a.size
    rng = np.random.default_rng(0)
    u, v = rng.integers(0, 256, (2, 37, 53)).astype(np.uint8)
    assert np.array_equal(_hist_u8(u, v), np.bincount((u.astype(np.int64) * 256 + v).ravel(), minlength=65536))
    assert threshold_otsu(u) == threshold_otsu(u.astype(np.int64))       # uint8 fast path == generic integer path


def test_mask_from_slide_luma_is_pil_L():
    from PIL import Image
    from digipathai_b200.tissue import mask_from_slide
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (40, 56, 3)).astype(np.uint8)
    img[:20] = rng.integers(0, 6, (20, 56, 3))                       # many values whose luma rounds to 0 or 1
    want = np.asarray(Image.fromarray(img).convert("L")).T
    got = mask_from_slide(ArraySlide(img), 0)
    assert got.shape == (56, 40) and set(np.unique(got)) <= {0, 255}
    assert np.array_equal(got > 0, want > 0) and (want == 0).any()
