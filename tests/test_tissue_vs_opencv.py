"""Pieces of the tissue mask (TissueMaskGenerationOS, DigiPathAI/helpers/utils.py:336-354) that are restated from
scikit-image -- which is absent here -- against OpenCV's independent implementations of the same definitions: Otsu's
threshold on the 8-bit R, G, B channels (three of the mask's four thresholds) and the HSV saturation plane."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _images():
    rng = np.random.default_rng(0)
    for trial in range(120):
        kind = trial % 4
        if kind == 0:
            a = rng.integers(0, 256, (64, 64))
        elif kind == 1:                                   # bimodal, like tissue vs background
            a = np.where(rng.random((80, 80)) < 0.4, rng.normal(90, 20, (80, 80)), rng.normal(200, 15, (80, 80)))
        elif kind == 2:                                   # mostly background with a small blob
            a = rng.normal(240, 3, (50, 70))
            a[10:30, 10:40] = rng.normal(150, 25, (20, 30))
        else:
            lo, hi = sorted(rng.integers(0, 256, 2))
            a = rng.integers(lo, hi + 1, (40, 40))
        a = np.clip(np.rint(a), 0, 255).astype(np.uint8)
        if a.min() != a.max():
            yield a


def test_otsu_on_8bit_channels_selects_the_same_pixels_as_opencv():
    from digipathai_b200 import tissue
    from oracle import pipeline_ref
    n = 0
    for a in _images():
        t_cv, _ = cv2.threshold(a, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        want = a > t_cv
        assert np.array_equal(a > tissue.threshold_otsu(a), want)
        assert np.array_equal(a > pipeline_ref._otsu(a), want)
        n += 1
    assert n >= 100


def test_saturation_plane_is_opencvs_hsv_saturation():
    from digipathai_b200 import tissue
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (96, 80, 3)).astype(np.uint8)
    img[:8] = 0                                           # black: saturation 0 by convention (max == 0)
    img[8:16] = 200                                       # grey: saturation 0
    s_cv = cv2.cvtColor(img.astype(np.float32) / 255.0, cv2.COLOR_RGB2HSV)[..., 1]
    s = tissue.saturation(img)
    from oracle import pipeline_ref
    assert np.abs(pipeline_ref._rgb2hsv_s(img) - s_cv).max() <= 5e-6
    assert s.shape == s_cv.shape
    assert np.abs(s - s_cv).max() <= 5e-6                 # OpenCV computes in float32
