"""Pins the oracle (and the product's host-side grid logic) against golden vectors produced by executing the
reference's own get_prediction / WSIStridedPatchDataset / BinMorphoProcessMaskOS (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from digipathai_b200.slide import synthetic_slide
from digipathai_b200.tissue import TileGrid, morpho_process
from oracle import pipeline_ref
from standin import StandInModel

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(G, "pipeline_golden.npz"))
    cfg = json.loads(str(z["config"]))
    slide = synthetic_slide(cfg["width"], cfg["height"], cfg["seed"], cfg["n_levels"])
    return z, cfg, slide


def test_oracle_dataset_matches_reference_dataset(gold):
    z, cfg, slide = gold
    ds = pipeline_ref.WSIStridedPatchDataset(slide, cfg["patch"], True, cfg["stride"], True)
    assert np.array_equal(ds._X_idcs, z["X_idcs"]) and np.array_equal(ds._Y_idcs, z["Y_idcs"])
    mask = np.unpackbits(z["mask"])[: np.prod(z["mask_shape"])].reshape(z["mask_shape"]).astype(bool)
    assert np.array_equal(ds._mask > 0, mask)
    assert ds._resolution == z["resolution"] and ds._level == z["level"] and ds._sampling_stride == z["factor"]
    img, x, y = ds[0]
    assert (x, y) == tuple(z["tile0_xy"]) and np.array_equal(img, z["tile0"])
    img, x, y = ds[len(ds) - 1]
    assert (x, y) == tuple(z["tile_last_xy"]) and np.array_equal(img, z["tile_last"])


def test_oracle_get_prediction_is_bit_identical_to_the_reference_loop(gold):
    z, cfg, slide = gold
    models = {f"m{i}": StandInModel(*p) for i, p in enumerate(cfg["models"])}
    status = {}
    thr, mean, var = pipeline_ref.getSegmentation(slide, models, cfg["patch"], cfg["stride"], cfg["batch"],
                                                  cfg["tta"], status)
    assert mean.shape == tuple(z["mean_shape"])
    assert np.array_equal(mean[::8, ::8], z["mean_sub"]) and np.array_equal(var[::8, ::8], z["var_sub"])
    assert sha(mean) == str(z["mean_sha"])
    assert sha(var) == str(z["var_sha"])
    assert sha(thr) == str(z["thr_sha"])
    assert status["progress"] == int(z["progress_last"])      # ensemble progress quirk (Segmentation.py:139)


def test_product_tile_grid_matches_reference_dataset(gold):
    z, cfg, slide = gold
    grid = TileGrid(slide, cfg["patch"], cfg["stride"], cfg["batch"])
    assert np.array_equal(grid.X_idcs, z["X_idcs"]) and np.array_equal(grid.Y_idcs, z["Y_idcs"])
    assert tuple(grid.all_coords[0]) == tuple(z["tile0_xy"])
    assert tuple(grid.all_coords[-1]) == tuple(z["tile_last_xy"])
    assert grid.n_batches == len(z["X_idcs"]) // cfg["batch"]            # drop_last
    assert len(grid.coords) == grid.n_batches * cfg["batch"]


def test_morphology_matches_reference_cv2_calls():
    z = np.load(os.path.join(G, "morph_golden.npz"))
    shape = tuple(z["shape"])
    m = np.unpackbits(z["mask"])[: shape[0] * shape[1]].reshape(shape)
    for lvl in (0, 3, 4):
        want = np.unpackbits(z[f"out_{lvl}"])[: shape[0] * shape[1]].reshape(shape)
        assert np.array_equal(pipeline_ref.BinMorphoProcessMaskOS(m.copy(), lvl), want)
        assert np.array_equal(morpho_process(m.copy(), lvl), want)
    assert int(z["level5_raises"]) == 1
    with pytest.raises(ValueError):
        morpho_process(m.copy(), 5)
    with pytest.raises(ValueError):
        pipeline_ref.BinMorphoProcessMaskOS(m.copy(), 5)
