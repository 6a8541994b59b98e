"""GPU parity of the reference-named CRF helpers (digipathai_b200/crf.py: post_process_crf, do_crf --
DigiPathAI/helpers/utils.py:548-603) against the mean-field oracles: ``method='exact'`` against oracle/crf_ref.py
(exact Gaussian filters), the default ``method='lattice'`` against oracle/lattice_ref.py (permutohedral lattice, what
pydensecrf evaluates).  Labels must be identical wherever the oracle's marginal is further than 5e-3 from the decision
boundary."""
import numpy as np
import pytest

from test_crf_oracle import _blob_mask

pytestmark = pytest.mark.gpu


def test_post_process_crf_matches_oracle():
    from digipathai_b200.crf import post_process_crf
    from oracle import crf_ref
    mask, img = _blob_mask(40, 48, (0, 1), seed=3)
    rng = np.random.default_rng(0)
    p1 = np.clip(np.where(mask == 1, 0.65, 0.25) + 0.3 * rng.standard_normal(mask.shape), 0, 1).astype(np.float32)
    probs = np.stack([1.0 - p1, p1], axis=-1)[None]              # [1, h, w, 2]: the reference squeezes
    from oracle import lattice_ref
    for method, ref in (("exact", crf_ref), ("lattice", lattice_ref)):
        got = post_process_crf(img, probs, 2, method=method)
        want, q = ref.dense_crf(img, p1)
        assert got.shape == mask.shape and got.dtype == np.int64
        sure = np.abs(q - 0.5) > 5e-3
        assert np.array_equal(got[sure], want[sure]), method
        assert (want != (p1 > 0.5)).sum() > 0
    with pytest.raises(ValueError):
        post_process_crf(img, probs, 3)


@pytest.mark.parametrize("values,zero_unsure", [((0, 255), False), ((0, 100, 255), True), ((0, 255), True)])
@pytest.mark.parametrize("enable_color", [False, True])
def test_do_crf_matches_oracle(values, zero_unsure, enable_color):
    from digipathai_b200.crf import do_crf
    from oracle import crf_ref
    mask, img = _blob_mask(32, 36, values, seed=7)
    got = do_crf(img if enable_color else None, mask, 2, enable_color=enable_color, zero_unsure=zero_unsure,
                 method="exact")
    want, Q = crf_ref.do_crf(img, mask, 2, enable_color, zero_unsure)
    assert got.shape == mask.shape
    sure = np.abs(Q[:, 1] - 0.5).reshape(mask.shape) > 5e-3
    assert sure.mean() > 0.9
    assert np.array_equal(got[sure], want[sure])
    assert set(np.unique(got)) <= set(values)
