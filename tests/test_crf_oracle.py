"""Properties of the exact DenseCRF mean-field oracle (oracle/crf_ref.py, model of DigiPathAI/helpers/utils.py:568-603)."""
import numpy as np

from oracle import crf_ref


def _img(h, w, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    p = rng.uniform(0, 1, (h, w)).astype(np.float32)
    return img, p


def test_zero_iterations_is_the_unary_argmax_and_clip():
    img, p = _img(6, 7)
    p[0, 0], p[0, 1] = 0.0, 1.0
    lab, q = crf_ref.dense_crf(img, p, n_iter=0)
    assert np.array_equal(lab, (p > 0.5).astype(int))
    # unary_from_softmax(clip=1e-5): a probability of exactly 0 becomes 1e-5 / (1 + 1e-5)
    assert abs(q[0, 0] - 1e-5 / (1.0 + 1e-5)) < 1e-9 and abs(q[0, 1] - 1.0 / (1.0 + 1e-5)) < 1e-9


def test_two_pixel_known_answer():
    # two pixels, identical colour: every kernel matrix is [[1, k], [k, 1]] with k = exp(-d^2 / 2); symmetric
    # normalisation divides by (1 + k); one mean-field step can be written out by hand
    img = np.full((1, 2, 3), 100, np.uint8)
    p = np.array([[0.9, 0.4]], np.float32)
    lab, q = crf_ref.dense_crf(img, p, n_iter=1)
    U = -np.log(np.clip(np.stack([1 - p.ravel(), p.ravel()], 1), 1e-5, 1)).astype(np.float32).astype(np.float64)
    Q0 = np.exp(-U) / np.exp(-U).sum(1, keepdims=True)
    t = -U.copy()
    for s, wgt in ((10.0, 3.0), (50.0, 10.0)):
        k = np.exp(-0.5 / s ** 2)
        K = np.array([[1, k], [k, 1]]) / (1 + k)
        t += wgt * K @ Q0
    want = np.exp(t[:, 1]) / np.exp(t).sum(1)
    assert np.allclose(q.ravel(), want, atol=1e-12)


def test_transpose_equivariance_and_smoothing():
    img, p = _img(9, 12, 3)
    lab, q = crf_ref.dense_crf(img, p, n_iter=4)
    lab_t, q_t = crf_ref.dense_crf(np.ascontiguousarray(img.transpose(1, 0, 2)), np.ascontiguousarray(p.T), n_iter=4)
    assert np.allclose(q_t.T, q, atol=1e-12) and np.array_equal(lab_t.T, lab)
    # uniform colour + a noisy majority: the CRF pulls the minority pixels over
    flat = np.full((10, 10, 3), 128, np.uint8)
    rng = np.random.default_rng(1)
    noisy = np.where(rng.uniform(size=(10, 10)) < 0.2, 0.35, 0.8).astype(np.float32)
    lab2, _ = crf_ref.dense_crf(flat, noisy)
    assert lab2.all()


# ---- do_crf (DigiPathAI/helpers/utils.py:548-566) ----

def test_unary_from_labels_known_answer():
    lab = np.array([[0, 1], [2, 1]])
    U = crf_ref.unary_from_labels(lab, 2, 0.7, zero_unsure=True)
    lo, hi, unsure = np.float32(-np.log(0.7)), np.float32(-np.log(0.3)), np.float32(-np.log(0.5))
    assert U.dtype == np.float32 and U.shape == (4, 2)
    assert np.array_equal(U, np.array([[unsure, unsure], [lo, hi], [hi, lo], [lo, hi]], np.float32))
    U = crf_ref.unary_from_labels(np.array([0, 1, 1]), 2, 0.7, zero_unsure=False)
    assert np.array_equal(U, np.array([[lo, hi], [hi, lo], [hi, lo]], np.float32))


def _blob_mask(h, w, values, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    inside = ((yy - h * 0.5) / (h * 0.3)) ** 2 + ((xx - w * 0.45) / (w * 0.3)) ** 2 < 1
    idx = np.where(inside, len(values) - 1, len(values) - 2)
    flip = rng.uniform(size=(h, w)) < 0.15
    idx = np.where(flip, rng.integers(0, len(values), (h, w)), idx)
    img = np.where(inside[..., None], (170, 90, 160), (235, 235, 238)) + rng.normal(0, 10, (h, w, 3))
    return np.asarray(values)[idx], np.clip(img, 0, 255).astype(np.uint8)


def test_do_crf_prior_mapping_is_the_label_unary():
    """The product hands ``dp_crf_tiles`` a label-1 probability instead of unary_from_labels' energy table
    (digipathai_b200/crf.py:_prior_from_mask): same marginals, for every flag combination."""
    from digipathai_b200.crf import _prior_from_mask
    for values, zero_unsure in (((0, 255), False), ((0, 255), True), ((0, 100, 255), True)):
        for enable_color in (False, True):
            mask, img = _blob_mask(14, 17, values, seed=len(values) + zero_unsure)
            want_map, want_q = crf_ref.do_crf(img, mask, 2, enable_color, zero_unsure)
            colors, p1 = _prior_from_mask(mask, 2, zero_unsure)
            assert np.array_equal(colors, values)
            lab, q1 = crf_ref.dense_crf(img, p1, n_iter=5, sdims_gauss=3.0, compat_gauss=3.0, sdims_bil=80.0,
                                        schan_bil=13.0, compat_bil=10.0 if enable_color else 0.0)
            assert np.abs(q1.ravel() - want_q[:, 1]).max() < 1e-6
            sure = np.abs(want_q[:, 1] - 0.5).reshape(mask.shape) > 1e-5
            got = lab.copy()
            for u in np.unique(got):
                got[got == u] = colors[u]
            assert np.array_equal(got[sure], want_map[sure])


def test_do_crf_smooths_and_keeps_the_reference_quirks():
    mask, img = _blob_mask(24, 24, (0, 255), seed=9)
    out, _ = crf_ref.do_crf(img, mask, 2, enable_color=False, zero_unsure=False)
    assert set(np.unique(out)) <= {0, 255}
    assert 0 < (out != mask).sum() < mask.size // 3          # speckle removed, blob kept
    assert (out == 255).sum() > mask.size // 10
    # zero_unsure=True on a binary mask: value 0 is "unsure", value 255 is class 0 -> everything maps to colors[0]
    out, _ = crf_ref.do_crf(img, mask, 2, enable_color=False, zero_unsure=True)
    assert (out == 0).all()


def test_do_crf_errors():
    import pytest
    from digipathai_b200.crf import _prior_from_mask
    with pytest.raises(ValueError):
        _prior_from_mask(np.zeros((4, 4), np.uint8), 3, True)
    with pytest.raises(IndexError):       # three classes into two labels: the reference's fancy index overflows too
        _prior_from_mask(np.arange(16).reshape(4, 4) % 3, 2, False)
    with pytest.raises(IndexError):
        crf_ref.unary_from_labels(np.arange(16) % 3, 2, 0.7, zero_unsure=False)
