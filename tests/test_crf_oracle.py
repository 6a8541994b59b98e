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
