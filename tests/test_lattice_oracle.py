"""CPU checks of the permutohedral-lattice restatement (oracle/lattice_ref.py) and of host-side helpers added in
round 2.  pydensecrf is not available, so the lattice is checked against what it approximates: the exact Gaussian
filter (oracle/crf_ref.py) -- a wrong scale factor, blur weight or slice constant shows up as a wrong effective
standard deviation or a low correlation."""
import numpy as np

from oracle import crf_ref, lattice_ref


def test_lattice_impulse_response_is_a_gaussian_of_the_requested_width():
    h = w = 40
    yy, xx = np.mgrid[0:h, 0:w]
    coords = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float32)
    for sd in (3.0, 6.0):
        L = lattice_ref.Permutohedral(coords / sd)
        v = np.zeros((h * w, 1), np.float32)
        v[(h // 2) * w + w // 2] = 1
        out = L.compute(v)[:, 0].reshape(h, w)
        row, x = out[h // 2], np.arange(w) - w // 2
        sigma = np.sqrt((row * x * x).sum() / row.sum())
        assert abs(sigma - sd) / sd < 0.12, (sd, sigma)                  # measured 3.12 / 6.41
        py, px = divmod(int(out.argmax()), w)                             # the lattice is not centred on pixels:
        tol = sd / 2 + 1                                                   # ... the peak sits within half a cell
        assert abs(py - h // 2) <= tol and abs(px - w // 2) <= tol and out.min() > -1e-6


def test_lattice_bilateral_filter_tracks_the_exact_filter_and_the_crf_labels_agree():
    rng = np.random.default_rng(0)
    h = w = 24
    img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    img[:, : w // 2] = (img[:, : w // 2] * 0.2 + 150).astype(np.uint8)
    img[:, w // 2:] = (img[:, w // 2:] * 0.2 + 40).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    coords = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float32)
    feats = np.concatenate([coords / 10.0, img.reshape(-1, 3).astype(np.float32) / 20.0], 1)
    K = crf_ref._kernel(feats.astype(np.float64))
    nrm = 1 / np.sqrt(K.sum(1) + 1e-20)
    q = rng.random((h * w, 2)).astype(np.float32)
    exact = nrm[:, None] * (K @ (nrm[:, None] * q))
    lat = lattice_ref.DenseKernel(feats).filter(q)
    assert np.corrcoef(exact.ravel(), lat.ravel())[0, 1] > 0.97
    assert abs(exact.mean() - lat.mean()) < 0.02
    p1 = rng.random((h, w)).astype(np.float32)
    a, b = lattice_ref.dense_crf(img, p1), crf_ref.dense_crf(img, p1)
    assert (a[0] == b[0]).mean() > 0.98 and np.abs(a[1] - b[1]).max() < 0.05


def test_lattice_filter_is_linear_and_symmetric_normalisation_is_applied_on_both_sides():
    rng = np.random.default_rng(1)
    feats = rng.random((200, 5)).astype(np.float32) * 4
    dk = lattice_ref.DenseKernel(feats)
    a, b = rng.random((200, 2)).astype(np.float32), rng.random((200, 2)).astype(np.float32)
    assert np.allclose(dk.filter(a + b), dk.filter(a) + dk.filter(b), atol=2e-5)
    ones = np.ones((200, 1), np.float32)
    assert np.allclose(dk.norm, 1 / np.sqrt(dk.lattice.compute(ones)[:, 0] + 1e-20), rtol=1e-6)


def test_saturation_lookup_table_is_the_literal_rgb2hsv_expression_for_every_max_min_pair():
    """the table must use (max/255 - min/255) / (max/255), the expression `saturation()` evaluates, not
    ((max - min)/255) / (max/255) -- they differ by one ulp for about half of the pairs."""
    from digipathai_b200.tissue import _saturation_lut, saturation
    lut = _saturation_lut()
    mx, mn = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    ok = mn <= mx
    img = np.stack([mx[ok], mn[ok], mn[ok]], -1).astype(np.uint8)[None]       # pixels with max = R, min = G = B
    want = saturation(img)[0]
    got = lut[mx[ok] * 256 + (mx[ok] - mn[ok])]
    assert np.array_equal(got, want)
