"""Exact mean-field inference of the fully connected CRF that ``post_process_crf`` configures
(DigiPathAI/helpers/utils.py:568-603; its call site in getSegmentation is commented out, Segmentation.py:327-331).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference delegates to pydensecrf (not vendored, not installed
here; setup.py:32 leaves it unpinned), whose public algorithm (Kraehenbuehl & Koltun, "Efficient Inference in
Fully Connected CRFs with Gaussian Edge Potentials") is restated here from its published source:

  * unary_from_softmax(sm, scale=None, clip=1e-5):  U = -log(clip(p, 1e-5, 1))                  (utils.py:574)
  * create_pairwise_gaussian(sdims=(10,10)): features = pixel coordinates / sdims               (utils.py:584)
  * create_pairwise_bilateral(sdims=(50,50), schan=(20,20,20)): (coordinates / sdims, rgb / schan) (utils.py:593)
  * addPairwiseEnergy(feats, compat=w, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC): Potts label
    compatibility; kernel k(f_i, f_j) = exp(-|f_i - f_j|^2 / 2) (the centre included); symmetric normalisation
    K~ = D^-1/2 K D^-1/2 with D = diag(K 1) (+1e-20)                                             (utils.py:586-600)
  * inference(10): Q <- softmax(-U); repeat 10 x: Q <- softmax(-U + sum_k w_k K~_k Q)            (utils.py:601)
  * MAP = argmax_l Q                                                                             (utils.py:602)

pydensecrf evaluates K~ Q with a permutohedral-lattice approximation of the Gaussian filter; this oracle (and the
CUDA kernels) evaluate the filter EXACTLY, i.e. the model the lattice approximates.  "Parity unpinned" against
pydensecrf's lattice output, and stated so in DESIGN.md.
"""
from __future__ import annotations

import numpy as np


def _kernel(feats: np.ndarray) -> np.ndarray:
    d2 = ((feats[:, None, :] - feats[None, :, :]) ** 2).sum(-1)
    return np.exp(-0.5 * d2)


def dense_crf(image_hw3: np.ndarray, p1_hw: np.ndarray, n_iter: int = 10, sdims_gauss=10.0, compat_gauss=3.0,
              sdims_bil=50.0, schan_bil=20.0, compat_bil=10.0):
    """image uint8 [h,w,3], p1 float [h,w] (probability of label 1) -> (labels int64 [h,w], q1 float64 [h,w])."""
    h, w = p1_hw.shape
    p = np.stack([1.0 - p1_hw.ravel(), p1_hw.ravel()], axis=1).astype(np.float64)
    U = -np.log(np.clip(p, 1e-5, 1.0))
    Q = _mean_field(image_hw3, U, n_iter, sdims_gauss, compat_gauss, sdims_bil, schan_bil, compat_bil)
    return Q.argmax(1).reshape(h, w), Q[:, 1].reshape(h, w)


def _mean_field(img, U, n_iter, sdims_gauss, compat_gauss, sdims_bil, schan_bil, compat_bil):
    """Marginals float64 [h*w, n_labels] after ``n_iter`` mean-field updates; ``U`` energies [h*w, n_labels]
    (stored as float32, as pydensecrf does); ``compat_bil=None`` leaves the bilateral term out."""
    h, w = img.shape[:2]
    yy, xx = np.mgrid[0:h, 0:w]
    coords = np.stack([yy.ravel(), xx.ravel()], axis=1).astype(np.float64)
    rgb = img.reshape(-1, 3).astype(np.float64)
    U = np.asarray(U).astype(np.float32).astype(np.float64)
    terms = [(coords / sdims_gauss, compat_gauss)]
    if compat_bil is not None:
        terms.append((np.concatenate([coords / sdims_bil, rgb / schan_bil], axis=1), compat_bil))
    kernels = []
    for feats, wgt in terms:
        K = _kernel(feats)
        kernels.append((K, 1.0 / np.sqrt(K.sum(1) + 1e-20), wgt))

    def softmax(t):
        e = np.exp(t - t.max(1, keepdims=True))
        return e / e.sum(1, keepdims=True)

    Q = softmax(-U)
    for _ in range(n_iter):
        t = -U
        for K, norm, wgt in kernels:
            t = t + wgt * (norm[:, None] * (K @ (norm[:, None] * Q)))
        Q = softmax(t)
    return Q


def unary_from_labels(labels: np.ndarray, n_labels: int, gt_prob: float, zero_unsure: bool = True) -> np.ndarray:
    """pydensecrf.utils.unary_from_labels restated (called at DigiPathAI/helpers/utils.py:553): energies
    float32 [n_pixels, n_labels].  The labelled class costs -log(gt_prob), every other class
    -log((1 - gt_prob) / (n_labels - 1)).  With ``zero_unsure`` label 0 means "unknown" (uniform energy
    -log(1 / n_labels)) and label k > 0 names class k - 1."""
    lab = np.asarray(labels).ravel().astype(np.int64)
    U = np.full((lab.size, n_labels), -np.log((1.0 - gt_prob) / (n_labels - 1)), dtype=np.float32)
    idx = np.arange(lab.size)
    if zero_unsure:
        known = lab > 0
        U[idx[known], lab[known] - 1] = -np.log(gt_prob)
        U[~known, :] = -np.log(1.0 / n_labels)
    else:
        U[idx, lab] = -np.log(gt_prob)
    return U


def do_crf(im, mask: np.ndarray, n_labels: int, enable_color: bool = False, zero_unsure: bool = True):
    """``do_crf`` (DigiPathAI/helpers/utils.py:548-566; never called by the reference) with exact filters:
    unary from the hard labels (gt_prob 0.7), Gaussian sxy 3 compat 3, optional bilateral sxy 80 srgb 13 compat 10,
    5 mean-field iterations, MAP index mapped back through the sorted distinct mask values one index after the
    other (utils.py:563-565: a later index can re-map pixels an earlier one already rewrote).
    Returns (MAP [h, w] in mask values, marginals float64 [h*w, n_labels])."""
    colors, labels = np.unique(mask, return_inverse=True)
    h, w = mask.shape[:2]
    U = unary_from_labels(labels, n_labels, 0.7, zero_unsure)
    img = np.zeros((h, w, 3), np.uint8) if im is None else np.asarray(im).astype(np.uint8)
    Q = _mean_field(img, U, 5, 3.0, 3.0, 80.0, 13.0, 10.0 if enable_color else None)
    MAP = Q.argmax(1).reshape(h, w)
    for u in np.unique(MAP):
        MAP[MAP == u] = colors[u]
    return MAP, Q
