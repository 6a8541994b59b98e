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
    yy, xx = np.mgrid[0:h, 0:w]
    coords = np.stack([yy.ravel(), xx.ravel()], axis=1).astype(np.float64)
    rgb = image_hw3.reshape(-1, 3).astype(np.float64)
    p = np.stack([1.0 - p1_hw.ravel(), p1_hw.ravel()], axis=1).astype(np.float64)
    U = -np.log(np.clip(p, 1e-5, 1.0)).astype(np.float32).astype(np.float64)   # pydensecrf stores the unary as float32
    kernels = []
    for feats, wgt in ((coords / sdims_gauss, compat_gauss),
                       (np.concatenate([coords / sdims_bil, rgb / schan_bil], axis=1), compat_bil)):
        K = _kernel(feats)
        norm = 1.0 / np.sqrt(K.sum(1) + 1e-20)
        kernels.append((K, norm, wgt))

    def softmax(t):
        t = t - t.max(1, keepdims=True)
        e = np.exp(t)
        return e / e.sum(1, keepdims=True)

    Q = softmax(-U)
    for _ in range(n_iter):
        t = -U
        for K, norm, wgt in kernels:
            t = t + wgt * (norm[:, None] * (K @ (norm[:, None] * Q)))
        Q = softmax(t)
    return Q.argmax(1).reshape(h, w), Q[:, 1].reshape(h, w)
