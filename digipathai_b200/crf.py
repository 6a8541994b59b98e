"""The reference's two fully connected CRF helpers on the B200 mean-field kernels (``dp_crf_tiles``).

Same names, argument meaning and return types as DigiPathAI/helpers/utils.py:548-603 (numpy in, numpy out), so a
caller of ``utils.post_process_crf`` / ``utils.do_crf`` can switch imports.  Default ``method='lattice'``: mean-field
inference with the Gaussian filters on the permutohedral lattice (``csrc/crf_lattice.cuh``), which is what the
reference gets from pydensecrf; ``method='exact'`` evaluates the same filters exactly (``csrc/crf.cuh``, all pairs,
~15x slower).  SURVEY.md 8(a) rows a13 / a13'; parity against a pydensecrf binary is unpinned.

The kernels hold two labels -- the path's masks are binary -- so ``num_cl`` / ``n_labels`` other than 2 raise.
These are tile-sized calls; ``getSegmentation(crf=True)`` tiles a slide itself.
"""
from __future__ import annotations

import numpy as np


def _run(image, p1, **kw):
    import torch
    from . import engine
    if not torch.cuda.is_available():
        raise RuntimeError("digipathai_b200.crf needs a CUDA device (there is no CPU fallback)")
    rgb = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8))[None].cuda()
    p = torch.from_numpy(np.ascontiguousarray(p1, dtype=np.float32))[None].cuda()
    with torch.cuda.device(p.device):
        lab = engine.dense_crf(rgb, p, **kw)
    return lab[0].cpu().numpy()


def post_process_crf(image, final_probabilities, num_cl, method="lattice"):
    """utils.py:568-603: ``image`` uint8 [h, w, 3], ``final_probabilities`` float [h, w, num_cl] (singleton axes
    squeezed away, as the reference does) -> int64 labels [h, w]."""
    if num_cl != 2:
        raise ValueError("the B200 CRF kernels hold two labels; num_cl must be 2")
    image = np.asarray(image)
    probs = np.asarray(final_probabilities).squeeze()
    if probs.ndim != 3 or probs.shape[2] != 2 or probs.shape[:2] != image.shape[:2]:
        raise ValueError(f"probabilities {probs.shape} do not match image {image.shape} with 2 classes")
    # two-label mean field only sees the unary difference U1 - U0 = log(clip(p0)) - log(clip(p1)); dp_crf_tiles
    # forms both energies from p1 and 1 - p1, which equals the reference's unary whenever the two channels sum to 1
    # (they are a softmax).
    return _run(image, probs[..., 1], method=method).astype(np.int64)


def _prior_from_mask(mask, n_labels, zero_unsure):
    """(sorted distinct mask values, float32 [h, w] label-1 probability standing for unary_from_labels' energies)."""
    if n_labels != 2:
        raise ValueError("the B200 CRF kernels hold two labels; n_labels must be 2")
    mask = np.asarray(mask)
    colors, labels = np.unique(mask, return_inverse=True)
    labels = labels.reshape(-1)
    cls = labels - 1 if zero_unsure else labels            # class index per pixel, -1 = unsure
    if cls.max() >= n_labels:
        raise IndexError(f"mask holds {len(colors)} distinct values: more classes than n_labels = {n_labels}")
    p1 = np.where(cls < 0, 0.5, np.where(cls == 1, 0.7, 0.3)).astype(np.float32)
    return colors, p1.reshape(mask.shape[:2])


def do_crf(im, mask, n_labels, enable_color=False, zero_unsure=True, method="lattice"):
    """utils.py:548-566 (never called by the reference): refine a hard label ``mask`` [h, w]; returns the MAP in
    the mask's own values.

    The unary of ``unary_from_labels(labels, n_labels, gt_prob=.7, zero_unsure)`` is a per-pixel two-entry energy
    table; mean-field inference depends only on softmax(-U), so it is handed to the kernels as the label-1
    probability: 0.7 / 0.3 for a labelled pixel, 0.5 for an "unsure" one (label 0 under ``zero_unsure``, where
    label k > 0 names class k - 1 -- quirk of the reference kept as is).  Pairwise terms: Gaussian sxy 3 compat 3,
    and with ``enable_color`` the bilateral sxy 80 srgb 13 compat 10; 5 iterations.
    """
    colors, p1 = _prior_from_mask(mask, n_labels, zero_unsure)
    h, w = p1.shape
    if enable_color:
        image = np.asarray(im).astype(np.uint8)
        compat_bilateral = 10.0
    else:
        image = np.zeros((h, w, 3), np.uint8)               # unused: the bilateral term carries weight 0
        compat_bilateral = 0.0
    MAP = _run(image, p1, n_iter=5, sdims_gauss=3.0, compat_gauss=3.0, sdims_bilateral=80.0, schan_bilateral=13.0,
               compat_bilateral=compat_bilateral, method=method).astype(np.int64)
    for u in np.unique(MAP):                                # index -> original value, one index after the other
        MAP[MAP == u] = colors[u]
    return MAP
