"""Test-time-augmentation bookkeeping: the reference's four transforms as elements of the dihedral group D4.

Reference: ``apply_tta`` / ``transform_prob`` in DigiPathAI/helpers/utils.py:487-522, driven by the loop at
DigiPathAI/Segmentation.py:150-158.  Two reference quirks are reproduced on purpose:

* the forward transforms are applied IN PLACE and therefore accumulate: pass ``k`` feeds the network
  ``T_k(T_{k-1}(...T_1(x)))`` but only ``T_k`` is inverted on the prediction;
* any string other than FLIP_LEFT_RIGHT / ROTATE_90 / ROTATE_180 / ROTATE_270 is an identity pass that still
  counts in the mean.

A transform ``T`` is stored as the 3-bit code of its *source map* ``out[i, j] = in[src_T(i, j)]``
(bit0 swap, bit1 mirror first coordinate, bit2 mirror second) -- the same encoding the CUDA kernels use
(``csrc/d4.cuh``).
"""
from __future__ import annotations

import numpy as np

IDENTITY = 0
CODES = {
    "DEFAULT": 0,
    "FLIP_LEFT_RIGHT": 4,   # np.fliplr: out[i, j] = in[i, P-1-j]
    "ROTATE_90": 5,         # np.rot90(k=1): out[i, j] = in[j, P-1-i]
    "ROTATE_180": 6,        # out[i, j] = in[P-1-i, P-1-j]
    "ROTATE_270": 3,        # np.rot90(k=3): out[i, j] = in[P-1-j, i]
}


def code_of(name: str) -> int:
    """Reference semantics: unknown names are identity passes (utils.py:490-499)."""
    return CODES.get(str(name), IDENTITY)


def src(code: int, i, j, P: int):
    u, v = (j, i) if code & 1 else (i, j)
    a = P - 1 - u if code & 2 else u
    b = P - 1 - v if code & 4 else v
    return a, b


def apply(code: int, arr: np.ndarray) -> np.ndarray:
    """out = T(arr) on axes (0, 1) of a square array."""
    P = arr.shape[0]
    ii, jj = np.meshgrid(np.arange(P), np.arange(P), indexing="ij")
    a, b = src(code, ii, jj, P)
    return arr[a, b]


def _table():
    P = 5
    base = np.arange(P * P).reshape(P, P)
    imgs = {c: apply(c, base) for c in range(8)}
    comp = np.zeros((8, 8), dtype=np.int64)  # comp[g2, g1] = code of "apply g1 first, then g2"
    inv = np.zeros(8, dtype=np.int64)
    for g1 in range(8):
        for g2 in range(8):
            out = apply(g2, imgs[g1])
            (hit,) = [c for c in range(8) if np.array_equal(imgs[c], out)]
            comp[g2, g1] = hit
            if hit == 0:
                inv[g1] = g2
    return comp, inv


_COMP, _INV = _table()


def compose(g2: int, g1: int) -> int:
    return int(_COMP[g2, g1])


def inverse(g: int) -> int:
    return int(_INV[g])


def pass_codes(tta_list):
    """[(tta_in, tta_out)] for every pass of one batch, reference order.

    ``tta_in`` is the cumulative transform the network input has undergone (Segmentation.py:151 mutates
    ``image_patches`` in place); ``tta_out`` is the transform that ``transform_prob`` undoes.
    The reference prepends 'DEFAULT' (Segmentation.py:95-99).
    """
    names = ["DEFAULT"] + ([] if tta_list is None else [str(t) for t in tta_list])
    out, cum = [], IDENTITY
    for nm in names:
        g = code_of(nm)
        cum = compose(g, cum)
        out.append((cum, g))
    return out
