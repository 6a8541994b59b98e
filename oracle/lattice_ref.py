"""Permutohedral-lattice Gaussian filter and the mean-field inference pydensecrf runs on it -- TEST INFRASTRUCTURE.

``post_process_crf`` (DigiPathAI/helpers/utils.py:568-603) and ``do_crf`` (utils.py:548-566) delegate to pydensecrf,
which is neither vendored nor installed here (setup.py:32, unpinned).  pydensecrf wraps Kraehenbuehl's densecrf
library, whose high-dimensional Gaussian filter is the permutohedral lattice of Adams, Baek & Davis ("Fast
High-Dimensional Filtering Using the Permutohedral Lattice", 2010).  This module restates that PUBLISHED algorithm --
densecrf's ``permutohedral.cpp`` (``Permutohedral::init`` / ``seqCompute``) and ``densecrf.cpp``
(``DenseKernel::initLattice`` / ``filter``, ``DenseCRF::inference``) -- from the papers and the library's public
source as recalled; it could not be checked against a pydensecrf binary: PARITY UNPINNED, said in DESIGN.md too.

  init(feature [N, d])
      scale_j = sqrt(2/3) (d+1) / sqrt((j+1)(j+2));   elevate f -> E in the hyperplane sum = 0 of R^(d+1)
      nearest remainder-0 point:  rem0_i = round(E_i / (d+1)) (d+1);  rank_i = order of E_i - rem0_i (descending),
      corrected by the coordinate sum so that the point lands in its simplex;  barycentric weights b_0..b_d
      simplex vertex r has key rem0 + canonical[r][rank]  (canonical[r][j] = r for j <= d-r else r-(d+1))
      lattice points are the distinct keys (hash table);  blur neighbours of a point along axis j: key -+ 1 in every
      coordinate, +-(d+1) in coordinate j
  compute(v)      splat  val[p] += b * v_i  over the d+1 vertices of every input point
                  blur   for j = 0..d:  val'[p] = val[p] + 0.5 (val[n1_j(p)] + val[n2_j(p)])     (missing neighbour = 0)
                  slice  out_i = alpha * sum_r b_r val[vertex_r(i)],   alpha = 1 / (1 + 2^-d)
  DenseKernel (NORMALIZE_SYMMETRIC)   norm = 1 / sqrt(compute(1) + 1e-20);  filter(Q) = norm * compute(norm * Q)
  inference(n)    Q = softmax(-U);  n times:  Q = softmax(-U + sum_k w_k filter_k(Q))             (Potts, weight w_k)
"""
from __future__ import annotations

import numpy as np


class Permutohedral:
    def __init__(self, feature: np.ndarray):
        f = np.asarray(feature, dtype=np.float32)
        N, d = f.shape
        self.N, self.d = N, d
        d1 = d + 1
        scale = np.array([1.0 / np.sqrt((j + 2.0) * (j + 1.0)) for j in range(d)], np.float32) * np.float32(np.sqrt(2.0 / 3.0) * d1)
        canonical = np.zeros((d1, d1), np.int32)
        for r in range(d1):
            canonical[r, : d1 - r] = r
            canonical[r, d1 - r:] = r - d1
        # ---- elevate (float32, as the library)
        E = np.zeros((N, d1), np.float32)
        sm = np.zeros(N, np.float32)
        for j in range(d, 0, -1):
            cf = f[:, j - 1] * scale[j - 1]
            E[:, j] = sm - np.float32(j) * cf
            sm = sm + cf
        E[:, 0] = sm
        # ---- closest remainder-0 point
        down = np.float32(1.0 / d1)
        rd = np.round(down * E).astype(np.int32)          # C round(): half away from zero; np.round is half-to-even --
        half = (np.abs(down * E - np.trunc(down * E)) == 0.5)
        rd = np.where(half, (np.trunc(down * E) + np.sign(E)).astype(np.int32), rd)   # ... corrected here
        rem0 = (rd * d1).astype(np.float32)
        ssum = rd.sum(1)
        # ---- rank of the differential (descending)
        diff = E - rem0
        rank = np.zeros((N, d1), np.int32)
        for i in range(d):
            for j in range(i + 1, d1):
                lt = diff[:, i] < diff[:, j]
                rank[:, i] += lt
                rank[:, j] += ~lt
        # ---- correct for points whose nearest remainder-0 point is off the hyperplane
        for i in range(d1):
            pos = (ssum > 0) & (rank[:, i] >= d1 - ssum)
            neg = (ssum < 0) & (rank[:, i] < -ssum)
            rem0[:, i] = np.where(pos, rem0[:, i] - d1, np.where(neg, rem0[:, i] + d1, rem0[:, i]))
            rank[:, i] = np.where(pos, rank[:, i] + ssum - d1, np.where(neg, rank[:, i] + d1 + ssum, rank[:, i] + ssum))
        # ---- barycentric coordinates
        bary = np.zeros((N, d + 2), np.float32)
        idx = np.arange(N)
        for i in range(d1):
            v = (E[:, i] - rem0[:, i]) * down
            np.add.at(bary, (idx, d - rank[:, i]), v)
            np.add.at(bary, (idx, d1 - rank[:, i]), -v)
        bary[:, 0] += np.float32(1.0) + bary[:, d1]
        self.bary = bary[:, :d1].copy()
        # ---- lattice points (hash table)
        table = {}
        keys = []
        offset = np.zeros((N, d1), np.int64)
        rem0i = rem0.astype(np.int32)
        for r in range(d1):
            key = rem0i[:, :d] + canonical[r][rank[:, :d]]
            for n in range(N):
                k = tuple(int(x) for x in key[n])
                o = table.get(k)
                if o is None:
                    o = len(keys)
                    table[k] = o
                    keys.append(k)
                offset[n, r] = o
        self.offset = offset
        M = self.M = len(keys)
        karr = np.array(keys, np.int64).reshape(M, d)
        # ---- blur neighbours (index M = "missing", reads zero)
        n1 = np.full((d1, M), M, np.int64)
        n2 = np.full((d1, M), M, np.int64)
        for j in range(d1):
            a, b = karr - 1, karr + 1
            if j < d:
                a[:, j] = karr[:, j] + d
                b[:, j] = karr[:, j] - d
            for i in range(M):
                n1[j, i] = table.get(tuple(int(x) for x in a[i]), M)
                n2[j, i] = table.get(tuple(int(x) for x in b[i]), M)
        self.n1, self.n2 = n1, n2

    def compute(self, v: np.ndarray) -> np.ndarray:
        """v float [N, c] -> filtered float32 [N, c]."""
        v = np.asarray(v, np.float32)
        N, d, M = self.N, self.d, self.M
        c = v.shape[1]
        val = np.zeros((M + 1, c), np.float64)          # the library sums in float32; float64 here bounds its order noise
        for r in range(d + 1):
            np.add.at(val, self.offset[:, r], self.bary[:, r:r + 1].astype(np.float64) * v)
        val = val.astype(np.float32)
        for j in range(d + 1):
            new = val.copy()
            new[:M] = val[:M] + np.float32(0.5) * (val[self.n1[j]] + val[self.n2[j]])
            new[M] = 0
            val = new
        alpha = np.float32(1.0 / (1.0 + 2.0 ** (-d)))
        out = np.zeros((N, c), np.float32)
        for r in range(d + 1):
            out += self.bary[:, r:r + 1] * val[self.offset[:, r]] * alpha
        return out


class DenseKernel:
    """DIAG_KERNEL features (already divided by their standard deviations), NORMALIZE_SYMMETRIC."""

    def __init__(self, feature: np.ndarray):
        self.lattice = Permutohedral(feature)
        ones = np.ones((feature.shape[0], 1), np.float32)
        self.norm = (1.0 / np.sqrt(self.lattice.compute(ones)[:, 0] + np.float32(1e-20))).astype(np.float32)

    def filter(self, q: np.ndarray) -> np.ndarray:
        return self.norm[:, None] * self.lattice.compute(self.norm[:, None] * q.astype(np.float32))


def _softmax(t):
    e = np.exp(t - t.max(1, keepdims=True))
    return (e / e.sum(1, keepdims=True)).astype(np.float32)


def mean_field(img_hw3, U, n_iter, sdims_gauss, compat_gauss, sdims_bil, schan_bil, compat_bil):
    """Marginals float32 [h*w, n_labels]; ``U`` energies [h*w, n_labels]; ``compat_bil=None`` drops the bilateral term.
    Feature order follows pydensecrf.utils.create_pairwise_gaussian / create_pairwise_bilateral (coordinates first:
    the image's axis 0, then axis 1; then the colour channels)."""
    h, w = img_hw3.shape[:2]
    yy, xx = np.mgrid[0:h, 0:w]
    coords = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float32)
    rgb = img_hw3.reshape(-1, 3).astype(np.float32)
    U = np.asarray(U, np.float32)
    kernels = [(DenseKernel(coords / np.float32(sdims_gauss)), np.float32(compat_gauss))]
    if compat_bil is not None:
        kernels.append((DenseKernel(np.concatenate([coords / np.float32(sdims_bil), rgb / np.float32(schan_bil)], 1)),
                        np.float32(compat_bil)))
    Q = _softmax(-U)
    for _ in range(n_iter):
        t = -U
        for k, wgt in kernels:
            t = t + wgt * k.filter(Q)
        Q = _softmax(t)
    return Q


def dense_crf(image_hw3, p1_hw, n_iter=10, sdims_gauss=10.0, compat_gauss=3.0, sdims_bil=50.0, schan_bil=20.0,
              compat_bil=10.0):
    """``post_process_crf`` on the lattice: image uint8 [h, w, 3], p1 [h, w] -> (labels [h, w], q1 float32 [h, w])."""
    h, w = p1_hw.shape
    p1 = np.asarray(p1_hw, np.float32).ravel()
    p = np.stack([np.float32(1.0) - p1, p1], 1)
    U = -np.log(np.clip(p, np.float32(1e-5), np.float32(1.0)))
    Q = mean_field(np.asarray(image_hw3), U, n_iter, sdims_gauss, compat_gauss, sdims_bil, schan_bil, compat_bil)
    return Q.argmax(1).reshape(h, w), Q[:, 1].reshape(h, w)
