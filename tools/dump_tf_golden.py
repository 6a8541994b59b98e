#!/usr/bin/env python
"""Golden-vector dumper for real-model parity (second half).  Runs WHERE THE REFERENCE RUNS
(TensorFlow 1.x + Keras + the DigiPathAI package importable) -- none of that exists in the build image, so this
script has never been executed there; it only calls the reference's own public entry points:

    load_trained_models(model, path, patch_size)     DigiPathAI/helpers/utils.py:427-448
    Model.predict(image_patches, batch_size=..., verbose=0, steps=None)   as at DigiPathAI/Segmentation.py:154-156

Usage (one file per checkpoint; keep them small -- a handful of tiles):

    python tools/dump_tf_golden.py dense ~/.DigiPathAI/digestpath_models/digestpath_densenet.h5 \\
           tests/golden/tf_dense_digestpath.npz [--tiles 4] [--seed 0] [--slide slide.tiff --xy 10240,20480 ...]

The output holds the uint8 tiles (synthetic H&E-like noise, or crops of a real slide), the reference's float32
[n, 256, 256, 2] softmax, and provenance strings.  On the B200 side convert the same ``.h5`` with tools/h5_to_npz.py
and drop both next to each other; tests/test_gpu_tf_golden.py picks up every ``tests/golden/tf_<model>_<set>.npz``
that has a matching ``<set>_<model file>.npz`` weight file and checks the CUDA path against it (BASELINE tolerance
discussion: DESIGN.md section 5).
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def synthetic_tiles(n: int, seed: int) -> np.ndarray:
    """uint8 [n, 256, 256, 3]: pink tissue-like blobs on a bright background (SURVEY.md 8(d) config 1 recipe)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:256, 0:256]
    out = np.empty((n, 256, 256, 3), np.uint8)
    for i in range(n):
        cy, cx = rng.uniform(64, 192, 2)
        ry, rx = rng.uniform(40, 110, 2)
        inside = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        img = 240.0 + 3.0 * rng.standard_normal((256, 256, 3))
        tis = np.array([170.0, 90.0, 160.0]) + 20.0 * rng.standard_normal((256, 256, 3))
        img[inside] = tis[inside]
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", choices=["dense", "inception", "deeplabv3"])
    ap.add_argument("h5")
    ap.add_argument("out")
    ap.add_argument("--tiles", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--slide", default=None, help="crop the tiles from this slide (openslide) instead of synthesising")
    ap.add_argument("--xy", nargs="*", default=[], help="level-0 tile origins 'x,y' for --slide")
    args = ap.parse_args()

    if args.slide:
        import openslide
        s = openslide.OpenSlide(args.slide)
        tiles = np.stack([np.asarray(s.read_region(tuple(int(v) for v in xy.split(",")), 0, (256, 256)).convert("RGB"))
                          for xy in args.xy])
    else:
        tiles = synthetic_tiles(args.tiles, args.seed)
    # the dataset's tile orientation and normalisation (loaders/dataloader.py:357-358,383-388)
    x = (np.transpose(tiles, (0, 2, 1, 3)).astype(np.float32) - 128.0) / 128.0

    from DigiPathAI.helpers.utils import load_trained_models        # the reference itself
    model = load_trained_models(args.model, args.h5, 256)
    probs = model.predict(x, batch_size=len(x), verbose=0, steps=None)
    import tensorflow as tf
    import keras
    np.savez_compressed(args.out, tiles=tiles, probs=np.asarray(probs, np.float32), model=args.model,
                        h5=args.h5, tf_version=tf.__version__, keras_version=keras.__version__,
                        note="tiles are in image layout [n, y, x, c]; the network saw them transposed to [x, y, c] "
                             "and normalised (v - 128) / 128; probs are in the network's [x, y] orientation")
    print("wrote", args.out, probs.shape, float(probs[..., 1].mean()))


if __name__ == "__main__":
    sys.exit(main())
