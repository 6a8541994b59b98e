#!/usr/bin/env python
"""Convert the reference's Keras ``.h5`` weight files (DigiPathAI/helpers/utils.py:58-98 downloads them,
:427-448 loads them with ``model.load_weights``) into the flat ``.npz`` that digipathai_b200 reads
(SURVEY.md 8(f) N4).  Uses ``h5py`` where it is installed and digipathai_b200/h5lite.py (pure Python) otherwise; ``load_trained_models`` reads ``.h5`` files directly, so the conversion is optional:

    python tools/h5_to_npz.py dense      ~/.DigiPathAI/digestpath_models/digestpath_densenet.h5   digestpath_densenet.npz
    python tools/h5_to_npz.py inception  .../digestpath_inception.h5                              digestpath_inception.npz
    python tools/h5_to_npz.py deeplabv3  .../digestpath_deeplabv3.h5                              digestpath_deeplabv3.npz

Keras names the reference's unnamed layers ``conv2d_<k>`` / ``batch_normalization_<k>`` with a process-global
counter, so the numeric suffixes in a file depend on what was built before it in the saving process; what is stable
is their ORDER.  ``map_layers`` therefore sorts the unnamed layers by suffix and assigns them positionally to the
creation-order names the graph builders use (models/densenet.py, models/inception.py); explicitly named layers map
by name.  ``map_layers`` is pure (dict in, dict out) and unit-tested without h5py (tests/test_h5_mapping.py).
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


from digipathai_b200.keras_h5 import _classify, _suffix, map_layers  # noqa: E402,F401  (the mapping lives in the package)


def read_h5(path: str) -> dict:
    """{layer name: {weight name: array}} -- through h5py where it exists, else the package's own reader."""
    try:
        import h5py
    except ImportError:
        from digipathai_b200.h5lite import read_keras_weights
        return read_keras_weights(path)
    layers = {}
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs["layer_names"]]
        for ln in names:
            wn = [n.decode() if isinstance(n, bytes) else n for n in g[ln].attrs["weight_names"]]
            if wn:
                layers[ln] = {n: np.asarray(g[ln][n]) for n in wn}
    return layers


def main():
    if len(sys.argv) != 4:
        raise SystemExit(__doc__)
    model, src, dst = sys.argv[1:]
    from digipathai_b200.Segmentation import save_npz
    save_npz(map_layers(model, read_h5(src)), dst)
    print(f"wrote {dst}")


if __name__ == "__main__":
    main()
