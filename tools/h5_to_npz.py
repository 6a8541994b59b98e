#!/usr/bin/env python
"""Convert the reference's Keras ``.h5`` weight files (DigiPathAI/helpers/utils.py:58-98 downloads them,
:427-448 loads them with ``model.load_weights``) into the flat ``.npz`` that digipathai_b200 reads
(SURVEY.md 8(f) N4).  Needs ``h5py`` -- not available in the build image, so run it wherever the reference runs:

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


def _classify(wdict):
    """{'kernel:0': a, ...} (weight-name suffix -> array) -> ('conv'|'dwconv'|'bn', payload)."""
    keys = {k.split("/")[-1].split(":")[0]: v for k, v in wdict.items()}
    if "depthwise_kernel" in keys:
        return "dwconv", {"kernel": np.asarray(keys["depthwise_kernel"], np.float32)}
    if "kernel" in keys:
        out = {"kernel": np.asarray(keys["kernel"], np.float32)}
        if "bias" in keys:
            out["bias"] = np.asarray(keys["bias"], np.float32)
        return "conv", out
    if "moving_mean" in keys:
        mean = np.asarray(keys["moving_mean"], np.float32)
        gamma = np.asarray(keys["gamma"], np.float32) if "gamma" in keys else np.ones_like(mean)   # scale=False
        beta = np.asarray(keys["beta"], np.float32) if "beta" in keys else np.zeros_like(mean)
        return "bn", (gamma, beta, mean, np.asarray(keys["moving_variance"], np.float32))
    return None, None


def _suffix(name):
    m = re.fullmatch(r"(conv2d|batch_normalization)(?:_(\d+))?", name)
    return (m.group(1), int(m.group(2) or 0)) if m else None


def map_layers(model: str, layers: dict) -> dict:
    """layers: Keras layer name -> {weight name -> array}.  Returns the weight dict of the given graph builder."""
    if model == "dense":
        from digipathai_b200.models.densenet import DECODER, layer_shapes
        unnamed_conv = [n + "_conv" for n, _, _ in DECODER] + ["head"]
        unnamed_bn = [n + "_norm" for n, _, _ in DECODER]
        bias_name = lambda n: n + "_bias"
    elif model == "inception":
        from digipathai_b200.models.inception import layer_shapes
        shapes = layer_shapes()
        unnamed_conv = [n for n, k, _ in shapes if k == "conv" and _suffix(n)]
        unnamed_bn = [n for n, k, _ in shapes if k.startswith("bn") and _suffix(n)]
        bias_name = lambda n: n + "/bias"
    elif model == "deeplabv3":
        from digipathai_b200.models.deeplab import layer_shapes
        unnamed_conv, unnamed_bn = [], []
        bias_name = lambda n: n + "/bias"
    else:
        raise ValueError(model)
    want = {n: (k, s) for n, k, s in layer_shapes()}
    out, convs, bns = {}, [], []
    for name, wd in layers.items():
        kind, payload = _classify(wd)
        if kind is None:
            continue
        sfx = _suffix(name)
        if sfx and (unnamed_conv or unnamed_bn):
            (convs if kind == "conv" else bns).append((sfx[1], payload))
            continue
        if kind == "bn":
            out[name] = payload
        else:
            out[name] = payload["kernel"]
            if "bias" in payload:
                out[bias_name(name)] = payload["bias"]
    for pool, names in ((convs, unnamed_conv), (bns, unnamed_bn)):
        pool.sort(key=lambda t: t[0])
        if len(pool) != len(names):
            raise ValueError(f"{model}: file has {len(pool)} unnamed layers of this kind, the graph has {len(names)}")
        for (_, payload), n in zip(pool, names):
            if isinstance(payload, tuple):
                out[n] = payload
            else:
                out[n] = payload["kernel"]
                if "bias" in payload:
                    out[bias_name(n)] = payload["bias"]
    missing = [n for n in want if n not in out]
    if missing:
        raise ValueError(f"{model}: {len(missing)} tensors missing after mapping, e.g. {missing[:5]}")
    for n, (k, s) in want.items():
        got = out[n][0].shape if isinstance(out[n], tuple) else out[n].shape
        if tuple(got) != tuple(s):
            raise ValueError(f"{model}: tensor {n} has shape {tuple(got)}, the graph expects {tuple(s)}")
    return {n: out[n] for n in want}


def read_h5(path: str) -> dict:
    import h5py
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
