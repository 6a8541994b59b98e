"""Keras ``.h5`` checkpoint -> the weight dict the graph builders take (row a2: ``load_trained_models`` of
DigiPathAI/helpers/utils.py:427-448 calls ``model.load_weights(path)`` on the files of utils.py:58-98).

``read_keras_weights`` (h5lite.py, pure Python: no h5py in this image) yields {layer name: {weight name: array}};
``map_layers`` turns that into the flat dict of models/*.py.  Keras names the reference's unnamed layers
``conv2d_<k>`` / ``batch_normalization_<k>`` with a process-global counter, so the numeric suffixes in a file depend on
what was built before it in the saving process; what is stable is their ORDER.  ``map_layers`` therefore sorts the
unnamed layers by suffix and assigns them positionally to the creation-order names the graph builders use; explicitly
named layers map by name.  Pure dict-in / dict-out, unit-tested in tests/test_h5_mapping.py and tests/test_h5lite.py.
"""
from __future__ import annotations

import re

import numpy as np


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



def load_keras_h5(model: str, path: str) -> dict:
    """``model`` in {'dense', 'inception', 'deeplabv3'}; ``path`` a Keras weight file -> weight dict."""
    from .h5lite import read_keras_weights
    return map_layers(model, read_keras_weights(path))
