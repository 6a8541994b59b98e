"""Keras-layer -> weight-dict mapping of tools/h5_to_npz.py, without h5py: synthetic layer
dicts shaped like a Keras .h5 (unnamed layers numbered from an arbitrary offset, scale=False BNs without gamma)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import h5_to_npz  # noqa: E402


def _as_keras_layers(weights, shapes, offset, unnamed_conv, unnamed_bn, bias_sep):
    layers, ci, bi = {}, 0, 0
    for name, kind, _ in shapes:
        if kind == "bias":
            continue
        if kind in ("conv", "dwconv"):
            lname = name
            if name in unnamed_conv:
                lname = f"conv2d_{offset + ci}"
                ci += 1
            wd = {f"{lname}/{'depthwise_kernel' if kind == 'dwconv' else 'kernel'}:0": weights[name]}
            if name + bias_sep in weights:
                wd[f"{lname}/bias:0"] = weights[name + bias_sep]
            layers[lname] = wd
        else:
            lname = name
            if name in unnamed_bn:
                lname = f"batch_normalization_{offset + bi}"
                bi += 1
            g, b, m, v = weights[name]
            wd = {f"{lname}/beta:0": b, f"{lname}/moving_mean:0": m, f"{lname}/moving_variance:0": v}
            if kind != "bn_noscale":
                wd[f"{lname}/gamma:0"] = g
            layers[lname] = wd
    return dict(reversed(list(layers.items())))     # file order must not matter


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if isinstance(a[k], tuple):
            assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k])), k
        else:
            assert np.array_equal(a[k], b[k]), k


def test_inception_positional_mapping_survives_a_shifted_counter():
    from digipathai_b200.models import inception as IN
    w, shapes = IN.init_inception_weights(1), IN.layer_shapes()
    uc = {n for n, k, _ in shapes if k == "conv" and h5_to_npz._suffix(n)}
    ub = {n for n, k, _ in shapes if k.startswith("bn") and h5_to_npz._suffix(n)}
    layers = _as_keras_layers(w, shapes, 131, uc, ub, "/bias")      # e.g. saved after a DenseNet was built
    _same(h5_to_npz.map_layers("inception", layers), w)


def test_densenet_and_deeplab_mapping():
    from digipathai_b200.models import deeplab as DL
    from digipathai_b200.models import densenet as DN
    w, shapes = DN.init_densenet_weights(2), DN.layer_shapes()
    uc = {n + "_conv" for n, _, _ in DN.DECODER} | {"head"}
    ub = {n + "_norm" for n, _, _ in DN.DECODER}
    # the DenseNet dict keeps biases under "<name>_bias" (dec6a_conv_bias, head_bias)
    layers = _as_keras_layers(w, [s for s in shapes], 7, uc, ub, "_bias")
    _same(h5_to_npz.map_layers("dense", layers), w)
    wd, sd = DL.init_deeplab_weights(3), DL.layer_shapes()
    _same(h5_to_npz.map_layers("deeplabv3", _as_keras_layers(wd, sd, 1, set(), set(), "/bias")), wd)


def test_wrong_file_is_rejected():
    from digipathai_b200.models import densenet as DN
    w, shapes = DN.init_densenet_weights(2), DN.layer_shapes()
    uc = {n + "_conv" for n, _, _ in DN.DECODER} | {"head"}
    ub = {n + "_norm" for n, _, _ in DN.DECODER}
    layers = _as_keras_layers(w, shapes, 7, uc, ub, "_bias")
    layers.pop("conv2d_9")
    with pytest.raises(ValueError, match="unnamed layers"):
        h5_to_npz.map_layers("dense", layers)
    with pytest.raises(ValueError):
        h5_to_npz.map_layers("inception", layers)
