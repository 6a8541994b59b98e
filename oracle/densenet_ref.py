"""fp32 PyTorch-CPU restatement of ``unet_densenet121`` (DigiPathAI/models/densenet.py:37-159).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Layer-for-layer what Keras would execute for the reference
graph, on NCHW float32 tensors; weights come in as the dict described in digipathai_b200/models/densenet.py
(Keras layer names, HWIO kernels, BN = (gamma, beta, moving_mean, moving_var)).

Keras semantics encoded here (documented behaviour, not verifiable against a TF binary in this image):
  * Conv2D padding='same', stride 1, 3x3  -> symmetric zero pad 1           (densenet.py:41,70)
  * ZeroPadding2D(3) + Conv2D(7, strides=2, 'valid')                        (densenet.py:116-117)
  * ZeroPadding2D(1) + MaxPooling2D(3, strides=2, 'valid')                  (densenet.py:122-123)
  * AveragePooling2D(2, strides=2)                                          (densenet.py:107)
  * BatchNormalization inference: gamma*(x-mean)/sqrt(var+eps)+beta, eps 1.001e-5 encoder / 1e-3 decoder
  * UpSampling2D() = 2x nearest                                             (densenet.py:138)
  * softmax over the channel axis                                           (densenet.py:156)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BLOCKS = (6, 12, 24, 16)
EPS_ENC = 1.001e-5
EPS_DEC = 1e-3


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _conv(x, k_hwio, stride=1, pad=0, bias=None):
    w = _t(k_hwio).permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
    return F.conv2d(x, w, None if bias is None else _t(bias), stride=stride, padding=pad)


class _BN:
    """Inference BN; in calibration mode it first overwrites the moving statistics with batch statistics."""

    def __init__(self, weights, calibrate):
        self.w, self.cal = weights, calibrate

    def __call__(self, x, name, eps):
        g, b, mu, var = self.w[name]
        if self.cal:
            mu = x.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
            var = x.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32)
            var = np.maximum(var, 1e-4)
            self.w[name] = (g, b, mu, var)
        sh = (1, -1, 1, 1)
        return (x - _t(mu).view(sh)) / torch.sqrt(_t(var).view(sh) + eps) * _t(g).view(sh) + _t(b).view(sh)


def forward(weights: dict, x_nhwc: np.ndarray, calibrate: bool = False, taps: dict | None = None,
            perturb: dict | None = None) -> np.ndarray:
    """float32 [B,P,P,3] in [-1,1] -> float32 [B,P,P,2] softmax, exactly ``Model.predict`` of the reference graph.

    ``taps`` (optional dict) receives intermediate NHWC activations by name for layer-level parity tests.
    ``perturb`` (optional dict name -> callable(NCHW tensor) -> tensor) replaces the activation recorded under that
    name before the graph continues (conditioning studies: tools/parity_study.py).
    """
    bn = _BN(weights, calibrate)
    rec = (lambda n, t: taps.__setitem__(n, t.permute(0, 2, 3, 1).contiguous().numpy())) if taps is not None else (lambda n, t: None)
    hook = (lambda n, t: perturb[n](t) if n in perturb else t) if perturb else (lambda n, t: t)
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2).contiguous()
        x = F.pad(x, (3, 3, 3, 3))
        x = _conv(x, weights["conv1/conv"], stride=2)
        x = F.relu(bn(x, "conv1/bn", EPS_ENC))
        x = hook("conv1", x)
        conv1 = x
        rec("conv1", x)
        x = F.max_pool2d(F.pad(x, (1, 1, 1, 1)), 3, stride=2)
        rec("pool1", x)
        skips = {}
        for b, n in zip((2, 3, 4, 5), BLOCKS):
            for i in range(1, n + 1):
                p = f"conv{b}_block{i}"
                x1 = F.relu(bn(x, p + "_0_bn", EPS_ENC))
                x1 = _conv(x1, weights[p + "_1_conv"])
                x1 = F.relu(bn(x1, p + "_1_bn", EPS_ENC))
                if i == 1:
                    rec(p + "_bottleneck", x1)
                x1 = _conv(x1, weights[p + "_2_conv"], pad=1)
                x = torch.cat([x, hook(p, x1)], dim=1)
            x = hook(f"conv{b}", x)
            skips[b] = x
            rec(f"conv{b}", x)
            if b < 5:
                x = F.relu(bn(x, f"pool{b}_bn", EPS_ENC))
                x = _conv(x, weights[f"pool{b}_conv"])
                x = F.avg_pool2d(x, 2, stride=2)
                rec(f"pool{b}", x)
        x = bn(x, "bn", EPS_ENC)
        rec("conv5", x)

        def block(x, name):
            x = _conv(x, weights[name + "_conv"], pad=1, bias=weights[name + "_conv_bias"])
            x = hook(name, F.relu(bn(x, name + "_norm", EPS_DEC)))
            rec(name, x)
            return x

        up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
        x = block(up(x), "dec6a")
        x = block(torch.cat([x, skips[4]], dim=1), "dec6b")
        x = block(up(x), "dec7a")
        x = block(torch.cat([x, skips[3]], dim=1), "dec7b")
        x = block(up(x), "dec8a")
        x = block(torch.cat([x, skips[2]], dim=1), "dec8b")
        x = block(up(x), "dec9a")
        x = block(torch.cat([x, conv1], dim=1), "dec9b")
        x = block(up(x), "dec10a")
        x = block(x, "dec10b")
        z = _conv(x, weights["head"], bias=weights["head_bias"])
        y = F.softmax(z, dim=1)
        return y.permute(0, 2, 3, 1).contiguous().numpy()


def calibrate_bn(weights: dict, calib_nhwc: np.ndarray) -> dict:
    """Replace every BN's moving statistics by the statistics of one oracle pass over ``calib_nhwc`` (in place).

    Stands in for the training-time moving averages the real ``.h5`` files carry: keeps activations O(1)
    through all 121 layers so that the softmax is informative and fp16 storage is exercised realistically.
    """
    forward(weights, calib_nhwc, calibrate=True)
    return weights


class OracleModel:
    """Minimal stand-in for the Keras ``Model`` the reference loop calls (Segmentation.py:154-156)."""

    def __init__(self, weights: dict):
        self.weights = weights

    def predict(self, image_patches, batch_size=None, verbose=0, steps=None):
        return forward(self.weights, np.asarray(image_patches, dtype=np.float32))
