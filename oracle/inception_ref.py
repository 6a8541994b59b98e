"""fp32 PyTorch-CPU restatement of ``get_inception_resnet_v2_unet_softmax``
(DigiPathAI/models/inception.py:36-271).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Layer-for-layer what Keras would execute for the reference graph,
on NCHW float32 tensors.  Weight dict keys follow Keras' automatic naming in creation order, which is what the
reference's ``.h5`` files carry for its unnamed layers (``conv2d_bn(..., name=None)``, inception.py:50-87):
``conv2d_<i>`` (HWIO kernel), ``conv2d_<i>/bias`` (only where ``use_bias=True``), ``batch_normalization_<i>``
= (gamma, beta, moving_mean, moving_var); the explicitly named layers keep their names (``block35_3_conv``,
``conv_7b``, ``conv_7b_bn``).  ``layer_plan()`` is that creation order.

Keras semantics encoded here (documented behaviour, not verifiable against a TF binary in this image;
"parity unpinned" for them, as for the DenseNet graph):
  * every Conv2D of this graph is padding='same' (inception.py:40,72); TensorFlow pads
    total = max((ceil(n/s)-1)*s + k - n, 0), floor(total/2) in front, the rest behind -- for the stride-2 3x3
    convs on even maps that is 0 in front and 1 behind
  * MaxPooling2D(3, strides=2, padding='same'): same padding rule, padded cells never win
  * AveragePooling2D(3, strides=1, padding='same'): mean over the valid cells only
  * conv2d_bn: BatchNormalization(scale=False) (gamma == 1), Keras default epsilon 1e-3, only when use_bias=False
    (inception.py:76-83); conv_block (decoder): Conv2D with bias + BatchNormalization (gamma, eps 1e-3) + ReLU
  * block output: x + scale * up, then ReLU except for block8_10 (inception.py:152-160,232-236)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-3


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv_same(x, k_hwio, stride=1, bias=None):
    kh, kw = k_hwio.shape[:2]
    pt, pb = _same_pad(x.shape[2], kh, stride)
    pl, pr = _same_pad(x.shape[3], kw, stride)
    w = _t(k_hwio).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, None if bias is None else _t(bias), stride=stride)


def _maxpool_same(x):
    pt, pb = _same_pad(x.shape[2], 3, 2)
    pl, pr = _same_pad(x.shape[3], 3, 2)
    return F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), 3, stride=2)


class _Net:
    """Walks the reference graph once; the same walk serves the forward pass, BN calibration and (with
    ``x=None``) the enumeration of layer names/shapes in Keras creation order."""

    def __init__(self, weights, calibrate=False, plan=None):
        self.w, self.cal, self.plan = weights, calibrate, plan
        self.n_conv = self.n_bn = 0

    # inception.py:50-87
    def conv2d_bn(self, x, cin, filters, kernel, strides=1, activation="relu", use_bias=False, name=None):
        kh, kw = (kernel, kernel) if isinstance(kernel, int) else kernel
        if name is None:                      # Keras consumes an automatic index only for unnamed layers
            self.n_conv += 1
        cname = name or f"conv2d_{self.n_conv}"
        bname = None
        if not use_bias:
            if name is None:
                self.n_bn += 1
            bname = (name + "_bn") if name else f"batch_normalization_{self.n_bn}"
        if self.plan is not None:
            self.plan.append((cname, "conv", (kh, kw, cin, filters)))
            if use_bias:
                self.plan.append((cname + "/bias", "bias", (filters,)))
            else:
                self.plan.append((bname, "bn_noscale", (filters,)))
            return None
        x = _conv_same(x, self.w[cname], strides, self.w[cname + "/bias"] if use_bias else None)
        if not use_bias:
            x = self.bn(x, bname)
        if activation is not None:
            x = F.relu(x)
        return x

    # inception.py:36-48
    def conv_block(self, x, cin, filters):
        self.n_conv += 1
        self.n_bn += 1
        cname, bname = f"conv2d_{self.n_conv}", f"batch_normalization_{self.n_bn}"
        if self.plan is not None:
            self.plan += [(cname, "conv", (3, 3, cin, filters)), (cname + "/bias", "bias", (filters,)),
                          (bname, "bn", (filters,))]
            return None
        x = _conv_same(x, self.w[cname], 1, self.w[cname + "/bias"])
        return F.relu(self.bn(x, bname))

    def bn(self, x, name):
        g, b, mu, var = self.w[name]
        if self.cal:
            mu = x.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
            var = np.maximum(x.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32), 1e-4)
            self.w[name] = (g, b, mu, var)
        sh = (1, -1, 1, 1)
        return (x - _t(mu).view(sh)) / torch.sqrt(_t(var).view(sh) + EPS) * _t(g).view(sh) + _t(b).view(sh)

    # inception.py:89-162
    def block(self, x, cin, scale, block_type, idx, activation="relu"):
        c = self.conv2d_bn
        dry = self.plan is not None
        if block_type == "block35":
            b0 = c(x, cin, 32, 1)
            b1 = c(x, cin, 32, 1)
            b1 = c(b1, 32, 32, 3)
            b2 = c(x, cin, 32, 1)
            b2 = c(b2, 32, 48, 3)
            b2 = c(b2, 48, 64, 3)
            branches, cm = [b0, b1, b2], 128
        elif block_type == "block17":
            b0 = c(x, cin, 192, 1)
            b1 = c(x, cin, 128, 1)
            b1 = c(b1, 128, 160, (1, 7))
            b1 = c(b1, 160, 192, (7, 1))
            branches, cm = [b0, b1], 384
        else:
            b0 = c(x, cin, 192, 1)
            b1 = c(x, cin, 192, 1)
            b1 = c(b1, 192, 224, (1, 3))
            b1 = c(b1, 224, 256, (3, 1))
            branches, cm = [b0, b1], 448
        mixed = None if dry else torch.cat(branches, dim=1)
        up = c(mixed, cm, cin, 1, activation=None, use_bias=True, name=f"{block_type}_{idx}_conv")
        if dry:
            return None
        x = x + up * scale
        return F.relu(x) if activation is not None else x

    # inception.py:165-271
    def run(self, x, taps=None):
        c = self.conv2d_bn
        dry = self.plan is not None
        cat = (lambda ts: None) if dry else (lambda ts: torch.cat(ts, dim=1))
        mp = (lambda t: None) if dry else _maxpool_same
        rec = (lambda n, t: None) if (dry or taps is None) else (
            lambda n, t: taps.__setitem__(n, t.permute(0, 2, 3, 1).contiguous().numpy()))
        x = c(x, 3, 32, 3, strides=2)
        rec("stem1", x)
        x = c(x, 32, 32, 3)
        x = c(x, 32, 64, 3)
        conv1 = x
        rec("conv1", x)
        x = mp(x)
        x = c(x, 64, 80, 1)
        x = c(x, 80, 192, 3)
        conv2 = x
        rec("conv2", x)
        x = mp(x)
        rec("pool2", x)
        # mixed_5b
        b0 = c(x, 192, 96, 1)
        b1 = c(x, 192, 48, 1)
        b1 = c(b1, 48, 64, 5)
        b2 = c(x, 192, 64, 1)
        b2 = c(b2, 64, 96, 3)
        b2 = c(b2, 96, 96, 3)
        bp = None if dry else F.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False)
        bp = c(bp, 192, 64, 1)
        x = cat([b0, b1, b2, bp])
        rec("mixed_5b", x)
        for i in range(1, 11):
            x = self.block(x, 320, 0.17, "block35", i)
        conv3 = x
        rec("conv3", x)
        # mixed_6a
        b0 = c(x, 320, 384, 3, strides=2)
        b1 = c(x, 320, 256, 1)
        b1 = c(b1, 256, 256, 3)
        b1 = c(b1, 256, 384, 3, strides=2)
        x = cat([b0, b1, mp(x)])
        rec("mixed_6a", x)
        for i in range(1, 21):
            x = self.block(x, 1088, 0.1, "block17", i)
        conv4 = x
        rec("conv4", x)
        # mixed_7a
        b0 = c(x, 1088, 256, 1)
        b0 = c(b0, 256, 384, 3, strides=2)
        b1 = c(x, 1088, 256, 1)
        b1 = c(b1, 256, 288, 3, strides=2)
        b2 = c(x, 1088, 256, 1)
        b2 = c(b2, 256, 288, 3)
        b2 = c(b2, 288, 320, 3, strides=2)
        x = cat([b0, b1, b2, mp(x)])
        rec("mixed_7a", x)
        for i in range(1, 10):
            x = self.block(x, 2080, 0.2, "block8", i)
        x = self.block(x, 2080, 1.0, "block8", 10, activation=None)
        rec("block8_10", x)
        x = c(x, 2080, 1536, 1, name="conv_7b")
        rec("conv5", x)
        up = (lambda t: None) if dry else (lambda t: F.interpolate(t, scale_factor=2, mode="nearest"))
        b = self.conv_block
        x = b(up(x), 1536, 320)
        x = b(cat([x, conv4]), 320 + 1088, 320)
        rec("conv6", x)
        x = b(up(x), 320, 256)
        x = b(cat([x, conv3]), 256 + 320, 256)
        rec("conv7", x)
        x = b(up(x), 256, 128)
        x = b(cat([x, conv2]), 128 + 192, 128)
        rec("conv8", x)
        x = b(up(x), 128, 96)
        x = b(cat([x, conv1]), 96 + 64, 96)
        rec("conv9", x)
        x = b(up(x), 96, 64)
        x = b(x, 64, 64)
        rec("conv10", x)
        self.n_conv += 1
        hname = f"conv2d_{self.n_conv}"
        if dry:
            self.plan += [(hname, "conv", (1, 1, 64, 2)), (hname + "/bias", "bias", (2,))]
            return None
        z = _conv_same(x, self.w[hname], 1, self.w[hname + "/bias"])
        return F.softmax(z, dim=1)


def layer_plan():
    """Ordered (name, kind, shape) of every weight tensor, in Keras creation order."""
    plan = []
    _Net(None, plan=plan).run(None)
    return plan


def forward(weights: dict, x_nhwc: np.ndarray, calibrate: bool = False, taps: dict | None = None) -> np.ndarray:
    """float32 [B,P,P,3] in [-1,1] -> float32 [B,P,P,2] softmax (``Model.predict`` of the reference graph)."""
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2).contiguous()
        y = _Net(weights, calibrate).run(x, taps)
        return y.permute(0, 2, 3, 1).contiguous().numpy()


def calibrate_bn(weights: dict, calib_nhwc: np.ndarray) -> dict:
    """Moving statistics := batch statistics of one oracle pass (stands in for trained moving averages)."""
    forward(weights, calib_nhwc, calibrate=True)
    return weights


class OracleModel:
    """Stand-in for the Keras ``Model`` the reference loop calls (Segmentation.py:154-156)."""

    def __init__(self, weights: dict):
        self.weights = weights

    def predict(self, image_patches, batch_size=None, verbose=0, steps=None):
        return forward(self.weights, np.asarray(image_patches, dtype=np.float32))
