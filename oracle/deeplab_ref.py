"""fp32 PyTorch-CPU restatement of ``Deeplabv3(input_shape=(P,P,3), classes=2, backbone='xception', OS=16,
activation='softmax')`` (DigiPathAI/models/deeplabv3.py:52-160,219-482 as built at helpers/utils.py:438-446).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every weighted layer of this graph is explicitly named in the
reference, so the weight dict uses those names: ``<name>`` = HWIO conv kernel (depthwise kernels are
[3,3,C,1]), ``<name>/bias`` (only ``custom_logits_semantic``), ``<name>_BN`` = (gamma, beta, mean, var).
``layer_plan()`` lists them in creation order.

Keras / TF semantics encoded here (documented behaviour, not verifiable against a TF binary in this image;
"parity unpinned" for them):
  * Conv2D / DepthwiseConv2D padding='same', stride 1, dilation r: symmetric zero pad r        (:65,75,98-104)
  * stride 2: explicit ZeroPadding2D((pad_beg, pad_end)) with kernel_size_effective - 1 total, then 'valid'
    (:66-71,105-116) -- 1 / 1 for a 3x3, nothing for the 1x1 shortcut
  * first conv ``entry_flow_conv1_1``: Conv2D(3x3, strides 2, padding='same') -> TF pads 0 in front, 1 behind
  * SepConv_BN: [ReLU] -> depthwise -> BN -> [ReLU] -> pointwise -> BN -> [ReLU]; the first ReLU only when
    depth_activation is False, the other two only when it is True                               (:73-85)
  * BatchNormalization eps: 1e-3 (Keras default) in the backbone, 1e-5 in ASPP / decoder       (:74,382-)
  * GlobalAveragePooling2D + 1x1 conv + BN + ReLU, then bilinear resize of the 1x1 map = broadcast (:378-388)
  * tf.compat.v1.image.resize(bilinear, align_corners=True): src = dst * (in - 1) / (out - 1)  (:385,420,443)
  * Dropout(0.1) is the identity at inference                                                   (:413)
  * softmax over the 2 classes after the final resize                                           (:455-456)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

ATROUS = (6, 12, 18)
VAR_FLOOR = 1e-2   # calibration only: channels whose batch variance is tiny are not blown up to unit variance


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class _Net:
    def __init__(self, weights, calibrate=False, plan=None):
        self.w, self.cal, self.plan = weights, calibrate, plan

    # ---- primitives
    def bn(self, x, name, eps, c):
        if self.plan is not None:
            self.plan.append((name, "bn", (c,)))
            return None
        g, b, mu, var = self.w[name]
        if self.cal:
            mu = x.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
            var = np.maximum(x.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32), VAR_FLOOR)
            self.w[name] = (g, b, mu, var)
        sh = (1, -1, 1, 1)
        return (x - _t(mu).view(sh)) / torch.sqrt(_t(var).view(sh) + eps) * _t(g).view(sh) + _t(b).view(sh)

    def conv(self, x, name, cin, cout, k=1, stride=1, rate=1, same_tf=False, bias=False):
        """Conv2D as _conv2d_same builds it (deeplabv3.py:88-116); same_tf = plain padding='same' (first conv)."""
        if self.plan is not None:
            self.plan.append((name, "conv", (k, k, cin, cout)))
            if bias:
                self.plan.append((name + "/bias", "bias", (cout,)))
            return None
        w = _t(self.w[name]).permute(3, 2, 0, 1).contiguous()
        b = _t(self.w[name + "/bias"]) if bias else None
        ke = k + (k - 1) * (rate - 1)
        if stride == 1:
            p = (ke - 1) // 2
            return F.conv2d(x, w, b, padding=p, dilation=rate)
        if same_tf:
            n = x.shape[2]
            total = max((-(-n // stride) - 1) * stride + ke - n, 0)
            x = F.pad(x, (total // 2, total - total // 2, total // 2, total - total // 2))
        else:
            total = ke - 1
            x = F.pad(x, (total // 2, total - total // 2, total // 2, total - total // 2))
        return F.conv2d(x, w, b, stride=stride, dilation=rate)

    def dwconv(self, x, name, c, stride=1, rate=1):
        if self.plan is not None:
            self.plan.append((name, "dwconv", (3, 3, c, 1)))
            return None
        w = _t(self.w[name]).permute(2, 3, 0, 1).contiguous()      # [3,3,C,1] -> [C,1,3,3]
        ke = 3 + 2 * (rate - 1)
        if stride == 1:
            return F.conv2d(x, w, None, padding=(ke - 1) // 2, dilation=rate, groups=c)
        total = ke - 1
        x = F.pad(x, (total // 2, total - total // 2, total // 2, total - total // 2))
        return F.conv2d(x, w, None, stride=stride, dilation=rate, groups=c)

    def relu(self, x):
        return None if self.plan is not None else F.relu(x)

    # deeplabv3.py:52-85
    def sepconv_bn(self, x, cin, filters, prefix, stride=1, rate=1, depth_activation=False, eps=1e-3):
        if not depth_activation:
            x = self.relu(x)
        x = self.dwconv(x, prefix + "_depthwise", cin, stride, rate)
        x = self.bn(x, prefix + "_depthwise_BN", eps, cin)
        if depth_activation:
            x = self.relu(x)
        x = self.conv(x, prefix + "_pointwise", cin, filters)
        x = self.bn(x, prefix + "_pointwise_BN", eps, filters)
        if depth_activation:
            x = self.relu(x)
        return x

    # deeplabv3.py:119-158
    def xception_block(self, inputs, cin, depth_list, prefix, skip_type, stride, rate=1, depth_activation=False,
                       return_skip=False):
        residual, c, skip = inputs, cin, None
        for i in range(3):
            residual = self.sepconv_bn(residual, c, depth_list[i], f"{prefix}_separable_conv{i + 1}",
                                       stride=stride if i == 2 else 1, rate=rate, depth_activation=depth_activation)
            c = depth_list[i]
            if i == 1:
                skip = residual
        if skip_type == "conv":
            sc = self.conv(inputs, prefix + "_shortcut", cin, depth_list[-1], k=1, stride=stride)
            sc = self.bn(sc, prefix + "_shortcut_BN", 1e-3, depth_list[-1])
            out = None if self.plan is not None else residual + sc
        elif skip_type == "sum":
            out = None if self.plan is not None else residual + inputs
        else:
            out = residual
        return (out, skip) if return_skip else out

    # deeplabv3.py:283-460 (xception, OS=16, classes=2)
    def run(self, x, taps=None):
        dry = self.plan is not None
        rec = (lambda n, t: None) if (dry or taps is None) else (
            lambda n, t: taps.__setitem__(n, t.permute(0, 2, 3, 1).contiguous().numpy()))
        x = self.conv(x, "entry_flow_conv1_1", 3, 32, k=3, stride=2, same_tf=True)
        x = self.relu(self.bn(x, "entry_flow_conv1_1_BN", 1e-3, 32))
        rec("conv1_1", x)
        x = self.conv(x, "entry_flow_conv1_2", 32, 64, k=3)
        x = self.relu(self.bn(x, "entry_flow_conv1_2_BN", 1e-3, 64))
        rec("conv1_2", x)
        x = self.xception_block(x, 64, [128, 128, 128], "entry_flow_block1", "conv", 2)
        rec("block1", x)
        x, skip1 = self.xception_block(x, 128, [256, 256, 256], "entry_flow_block2", "conv", 2, return_skip=True)
        rec("block2", x)
        rec("skip1", skip1)
        x = self.xception_block(x, 256, [728, 728, 728], "entry_flow_block3", "conv", 2)
        rec("block3", x)
        for i in range(16):
            x = self.xception_block(x, 728, [728, 728, 728], f"middle_flow_unit_{i + 1}", "sum", 1)
        rec("middle", x)
        x = self.xception_block(x, 728, [728, 1024, 1024], "exit_flow_block1", "conv", 1, rate=1)
        rec("exit1", x)
        x = self.xception_block(x, 1024, [1536, 1536, 2048], "exit_flow_block2", "none", 1, rate=2,
                                depth_activation=True)
        rec("exit2", x)
        # ---- ASPP (deeplabv3.py:375-413)
        b4 = None if dry else x.mean(dim=(2, 3), keepdim=True)
        b4 = self.conv(b4, "image_pooling", 2048, 256)
        b4 = self.relu(self.bn_1x1(b4, "image_pooling_BN", 1e-5, 256))
        if not dry:
            b4 = b4.expand(-1, -1, x.shape[2], x.shape[3])
        b0 = self.conv(x, "aspp0", 2048, 256)
        b0 = self.relu(self.bn(b0, "aspp0_BN", 1e-5, 256))
        bs = [self.sepconv_bn(x, 2048, 256, f"aspp{i + 1}", rate=ATROUS[i], depth_activation=True, eps=1e-5)
              for i in range(3)]
        x = None if dry else torch.cat([b4, b0] + bs, dim=1)
        x = self.conv(x, "concat_projection", 1280, 256)
        x = self.relu(self.bn(x, "concat_projection_BN", 1e-5, 256))
        rec("aspp", x)
        # ---- decoder (deeplabv3.py:415-433)
        if not dry:
            x = F.interpolate(x, size=(skip1.shape[2], skip1.shape[3]), mode="bilinear", align_corners=True)
        d = self.conv(skip1, "feature_projection0", 256, 48)
        d = self.relu(self.bn(d, "feature_projection0_BN", 1e-5, 48))
        x = None if dry else torch.cat([x, d], dim=1)
        x = self.sepconv_bn(x, 304, 256, "decoder_conv0", depth_activation=True, eps=1e-5)
        x = self.sepconv_bn(x, 256, 256, "decoder_conv1", depth_activation=True, eps=1e-5)
        rec("decoder", x)
        x = self.conv(x, "custom_logits_semantic", 256, 2, bias=True)
        if dry:
            return None
        x = F.interpolate(x, size=self.in_hw, mode="bilinear", align_corners=True)
        return F.softmax(x, dim=1)

    def bn_1x1(self, x, name, eps, c):
        """BN on the pooled 1x1 map: calibration statistics over the batch only would be degenerate for tiny
        batches, so calibration leaves this BN at its initial statistics."""
        if self.plan is not None or not self.cal:
            return self.bn(x, name, eps, c)
        cal, self.cal = self.cal, False
        try:
            return self.bn(x, name, eps, c)
        finally:
            self.cal = cal


def layer_plan():
    plan = []
    _Net(None, plan=plan).run(None)
    return plan


def forward(weights: dict, x_nhwc: np.ndarray, calibrate: bool = False, taps: dict | None = None) -> np.ndarray:
    """float32 [B,P,P,3] in [-1,1] -> float32 [B,P,P,2] softmax (``Model.predict`` of the reference graph)."""
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2).contiguous()
        net = _Net(weights, calibrate)
        net.in_hw = (x.shape[2], x.shape[3])
        y = net.run(x, taps)
        return y.permute(0, 2, 3, 1).contiguous().numpy()


def calibrate_bn(weights: dict, calib_nhwc: np.ndarray) -> dict:
    forward(weights, calib_nhwc, calibrate=True)
    return weights


class OracleModel:
    """Stand-in for the Keras ``Model`` the reference loop calls (Segmentation.py:154-156)."""

    def __init__(self, weights: dict):
        self.weights = weights

    def predict(self, image_patches, batch_size=None, verbose=0, steps=None):
        return forward(self.weights, np.asarray(image_patches, dtype=np.float32))
