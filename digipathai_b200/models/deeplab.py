"""DeepLabv3+ with the modified Xception backbone, OS = 16, 2 classes, softmax (reference:
DigiPathAI/models/deeplabv3.py:52-160,219-482 as instantiated at DigiPathAI/helpers/utils.py:438-446) as a layer
program for csrc/runtime.cu.

Weights are exchanged as a dict keyed by the reference's explicit Keras layer names: ``<name>`` = HWIO kernel
(depthwise kernels [3,3,C,1]), ``<name>_BN`` = (gamma, beta, moving_mean, moving_var),
``custom_logits_semantic/bias``.

How the graph maps onto the kernels:
  * every pointwise / 1x1 conv (96 % of the MACs) is the tcgen05 1x1 conv kernel with its BatchNorm folded into
    the fp16 weights; the Xception units' additions (``layers.add``, deeplabv3.py:150-153) ride in the closing
    pointwise conv's residual epilogue: 'sum' units update the running tensor in place, 'conv' units first write
    the stride-2 1x1 shortcut (element-strided TMA boxes) into the output buffer and then accumulate onto it;
  * depthwise 3x3 convs (stride 1/2, dilation 1/2/6/12/18) are an HBM-bound CUDA-core kernel (0.19 of 12.8
    GMAC: not tensor-core work) that also applies the pre-ReLU / folded BN / post-ReLU around them;
  * 728-channel tensors are carried as 736 channels (zero weights, zero activations in the 8 pad channels) so
    that every Cout is a multiple of 16;
  * ASPP image pooling = global-average-pool kernel + 1x1 conv on the [B,1,1,2048] map + broadcast into the
    concat buffer (a bilinear resize of a 1x1 map is a broadcast); the five ASPP branches write channel ranges
    of one 1280-channel buffer;
  * ``tf.image.resize(bilinear, align_corners=True)`` is one kernel; the last one is fused with the head: the
    2-class softmax's channel 1 is sigmoid(z1 - z0), bilinear interpolation is linear, so the 1x1 logits conv
    collapses to one fp32 dot product per 64x64 pixel and the head kernel interpolates that difference to
    PxP, applies the sigmoid and the inverse TTA, and writes the fp32 probabilities.
"""
from __future__ import annotations

import numpy as np

from ..program import (KIND_1X1, KIND_3X3, KIND_STEM4, KIND_TAPS, OP_BCAST, OP_CONV, OP_DWCONV, OP_GAP, OP_HEAD_DOT,
                       OP_HEAD_RESIZE, OP_RESIZE, OP_STEM_S2D, Op, Program, bn_affine, pack_conv_weights,
                       pack_stem4_weights, wdtype, weight_precision)

ATROUS = (6, 12, 18)


def _pad16(c):
    return (c + 15) // 16 * 16


class _Builder:
    def __init__(self, weights, patch):
        self.w, self.P = weights, patch
        self.pr = Program(patch=patch)
        self.shapes = []
        self.macs = 0
        self._bufs = {}

    def buf(self, name, hw, c):
        if name not in self._bufs:
            self._bufs[name] = self.pr.add_buf(name, hw, hw, c)
        return self._bufs[name]

    # ---- layers
    def conv(self, name, ib, ioff, cin, ob, ooff, cout, hw_out, bn_eps=None, relu=False, k=1, stride=1, residual=False,
             kind=None):
        """Conv2D(k x k) [+ BN folded] [+ ReLU]; cin/cout are the reference's counts, buffers may be padded."""
        self.shapes.append((name, "conv", (k, k, cin, cout)))
        if bn_eps is not None:
            self.shapes.append((name + "_BN", "bn", (cout,)))
        self.macs += hw_out * hw_out * k * k * cin * cout
        if self.w is None:
            return
        kern = np.asarray(self.w[name], np.float32)
        shift = np.zeros(cout, np.float32)
        if bn_eps is not None:
            s, shift = bn_affine(*self.w[name + "_BN"], bn_eps)
            kern = kern * s
        cin_p, cout_p = (64 if kind == KIND_STEM4 else _pad16(cin)), _pad16(cout)
        if kind == KIND_STEM4:
            wp = pack_stem4_weights(kern, pad=0)
        else:
            kp = np.zeros((k, k, cin_p, cout_p), np.float32)
            kp[:, :, :cin, :cout] = kern
            if kind is None:
                kind = KIND_1X1 if (k, stride) == (1, 1) else (KIND_3X3 if (k, stride) == (3, 1) else KIND_TAPS)
            wp = pack_conv_weights(kp, kind)
        sh = np.zeros(cout_p, np.float32)
        sh[:cout] = shift
        op = Op(OP_CONV, in_buf=ib, in_choff=ioff, cin=cin_p, out_buf=ob, out_choff=ooff, cout=cout_p, kind=kind,
                relu=int(relu), epi_shift=sh, w=wp, residual=int(residual), name=name)
        if kind == KIND_TAPS:
            op.kh = op.kw = k
            op.stride = stride
        self.pr.ops.append(op)

    def dwconv(self, name, ib, ioff, c, ob, ooff, hw_out, bn_eps, pre_relu, post_relu, stride=1, rate=1):
        """[ReLU] -> DepthwiseConv2D(3x3) -> BN [-> ReLU] (deeplabv3.py:73-79)."""
        self.shapes += [(name, "dwconv", (3, 3, c, 1)), (name + "_BN", "bn", (c,))]
        self.macs += hw_out * hw_out * 9 * c
        if self.w is None:
            return
        kern = np.asarray(self.w[name], np.float32)[:, :, :, 0]            # [3,3,C]
        s, shift = bn_affine(*self.w[name + "_BN"], bn_eps)
        cp = _pad16(c)
        wp = np.zeros((9, cp), np.float32)
        wp[:, :c] = (kern * s).reshape(9, c)
        sh = np.zeros(cp, np.float32)
        sh[:c] = shift
        self.pr.ops.append(Op(OP_DWCONV, in_buf=ib, in_choff=ioff, cin=cp, out_buf=ob, out_choff=ooff, cout=cp,
                              relu=int(post_relu), pro=int(pre_relu), stride=stride, rate=rate,
                              w=wp.astype(wdtype()), epi_shift=sh, name=name))

    def sepconv(self, prefix, ib, ioff, cin, tb, ob, ooff, filters, hw_in, stride=1, rate=1, depth_activation=False,
                eps=1e-3, residual=False):
        """SepConv_BN (deeplabv3.py:52-85): depthwise into scratch buffer `tb`, pointwise into (ob, ooff)."""
        hw_out = hw_in // stride
        self.dwconv(prefix + "_depthwise", ib, ioff, cin, tb, 0, hw_out, eps, not depth_activation, depth_activation,
                    stride, rate)
        self.conv(prefix + "_pointwise", tb, 0, cin, ob, ooff, filters, hw_out, bn_eps=eps, relu=depth_activation,
                  residual=residual)

    def simple(self, typ, ib, ioff, c, ob, ooff, name, **kw):
        if self.w is None:
            return
        self.pr.ops.append(Op(typ, in_buf=ib, in_choff=ioff, cin=c, out_buf=ob, out_choff=ooff, cout=c, name=name, **kw))


def _build(weights, patch):
    if patch < 64 or patch & (patch - 1):
        raise ValueError("patch_size must be a power of two >= 64 for the B200 tile kernels")
    P = patch
    b = _Builder(weights, P)
    h1, h2, h3, h4 = P // 2, P // 4, P // 8, P // 16
    C7 = 736                                      # 728 carried as 736 (see module docstring)

    # ---- entry flow (deeplabv3.py:283-300)
    S0 = b.buf("stem_s2d", h1, 64)
    if weights is not None:
        b.pr.ops.append(Op(OP_STEM_S2D, out_buf=S0, cout=64, name="stem_s2d"))
    A1, A2 = b.buf("A1", h1, 32), b.buf("A2", h1, 64)
    b.conv("entry_flow_conv1_1", S0, 0, 3, A1, 0, 32, h1, bn_eps=1e-3, relu=True, k=3, stride=2, kind=KIND_STEM4)
    b.conv("entry_flow_conv1_2", A1, 0, 32, A2, 0, 64, h1, bn_eps=1e-3, relu=True, k=3)

    def entry_block(prefix, X, cin, hw, c, out_name, skip_name=None):
        """_xception_block(skip_connection_type='conv', stride=2): returns the output buffer ([hw/2] x c)."""
        cp = C7 if c == 728 else c
        cinp = C7 if cin == 728 else cin
        D1 = b.buf(f"{prefix}_d1", hw, cinp)
        T1 = b.buf(f"{prefix}_t1", hw, cp)
        D2 = b.buf(f"{prefix}_d2", hw, cp)
        T2 = b.buf(skip_name or f"{prefix}_t2", hw, cp)
        D3 = b.buf(f"{prefix}_d3", hw // 2, cp)
        Y = b.buf(out_name, hw // 2, cp)
        b.sepconv(prefix + "_separable_conv1", X, 0, cin, D1, T1, 0, c, hw)
        b.sepconv(prefix + "_separable_conv2", T1, 0, c, D2, T2, 0, c, hw)
        # shortcut first (into Y), then the closing pointwise conv accumulates onto it
        b.dwconv(prefix + "_separable_conv3_depthwise", T2, 0, c, D3, 0, hw // 2, 1e-3, True, False, stride=2)
        # creation order in the reference: separable_conv3 (dw, pw) then shortcut; names are explicit, order free
        b.shapes.append(None)  # placeholder swapped below to keep layer_shapes() in creation order
        mark = len(b.shapes) - 1
        b.conv(prefix + "_shortcut", X, 0, cin, Y, 0, c, hw // 2, bn_eps=1e-3, k=1, stride=2)
        sc = b.shapes[mark + 1:]
        del b.shapes[mark:]
        b.conv(prefix + "_separable_conv3_pointwise", D3, 0, c, Y, 0, c, hw // 2, bn_eps=1e-3, residual=True)
        b.shapes += sc
        return Y

    B1 = entry_block("entry_flow_block1", A2, 64, h1, 128, "B1")
    B2 = entry_block("entry_flow_block2", B1, 128, h2, 256, "B2", skip_name="skip1")
    SK = b.buf("skip1", h2, 256)
    X = entry_block("entry_flow_block3", B2, 256, h3, 728, "X")

    # ---- middle flow: 16 x [728, 728, 728] 'sum' units, in place on X (deeplabv3.py:302-305)
    M1, M2 = b.buf("M1", h4, C7), b.buf("M2", h4, C7)
    for i in range(16):
        p = f"middle_flow_unit_{i + 1}"
        b.sepconv(p + "_separable_conv1", X, 0, 728, M1, M2, 0, 728, h4)
        b.sepconv(p + "_separable_conv2", M2, 0, 728, M1, M2, 0, 728, h4)
        b.sepconv(p + "_separable_conv3", M2, 0, 728, M1, X, 0, 728, h4, residual=True)

    # ---- exit flow (deeplabv3.py:307-312)
    E1, E2, E3 = b.buf("E1", h4, 1024), b.buf("E2", h4, 1024), b.buf("E3", h4, 1024)
    p = "exit_flow_block1"
    b.sepconv(p + "_separable_conv1", X, 0, 728, M1, M2, 0, 728, h4)
    b.sepconv(p + "_separable_conv2", M2, 0, 728, M1, E1, 0, 1024, h4)
    b.dwconv(p + "_separable_conv3_depthwise", E1, 0, 1024, E2, 0, h4, 1e-3, True, False)
    b.shapes.append(None)
    mark = len(b.shapes) - 1
    b.conv(p + "_shortcut", X, 0, 728, E3, 0, 1024, h4, bn_eps=1e-3)
    sc = b.shapes[mark + 1:]
    del b.shapes[mark:]
    b.conv(p + "_separable_conv3_pointwise", E2, 0, 1024, E3, 0, 1024, h4, bn_eps=1e-3, residual=True)
    b.shapes += sc
    F1, F2, G = b.buf("F1", h4, 1536), b.buf("F2", h4, 1536), b.buf("G", h4, 2048)
    p = "exit_flow_block2"
    b.sepconv(p + "_separable_conv1", E3, 0, 1024, E2, F1, 0, 1536, h4, rate=2, depth_activation=True)
    b.sepconv(p + "_separable_conv2", F1, 0, 1536, F2, F1, 0, 1536, h4, rate=2, depth_activation=True)
    b.sepconv(p + "_separable_conv3", F1, 0, 1536, F2, G, 0, 2048, h4, rate=2, depth_activation=True)

    # ---- ASPP (deeplabv3.py:375-413): CC = [image pooling 256 | aspp0 256 | aspp1 256 | aspp2 256 | aspp3 256]
    CC, GP, GQ, H1 = b.buf("CC", h4, 1280), b.buf("GP", 1, 2048), b.buf("GQ", 1, 256), b.buf("H1", h4, 2048)
    b.simple(OP_GAP, G, 0, 2048, GP, 0, "image_pooling_gap")
    b.conv("image_pooling", GP, 0, 2048, GQ, 0, 256, 1, bn_eps=1e-5, relu=True)
    b.simple(OP_BCAST, GQ, 0, 256, CC, 0, "image_pooling_broadcast")
    b.conv("aspp0", G, 0, 2048, CC, 256, 256, h4, bn_eps=1e-5, relu=True)
    for i in range(3):
        b.sepconv(f"aspp{i + 1}", G, 0, 2048, H1, CC, 512 + 256 * i, 256, h4, rate=ATROUS[i], depth_activation=True,
                  eps=1e-5)
    PJ = b.buf("PJ", h4, 256)
    b.conv("concat_projection", CC, 0, 1280, PJ, 0, 256, h4, bn_eps=1e-5, relu=True)

    # ---- decoder (deeplabv3.py:415-433): DC = [upsampled ASPP 256 | feature_projection0 48]
    DC, D0, D1, D2, D3 = (b.buf("DC", h2, 304), b.buf("D0", h2, 304), b.buf("D1", h2, 256), b.buf("D2", h2, 256),
                          b.buf("D3", h2, 256))
    b.simple(OP_RESIZE, PJ, 0, 256, DC, 0, "decoder_resize")
    b.conv("feature_projection0", SK, 0, 256, DC, 256, 48, h2, bn_eps=1e-5, relu=True)
    b.sepconv("decoder_conv0", DC, 0, 304, D0, D1, 0, 256, h2, depth_activation=True, eps=1e-5)
    b.sepconv("decoder_conv1", D1, 0, 256, D2, D3, 0, 256, h2, depth_activation=True, eps=1e-5)

    # ---- logits + final resize + softmax channel 1 (deeplabv3.py:440-456)
    b.shapes += [("custom_logits_semantic", "conv", (1, 1, 256, 2)), ("custom_logits_semantic/bias", "bias", (2,))]
    b.macs += h2 * h2 * 256 * 2
    LG = b.buf("LG", h2, 8)      # one fp32 logit difference per pixel, stored in the first 4 bytes of an 8-half pixel
    if weights is not None:
        hk = np.asarray(weights["custom_logits_semantic"], np.float32)[0, 0]
        hb = np.asarray(weights["custom_logits_semantic/bias"], np.float32)
        b.pr.ops.append(Op(OP_HEAD_DOT, in_buf=D3, cin=256, out_buf=LG, cout=8,
                           head_w=(hk[:, 1] - hk[:, 0]).astype(np.float32), head_b=float(hb[1] - hb[0]),
                           name="custom_logits_semantic"))
        b.pr.ops.append(Op(OP_HEAD_RESIZE, in_buf=LG, cin=8, out_buf=LG, cout=8, head=1, name="final_resize_softmax"))
    return b


def layer_shapes():
    """(name, kind, shape) of every weight tensor of the reference graph."""
    return [s for s in _build(None, 256).shapes if s is not None]


def reference_macs_per_tile(patch: int = 256) -> int:
    """Conv + depthwise MACs of one forward of the reference graph."""
    return _build(None, patch).macs


def init_deeplab_weights(seed: int = 0) -> dict:
    """Seeded random-init weights (He-normal); BN moving statistics start at (0, 1), tests calibrate them."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, kind, shape in layer_shapes():
        if kind in ("conv", "dwconv"):
            fan_in = shape[0] * shape[1] * (shape[2] if kind == "conv" else 1)
            w[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif kind == "bias":
            w[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:
            w[name] = (rng.uniform(0.8, 1.2, shape).astype(np.float32), (0.1 * rng.standard_normal(shape)).astype(np.float32),
                       np.zeros(shape, np.float32), np.ones(shape, np.float32))
    return w


def deeplabv3plus_xception_program(weights: dict, patch: int = 256, precision: str = "fp16") -> Program:
    with weight_precision(precision):
        pr = _build(weights, patch).pr
    pr.precision = precision
    return pr
