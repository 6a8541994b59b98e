"""Inception-ResNet-v2 encoder / U-Net decoder (reference: DigiPathAI/models/inception.py:36-271) as a layer
program for csrc/runtime.cu.

Weights are exchanged as a dict keyed the way Keras names the reference's layers: the unnamed
``conv2d_bn(..., name=None)`` / ``conv_block(prefix=None)`` layers get Keras' automatic ``conv2d_<i>`` /
``batch_normalization_<i>`` names in creation order (inception.py:72-83,40-44), the named ones keep
``block35_<i>_conv`` / ``block17_<i>_conv`` / ``block8_<i>_conv`` / ``conv_7b`` / ``conv_7b_bn``.  Conv kernels are
HWIO float32, biases ``<name>/bias``, BN entries (gamma, beta, moving_mean, moving_var) with gamma == 1 where the
reference builds ``BatchNormalization(scale=False)``.

How the graph maps onto the kernels (every `Concatenate` is a channel range of one buffer -- no concat copies):
  * sibling 1x1 convs that read the same tensor (the first conv of every branch) are ONE 1x1 conv with the
    branches' filters stacked along Cout (block35: 320->96, block17: 1088->320, block8: 2080->384,
    mixed_7a: 1088->768); the scratch buffer of a block is laid out so that the `mixed` tensor the block's
    closing 1x1 conv reads is one contiguous channel range;
  * the closing conv of a residual block, ``x + scale * (conv(mixed) + bias)`` then ReLU
    (inception.py:142-160), runs as a 1x1 conv whose epilogue adds the fp16 residual it is about to overwrite:
    ``scale`` is folded into the weights/bias, the running tensor ``x`` lives in place inside the decoder's
    concat buffer (conv3 / conv4 skips) so the skip connections cost nothing;
  * 1x7 / 7x1 / 1x3 / 3x1 / 5x5 and the stride-2 3x3 convs use the generic tap-table mode (KIND_TAPS):
    halo mode with asymmetric halos on 16x16+ maps, one shifted (for stride 2: element-strided) TMA box per tap
    otherwise; padding is TensorFlow's 'same' rule (0 in front / 1 behind for stride 2 on even maps);
  * the 3x3/2 stem conv on the RGB tile reuses the DenseNet stem path: 4x4 window on the 2x2 space-to-depth
    image gathered by stem_s2d_kernel, taps outside the 3x3 kernel are zero weights;
  * decoder: identical to the DenseNet U-Net's (sub-pixel up-convs, head fused into the last conv's epilogue).
"""
from __future__ import annotations

import numpy as np

from ..program import (KIND_1X1, KIND_3X3, KIND_STEM4, KIND_TAPS, KIND_UP2, OP_AVGPOOL3, OP_CONV, OP_MAXPOOL,
                       OP_STEM_S2D, POOL_TF_SAME, Op, Program, bn_affine, pack_conv_weights, pack_stem4_weights,
                       weight_precision)

EPS = 1e-3  # Keras BatchNormalization default; the reference never overrides it in this file


class _Conv:
    """One Conv2D(+BN) of the reference graph: names in creation order, geometry, folded weights."""

    def __init__(self, cname, bname, has_bias, has_bn, bn_scale, kh, kw, cin, cout, stride, relu):
        self.cname, self.bname, self.has_bias, self.has_bn, self.bn_scale = cname, bname, has_bias, has_bn, bn_scale
        self.kh, self.kw, self.cin, self.cout, self.stride, self.relu = kh, kw, cin, cout, stride, relu

    def shapes(self):
        out = [(self.cname, "conv", (self.kh, self.kw, self.cin, self.cout))]
        if self.has_bias:
            out.append((self.cname + "/bias", "bias", (self.cout,)))
        if self.has_bn:
            out.append((self.bname, "bn" if self.bn_scale else "bn_noscale", (self.cout,)))
        return out

    def folded(self, weights, extra_scale=1.0):
        """-> (HWIO fp32 kernel with BN scale folded in, fp32 shift) so that y = conv(x, k) + shift."""
        k = np.asarray(weights[self.cname], np.float32)
        assert k.shape == (self.kh, self.kw, self.cin, self.cout), (self.cname, k.shape)
        shift = np.zeros(self.cout, np.float32)
        if self.has_bias:
            shift = np.asarray(weights[self.cname + "/bias"], np.float32).copy()
        if self.has_bn:
            s, t = bn_affine(*weights[self.bname], EPS)
            k = k * s
            shift = (shift * s + t).astype(np.float32)
        if extra_scale != 1.0:
            k = k * np.float32(extra_scale)
            shift = (shift * np.float32(extra_scale)).astype(np.float32)
        return k.astype(np.float32), shift.astype(np.float32)


class _Builder:
    def __init__(self, weights, patch):
        self.w, self.P = weights, patch
        self.pr = Program(patch=patch)
        self.n_conv = self.n_bn = 0
        self.decl = []          # every _Conv in creation order
        self.macs = 0           # conv MACs of the reference graph per tile

    # ---- declaration (creation order == Keras naming order)
    def conv2d_bn(self, cin, filters, kernel, hw_out, strides=1, activation="relu", use_bias=False, name=None):
        kh, kw = (kernel, kernel) if isinstance(kernel, int) else kernel
        if name is None:
            self.n_conv += 1
        cname = name or f"conv2d_{self.n_conv}"
        bname = None
        if not use_bias:
            if name is None:
                self.n_bn += 1
            bname = (name + "_bn") if name else f"batch_normalization_{self.n_bn}"
        c = _Conv(cname, bname, use_bias, not use_bias, False, kh, kw, cin, filters, strides, activation is not None)
        self.decl.append(c)
        self.macs += hw_out * hw_out * kh * kw * cin * filters
        return c

    def conv_block(self, cin, filters, hw_out):
        self.n_conv += 1
        self.n_bn += 1
        c = _Conv(f"conv2d_{self.n_conv}", f"batch_normalization_{self.n_bn}", True, True, True, 3, 3, cin, filters, 1, True)
        self.decl.append(c)
        self.macs += hw_out * hw_out * 9 * cin * filters
        return c

    # ---- emission
    def emit(self, convs, ib, ioff, ob, ooff, residual_scale=None, kind=None, head=None, name=None):
        """One OP_CONV computing the listed sibling convs (same input, same geometry), filters stacked in order."""
        if self.w is None:
            return
        c0 = convs[0]
        ks, shs = zip(*[c.folded(self.w, residual_scale if residual_scale is not None else 1.0) for c in convs])
        k = np.concatenate(ks, axis=3)
        sh = np.concatenate(shs)
        assert all((c.kh, c.kw, c.cin, c.stride, c.relu) == (c0.kh, c0.kw, c0.cin, c0.stride, c0.relu) for c in convs)
        if kind is None:
            if (c0.kh, c0.kw, c0.stride) == (1, 1, 1):
                kind = KIND_1X1
            elif (c0.kh, c0.kw, c0.stride) == (3, 3, 1):
                kind = KIND_3X3
            else:
                kind = KIND_TAPS
        if kind == KIND_STEM4:
            wp = pack_stem4_weights(k, pad=0)
            cin = 64
        else:
            wp = pack_conv_weights(k, kind)
            cin = c0.cin
        op = Op(OP_CONV, in_buf=ib, in_choff=ioff, cin=cin, out_buf=ob, out_choff=ooff, cout=k.shape[3], kind=kind,
                relu=int(c0.relu), epi_shift=sh, w=wp, name=name or "+".join(c.cname for c in convs))
        if kind == KIND_TAPS:
            op.kh, op.kw, op.stride = c0.kh, c0.kw, c0.stride
        if residual_scale is not None:
            op.residual = 1
        if head is not None:
            hk = np.asarray(self.w[head.cname], np.float32)[0, 0]
            hb = np.asarray(self.w[head.cname + "/bias"], np.float32)
            op.head = 1
            op.head_w = (hk[:, 1] - hk[:, 0]).astype(np.float32)
            op.head_b = float(hb[1] - hb[0])
        self.pr.ops.append(op)

    def pool(self, typ, ib, ioff, c, ob, ooff, name):
        if self.w is None:
            return
        self.pr.ops.append(Op(typ, in_buf=ib, in_choff=ioff, cin=c, out_buf=ob, out_choff=ooff, cout=c,
                              pool=POOL_TF_SAME if typ == OP_MAXPOOL else 0, name=name))


def _build(weights, patch):
    if patch < 64 or patch & (patch - 1):
        raise ValueError("patch_size must be a power of two >= 64 for the B200 tile kernels")
    P = patch
    b = _Builder(weights, P)
    pr = b.pr
    h1, h2, h3, h4, h5 = P // 2, P // 4, P // 8, P // 16, P // 32
    buf = lambda name, hw, c: pr.add_buf(name, hw, hw, c)
    S0 = buf("stem_s2d", h1, 64)
    S1, S2 = buf("S1", h1, 32), buf("S2", h1, 32)
    D1 = buf("D1", h1, 160)      # [dec9a 96 | conv1 64]
    P1, S3 = buf("P1", h2, 64), buf("S3", h2, 80)
    D2 = buf("D2", h2, 320)      # [dec8a 128 | conv2 192]
    P2 = buf("P2", h3, 192)
    T5, T5b, T5p = buf("T5", h3, 112), buf("T5b", h3, 96), buf("T5p", h3, 192)
    D3 = buf("D3", h3, 576)      # [dec7a 256 | block35 running tensor / conv3 320]
    M35 = buf("M35", h3, 240)    # [b1a 32 | b2a 32 | b0 32 | b1 32 | b2 64 | b2b 48]; mixed = [64:192)
    T6a, T6b = buf("T6a", h3, 256), buf("T6b", h3, 256)
    D4 = buf("D4", h4, 1408)     # [dec6a 320 | block17 running tensor / conv4 1088]
    M17 = buf("M17", h4, 672)    # [b1a 128 | b0 192 | b1 192 | b1b 160]; mixed = [128:512)
    T7, T7b = buf("T7", h4, 768), buf("T7b", h4, 288)
    X8 = buf("X8", h5, 2080)     # block8 running tensor
    M8 = buf("M8", h5, 864)      # [b1a 192 | b0 192 | b1 256 | b1b 224]; mixed = [192:640)
    C5 = buf("C5", h5, 1536)
    E6, E7, E8, E9, E10 = buf("E6", h4, 320), buf("E7", h3, 256), buf("E8", h2, 128), buf("E9", h1, 96), buf("E10", P, 64)
    c = b.conv2d_bn

    # ---- stem (inception.py:171-183)
    if weights is not None:
        pr.ops.append(Op(OP_STEM_S2D, out_buf=S0, cout=64, name="stem_s2d"))
    b.emit([c(3, 32, 3, h1, strides=2)], S0, 0, S1, 0, kind=KIND_STEM4)
    b.emit([c(32, 32, 3, h1)], S1, 0, S2, 0)
    b.emit([c(32, 64, 3, h1)], S2, 0, D1, 96)                   # conv1
    b.pool(OP_MAXPOOL, D1, 96, 64, P1, 0, "pool1")
    b.emit([c(64, 80, 1, h2)], P1, 0, S3, 0)
    b.emit([c(80, 192, 3, h2)], S3, 0, D2, 128)                 # conv2
    b.pool(OP_MAXPOOL, D2, 128, 192, P2, 0, "pool2")

    # ---- mixed_5b (inception.py:185-196) -> D3[256:576) = [b0 96 | b1 64 | b2 96 | pool 64]
    X3 = 256
    b0 = c(192, 96, 1, h3)
    b1a = c(192, 48, 1, h3)
    b1 = c(48, 64, 5, h3)
    b2a = c(192, 64, 1, h3)
    b2b = c(64, 96, 3, h3)
    b2 = c(96, 96, 3, h3)
    bp = c(192, 64, 1, h3)
    b.emit([b0], P2, 0, D3, X3)
    b.emit([b1a, b2a], P2, 0, T5, 0)
    b.emit([b1], T5, 0, D3, X3 + 96)
    b.emit([b2b], T5, 48, T5b, 0)
    b.emit([b2], T5b, 0, D3, X3 + 160)
    b.pool(OP_AVGPOOL3, P2, 0, 192, T5p, 0, "mixed_5b_avgpool")
    b.emit([bp], T5p, 0, D3, X3 + 256)

    # ---- 10 x block35 (inception.py:199-203, 118-125), in place on D3[256:576)
    for i in range(1, 11):
        b0 = c(320, 32, 1, h3)
        b1a = c(320, 32, 1, h3)
        b1 = c(32, 32, 3, h3)
        b2a = c(320, 32, 1, h3)
        b2b = c(32, 48, 3, h3)
        b2 = c(48, 64, 3, h3)
        up = c(128, 320, 1, h3, activation="relu", use_bias=True, name=f"block35_{i}_conv")
        b.emit([b1a, b2a, b0], D3, X3, M35, 0)
        b.emit([b1], M35, 0, M35, 96)
        b.emit([b2b], M35, 32, M35, 192)
        b.emit([b2], M35, 192, M35, 128)
        b.emit([up], M35, 64, D3, X3, residual_scale=0.17)

    # ---- mixed_6a (inception.py:205-212) -> D4[320:1408) = [b0 384 | b1 384 | maxpool 320]
    X4 = 320
    b0 = c(320, 384, 3, h4, strides=2)
    b1a = c(320, 256, 1, h3)
    b1b = c(256, 256, 3, h3)
    b1 = c(256, 384, 3, h4, strides=2)
    b.emit([b0], D3, X3, D4, X4)
    b.emit([b1a], D3, X3, T6a, 0)
    b.emit([b1b], T6a, 0, T6b, 0)
    b.emit([b1], T6b, 0, D4, X4 + 384)
    b.pool(OP_MAXPOOL, D3, X3, 320, D4, X4 + 768, "mixed_6a_pool")

    # ---- 20 x block17 (inception.py:214-219, 126-131), in place on D4[320:1408)
    for i in range(1, 21):
        b0 = c(1088, 192, 1, h4)
        b1a = c(1088, 128, 1, h4)
        b1b = c(128, 160, (1, 7), h4)
        b1 = c(160, 192, (7, 1), h4)
        up = c(384, 1088, 1, h4, activation="relu", use_bias=True, name=f"block17_{i}_conv")
        b.emit([b1a, b0], D4, X4, M17, 0)
        b.emit([b1b], M17, 0, M17, 512)
        b.emit([b1], M17, 512, M17, 320)
        b.emit([up], M17, 128, D4, X4, residual_scale=0.1)

    # ---- mixed_7a (inception.py:221-232) -> X8 = [b0 384 | b1 288 | b2 320 | maxpool 1088]
    b0a = c(1088, 256, 1, h4)
    b0 = c(256, 384, 3, h5, strides=2)
    b1a = c(1088, 256, 1, h4)
    b1 = c(256, 288, 3, h5, strides=2)
    b2a = c(1088, 256, 1, h4)
    b2b = c(256, 288, 3, h4)
    b2 = c(288, 320, 3, h5, strides=2)
    b.emit([b0a, b1a, b2a], D4, X4, T7, 0)
    b.emit([b0], T7, 0, X8, 0)
    b.emit([b1], T7, 256, X8, 384)
    b.emit([b2b], T7, 512, T7b, 0)
    b.emit([b2], T7b, 0, X8, 672)
    b.pool(OP_MAXPOOL, D4, X4, 1088, X8, 992, "mixed_7a_pool")

    # ---- 10 x block8 (inception.py:234-244, 132-137), in place on X8; the last one has scale 1 and no ReLU
    for i in range(1, 11):
        b0 = c(2080, 192, 1, h5)
        b1a = c(2080, 192, 1, h5)
        b1b = c(192, 224, (1, 3), h5)
        b1 = c(224, 256, (3, 1), h5)
        up = c(448, 2080, 1, h5, activation="relu" if i < 10 else None, use_bias=True, name=f"block8_{i}_conv")
        b.emit([b1a, b0], X8, 0, M8, 0)
        b.emit([b1b], M8, 0, M8, 640)
        b.emit([b1], M8, 640, M8, 384)
        b.emit([up], M8, 192, X8, 0, residual_scale=0.2 if i < 10 else 1.0)

    # ---- conv_7b + decoder (inception.py:246-268)
    b.emit([c(2080, 1536, 1, h5, name="conv_7b")], X8, 0, C5, 0)
    d = b.conv_block
    b.emit([d(1536, 320, h4)], C5, 0, D4, 0, kind=KIND_UP2, name="dec6a")
    b.emit([d(1408, 320, h4)], D4, 0, E6, 0, name="dec6b")
    b.emit([d(320, 256, h3)], E6, 0, D3, 0, kind=KIND_UP2, name="dec7a")
    b.emit([d(576, 256, h3)], D3, 0, E7, 0, name="dec7b")
    b.emit([d(256, 128, h2)], E7, 0, D2, 0, kind=KIND_UP2, name="dec8a")
    b.emit([d(320, 128, h2)], D2, 0, E8, 0, name="dec8b")
    b.emit([d(128, 96, h1)], E8, 0, D1, 0, kind=KIND_UP2, name="dec9a")
    b.emit([d(160, 96, h1)], D1, 0, E9, 0, name="dec9b")
    b.emit([d(96, 64, P)], E9, 0, E10, 0, kind=KIND_UP2, name="dec10a")
    last = d(64, 64, P)
    b.n_conv += 1
    head = _Conv(f"conv2d_{b.n_conv}", None, True, False, False, 1, 1, 64, 2, 1, False)
    b.decl.append(head)
    b.macs += P * P * 64 * 2
    b.emit([last], E10, 0, E10, 0, head=head, name="dec10b")
    return b


def layer_shapes():
    """Ordered (name, kind, shape) of every weight tensor of get_inception_resnet_v2_unet_softmax."""
    out = []
    for cv in _build(None, 256).decl:
        out += cv.shapes()
    return out


def reference_macs_per_tile(patch: int = 256) -> int:
    """Conv MACs of one forward of the reference graph (no merging / sub-pixel rewrite)."""
    return _build(None, patch).macs


def init_inception_weights(seed: int = 0) -> dict:
    """Seeded random-init weights of the reference architecture (He-normal convs, inception.py:40).  BN moving
    statistics start at (0, 1); tests calibrate them with an oracle pass."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, kind, shape in layer_shapes():
        if kind == "conv":
            fan_in = shape[0] * shape[1] * shape[2]
            w[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif kind == "bias":
            w[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:
            gamma = (rng.uniform(0.8, 1.2, shape) if kind == "bn" else np.ones(shape)).astype(np.float32)
            beta = (0.1 * rng.standard_normal(shape)).astype(np.float32)
            w[name] = (gamma, beta, np.zeros(shape, np.float32), np.ones(shape, np.float32))
    return w


def inception_resnet_v2_unet_program(weights: dict, patch: int = 256, precision: str = "fp16") -> Program:
    with weight_precision(precision):
        pr = _build(weights, patch).pr
    pr.precision = precision
    return pr
