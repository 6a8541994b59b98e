"""DenseNet-121 encoder / U-Net decoder (reference: DigiPathAI/models/densenet.py:37-159) as a layer program.

Weights are exchanged as a dict keyed by the reference's Keras layer names -- ``conv1/conv``, ``conv1/bn``,
``conv{2..5}_block{i}_{0_bn,1_conv,1_bn,2_conv}``, ``pool{2,3,4}_{bn,conv}``, ``bn`` -- plus
``dec{6..10}{a,b}_{conv,norm}`` and ``head`` for the decoder layers the reference leaves unnamed
(``conv_block(..., prefix=None)``, densenet.py:37-48,138-156; Keras numbers them by creation order, which is
exactly this order).  Conv kernels are HWIO float32; BN entries are (gamma, beta, moving_mean, moving_var).

Buffer plan (per image; every `concatenate` of the reference is a channel range of one buffer, so no concat
copies exist): D1 = [dec9a 96 | conv1 64] @128^2, D2 = [dec8a 128 | block2 256] @64^2,
D3 = [dec7a 256 | block3 512] @32^2, D4 = [dec6a 320 | block4 1024] @16^2, D5 = block5 1024 @8^2.
The stem conv runs in halo mode, which needs maps at least 16 rows high: patch >= 64.
"""
from __future__ import annotations

import numpy as np

from ..program import (KIND_1X1, KIND_3X3, KIND_STEM4, KIND_UP2, OP_BNPOOL, OP_CONV, OP_DENSE_LAYER, OP_MAXPOOL,
                       OP_STEM_S2D,
                       PRO_AFFINE_RELU, Op, Program, bn_affine, pack_conv_weights, pack_stem4_weights, pad64,
                       weight_precision)

DENSENET_BLOCKS = (6, 12, 24, 16)
GROWTH = 32
EPS_ENC = 1.001e-5   # densenet.py:61,67,101,118,134
EPS_DEC = 1e-3       # Keras BatchNormalization default, densenet.py:44
DECODER = (  # name, cin, cout  (densenet.py:138-155)
    ("dec6a", 1024, 320), ("dec6b", 1344, 320), ("dec7a", 320, 256), ("dec7b", 768, 256),
    ("dec8a", 256, 128), ("dec8b", 384, 128), ("dec9a", 128, 96), ("dec9b", 160, 96),
    ("dec10a", 96, 64), ("dec10b", 64, 64),
)


def layer_shapes():
    """Ordered (name, kind, shape) of every weight tensor of unet_densenet121."""
    out = [("conv1/conv", "conv", (7, 7, 3, 64)), ("conv1/bn", "bn", (64,))]
    c = 64
    for b, n in zip((2, 3, 4, 5), DENSENET_BLOCKS):
        for i in range(1, n + 1):
            p = f"conv{b}_block{i}"
            out += [(p + "_0_bn", "bn", (c,)), (p + "_1_conv", "conv", (1, 1, c, 4 * GROWTH)),
                    (p + "_1_bn", "bn", (4 * GROWTH,)), (p + "_2_conv", "conv", (3, 3, 4 * GROWTH, GROWTH))]
            c += GROWTH
        if b < 5:
            out += [(f"pool{b}_bn", "bn", (c,)), (f"pool{b}_conv", "conv", (1, 1, c, c // 2))]
            c //= 2
    out.append(("bn", "bn", (c,)))
    for name, cin, cout in DECODER:
        out += [(name + "_conv", "conv", (3, 3, cin, cout)), (name + "_conv_bias", "bias", (cout,)),
                (name + "_norm", "bn", (cout,))]
    out += [("head", "conv", (1, 1, 64, 2)), ("head_bias", "bias", (2,))]
    return out


def init_densenet_weights(seed: int = 0) -> dict:
    """Seeded random-init weights of the reference architecture (He-normal convs as in densenet.py:41).

    BatchNorm gamma/beta are mildly randomised so that the affine paths are exercised; moving statistics start
    at (0, 1).  Tests calibrate them with an oracle pass (oracle/densenet_ref.py:calibrate_bn) so that
    activations stay O(1) through all 121 layers; for throughput measurements the values are irrelevant.
    """
    rng = np.random.default_rng(seed)
    w = {}
    for name, kind, shape in layer_shapes():
        if kind == "conv":
            fan_in = shape[0] * shape[1] * shape[2]
            w[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif kind == "bias":
            w[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:
            gamma = rng.uniform(0.8, 1.2, shape).astype(np.float32)
            beta = (0.1 * rng.standard_normal(shape)).astype(np.float32)
            w[name] = (gamma, beta, np.zeros(shape, np.float32), np.ones(shape, np.float32))
    return w


def densenet121_unet_program(weights: dict, patch: int = 256, fuse_dense: bool = True,
                             precision: str = "fp16") -> Program:
    """``precision``: 'fp16' = the tensor-core configuration (fp16 weights / activations, fp32 accumulation);
    'fp32' = same program with un-rounded weights for the runtime's fp32 kernels (program.py)."""
    with weight_precision(precision):
        pr = _build(weights, patch, fuse_dense)
    pr.precision = precision
    return pr


def _build(weights: dict, patch: int, fuse_dense: bool) -> Program:
    if patch < 64 or patch & (patch - 1):
        raise ValueError("patch_size must be a power of two >= 64 for the B200 tile kernels")
    P = patch
    pr = Program(patch=P)
    S = pr.add_buf("stem_s2d", P // 2, P // 2, 64)
    D1 = pr.add_buf("D1", P // 2, P // 2, 160)
    D2 = pr.add_buf("D2", P // 4, P // 4, 384)
    D3 = pr.add_buf("D3", P // 8, P // 8, 768)
    D4 = pr.add_buf("D4", P // 16, P // 16, 1344)
    D5 = pr.add_buf("D5", P // 32, P // 32, 1024)
    T = {b: pr.add_buf(f"T{b}", P // (2 ** b), P // (2 ** b), 128) for b in (2, 3, 4, 5)}
    Q = {2: pr.add_buf("Q2", P // 8, P // 8, 256), 3: pr.add_buf("Q3", P // 16, P // 16, 512),
         4: pr.add_buf("Q4", P // 32, P // 32, 1024)}
    N5 = pr.add_buf("N5", P // 32, P // 32, 1024)
    E6 = pr.add_buf("E6", P // 16, P // 16, 320)
    E7 = pr.add_buf("E7", P // 8, P // 8, 256)
    E8 = pr.add_buf("E8", P // 4, P // 4, 128)
    E9 = pr.add_buf("E9", P // 2, P // 2, 96)
    E10 = pr.add_buf("E10", P, P, 64)

    ops = pr.ops
    # ---- stem: pad3 + conv7x7/2 + BN + ReLU (densenet.py:116-120) as a 4x4 conv on the space-to-depth image;
    # the gather kernel unrolls the 4 column taps into 64 channels, the tensor-core kernel does the 4 row taps.
    # BatchNorm scales that follow a conv are folded into its weights in fp32 before the single rounding to
    # fp16 (same relative error as rounding the raw weights); the epilogue then only adds the shift.
    ops.append(Op(OP_STEM_S2D, out_buf=S, cout=64, name="stem_s2d"))
    sc, sh = bn_affine(*weights["conv1/bn"], EPS_ENC)
    ops.append(Op(OP_CONV, in_buf=S, cin=64, out_buf=D1, out_choff=96, cout=64, kind=KIND_STEM4, relu=1,
                  w=pack_stem4_weights(weights["conv1/conv"] * sc), epi_shift=sh, name="conv1"))
    # ---- pad1 + maxpool3/2 (densenet.py:122-123) straight into block2's concat buffer
    ops.append(Op(OP_MAXPOOL, in_buf=D1, in_choff=96, cin=64, out_buf=D2, out_choff=128, cout=64, name="pool1"))

    place = {2: (D2, 128, 64), 3: (D3, 256, 128), 4: (D4, 320, 256), 5: (D5, 0, 512)}  # buffer, base, C0
    for b, n in zip((2, 3, 4, 5), DENSENET_BLOCKS):
        D, base, c0 = place[b]
        for i in range(1, n + 1):
            p = f"conv{b}_block{i}"
            c = c0 + GROWTH * (i - 1)
            ps, psh = bn_affine(*weights[p + "_0_bn"], EPS_ENC)
            es, esh = bn_affine(*weights[p + "_1_bn"], EPS_ENC)
            if fuse_dense and (P // 2 ** b) >= 8:
                # whole dense layer in one kernel, bottleneck kept in shared memory (csrc/dense_layer.cuh)
                ops.append(Op(OP_DENSE_LAYER, in_buf=D, in_choff=base, cin=c, out_buf=D, out_choff=base + c,
                              cout=GROWTH, mid_buf=T[b], safe_cin=(c - GROWTH if i > 1 else 0), pro=PRO_AFFINE_RELU, pro_scale=pad64(ps), pro_shift=pad64(psh),
                              epi_shift=esh, w=pack_conv_weights(weights[p + "_1_conv"] * es, KIND_1X1),
                              w2=pack_conv_weights(weights[p + "_2_conv"], KIND_3X3), name=p))
                continue
            # BN-ReLU (pre-activation, in the A-tile prologue) -> 1x1 -> BN-ReLU (epilogue)   densenet.py:59-69
            ops.append(Op(OP_CONV, in_buf=D, in_choff=base, cin=c, out_buf=T[b], cout=128, kind=KIND_1X1, relu=1,
                          pro=PRO_AFFINE_RELU, pro_scale=pad64(ps), pro_shift=pad64(psh), epi_shift=esh,
                          w=pack_conv_weights(weights[p + "_1_conv"] * es, KIND_1X1), name=p + "_1_conv"))
            # 3x3 -> its 32 channels land at the tail of the concat buffer               densenet.py:70-74
            ops.append(Op(OP_CONV, in_buf=T[b], cin=128, out_buf=D, out_choff=base + c, cout=GROWTH, kind=KIND_3X3,
                          w=pack_conv_weights(weights[p + "_2_conv"], KIND_3X3), name=p + "_2_conv"))
        c = c0 + GROWTH * n
        if b < 5:
            # transition: BN-ReLU-1x1-avgpool (densenet.py:101-107); pool and 1x1 commute, pool first = 4x fewer MACs
            ts, tsh = bn_affine(*weights[f"pool{b}_bn"], EPS_ENC)
            ops.append(Op(OP_BNPOOL, in_buf=D, in_choff=base, cin=c, out_buf=Q[b], cout=c, relu=1, pool=1,
                          epi_scale=ts, epi_shift=tsh, name=f"pool{b}_bn_pool"))
            nD, nbase, nc0 = place[b + 1]
            assert nc0 == c // 2
            ops.append(Op(OP_CONV, in_buf=Q[b], cin=c, out_buf=nD, out_choff=nbase, cout=c // 2, kind=KIND_1X1,
                          w=pack_conv_weights(weights[f"pool{b}_conv"], KIND_1X1), name=f"pool{b}_conv"))
    # ---- final BN without ReLU (densenet.py:134); cannot be folded into dec6a because zero padding follows it
    fs, fsh = bn_affine(*weights["bn"], EPS_ENC)
    ops.append(Op(OP_BNPOOL, in_buf=D5, cin=1024, out_buf=N5, cout=1024, relu=0, pool=0, epi_scale=fs, epi_shift=fsh,
                  name="bn"))

    def dec(name, kind, ib, ioff, cin, ob, ooff, cout, head=False):
        g, be, mu, var = weights[name + "_norm"]
        s, t = bn_affine(g, be, mu, var, EPS_DEC)
        t = (t + weights[name + "_conv_bias"].astype(np.float32) * s).astype(np.float32)  # conv bias folded
        o = Op(OP_CONV, in_buf=ib, in_choff=ioff, cin=cin, out_buf=ob, out_choff=ooff, cout=cout, kind=kind, relu=1,
               epi_shift=t, w=pack_conv_weights(weights[name + "_conv"] * s, kind), name=name)
        if head:
            # Conv2D(2, 1x1, softmax) (densenet.py:156): only channel 1 is consumed downstream
            # (Segmentation.py:167), and softmax(z)[1] == sigmoid(z1 - z0).
            hk = weights["head"][0, 0].astype(np.float32)  # [64, 2]
            hb = weights["head_bias"].astype(np.float32)
            o.head = 1
            o.head_w = (hk[:, 1] - hk[:, 0]).astype(np.float32)
            o.head_b = float(hb[1] - hb[0])
        ops.append(o)

    dec("dec6a", KIND_UP2, N5, 0, 1024, D4, 0, 320)
    dec("dec6b", KIND_3X3, D4, 0, 1344, E6, 0, 320)
    dec("dec7a", KIND_UP2, E6, 0, 320, D3, 0, 256)
    dec("dec7b", KIND_3X3, D3, 0, 768, E7, 0, 256)
    dec("dec8a", KIND_UP2, E7, 0, 256, D2, 0, 128)
    dec("dec8b", KIND_3X3, D2, 0, 384, E8, 0, 128)
    dec("dec9a", KIND_UP2, E8, 0, 128, D1, 0, 96)
    dec("dec9b", KIND_3X3, D1, 0, 160, E9, 0, 96)
    dec("dec10a", KIND_UP2, E9, 0, 96, E10, 0, 64)
    dec("dec10b", KIND_3X3, E10, 0, 64, E10, 0, 64, head=True)
    return pr


def reference_macs_per_tile(patch: int = 256) -> int:
    """Conv MACs of one unet_densenet121 forward as the reference graph states it (no sub-pixel rewrite)."""
    P = patch
    macs = (P // 2) ** 2 * 147 * 64
    c, hw = 64, (P // 4) ** 2
    for b, n in zip((2, 3, 4, 5), DENSENET_BLOCKS):
        for i in range(n):
            macs += hw * (c * 128 + 9 * 128 * 32)
            c += 32
        if b < 5:
            macs += hw * c * (c // 2)
            c //= 2
            hw //= 4
    for name, cin, cout in DECODER:
        hw = hw * 4 if name.endswith("a") else hw
        macs += hw * 9 * cin * cout
    macs += P * P * 64 * 2
    return macs
