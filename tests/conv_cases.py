"""Single-op layer programs used by the GPU parity tests and the bring-up script (test infrastructure)."""
from __future__ import annotations

import numpy as np

from digipathai_b200.program import (KIND_1X1, KIND_3X3, KIND_TAPS, KIND_UP2, OP_CONV, OP_DENSE_LAYER,
                                     PRO_AFFINE_RELU, Op, Program, pack_conv_weights, pad64)

# name: (kind, H, W, Cin, Cout, B, in_ctot, in_choff, out_ctot, out_choff, prologue, relu, head)
CASES = {
    "d_1x1_plain":   (KIND_1X1, 16, 16, 64, 128, 2, 64, 0, 128, 0, False, True, False),
    "d_1x1_pro_tail": (KIND_1X1, 16, 16, 96, 128, 3, 160, 32, 128, 0, True, True, False),
    "d_1x1_stem":    (KIND_1X1, 32, 32, 160, 64, 2, 160, 0, 160, 96, False, True, False),
    "d_1x1_wide":    (KIND_1X1, 8, 8, 1024, 512, 4, 1024, 0, 1024, 0, False, False, False),
    "d_1x1_msub":    (KIND_1X1, 64, 64, 64, 128, 12, 64, 0, 128, 0, True, True, False),
    "t_3x3_8x8":     (KIND_3X3, 8, 8, 128, 32, 5, 128, 0, 1024, 512, False, False, False),
    "t_up2_8x8":     (KIND_UP2, 8, 8, 128, 320, 4, 128, 0, 384, 0, False, True, False),
    "h_3x3_16":      (KIND_3X3, 16, 16, 128, 32, 3, 128, 0, 1344, 320, False, False, False),
    "h_3x3_64":      (KIND_3X3, 64, 64, 64, 64, 2, 64, 0, 64, 0, False, True, False),
    "h_3x3_big":     (KIND_3X3, 32, 32, 768, 256, 6, 768, 0, 256, 0, False, True, False),
    "h_3x3_n320":    (KIND_3X3, 16, 16, 96, 320, 2, 96, 0, 320, 0, False, True, False),
    "h_up2_16":      (KIND_UP2, 16, 16, 64, 64, 2, 64, 0, 96, 32, False, True, False),
    "h_up2_n256":    (KIND_UP2, 16, 16, 320, 256, 2, 320, 0, 768, 0, False, True, False),
    "h_up2_n96":     (KIND_UP2, 32, 32, 128, 96, 2, 128, 0, 160, 0, False, True, False),
    "h_head":        (KIND_3X3, 64, 64, 64, 64, 3, 64, 0, 64, 0, False, True, True),
}


# Inception-ResNet-v2 shapes (models/inception.py): generic tap tables, stride 2, residual / clipped-N epilogues,
# channel tails of 16 / 48, N tiles that are odd multiples of 16.
# name: (kind, H, W, Cin, Cout, B, in_ctot, in_choff, out_ctot, out_choff, relu, (kh, kw, stride), residual)
TAP_CASES = {
    "g_1x7_16":        (KIND_TAPS, 16, 16, 128, 160, 3, 672, 0, 672, 512, True, (1, 7, 1), False),
    "g_7x1_16":        (KIND_TAPS, 16, 16, 160, 192, 3, 672, 512, 672, 320, True, (7, 1, 1), False),
    "g_5x5_32":        (KIND_TAPS, 32, 32, 48, 64, 2, 112, 0, 576, 352, True, (5, 5, 1), False),
    "g_3x3_as_taps":   (KIND_TAPS, 16, 16, 128, 32, 3, 128, 0, 64, 32, False, (3, 3, 1), False),
    "g_1x3_8":         (KIND_TAPS, 8, 8, 192, 224, 4, 864, 0, 864, 640, True, (1, 3, 1), False),
    "g_3x1_8":         (KIND_TAPS, 8, 8, 224, 256, 3, 864, 640, 864, 384, True, (3, 1, 1), False),
    "g_1x7_4":         (KIND_TAPS, 4, 4, 128, 160, 9, 672, 0, 672, 512, True, (1, 7, 1), False),
    "g_3x3s2_32":      (KIND_TAPS, 32, 32, 320, 384, 3, 576, 256, 1408, 320, True, (3, 3, 2), False),
    "g_3x3s2_16":      (KIND_TAPS, 16, 16, 256, 288, 5, 768, 256, 2080, 384, True, (3, 3, 2), False),
    "g_3x3s2_8":       (KIND_TAPS, 8, 8, 64, 64, 8, 64, 0, 64, 0, True, (3, 3, 2), False),
    "d_res_1088":      (KIND_1X1, 16, 16, 384, 1088, 3, 672, 128, 1408, 320, True, (1, 1, 1), True),
    "d_res_2080_lin":  (KIND_1X1, 8, 8, 448, 2080, 4, 864, 192, 2080, 0, False, (1, 1, 1), True),
    "d_res_320":       (KIND_1X1, 32, 32, 128, 320, 2, 240, 64, 576, 256, True, (1, 1, 1), True),
    "d_1x1_n80":       (KIND_1X1, 64, 64, 64, 80, 2, 64, 0, 80, 0, True, (1, 1, 1), False),
    "d_1x1_n112":      (KIND_1X1, 32, 32, 192, 112, 2, 192, 0, 112, 0, True, (1, 1, 1), False),
    "d_1x1_n768":      (KIND_1X1, 16, 16, 1088, 768, 2, 1408, 320, 768, 0, True, (1, 1, 1), False),
    "h_3x3_c80_n192":  (KIND_3X3, 64, 64, 80, 192, 2, 80, 0, 320, 128, True, (3, 3, 1), False),
    "h_3x3_c32":       (KIND_3X3, 32, 32, 32, 32, 2, 240, 0, 240, 96, True, (3, 3, 1), False),
    "h_3x3_c48_n64":   (KIND_3X3, 32, 32, 48, 64, 2, 240, 192, 240, 128, True, (3, 3, 1), False),
    "h_3x3_n288":      (KIND_3X3, 16, 16, 256, 288, 2, 768, 512, 288, 0, True, (3, 3, 1), False),
    "t_up2_8_c1536":   (KIND_UP2, 8, 8, 1536, 320, 2, 1536, 0, 1408, 0, True, (3, 3, 1), False),
}


def build_tap_case(name: str, seed: int = 0):
    """-> (program, input fp16 [B,H,W,in_ctot], initial output-buffer content fp16, B)."""
    kind, H, W, cin, cout, B, ictot, ioff, octot, ooff, relu, (kh, kw, stride), residual = TAP_CASES[name]
    rng = np.random.default_rng(seed)
    pr = Program(patch=64)
    ib = pr.add_buf("in", H, W, ictot)
    up = 2 if kind == KIND_UP2 else 1
    OH, OW = (H * up, W * up) if stride == 1 else (H // 2, W // 2)
    ob = pr.add_buf("out", OH, OW, octot)
    kern = (rng.standard_normal((kh, kw, cin, cout)) * np.sqrt(2.0 / (kh * kw * cin))).astype(np.float32)
    op = Op(OP_CONV, in_buf=ib, in_choff=ioff, cin=cin, out_buf=ob, out_choff=ooff, cout=cout, kind=kind,
            relu=int(relu), w=pack_conv_weights(kern, kind), name=name, residual=int(residual),
            epi_shift=(0.2 * rng.standard_normal(cout)).astype(np.float32))
    if kind == KIND_TAPS:
        op.kh, op.kw, op.stride = kh, kw, stride
    pr.ops.append(op)
    x = rng.standard_normal((B, H, W, ictot)).astype(np.float16)
    out0 = rng.standard_normal((B, OH, OW, octot)).astype(np.float16)
    return pr, x, out0, B


def build_case(name: str, seed: int = 0):
    kind, H, W, cin, cout, B, ictot, ioff, octot, ooff, pro, relu, head = CASES[name]
    rng = np.random.default_rng(seed)
    patch = H if head else 64
    pr = Program(patch=patch)
    ib = pr.add_buf("in", H, W, ictot)
    up = 2 if kind == KIND_UP2 else 1
    ob = pr.add_buf("out", H * up, W * up, octot)
    k = {KIND_1X1: 1, KIND_3X3: 3, KIND_UP2: 3}[kind]
    kern = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    op = Op(OP_CONV, in_buf=ib, in_choff=ioff, cin=cin, out_buf=ob, out_choff=ooff, cout=cout, kind=kind,
            relu=int(relu), w=pack_conv_weights(kern, kind), name=name,
            epi_scale=rng.uniform(0.5, 1.5, cout).astype(np.float32),
            epi_shift=(0.2 * rng.standard_normal(cout)).astype(np.float32))
    if pro:
        op.pro = PRO_AFFINE_RELU
        op.pro_scale = pad64(rng.uniform(0.5, 1.5, cin).astype(np.float32))
        op.pro_shift = pad64((0.3 * rng.standard_normal(cin)).astype(np.float32))
    if head:
        op.head = 1
        op.head_w = (rng.standard_normal(cout) * 0.3).astype(np.float32)
        op.head_b = 0.1
    pr.ops.append(op)
    x = rng.standard_normal((B, H, W, ictot)).astype(np.float16)
    return pr, x, B


# fused dense layer: name -> (H, W, Cin, B, ctot, base)
DENSE_CASES = {
    "dl_16x16_c256": (16, 16, 256, 3, 1344, 320),
    "dl_32x32_c160": (32, 32, 160, 2, 768, 256),     # channel tail chunk (160 = 2.5 x 64)
    "dl_64x64_c64":  (64, 64, 64, 2, 384, 128),
    "dl_16x16_c992": (16, 16, 992, 5, 1344, 320),    # last layer of block 4, odd batch
    "dl_8x8_c512":   (8, 8, 512, 4, 1024, 0),        # block 5: 8-row regions
    "dl_64x64_c96_b12": (64, 64, 96, 12, 384, 128),  # enough regions to keep 16-row regions (RH = 16)
}


def build_dense_case(name: str, seed: int = 0):
    H, W, cin, B, ctot, base = DENSE_CASES[name]
    rng = np.random.default_rng(seed)
    pr = Program(patch=64)
    D = pr.add_buf("D", H, W, ctot)
    T = pr.add_buf("T", H, W, 128)
    k1 = (rng.standard_normal((1, 1, cin, 128)) * np.sqrt(2.0 / cin)).astype(np.float32)
    k3 = (rng.standard_normal((3, 3, 128, 32)) * np.sqrt(2.0 / (9 * 128))).astype(np.float32)
    op = Op(OP_DENSE_LAYER, in_buf=D, in_choff=base, cin=cin, out_buf=D, out_choff=base + cin, cout=32, mid_buf=T,
            pro=PRO_AFFINE_RELU, pro_scale=pad64(rng.uniform(0.5, 1.5, cin).astype(np.float32)),
            pro_shift=pad64((0.3 * rng.standard_normal(cin)).astype(np.float32)),
            epi_shift=(0.2 * rng.standard_normal(128)).astype(np.float32),
            w=pack_conv_weights(k1, KIND_1X1), w2=pack_conv_weights(k3, KIND_3X3), name=name)
    pr.ops.append(op)
    x = rng.standard_normal((B, H, W, ctot)).astype(np.float16)
    return pr, x, B
