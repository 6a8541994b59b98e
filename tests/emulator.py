"""CPU emulator of the layer program (digipathai_b200/program.py) -- test infrastructure.

Executes exactly what csrc/runtime.cu executes (same tap tables, same packed fp16 weights, same buffer plan,
same fp16 rounding points) with torch-CPU fp32 arithmetic.  It validates the graph builder / weight packer
against the oracle without a GPU, and predicts the fp16-storage error the CUDA path should show.
"""
from __future__ import annotations

import numpy as np
import torch

from digipathai_b200 import tta
from digipathai_b200.program import (KIND_1X1, KIND_3X3, KIND_STEM4, KIND_TAPS, KIND_UP2, OP_AVGPOOL3, OP_BCAST,
                                     OP_BNPOOL, OP_CONV, OP_DWCONV, OP_GAP, OP_HEAD_DOT, OP_HEAD_RESIZE, OP_MAXPOOL,
                                     OP_DENSE_LAYER, OP_RESIZE, OP_STEM_IM2COL, OP_STEM_S2D, POOL_TF_SAME, Program,
                                     tap_offsets)


def entries(kind, op=None, H=0, W=0):
    if kind == KIND_TAPS:
        return [(dy, dx, 0) for dy, dx in tap_offsets(op.kh, op.kw, op.stride, H, W)]
    if kind == KIND_1X1:
        return [(0, 0, 0)]
    if kind == KIND_3X3:
        return [(ky - 1, kx - 1, 0) for ky in range(3) for kx in range(3)]
    if kind == KIND_STEM4:
        return [(t - 2, 0, 0) for t in range(4)]
    out = []
    for ph in range(4):
        a, b = ph >> 1, ph & 1
        for t in range(4):
            ty, tx = t >> 1, t & 1
            out.append((a - 1 + ty, b - 1 + tx, ph))
    return out


def _shift(x, dy, dx):
    """x[n,h,w,c] -> y[n,h,w,c] = x[n,h+dy,w+dx,c] with zero fill."""
    n, h, w, c = x.shape
    y = torch.zeros_like(x)
    hs, he = max(0, -dy), min(h, h - dy)
    ws, we = max(0, -dx), min(w, w - dx)
    y[:, hs:he, ws:we] = x[:, hs + dy:he + dy, ws + dx:we + dx]
    return y


def conv_eval(op, x: torch.Tensor) -> torch.Tensor:
    """x fp32 [n,h,w,cin] (already pre-activated / quantised) -> fp32 conv output incl. BN affine, BEFORE the
    residual add / ReLU / output rounding.  Stride-2 'same' convs (KIND_TAPS) subsample the stride-1 tap sums."""
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    w = torch.from_numpy(op.w.astype(np.float32))  # [e, co, ci]
    n, h, wd, _ = x.shape
    ents = entries(op.kind, op, h, wd)
    ng = 4 if op.kind == KIND_UP2 else 1
    s = op.stride if op.kind == KIND_TAPS else 1
    acc = torch.zeros(ng, n, -(-h // s), -(-wd // s), op.cout)
    for e, (dy, dx, g) in enumerate(ents):
        acc[g] += (_shift(x, dy, dx)[:, ::s, ::s] @ w[e].T)
    sc = f(op.epi_scale) if op.epi_scale is not None else torch.ones(op.cout)
    sh = f(op.epi_shift) if op.epi_shift is not None else torch.zeros(op.cout)
    acc = acc * sc + sh
    if op.kind == KIND_UP2:
        y = torch.zeros(n, 2 * h, 2 * wd, op.cout)
        for g in range(4):
            y[:, (g >> 1)::2, (g & 1)::2] = acc[g]
        return y
    return acc[0]


def maxpool_same(x_nchw: torch.Tensor) -> torch.Tensor:
    """MaxPooling2D(3, strides=2, padding='same') with TensorFlow's padding rule (padded cells never win)."""
    h, w = x_nchw.shape[2:]
    from digipathai_b200.program import same_pad_before
    ph, pw = same_pad_before(h, 3, 2), same_pad_before(w, 3, 2)
    oh, ow = -(-h // 2), -(-w // 2)
    pad = (pw, (ow - 1) * 2 + 3 - w - pw, ph, (oh - 1) * 2 + 3 - h - ph)
    return torch.nn.functional.max_pool2d(torch.nn.functional.pad(x_nchw, pad, value=float("-inf")), 3, stride=2)


def dwconv_eval(op, x: torch.Tensor) -> torch.Tensor:
    """[ReLU] -> depthwise 3x3 (stride, dilation; explicit symmetric padding as deeplabv3.py:62-71) + shift [-> ReLU]
    on fp32 [n,h,w,c]; weights are the packed fp16 [9][C] with the BN scale folded in."""
    c = op.cin
    if op.pro:
        x = torch.relu(x)
    w = torch.from_numpy(op.w.astype(np.float32)).reshape(3, 3, c).permute(2, 0, 1).unsqueeze(1).contiguous()
    pad = op.rate  # (3 + 2 (rate - 1) - 1) / 2, both strides
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, None, stride=op.stride, padding=pad, dilation=op.rate,
                                   groups=c).permute(0, 2, 3, 1)
    y = y + torch.from_numpy(np.asarray(op.epi_shift, np.float32))
    return torch.relu(y) if op.relu else y


def resize_bilinear_ac(x_nhwc: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """tf.compat.v1.image.resize(method='bilinear', align_corners=True)."""
    return torch.nn.functional.interpolate(x_nhwc.permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear",
                                           align_corners=True).permute(0, 2, 3, 1)


def stem_im2col(tiles_u8: np.ndarray, code: int) -> torch.Tensor:
    """tiles uint8 [B,P,P,3] (reference [x,y,c] orientation) -> fp32 [B,P/2,P/2,160]."""
    B, P = tiles_u8.shape[:2]
    x = torch.from_numpy(np.stack([tta.apply(code, t) for t in tiles_u8]).astype(np.float32))
    x = (x - 128.0) / 128.0
    xp = torch.zeros(B, P + 6, P + 6, 3)
    xp[:, 3:P + 3, 3:P + 3] = x
    cols = []
    for ky in range(7):
        for kx in range(7):
            cols.append(xp[:, ky:ky + P:2, kx:kx + P:2, :])
    out = torch.cat(cols, dim=-1)
    return torch.cat([out, torch.zeros(B, P // 2, P // 2, 160 - 147)], dim=-1)


def stem_s2d(tiles_u8: np.ndarray, code: int) -> torch.Tensor:
    """tiles uint8 [B,P,P,3] -> fp32 [B,P/2,P/2,64] exactly as stem_s2d_kernel lays it out."""
    B, P = tiles_u8.shape[:2]
    x = torch.from_numpy(np.stack([tta.apply(code, t) for t in tiles_u8]).astype(np.float32))
    x = (x - 128.0) / 128.0
    OH = P // 2
    out = torch.zeros(B, OH, OH, 64)
    for dq in range(4):
        for a in range(2):
            for b in range(2):
                # columns jq = q + dq - 2 in [0, OH)
                q_lo, q_hi = max(0, 2 - dq), min(OH, OH + 2 - dq)
                src = x[:, a::2, b::2, :]                      # [B, OH, OH, 3] indexed [r][jq]
                ch = dq * 16 + (a * 2 + b) * 3
                out[:, :, q_lo:q_hi, ch:ch + 3] = src[:, :, q_lo + dq - 2:q_hi + dq - 2, :]
    return out


def run(prog: Program, tiles_u8: np.ndarray, tta_in: int = 0, tta_out: int = 0, fp16_storage: bool = True,
        keep: bool = False):
    """Returns probs float32 [B,P,P] (and the buffer dict when keep=True)."""
    B, P = tiles_u8.shape[0], prog.patch
    q = (lambda t: t.half().float()) if fp16_storage else (lambda t: t)
    bufs = [torch.zeros(B, h, w, c) for (h, w, c) in prog.bufs]
    probs = None
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    for op in prog.ops:
        if op.type == OP_STEM_IM2COL:
            bufs[op.out_buf][:] = q(stem_im2col(tiles_u8, tta_in))
        elif op.type == OP_STEM_S2D:
            bufs[op.out_buf][:] = q(stem_s2d(tiles_u8, tta_in))
        elif op.type == OP_MAXPOOL:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin].permute(0, 3, 1, 2)
            if op.pool == POOL_TF_SAME:
                y = maxpool_same(x)
            else:
                y = torch.nn.functional.max_pool2d(torch.nn.functional.pad(x, (1, 1, 1, 1)), 3, stride=2)
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = y.permute(0, 2, 3, 1)
        elif op.type == OP_AVGPOOL3:
            # AveragePooling2D(3, strides=1, padding='same'): mean over the VALID cells of each window
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin].permute(0, 3, 1, 2)
            y = torch.nn.functional.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False)
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = q(y.permute(0, 2, 3, 1))
        elif op.type == OP_DWCONV:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = q(dwconv_eval(op, x))
        elif op.type == OP_GAP:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = q(x.mean(dim=(1, 2), keepdim=True))
        elif op.type == OP_BCAST:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = x   # broadcasts over H, W
        elif op.type == OP_RESIZE:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            oh, ow = prog.bufs[op.out_buf][:2]
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = q(resize_bilinear_ac(x, oh, ow))
        elif op.type == OP_HEAD_DOT:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            head_logit = x @ f(op.head_w) + op.head_b          # fp32 [n,h,w], kept in fp32 (not a fp16 buffer)
        elif op.type == OP_HEAD_RESIZE:
            z = resize_bilinear_ac(head_logit.unsqueeze(-1), P, P)[..., 0]
            pr = torch.sigmoid(z).numpy()
            probs = np.stack([tta.apply(tta.inverse(tta_out), t) for t in pr])
        elif op.type == OP_BNPOOL:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            y = x * f(op.epi_scale) + f(op.epi_shift)
            if op.relu:
                y = torch.relu(y)
            if op.pool:
                y = torch.nn.functional.avg_pool2d(y.permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1)
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cin] = q(y)
        elif op.type == OP_DENSE_LAYER:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            x = q(torch.relu(x * f(op.pro_scale[:op.cin]) + f(op.pro_shift[:op.cin])))
            t = q(torch.relu(x @ torch.from_numpy(op.w[0].astype(np.float32)).T + f(op.epi_shift)))
            w2 = torch.from_numpy(op.w2.astype(np.float32))
            y = torch.zeros(*t.shape[:3], op.cout)
            for e, (dy, dx, g) in enumerate(entries(KIND_3X3)):
                y += _shift(t, dy, dx) @ w2[e].T
            bufs[op.mid_buf][:] = t
            bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cout] = q(y)
        elif op.type == OP_CONV:
            x = bufs[op.in_buf][..., op.in_choff:op.in_choff + op.cin]
            if op.pro:
                x = x * f(op.pro_scale[:op.cin]) + f(op.pro_shift[:op.cin])
                if op.pro == 2:
                    x = torch.relu(x)
                x = q(x)
            y = conv_eval(op, x)
            if op.residual:
                y = y + bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cout]
            if op.relu:
                y = torch.relu(y)
            if op.head:
                z = y @ f(op.head_w) + op.head_b
                pr = torch.sigmoid(z).numpy()
                probs = np.stack([tta.apply(tta.inverse(tta_out), t) for t in pr])
            else:
                bufs[op.out_buf][..., op.out_choff:op.out_choff + op.cout] = q(y)
        else:
            raise ValueError(op.type)
    return (probs, bufs) if keep else probs
