"""Layer program + flat model container ("DPB1").

The reference builds Keras graphs in Python (DigiPathAI/models/*.py) and loads ``.h5`` weights into them
(DigiPathAI/helpers/utils.py:427-448).  Here a graph builder emits a *layer program* -- a list of ops over
NHWC fp16 buffers -- and this module packs it, together with the fp16 weights and the folded BatchNorm
affine terms, into one flat little-endian container that ``dp_model_create`` (csrc/runtime.cu) executes.

Container layout (must match ``BlobHeader`` / ``BlobBuf`` / ``BlobOp`` in csrc/runtime.cu)::

    header  72 B : 'DPB1' | u32 version=1 | u32 n_bufs | u32 n_ops | u32 patch | u32 precision | u64 data_off
                   | u64 total_bytes | 4 x u64 0        (precision 0: fp16 weights + activations, 1: fp32 both, 2: fp32 both + 3xTF32 convs)
    n_bufs x 16 B: i32 H, W, C, 0                      (per-image activation buffer geometry)
    n_ops  x 128 B: 12 x i32 (type in_buf in_choff cin out_buf out_choff cout kind relu pro head pool)
                   | f32 head_b | 3 x i32 0 | 8 x i64 offsets into the data section (-1 = absent):
                     w, epi_scale, epi_shift, pro_scale, pro_shift, head_w, w2, 0
                   (rsv i32 #0 = mid_buf, #1 = safe_cin for OP_DENSE_LAYER;
                    OP_CONV: #0 = kh | kw << 8 | stride << 16 for KIND_TAPS, #1 = residual flag)
    data section : 256-byte aligned arrays (weights [entries][Cout][Cin] in the program's precision; fp32 vectors)

Precision.  ``precision='fp16'`` (default) is the tensor-core configuration BASELINE.json names: fp16 weights and
activations, fp32 accumulation.  ``precision='fp32'`` packs the same program with un-rounded fp32 weights; the
runtime then keeps fp32 activations and runs its fp32 kernels (csrc/precise.cuh) -- the mode that meets the 1e-3
probability tolerance on any weight set, at a fraction of the throughput.  ``precision='tf32x3'`` is the same fp32
program with its convs on the tensor cores: every operand split into two TF32 numbers, three tcgen05 MMAs per product
(csrc/precise_tc.cuh; ~2^-21 relative per product instead of fp32's 2^-24).
"""
from __future__ import annotations

import struct
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

OP_STEM_IM2COL, OP_MAXPOOL, OP_CONV, OP_BNPOOL, OP_STEM_S2D, OP_DENSE_LAYER, OP_AVGPOOL3 = 1, 2, 3, 4, 5, 6, 7
# DeepLabv3+ (models/deeplab.py): depthwise 3x3 (stride / dilation), global average pool, broadcast of a 1x1 map,
# bilinear align_corners resize, logit-difference dot product, and resize + sigmoid + inverse-TTA head
OP_DWCONV, OP_GAP, OP_BCAST, OP_RESIZE, OP_HEAD_DOT, OP_HEAD_RESIZE = 8, 9, 10, 11, 12, 13
KIND_1X1, KIND_3X3, KIND_UP2, KIND_STEM4, KIND_TAPS = 1, 3, 4, 5, 6
POOL_PAD1_ZERO, POOL_TF_SAME = 0, 1   # OP_MAXPOOL `pool` field: ZeroPadding2D(1)+valid (densenet.py:122-123) / padding='same'
PRO_NONE, PRO_AFFINE, PRO_AFFINE_RELU = 0, 1, 2
PRECISIONS = {"fp16": 0, "fp32": 1, "tf32x3": 2}

_weight_dtype = [np.float16]


@contextmanager
def weight_precision(precision: str):
    """Graph builders run inside this context: the pack_* helpers then round weights to the program's storage type
    (fp16: one rounding after the BN scale has been folded in fp32; fp32: none)."""
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    _weight_dtype.append(np.float16 if precision == "fp16" else np.float32)
    try:
        yield
    finally:
        _weight_dtype.pop()


def wdtype():
    """Storage type of packed weights in the active ``weight_precision`` context (fp16 outside any)."""
    return _weight_dtype[-1]


@dataclass
class Op:
    type: int
    in_buf: int = 0
    in_choff: int = 0
    cin: int = 0
    out_buf: int = 0
    out_choff: int = 0
    cout: int = 0
    kind: int = 0
    relu: int = 0
    pro: int = 0
    head: int = 0
    pool: int = 0
    head_b: float = 0.0
    w: Optional[np.ndarray] = None          # fp16 / fp32 [entries, cout, cin]
    epi_scale: Optional[np.ndarray] = None  # fp32 [cout]  (BNPOOL: [cin])
    epi_shift: Optional[np.ndarray] = None
    pro_scale: Optional[np.ndarray] = None  # fp32 [ceil(cin/64)*64]
    pro_shift: Optional[np.ndarray] = None
    head_w: Optional[np.ndarray] = None     # fp32 [cout]
    w2: Optional[np.ndarray] = None         # OP_DENSE_LAYER: fp16 [9, 32, 128] 3x3 weights (w = [1, 128, cin])
    mid_buf: int = 0                        # OP_DENSE_LAYER: bottleneck buffer (only the debug path writes it)
    safe_cin: int = 0                       # OP_DENSE_LAYER: leading input channels NOT written by the preceding op
    kh: int = 0                             # KIND_TAPS: kernel height / width / stride ('same' padding, TF rule)
    kw: int = 0
    stride: int = 1
    residual: int = 0                       # OP_CONV: out = act(out_old + conv + shift), in place (inception.py:152-160)
    rate: int = 1                           # OP_DWCONV: dilation rate (stride in `stride`, pre-ReLU in `pro`, post-ReLU in `relu`)
    name: str = ""


@dataclass
class Program:
    patch: int
    bufs: List[tuple] = field(default_factory=list)  # (H, W, C)
    ops: List[Op] = field(default_factory=list)
    buf_names: List[str] = field(default_factory=list)
    precision: str = "fp16"

    def add_buf(self, name: str, h: int, w: int, c: int) -> int:
        assert c % 8 == 0, "channel strides must keep 16-byte alignment"
        self.bufs.append((h, w, c))
        self.buf_names.append(name)
        return len(self.bufs) - 1

    def buf(self, name: str) -> int:
        return self.buf_names.index(name)

    def op_index(self, name: str) -> int:
        return [o.name for o in self.ops].index(name)


# ----------------------------------------------------------------------------------------------- packing
_UP2_ROWS = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}


def pack_conv_weights(k_hwio: np.ndarray, kind: int) -> np.ndarray:
    """HWIO Keras kernel -> fp16 [entries][Cout][Cin] in the tap order of csrc/runtime.cu:fill_entries.

    KIND_UP2 implements ``UpSampling2D()`` followed by a 3x3 'same' conv (densenet.py:138-155) as four
    sub-pixel phases: output pixel (2h+a, 2w+b) only ever sees input rows {h+a-1, h+a} and columns
    {w+b-1, w+b}, so the 3x3 kernel collapses to a 2x2 kernel per phase whose taps are sums of the original
    taps (summed in fp32, rounded once to fp16).
    """
    k = np.asarray(k_hwio, dtype=np.float32)
    if kind == KIND_1X1:
        assert k.shape[:2] == (1, 1)
        w = k[0, 0].T[None]
    elif kind == KIND_3X3:
        assert k.shape[:2] == (3, 3)
        w = np.stack([k[ky, kx].T for ky in range(3) for kx in range(3)])
    elif kind == KIND_UP2:
        assert k.shape[:2] == (3, 3)
        ents = []
        for ph in range(4):
            a, b = ph >> 1, ph & 1
            for t in range(4):
                ty, tx = t >> 1, t & 1
                acc = np.zeros(k.shape[2:], dtype=np.float32)
                for ky in _UP2_ROWS[(a, ty)]:
                    for kx in _UP2_ROWS[(b, tx)]:
                        acc += k[ky, kx]
                ents.append(acc.T)
        w = np.stack(ents)
    elif kind == KIND_TAPS:
        kh, kw = k.shape[:2]
        w = np.stack([k[ky, kx].T for ky in range(kh) for kx in range(kw)])
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(w).astype(wdtype())


def same_pad_before(size: int, k: int, stride: int) -> int:
    """Leading padding of TensorFlow's padding='same' along one axis (extra cell goes after)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2


def tap_offsets(kh: int, kw: int, stride: int, H: int, W: int):
    """(dy, dx) of every tap of a 'same' conv, ky-major: input pixel = stride * output pixel + (dy, dx)."""
    ph, pw = same_pad_before(H, kh, stride), same_pad_before(W, kw, stride)
    return [(ky - ph, kx - pw) for ky in range(kh) for kx in range(kw)]


def pack_stem_weights(k_hwio: np.ndarray, kpad: int = 160) -> np.ndarray:
    """7x7x3xCout stem kernel -> fp16 [1][Cout][kpad]; column (ky*7+kx)*3+c matches stem_im2col_kernel."""
    k = np.asarray(k_hwio, dtype=np.float32)
    kh, kw, ci, co = k.shape
    w = np.zeros((1, co, kpad), dtype=np.float32)
    w[0, :, : kh * kw * ci] = k.reshape(kh * kw * ci, co).T
    return w.astype(wdtype())


def pack_stem4_weights(k_hwio: np.ndarray, pad: int = 3) -> np.ndarray:
    """k x k x 3 x Cout stride-2 stem kernel (k <= 7) -> fp16 [4 row taps][Cout][64] for the space-to-depth stem
    (stem_s2d_kernel).  ``pad`` = leading zero padding: 3 for ZeroPadding2D(3)+7x7/2 (densenet.py:116-117), 0 for
    a 3x3/2 padding='same' conv on an even-sized tile (inception.py:174).

    Row tap t <-> dr = t-2; input channel dq*16 + (a*2+b)*3 + c multiplies original tap
    (ky, kx) = (2*dr + a + pad, 2*(dq-2) + b + pad); combinations that fall outside the kernel are zero.
    """
    k = np.asarray(k_hwio, dtype=np.float32)
    assert k.shape[2] == 3 and k.shape[0] == k.shape[1]
    K = k.shape[0]
    assert pad <= 4 and K - 1 - pad <= 3, "kernel does not fit the 4x4 space-to-depth window"
    co = k.shape[3]
    w = np.zeros((4, co, 64), dtype=np.float32)
    for t in range(4):
        for dq in range(4):
            for a in range(2):
                for b in range(2):
                    ky, kx = 2 * (t - 2) + a + pad, 2 * (dq - 2) + b + pad
                    if 0 <= ky < K and 0 <= kx < K:
                        for c in range(3):
                            w[t, :, dq * 16 + (a * 2 + b) * 3 + c] = k[ky, kx, c, :]
    return w.astype(wdtype())


def bn_affine(gamma, beta, mean, var, eps):
    """Inference BatchNorm as y = scale*x + shift (Keras: gamma*(x-mean)/sqrt(var+eps)+beta)."""
    scale = (np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps))
    shift = np.asarray(beta, np.float64) - np.asarray(mean, np.float64) * scale
    return scale.astype(np.float32), shift.astype(np.float32)


def pad64(v: np.ndarray) -> np.ndarray:
    n = (len(v) + 63) // 64 * 64
    out = np.zeros(n, dtype=np.float32)
    out[: len(v)] = v
    return out


def serialize(prog: Program) -> bytes:
    data = bytearray()
    wdt = np.float16 if prog.precision == "fp16" else np.float32
    for o in prog.ops:
        for a in (o.w, o.w2):
            assert a is None or a.dtype == wdt, (o.name, a.dtype, prog.precision)

    def put(arr: Optional[np.ndarray], dtype) -> int:
        if arr is None:
            return -1
        a = np.ascontiguousarray(arr, dtype=dtype)
        pad = (-len(data)) % 256
        data.extend(b"\0" * pad)
        off = len(data)
        data.extend(a.tobytes())
        return off

    op_recs = []
    for o in prog.ops:
        if o.type == OP_CONV:
            ents = {KIND_1X1: 1, KIND_3X3: 9, KIND_UP2: 16, KIND_STEM4: 4, KIND_TAPS: o.kh * o.kw}[o.kind]
            if o.kind == KIND_TAPS:
                assert 1 <= ents <= 32 and o.stride in (1, 2), o.name
            assert o.w is not None and o.w.shape == (ents, o.cout, o.cin), (o.name, o.w.shape, (ents, o.cout, o.cin))
            assert o.cout % 16 == 0 and o.cin % 8 == 0, o.name
            if o.pro:
                assert o.pro_scale is not None and len(o.pro_scale) == (o.cin + 63) // 64 * 64, o.name
            if o.head:
                assert o.head_w is not None and len(o.head_w) == o.cout, o.name
        offs = [
            put(o.w, wdt), put(o.epi_scale, np.float32), put(o.epi_shift, np.float32),
            put(o.pro_scale, np.float32), put(o.pro_shift, np.float32), put(o.head_w, np.float32),
            put(o.w2, wdt), 0,
        ]
        if o.type == OP_DWCONV:
            assert o.w is not None and o.w.shape == (9, o.cin) and o.epi_shift is not None and o.cin % 8 == 0, o.name
            assert o.stride in (1, 2) and 1 <= o.rate < 256, o.name
        if o.type == OP_DENSE_LAYER:
            assert o.w is not None and o.w.shape == (1, 128, o.cin) and o.w2 is not None and o.w2.shape == (9, 32, 128)
            assert o.cout == 32 and o.in_buf == o.out_buf and o.pro_scale is not None and o.epi_shift is not None
        op_recs.append(
            struct.pack(
                "<12if3i8q", o.type, o.in_buf, o.in_choff, o.cin, o.out_buf, o.out_choff, o.cout, o.kind, o.relu,
                o.pro, o.head, o.pool, float(o.head_b),
                (o.kh | (o.kw << 8) | (o.stride << 16)) if o.type == OP_CONV else
                ((o.stride | (o.rate << 8)) if o.type == OP_DWCONV else o.mid_buf),
                o.residual if o.type == OP_CONV else o.safe_cin, 0, *offs,
            )
        )
    buf_recs = [struct.pack("<4i", h, w, c, 0) for (h, w, c) in prog.bufs]
    tab = 72 + 16 * len(buf_recs) + 128 * len(op_recs)
    data_off = (tab + 255) // 256 * 256
    total = data_off + len(data)
    header = struct.pack("<4s5I2Q4Q", b"DPB1", 1, len(buf_recs), len(op_recs), prog.patch, PRECISIONS[prog.precision], data_off, total,
                         0, 0, 0, 0)
    assert len(header) == 72
    blob = bytearray(header)
    for r in buf_recs:
        blob.extend(r)
    for r in op_recs:
        assert len(r) == 128
        blob.extend(r)
    blob.extend(b"\0" * (data_off - len(blob)))
    blob.extend(data)
    assert len(blob) == total
    return bytes(blob)
