"""Accuracy of one conv in the fp32 FMA mode and the 3xTF32 tensor-core mode against a float64 evaluation, as a
function of the reduction length K = taps x Cin (1x1 and 3x3 convs, positive inputs so that sums do not cancel).
    python tools/tf32x3_accuracy.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from digipathai_b200.engine import TileModel
from digipathai_b200.program import KIND_1X1, KIND_3X3, OP_CONV, Op, Program, pack_conv_weights, weight_precision


def case(kind, cin, cout, H, prec, seed=0, positive=True):
    rng = np.random.default_rng(seed)
    k = 1 if kind == KIND_1X1 else 3
    kern = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    x = rng.standard_normal((2, H, H, cin)).astype(np.float32)
    if positive:
        kern, x = np.abs(kern), np.abs(x)
    with weight_precision(prec):
        pr = Program(patch=64)
        pr.precision = prec
        ib = pr.add_buf("in", H, H, cin)
        ob = pr.add_buf("out", H, H, cout)
        pr.ops.append(Op(OP_CONV, in_buf=ib, in_choff=0, cin=cin, out_buf=ob, out_choff=0, cout=cout, kind=kind, relu=0,
                         w=pack_conv_weights(kern, kind), name="c", epi_shift=np.zeros(cout, np.float32)))
    m = TileModel(pr, device=0, max_batch=2)
    m.write_buffer(0, x)
    m.write_buffer(1, np.zeros((2, H, H, cout), np.float32))
    m.run_ops(2, 0, 1)
    torch.cuda.synchronize()
    got = m.read_buffer(1, 2).astype(np.float64)
    m.close()
    xp = np.pad(x.astype(np.float64), ((0, 0), (k // 2, k // 2), (k // 2, k // 2), (0, 0)))
    ref = np.zeros((2, H, H, cout))
    for dy in range(k):
        for dx in range(k):
            ref += xp[:, dy:dy + H, dx:dx + H] @ kern[dy, dx].astype(np.float64)
    e = got - ref
    s = np.abs(ref).max()
    return np.abs(e).max() / s, e.mean() / s, np.sqrt((e ** 2).mean()) / s


for positive in (True, False):
    print("inputs", "positive" if positive else "signed")
    for kind, cins in ((KIND_1X1, (32, 128, 512, 2048)), (KIND_3X3, (32, 256, 1024))):
        for cin in cins:
            row = []
            for prec in ("fp32", "tf32x3"):
                mx, bias, rms = case(kind, cin, 128, 16, prec, positive=positive)
                row.append(f"{prec}: max {mx:.2e} bias {bias:+.2e} rms {rms:.2e}")
            print(f"  {'1x1' if kind == KIND_1X1 else '3x3'} Cin {cin:5d}  K {cin * (1 if kind == KIND_1X1 else 9):6d}   " + "   ".join(row), flush=True)
