#!/usr/bin/env python
"""Where the fp16 forward's deviation from the fp32 oracle comes from (task 1a/1b).  TEST TOOLING:
imports the oracle, so it lives under tools/ and is never imported by the product.

  python tools/parity_study.py cond   [--out profiles/r2_parity_conditioning.md]     (CPU only)
  python tools/parity_study.py drift  [--out profiles/r2_parity_drift.md] [--batch 4]  (needs the GPU)

cond:  conditioning of the stand-in network instance itself, measured on the fp32 oracle alone.  A relative
       perturbation of 2^-11 (one fp16 / TF32 rounding: uniform in +-2^-11) is injected into the fp32 oracle at one
       activation, or at every layer output, and the change of the output probability is recorded.  No kernel, no
       fp16 arithmetic, no layer program is involved: this is what ANY implementation that rounds activations to a
       10-bit mantissa once per layer must expect on this instance.
drift: per-stage deviation of the CUDA path (fp16 mode and fp32 mode) from the fp32 oracle's taps.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _setup(n_tiles=2, seed_w=0, seed_t=1):
    from digipathai_b200.models.densenet import init_densenet_weights
    from oracle import densenet_ref
    rng = np.random.default_rng(seed_t)
    tiles = rng.integers(0, 256, (n_tiles, 256, 256, 3)).astype(np.uint8)
    w = init_densenet_weights(seed_w)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    densenet_ref.calibrate_bn(w, x[:2])
    return w, tiles, x


def _stats(p, want):
    d = np.abs(p - want)
    mism = int(((p >= 0.3) != (want >= 0.3)).sum())
    return float(d.max()), float(d.mean()), mism


def cond(args):
    import torch
    from digipathai_b200.models.densenet import DENSENET_BLOCKS, DECODER
    from oracle import densenet_ref
    w, tiles, x = _setup(2)
    want = densenet_ref.forward(w, x)[..., 1]
    g = torch.Generator().manual_seed(5)
    eps = 2.0 ** -11

    def noise(t):
        return t * (1.0 + (torch.rand(t.shape, generator=g) * 2 - 1) * eps)

    def rnd16(t):
        return t.half().float()

    dense_names = [f"conv{b}_block{i}" for b, n in zip((2, 3, 4, 5), DENSENET_BLOCKS) for i in range(1, n + 1)]
    dec_names = [n for n, _, _ in DECODER]
    rows = []
    for label, names, fn in (
        ("2^-11 relative noise at conv1 (stem output) only", ["conv1"], noise),
        ("... at the conv2 block output only", ["conv2"], noise),
        ("... at the conv3 block output only", ["conv3"], noise),
        ("... at the conv4 block output only", ["conv4"], noise),
        ("... at the conv5 block output only", ["conv5"], noise),
        ("... at dec10a only (second-to-last conv)", ["dec10a"], noise),
        ("... at the 32 new channels of every dense layer (58 injections)", dense_names, noise),
        ("... at every decoder conv output (10 injections)", dec_names, noise),
        ("... at all 69 layer outputs", ["conv1"] + dense_names + dec_names, noise),
        ("round-to-fp16 (instead of noise) at all 69 layer outputs, fp32 weights and math", ["conv1"] + dense_names + dec_names, rnd16),
    ):
        got = densenet_ref.forward(w, x, perturb={n: fn for n in names})[..., 1]
        rows.append((label,) + _stats(got, want))
    # weights only: every conv kernel rounded to fp16 once, activations and math fp32
    w16 = {k: (v.astype(np.float16).astype(np.float32) if isinstance(v, np.ndarray) and v.ndim == 4 else v) for k, v in w.items()}
    rows.append(("conv kernels rounded to fp16, activations fp32",) + _stats(densenet_ref.forward(w16, x)[..., 1], want))
    z = np.log(want / (1 - want))
    lines = ["# Conditioning of the stand-in DenseNet-121 U-Net instance (fp32 oracle only)", "",
             "`python tools/parity_study.py cond` -- seed-0 He-normal weights, BN statistics calibrated on the 2 test tiles, "
             "2 tiles 256x256 (131 072 output pixels).  Output = softmax channel 1; logit z1-z0 has std "
             f"{z.std():.2f}; {int((np.abs(want - 0.3) < 1e-3).sum())} pixels lie within 1e-3 of the 0.3 threshold.", "",
             "| perturbation of the fp32 oracle | max abs dev of p | mean abs dev | label flips (p >= 0.3) |", "|---|---|---|---|"]
    for label, mx, mn, mm in rows:
        lines.append(f"| {label} | {mx:.2e} | {mn:.2e} | {mm} |")
    lines += ["",
              "Reading: a single 10-bit-mantissa rounding of ONE early activation already moves the output by more than "
              "the 1e-3 target; one rounding per layer (what fp16, bf16 or TF32 storage does by definition) gives the "
              "1e-2 level the fp16 CUDA path shows.  The deviation is a property of instance x number format, not of "
              "the kernels; only operands with > 10 mantissa bits (the library's `precision=fp32` mode) can meet 1e-3."]
    txt = "\n".join(lines) + "\n"
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


def drift(args):
    import torch
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.densenet import densenet121_unet_program
    from oracle import densenet_ref
    B = args.batch
    w, tiles, x = _setup(B)
    taps = {}
    want = densenet_ref.forward(w, x, taps=taps)[..., 1]
    # stage -> (buffer, channel range) in the layer program's buffer plan (models/densenet.py)
    stages = [("conv1", "D1", 96, 160), ("conv2", "D2", 128, 384), ("conv3", "D3", 256, 768), ("conv4", "D4", 320, 1344),
              ("conv5", "N5", 0, 1024), ("dec6a", "D4", 0, 320), ("dec6b", "E6", 0, 320), ("dec7a", "D3", 0, 256),
              ("dec7b", "E7", 0, 256), ("dec8a", "D2", 0, 128), ("dec8b", "E8", 0, 128), ("dec9a", "D1", 0, 96),
              ("dec9b", "E9", 0, 96), ("dec10a", "E10", 0, 64)]
    out = {}
    for prec in ("fp16", "fp32"):
        prog = densenet121_unet_program(w, 256, precision=prec)
        model = TileModel(prog, device=0, max_batch=B)
        got = model.forward_tile_batch(torch.from_numpy(tiles).cuda()).cpu().numpy()
        torch.cuda.synchronize()
        rows = []
        for name, buf, lo, hi in stages:
            if name == "dec10a" and prec == "fp16":
                rows.append((name, float("nan"), float("nan"), float("nan")))   # conv10b overwrites E10 in place
                continue
            a = model.read_buffer(prog.buf(buf), B).astype(np.float32)[..., lo:hi]
            b = taps[name]
            d = np.abs(a - b)
            rows.append((name, float(d.max()), float(d.mean()), float(np.abs(b).max())))
        rows.append(("p (output)",) + _stats(got, want)[:2] + (1.0,))
        out[prec] = (rows, _stats(got, want))
        model.close()
    lines = [f"# Per-stage drift of the CUDA forward from the fp32 oracle (batch {B}, 256x256 tiles)", "",
             "`python tools/parity_study.py drift` on a B200.  Columns: max / mean |CUDA - oracle| of the stage's tensor "
             "(absolute; the tensor's max |value| alongside).", "",
             "| stage | fp16 max | fp16 mean | fp32-mode max | fp32-mode mean | max abs value |", "|---|---|---|---|---|---|"]
    for r16, r32 in zip(out["fp16"][0], out["fp32"][0]):
        lines.append(f"| {r16[0]} | {r16[1]:.2e} | {r16[2]:.2e} | {r32[1]:.2e} | {r32[2]:.2e} | {r32[3]:.1f} |")
    for prec in ("fp16", "fp32"):
        mx, mn, mm = out[prec][1]
        band = int((np.abs(want - 0.3) < mx).sum())
        lines.append(f"\n{prec} mode: max|p - oracle| {mx:.3e}, mean {mn:.3e}, label mismatches {mm} "
                     f"(pixels inside the +-max band: {band})")
    txt = "\n".join(lines) + "\n"
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["cond", "drift"])
    ap.add_argument("--out", default="")
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    {"cond": cond, "drift": drift}[a.what](a)
