"""CPU checks of the graph builder / weight packer: the layer program, evaluated by the CPU emulator with the
same tap tables and packed weights the CUDA runtime uses, must reproduce the fp32 oracle graph."""
import numpy as np
import pytest

import emulator
from digipathai_b200 import tta
from digipathai_b200.models import densenet as DN
from digipathai_b200 import program as PG
from oracle import densenet_ref


@pytest.fixture(scope="module")
def small():
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 256, (2, 64, 64, 3)).astype(np.uint8)
    w = DN.init_densenet_weights(3)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    densenet_ref.calibrate_bn(w, x)
    return w, tiles, x


def _fp32_program(w, patch):
    """Program with un-rounded fp32 weights: isolates graph/packing logic from fp16 precision."""
    return DN.densenet121_unet_program(w, patch), DN.densenet121_unet_program(w, patch, precision="fp32")


def test_program_is_the_reference_graph(small):
    w, tiles, x = small
    prog, prog32 = _fp32_program(w, 64)
    want = densenet_ref.forward(w, x)[..., 1]
    got = emulator.run(prog32, tiles, fp16_storage=False)
    assert np.abs(got - want).max() < 5e-5      # fp32 both sides: sub-pixel rewrite, pool/conv commutation, BN folds
    got16 = emulator.run(prog, tiles, fp16_storage=True)
    assert np.abs(got16 - want).max() < 5e-2    # fp16 storage: the precision floor the CUDA path shares


def test_program_tta_codes(small):
    w, tiles, x = small
    _, prog32 = _fp32_program(w, 64)
    for seq in (['FLIP_LEFT_RIGHT', 'ROTATE_90'], ['ROTATE_270']):
        cin, cout = tta.pass_codes(seq)[-1]
        xin = np.stack([tta.apply(cin, t) for t in x])
        want = densenet_ref.forward(w, xin)[..., 1]
        want = np.stack([tta.apply(tta.inverse(cout), t) for t in want])
        got = emulator.run(prog32, tiles, tta_in=cin, tta_out=cout, fp16_storage=False)
        assert np.abs(got - want).max() < 5e-5


def test_mac_counts():
    assert DN.reference_macs_per_tile(256) == 21158166528
    shapes = DN.layer_shapes()
    convs = [s for s in shapes if s[1] == "conv"]
    assert len(convs) == 131 + 1 - 1                                # 130 named convs + head == 131
    params = sum(int(np.prod(s[2])) for s in convs)
    assert abs(params - 17.28e6) / 17.28e6 < 0.01                   # "17.28 M conv params" 


def test_up2_weight_rewrite_is_exact_on_small_case():
    rng = np.random.default_rng(0)
    k = rng.standard_normal((3, 3, 8, 16)).astype(np.float32)
    x = rng.standard_normal((1, 5, 6, 8)).astype(np.float32)
    import torch
    import torch.nn.functional as F
    up = F.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    want = F.conv2d(up, torch.from_numpy(k).permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1).numpy()
    w = PG.pack_conv_weights(k, PG.KIND_UP2).astype(np.float32)   # fp16-rounded
    k16 = np.zeros_like(k)
    got = np.zeros((1, 10, 12, 16), np.float32)
    xt = torch.from_numpy(x)
    for e, (dy, dx, g) in enumerate(emulator.entries(PG.KIND_UP2)):
        got[:, (g >> 1)::2, (g & 1)::2] += (emulator._shift(xt, dy, dx) @ torch.from_numpy(w[e].T)).numpy()
    assert np.abs(got - want).max() < 2e-2   # only fp16 weight rounding separates them


def test_stacked_tap_formulation_of_the_dense_layer_3x3():
    """The arithmetic the STACK variant of dense_layer_kernel performs (csrc/dense_layer.cuh header), restated in
    numpy with the kernel's own indexing -- M rows = (halo row, x), B = the three dy taps of one dx stacked along N,
    output = accumulator rows r, r + 8, r + 16 -- equals the plain 3x3 'same' convolution, for 14-row regions on a
    map they do not divide (32 rows) and 8-row regions."""
    import emulator
    rng = np.random.default_rng(0)
    w2 = rng.standard_normal((9, 32, 128)).astype(np.float32)                 # [tap = dy * 3 + dx][cout][k]
    taps = list(emulator.entries(3))
    assert [(dy, dx) for dy, dx, _ in taps] == [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    for H, W, RH in ((32, 16, 14), (16, 8, 8), (8, 8, 8)):
        t = rng.standard_normal((H, W, 128)).astype(np.float32)
        ref = np.zeros((H, W, 32), np.float32)
        tp = np.pad(t, ((1, 1), (1, 1), (0, 0)))
        for e, (dy, dx, _) in enumerate(taps):
            ref += tp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] @ w2[e].T
        out = np.full((H, W, 32), np.nan, np.float32)
        for h0 in range(0, H, RH):
            for w0 in range(0, W, 8):
                halo = np.zeros((RH + 2, 10, 128), np.float32)               # T tile, zero outside the image
                for hr in range(RH + 2):
                    for c in range(10):
                        ih, iw = h0 - 1 + hr, w0 - 1 + c
                        if 0 <= ih < H and 0 <= iw < W:
                            halo[hr, c] = t[ih, iw]
                acc = np.zeros((128, 96), np.float32)                         # D'
                for g in range(3):                                            # dx index
                    A = np.zeros((128, 128), np.float32)
                    for m in range((RH + 2) * 8):
                        A[m] = halo[m >> 3, (m & 7) + g]
                    B = np.concatenate([w2[j * 3 + g] for j in range(3)], axis=0)   # [96][k]
                    acc += A @ B.T
                for r in range(RH * 8):
                    h, w = h0 + (r >> 3), w0 + (r & 7)
                    if h < H and w < W:
                        out[h, w] = acc[r, 0:32] + acc[r + 8, 32:64] + acc[r + 16, 64:96]
        assert not np.isnan(out).any()
        assert np.abs(out - ref).max() < 1e-3 * np.abs(ref).max()
