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
