"""CPU checks of the Inception-ResNet-v2 U-Net graph builder: the layer program (merged sibling 1x1 convs, in-place
residual chain, generic tap tables, stride-2 'same' convs, folded BN / residual scales), evaluated by the CPU
emulator with the tap tables and packed weights the CUDA runtime uses, must reproduce the fp32 oracle graph
(oracle/inception_ref.py, a restatement of DigiPathAI/models/inception.py:36-271)."""
import types

import numpy as np
import pytest

import emulator
from digipathai_b200 import program as PG
from digipathai_b200.models import inception as IN
from oracle import inception_ref


@pytest.fixture(scope="module")
def small():
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 256, (3, 64, 64, 3)).astype(np.uint8)
    w = IN.init_inception_weights(3)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    inception_ref.calibrate_bn(w, x)
    return w, tiles, x


def _fp32_program(w, patch):
    return IN.inception_resnet_v2_unet_program(w, patch, precision="fp32")


def test_layer_names_follow_keras_creation_order():
    shapes = IN.layer_shapes()
    assert shapes == inception_ref.layer_plan()          # builder and oracle walk the graph identically
    convs = [s for s in shapes if s[1] == "conv"]
    assert len(convs) == 255                             # SURVEY.md a8': 255 convs
    names = [s[0] for s in shapes]
    assert names[0] == "conv2d_1" and "block35_10_conv" in names and "block17_20_conv" in names
    assert "block8_10_conv" in names and "conv_7b" in names and "conv_7b_bn" in names
    assert names[-2:] == ["conv2d_214", "conv2d_214/bias"]   # 214 unnamed convs + 40 block convs + conv_7b


def test_mac_count():
    assert IN.reference_macs_per_tile(256) == 28703326208   # 57.407 GFLOP per tile


def test_program_is_the_reference_graph(small):
    w, tiles, x = small
    taps = {}
    want = inception_ref.forward(w, x, taps=taps)[..., 1]
    prog32 = _fp32_program(w, 64)
    got, bufs = emulator.run(prog32, tiles, fp16_storage=False, keep=True)
    # skip tensors live in place inside the decoder's concat buffers
    for name, buf, off, c in (("conv3", "D3", 256, 320), ("conv4", "D4", 320, 1088), ("block8_10", "X8", 0, 2080),
                              ("conv5", "C5", 0, 1536)):
        # D3/D4 front ranges are overwritten by the decoder later; the running tensors are not
        a = bufs[prog32.buf(buf)][..., off:off + c].numpy()
        assert np.abs(a - taps[name]).max() <= 2e-4 * max(1.0, np.abs(taps[name]).max()), name
    assert np.abs(got - want).max() < 5e-4               # fp32 both sides; 40 residual blocks deep
    got16 = emulator.run(IN.inception_resnet_v2_unet_program(w, 64), tiles, fp16_storage=True)
    assert np.abs(got16 - want).mean() < 5e-2            # fp16 storage: the precision floor the CUDA path shares


def test_tf_same_padding_rule():
    assert PG.same_pad_before(32, 3, 2) == 0 and PG.same_pad_before(32, 3, 1) == 1
    assert PG.same_pad_before(16, 7, 1) == 3 and PG.same_pad_before(32, 5, 1) == 2
    assert PG.tap_offsets(3, 3, 2, 32, 32)[0] == (0, 0) and PG.tap_offsets(1, 7, 1, 16, 16)[0] == (0, -3)
    # stride-2 'same' conv of the emulator == explicit asymmetric padding + strided valid conv
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    k = rng.standard_normal((3, 3, 8, 16)).astype(np.float32)
    x = rng.standard_normal((2, 8, 8, 8)).astype(np.float32)
    op = PG.Op(PG.OP_CONV, cin=8, cout=16, kind=PG.KIND_TAPS, kh=3, kw=3, stride=2,
               w=PG.pack_conv_weights(k, PG.KIND_TAPS).astype(np.float32))
    got = emulator.conv_eval(op, torch.from_numpy(x)).numpy()
    xt = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (0, 1, 0, 1))
    want = F.conv2d(xt, torch.from_numpy(k).permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1).numpy()
    assert np.abs(got - want).max() < 2e-2               # fp16 weight rounding only
