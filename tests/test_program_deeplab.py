"""CPU checks of the DeepLabv3+ (Xception, OS 16) graph builder: the layer program evaluated by the CPU emulator must
reproduce the fp32 oracle graph (oracle/deeplab_ref.py, a restatement of DigiPathAI/models/deeplabv3.py:52-160,
219-482 as instantiated at helpers/utils.py:438-446)."""
import types

import numpy as np
import pytest

import emulator
from digipathai_b200 import program as PG
from digipathai_b200.models import deeplab as DL
from oracle import deeplab_ref


def _fp32_program(w, patch):
    return DL.deeplabv3plus_xception_program(w, patch, precision="fp32")


def test_layer_names_and_mac_count():
    shapes = DL.layer_shapes()
    assert shapes == deeplab_ref.layer_plan()
    names = [s[0] for s in shapes]
    for must in ("entry_flow_conv1_1", "entry_flow_block2_separable_conv2_pointwise_BN", "middle_flow_unit_16_separable_conv3_depthwise",
                 "exit_flow_block1_shortcut", "image_pooling", "aspp3_depthwise", "concat_projection", "feature_projection0",
                 "decoder_conv1_pointwise", "custom_logits_semantic/bias"):
        assert must in names, must
    macs = DL.reference_macs_per_tile(256)
    assert abs(macs - 12.80e9) / 12.80e9 < 0.01          # SURVEY.md a8'': ~12.80 GMAC per tile


def test_program_is_the_reference_graph():
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 256, (2, 128, 128, 3)).astype(np.uint8)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    w = deeplab_ref.calibrate_bn(DL.init_deeplab_weights(3), x)
    taps = {}
    want = deeplab_ref.forward(w, x, taps=taps)[..., 1]
    prog32 = _fp32_program(w, 128)
    got, bufs = emulator.run(prog32, tiles, fp16_storage=False, keep=True)
    for name, buf, c in (("conv1_2", "A2", 64), ("block1", "B1", 128), ("skip1", "skip1", 256), ("block2", "B2", 256),
                         ("exit1", "E3", 1024), ("exit2", "G", 2048), ("aspp", "PJ", 256), ("decoder", "D3", 256)):
        a = bufs[prog32.buf(buf)][..., :c].numpy()
        assert np.abs(a - taps[name]).max() <= 1e-3 * max(1.0, np.abs(taps[name]).max()), name
    assert np.abs(got - want).max() < 2e-3               # fp32 both sides, 65 separable convs deep
    # 728-channel tensors are carried as 736: the pad channels stay exactly zero
    assert np.abs(bufs[prog32.buf("X")][..., 728:]).max() == 0


def test_resize_matches_tf_align_corners_formula():
    import torch
    x = torch.arange(16, dtype=torch.float32).reshape(1, 4, 4, 1)
    y = emulator.resize_bilinear_ac(x, 13, 13)[0, :, :, 0].numpy()
    # src = dst * (in - 1) / (out - 1): corners map to corners, the centre to the centre
    assert y[0, 0] == 0 and y[12, 12] == 15 and abs(y[6, 6] - 7.5) < 1e-6 and abs(y[0, 4] - 1.0) < 1e-6
