"""GPU parity of the Inception-ResNet-v2 U-Net forward (through dp_forward_tiles) against the fp32 oracle and the
CPU emulator of the same layer program.

Tolerance: as for the DenseNet graph (tests/test_gpu_forward.py), fp16 storage of weights and activations through
a ~240-conv-deep random-init network cannot meet BASELINE.json's 1e-3: rounding the WEIGHTS alone to fp16 (fp32
activations, CPU emulator) already moves the probability by 4.7e-2 max / 4.8e-3 mean on this network.  The 1e-3
bound is asserted on the fp32 precision mode (tests/test_gpu_precision.py: 7e-5 measured).  Asserted here for the
fp16 mode: 1.5 x the measured 8.4e-2 / 6.5e-3 (worst of the plain and the TTA pass) = max-abs <= 1.26e-1 and
mean-abs <= 1e-2 against the oracle, label
mismatches only inside the +-max-abs band,
and per-buffer agreement with the emulator on the early (shallow) tensors where rounding has not accumulated --
that is what pins the stem, the TF-'same' pools and the tap tables end to end.  Kernel correctness proper is
asserted per conv shape in tests/test_gpu_conv.py::test_tap_conv_case at fp16-ulp tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAX_ABS, MEAN_ABS = 1.26e-1, 1.0e-2


@pytest.fixture(scope="module")
def setup():
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.inception import inception_resnet_v2_unet_program, init_inception_weights
    from oracle import inception_ref
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (2, 256, 256, 3)).astype(np.uint8)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    w = init_inception_weights(0)
    inception_ref.calibrate_bn(w, x)
    prog = inception_resnet_v2_unet_program(w, 256)
    emu, ebufs = emulator.run(prog, tiles, keep=True)
    model = TileModel(prog, device=0, max_batch=4)
    return dict(w=w, tiles=tiles, x=x, prog=prog, model=model, oracle=inception_ref.forward(w, x)[..., 1],
                emu=emu, ebufs=ebufs, torch=torch, ref=inception_ref)


def _check(got, want):
    d = np.abs(got - want)
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (d.max(), d.mean())
    mism = ((got >= 0.3) != (want >= 0.3))
    assert (np.abs(want - 0.3)[mism] <= d.max()).all()
    return d.max(), d.mean(), int(mism.sum())


def test_forward_matches_oracle_and_emulator(setup):
    s = setup
    t = s["torch"].from_numpy(s["tiles"]).cuda()
    got = s["model"].forward_tile_batch(t).cpu().numpy()
    e1, a1, m1 = _check(got, s["oracle"])
    e2, a2, _ = _check(got, s["emu"])
    print(f"\ninception forward 2 tiles: vs oracle max {e1:.3e} mean {a1:.3e} (label mismatches {m1}); "
          f"vs emulator max {e2:.3e} mean {a2:.3e}")
    # shallow tensors agree with the emulator to fp16 rounding: stem s2d gather, 3x3/2 stem conv, TF-same max
    # pools, 1x1 / 3x3 / 5x5 convs and the 3x3 average pool of mixed_5b
    prog = s["prog"]
    for name, tol in (("stem_s2d", 0.0), ("S1", 2e-3), ("S2", 4e-3), ("P1", 8e-3), ("S3", 8e-3), ("P2", 2e-2),
                      ("T5p", 2e-2), ("T5", 3e-2)):
        bi = prog.buf(name)
        a = s["model"].read_buffer(bi, 2).astype(np.float32)
        b = s["ebufs"][bi].numpy()
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (name, np.abs(a - b).max())
    s["model"].set_option("naive_conv", 1)
    naive = s["model"].forward_tile_batch(t).cpu().numpy()
    s["model"].set_option("naive_conv", 0)
    _check(naive, s["oracle"])
    _check(got, naive)


def test_tta_pass_and_determinism(setup):
    from digipathai_b200 import tta
    s = setup
    cin, cout = tta.pass_codes(['FLIP_LEFT_RIGHT', 'ROTATE_90'])[-1]
    xin = np.stack([tta.apply(cin, t) for t in s["x"]])
    want = s["ref"].forward(s["w"], xin)[..., 1]
    want = np.stack([np.rot90(t, 3) for t in want])
    t = s["torch"].from_numpy(s["tiles"]).cuda()
    got = s["model"].forward_tile_batch(t, cin, cout).cpu().numpy()
    _check(got, want)
    again = s["model"].forward_tile_batch(t, cin, cout).cpu().numpy()
    assert np.array_equal(got, again)
    # batch composition does not change a tile's result
    four = s["torch"].from_numpy(np.concatenate([s["tiles"], s["tiles"][::-1]])).cuda()
    g4 = s["model"].forward_tile_batch(four, cin, cout).cpu().numpy()
    assert np.array_equal(g4[:2], got) and np.array_equal(g4[2], got[1]) and np.array_equal(g4[3], got[0])


def test_executed_macs_accounting(setup):
    from digipathai_b200.models.inception import reference_macs_per_tile
    ref = reference_macs_per_tile(256)
    ex = setup["model"].executed_macs(1)
    # sub-pixel up-convs remove 5/9 of five decoder layers; padded N tiles / K tails / the zero-padded stem add a
    # little; executed work stays below the reference graph's
    assert 0.7 * ref < ex < 1.02 * ref, (ex, ref)
