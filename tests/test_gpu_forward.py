"""GPU parity of the whole DenseNet-121 U-Net forward (through dp_forward_tiles) against the fp32 oracle.

Tolerance.  BASELINE.json asks for 1e-3 max-abs on the probability; that bound is asserted on the library's fp32
precision mode in tests/test_gpu_precision.py (measured 2e-5 .. 5e-5).  THIS file covers the fp16 tensor-core mode
(the configuration BASELINE.json's configs[1] names), which cannot reach 1e-3 on this 121-layer random-init
stand-in network whatever the kernels do: profiles/r2_parity_conditioning.md shows the fp32 oracle itself moving by
2.3e-2 max / 2e-3 mean when ONE activation tensor (the stem output) is perturbed by a single 10-bit-mantissa
rounding, and profiles/r2_parity_drift.md shows the same 1000x depth amplification acting on the fp32 mode's own
rounding (1e-7 at conv1 -> 1e-4 at conv5).  Measured on the B200 (rounds 1-2): 2.0e-2 .. 2.2e-2 max, 2.2e-3 mean.
Asserted: 1.5 x measured = max-abs <= 3.3e-2 and mean-abs <= 3.3e-3 against the oracle, the same against the CPU
emulator of the fp16 program, and label (p >= 0.3) mismatches only inside the +-max-abs band.
Kernel correctness proper is asserted per layer in test_gpu_conv.py at fp16-ulp tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAX_ABS, MEAN_ABS = 3.3e-2, 3.3e-3     # 1.5 x the measured 2.2e-2 / 2.2e-3 (see above)


@pytest.fixture(scope="module")
def setup(calibrated_weights):
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.densenet import densenet121_unet_program
    from oracle import densenet_ref
    w, tiles = calibrated_weights
    rng = np.random.default_rng(7)
    tiles = np.concatenate([tiles, rng.integers(0, 256, (1, 256, 256, 3)).astype(np.uint8)])
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    prog = densenet121_unet_program(w, 256)
    model = TileModel(prog, device=0, max_batch=32)
    return dict(w=w, tiles=tiles, x=x, prog=prog, model=model, oracle=densenet_ref.forward(w, x)[..., 1],
                emu=emulator.run(prog, tiles), torch=torch, ref=densenet_ref)


def _check(got, want):
    d = np.abs(got - want)
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (d.max(), d.mean())
    mism = ((got >= 0.3) != (want >= 0.3))
    assert (np.abs(want - 0.3)[mism] <= d.max()).all()
    return d.max(), int(mism.sum())


def test_forward_matches_oracle_and_emulator(setup):
    s = setup
    t = s["torch"].from_numpy(s["tiles"]).cuda()
    got = s["model"].forward_tile_batch(t).cpu().numpy()
    e1, m1 = _check(got, s["oracle"])
    e2, _ = _check(got, s["emu"])
    print(f"\nforward 3 tiles: max|p-oracle| {e1:.3e} (label mismatches {m1}), max|p-emulator| {e2:.3e}")
    s["model"].set_option("naive_conv", 1)
    naive = s["model"].forward_tile_batch(t).cpu().numpy()
    s["model"].set_option("naive_conv", 0)
    _check(naive, s["oracle"])
    _check(got, naive)


def test_forward_tta_pass_is_cumulative_and_inverted(setup):
    from digipathai_b200 import tta
    s = setup
    cin, cout = tta.pass_codes(['FLIP_LEFT_RIGHT', 'ROTATE_90'])[-1]
    xin = np.stack([tta.apply(cin, t) for t in s["x"]])
    want = s["ref"].forward(s["w"], xin)[..., 1]
    want = np.stack([np.rot90(t, 3) for t in want])          # transform_prob('ROTATE_90')
    t = s["torch"].from_numpy(s["tiles"]).cuda()
    got = s["model"].forward_tile_batch(t, cin, cout).cpu().numpy()
    _check(got, want)
    # a wrong orientation is far outside the tolerance (the check has teeth)
    assert np.abs(got - np.stack([np.rot90(t, 1) for t in want])).max() > 0.2


def test_forward_from_slide_coords_equals_pregathered_tiles(setup):
    s = setup
    torch = s["torch"]
    rng = np.random.default_rng(3)
    slide_xy = rng.integers(0, 256, (700, 600, 3)).astype(np.uint8)
    coords = np.array([[0, 0], [444, 344], [17, 301], [128, 128], [300, 5]], np.int32)
    tiles = np.stack([slide_xy[x:x + 256, y:y + 256] for x, y in coords])
    a = s["model"].forward_tiles(torch.from_numpy(slide_xy).cuda(), torch.from_numpy(coords).cuda()).cpu().numpy()
    b = s["model"].forward_tile_batch(torch.from_numpy(tiles).cuda()).cpu().numpy()
    assert np.array_equal(a, b)


def test_batch_32_is_consistent_and_keras_style_predict(setup):
    s = setup
    tiles = np.concatenate([s["tiles"][:1]] * 31 + [s["tiles"][1:2]])
    p = s["model"].predict((tiles.astype(np.float32) - 128.0) / 128.0)
    assert p.shape == (32, 256, 256, 2) and np.allclose(p.sum(-1), 1.0, atol=1e-6)
    for i in range(1, 31):
        assert np.array_equal(p[i], p[0])
    _check(p[0, ..., 1], s["oracle"][0])
    _check(p[31, ..., 1], s["oracle"][1])
    with pytest.raises(ValueError):
        s["model"].predict(np.full((1, 256, 256, 3), 0.1234, np.float32))


def test_executed_macs_accounting(setup):
    from digipathai_b200.models.densenet import reference_macs_per_tile
    ref = reference_macs_per_tile(256)
    ex = setup["model"].executed_macs(1)
    # sub-pixel rewrite removes 5/9 of the five up-conv layers (-21 %); the fused dense layers recompute the 1x1
    # conv on their halo rows and, with 8-row regions, issue half-empty M = 128 MMAs for the 3x3 (+14 % overall);
    # the stem runs as a zero-padded 4x4 conv.  Executed work stays below the reference graph's.
    assert 0.75 * ref < ex < 1.0 * ref, (ex, ref)


def test_host_batch_pipeline_matches_direct_forward(setup):
    """engine.HostBatchPipeline (three-stream, double-buffered host front end) returns exactly what the direct call
    returns, for more batches than it has buffer slots."""
    from digipathai_b200.engine import HostBatchPipeline
    s = setup
    torch = s["torch"]
    rng = np.random.default_rng(4)
    batches = [rng.integers(0, 256, (3, 256, 256, 3)).astype(np.uint8) for _ in range(5)]
    want = [s["model"].forward_tile_batch(torch.from_numpy(b).cuda()).cpu().numpy() for b in batches]
    pipe = HostBatchPipeline(s["model"], batch=3)
    ins = [torch.from_numpy(b).pin_memory() for b in batches]
    outs = [torch.empty((3, 256, 256), dtype=torch.float32).pin_memory() for _ in batches]
    for i, o in zip(ins, outs):
        pipe.submit(i, o)
    pipe.drain()
    for o, w in zip(outs, want):
        assert np.array_equal(o.numpy(), w)


def test_early_tensors_agree_with_the_emulator_to_fp16_rounding(setup):
    """Before rounding noise has accumulated through the dense blocks, the CUDA path and the CPU emulator of the same
    layer program agree to fp16 rounding: pins the stem gather, the 4x4 space-to-depth stem conv, the zero-padded
    max pool, the first fused dense layers and the first transition end to end."""
    import emulator
    s = setup
    t = s["torch"].from_numpy(s["tiles"]).cuda()
    s["model"].forward_tile_batch(t)
    s["torch"].cuda.synchronize()
    _, ebufs = emulator.run(s["prog"], s["tiles"], keep=True)
    prog = s["prog"]
    n = len(s["tiles"])
    # (buffer, channel range, tolerance relative to the tensor's max): D1[96:160) = conv1, D2[128:384) = pool1 + block 2
    for name, lo, hi, tol in (("stem_s2d", 0, 64, 0.0), ("D1", 96, 160, 1e-3), ("D2", 128, 192, 1e-3),
                              ("D2", 192, 256, 4e-3), ("D2", 256, 384, 8e-3), ("Q2", 0, 256, 8e-3)):
        bi = prog.buf(name)
        a = s["model"].read_buffer(bi, n).astype(np.float32)[..., lo:hi]
        b = ebufs[bi].numpy()[..., lo:hi]
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (name, lo, np.abs(a - b).max(), np.abs(b).max())
