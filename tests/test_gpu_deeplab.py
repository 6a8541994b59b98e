"""GPU parity of the DeepLabv3+ path: the CUDA-core helper kernels (depthwise conv with stride / dilation, global
average pool, broadcast, bilinear align_corners resize, logit-difference head) op by op against the CPU
emulator at fp16-ulp tolerance, and the whole forward against the fp32 oracle.

End-to-end tolerance: this 65-separable-conv random-init network amplifies fp16 rounding far more than the two
U-Nets (the CPU emulator with fp16 WEIGHTS ONLY and fp32 activations already deviates from the fp32 oracle by
1.4e-1 max / 2.0e-2 mean); asserted are mean-abs <= 6e-2 against the oracle plus agreement of every entry-flow
tensor with the emulator to fp16 rounding, which is what pins the stem, the stride-2 depthwise convs, the strided
1x1 shortcuts and the residual epilogue end to end before the 16 middle-flow units amplify rounding noise.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_single(op_builder, in_shape, out_shape, B=3, seed=0, head=False, tta_out=0):
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.program import Program
    rng = np.random.default_rng(seed)
    pr = Program(patch=64)
    ib = pr.add_buf("in", *in_shape)
    ob = pr.add_buf("out", *out_shape)
    for op in op_builder(ib, ob, rng):
        pr.ops.append(op)
    x = rng.standard_normal((B,) + in_shape).astype(np.float16)
    m = TileModel(pr, device=0, max_batch=B)
    m.write_buffer(ib, x)
    m.write_buffer(ob, np.zeros((B,) + out_shape, np.float16))
    probs = torch.zeros((B, 64, 64), dtype=torch.float32, device="cuda") if head else None
    m.run_ops(B, 0, len(pr.ops), tta_out, probs)
    torch.cuda.synchronize()
    got = probs.cpu().numpy() if head else m.read_buffer(ob, B).astype(np.float32)
    m.close()
    return pr, x, got


@pytest.mark.parametrize("stride,rate,pre,post,hw,c", [(1, 1, 1, 0, 32, 64), (2, 1, 1, 0, 32, 128), (1, 2, 0, 1, 16, 136),
                                                       (1, 6, 0, 1, 16, 64), (1, 18, 0, 1, 16, 64), (1, 12, 0, 1, 8, 72)])
def test_depthwise_conv(stride, rate, pre, post, hw, c):
    import torch
    import emulator
    from digipathai_b200.program import OP_DWCONV, Op

    def build(ib, ob, rng):
        return [Op(OP_DWCONV, in_buf=ib, in_choff=8, cin=c, out_buf=ob, out_choff=16, cout=c, relu=post, pro=pre,
                   stride=stride, rate=rate, w=(rng.standard_normal((9, c)) / 3).astype(np.float16),
                   epi_shift=(0.2 * rng.standard_normal(c)).astype(np.float32), name="dw")]
    pr, x, got = _run_single(build, (hw, hw, c + 8), (hw // stride, hw // stride, c + 32))
    op = pr.ops[0]
    ref = emulator.dwconv_eval(op, torch.from_numpy(x.astype(np.float32))[..., 8:8 + c]).numpy()
    assert np.abs(got[..., 16:16 + c] - ref).max() <= np.abs(ref).max() * 2.0 ** -10
    assert np.abs(got[..., :16]).max() == 0 and np.abs(got[..., 16 + c:]).max() == 0


def test_global_pool_broadcast_and_resize():
    import torch
    import emulator
    from digipathai_b200.program import OP_BCAST, OP_GAP, OP_RESIZE, Op
    pr, x, got = _run_single(lambda ib, ob, rng: [Op(OP_GAP, in_buf=ib, cin=256, out_buf=ob, cout=256, name="gap")],
                             (16, 16, 256), (1, 1, 256))
    ref = x.astype(np.float32).mean(axis=(1, 2), keepdims=True)
    assert np.abs(got - ref).max() <= 2.0 ** -10
    pr, x, got = _run_single(lambda ib, ob, rng: [Op(OP_BCAST, in_buf=ib, cin=256, out_buf=ob, out_choff=64, cout=256, name="b")],
                             (1, 1, 256), (16, 16, 384))
    assert np.array_equal(got[..., 64:320], np.broadcast_to(x.astype(np.float32), (3, 16, 16, 256)))
    assert np.abs(got[..., :64]).max() == 0 and np.abs(got[..., 320:]).max() == 0
    pr, x, got = _run_single(lambda ib, ob, rng: [Op(OP_RESIZE, in_buf=ib, cin=256, out_buf=ob, cout=256, name="r")],
                             (16, 16, 256), (64, 64, 304))
    ref = emulator.resize_bilinear_ac(torch.from_numpy(x.astype(np.float32)), 64, 64).numpy()
    assert np.abs(got[..., :256] - ref).max() <= np.abs(ref).max() * 2.0 ** -10
    assert np.abs(got[..., 256:]).max() == 0


def test_logit_difference_head():
    import torch
    import emulator
    from digipathai_b200 import tta
    from digipathai_b200.program import OP_HEAD_DOT, OP_HEAD_RESIZE, Op
    hw = {}

    def build(ib, ob, rng):
        hw["w"] = (0.2 * rng.standard_normal(256)).astype(np.float32)
        return [Op(OP_HEAD_DOT, in_buf=ib, cin=256, out_buf=ob, cout=8, head_w=hw["w"], head_b=0.3, name="dot"),
                Op(OP_HEAD_RESIZE, in_buf=ob, cin=8, out_buf=ob, cout=8, head=1, name="head")]
    pr, x, got = _run_single(build, (16, 16, 256), (16, 16, 8), head=True, tta_out=5)
    z = torch.from_numpy(x.astype(np.float32)) @ torch.from_numpy(hw["w"]) + 0.3
    p = torch.sigmoid(emulator.resize_bilinear_ac(z.unsqueeze(-1), 64, 64)[..., 0]).numpy()
    ref = np.stack([tta.apply(tta.inverse(5), t) for t in p])
    assert np.abs(got - ref).max() <= 2e-5


@pytest.fixture(scope="module")
def full():
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.deeplab import deeplabv3plus_xception_program, init_deeplab_weights
    from oracle import deeplab_ref
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (2, 256, 256, 3)).astype(np.uint8)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    w = deeplab_ref.calibrate_bn(init_deeplab_weights(0), x)
    prog = deeplabv3plus_xception_program(w, 256)
    emu, ebufs = emulator.run(prog, tiles, keep=True)
    model = TileModel(prog, device=0, max_batch=2)
    return dict(tiles=tiles, prog=prog, model=model, oracle=deeplab_ref.forward(w, x)[..., 1], emu=emu, ebufs=ebufs,
                torch=torch)


def test_forward_matches_oracle_and_emulator(full):
    s = full
    t = s["torch"].from_numpy(s["tiles"]).cuda()
    got = s["model"].forward_tile_batch(t).cpu().numpy()
    d, de = np.abs(got - s["oracle"]), np.abs(got - s["emu"])
    print(f"\ndeeplab forward 2 tiles: vs oracle max {d.max():.3e} mean {d.mean():.3e}; vs emulator max {de.max():.3e} "
          f"mean {de.mean():.3e}")
    # fp16 mode, measured on the B200: 2.2e-1 max / 2.9e-2 mean against the fp32 oracle (the sixteen middle-flow units
    # amplify rounding 10x; the fp32 precision mode reaches 2.7e-4, tests/test_gpu_precision.py).  Bounds = 1.5 x measured.
    assert d.max() <= 3.3e-1 and d.mean() <= 4.4e-2 and de.max() <= 3.3e-1 and de.mean() <= 4.4e-2
    mism = (got >= 0.3) != (s["oracle"] >= 0.3)
    assert (np.abs(s["oracle"] - 0.3)[mism] <= d.max()).all()
    prog = s["prog"]
    for name, tol in (("stem_s2d", 0.0), ("A1", 1e-3), ("A2", 1e-3), ("entry_flow_block1_d1", 2e-3), ("B1", 2e-3),
                      ("skip1", 4e-3), ("B2", 4e-3), ("entry_flow_block3_d3", 8e-3)):
        bi = prog.buf(name)
        a = s["model"].read_buffer(bi, 2).astype(np.float32)
        b = s["ebufs"][bi].numpy()
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (name, np.abs(a - b).max())
    again = s["model"].forward_tile_batch(t).cpu().numpy()
    assert np.array_equal(got, again)
    s["model"].set_option("naive_conv", 1)
    naive = s["model"].forward_tile_batch(t).cpu().numpy()
    s["model"].set_option("naive_conv", 0)
    assert np.abs(naive - s["oracle"]).mean() <= 4.4e-2 and np.abs(naive - s["oracle"]).max() <= 3.3e-1


def test_executed_macs_accounting(full):
    from digipathai_b200.models.deeplab import reference_macs_per_tile
    ref = reference_macs_per_tile(256)
    ex = full["model"].executed_macs(1)
    # tensor-core MACs only (depthwise convs run on CUDA cores); 728 -> 736 padding and the zero-padded stem add a little
    assert 0.9 * ref < ex < 1.1 * ref, (ex, ref)
