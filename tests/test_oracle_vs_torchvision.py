"""The DenseNet-121 encoder of the oracle (oracle/densenet_ref.py: stem, four dense blocks, three transitions, final BN --
DigiPathAI/models/densenet.py:110-134, i.e. keras.applications' DenseNet121 body) against an INDEPENDENT implementation
that exists in this image: ``torchvision.models.densenet121``.  The oracle's weights are loaded into torchvision's module
tree by name; every stage output must agree to fp32 rounding.  This pins the encoder's topology and layer semantics
(pre-activation order BN-ReLU-conv, bottleneck width 4 x growth, concatenation order, 7x7/2 stem with padding 3,
zero-padded 3/2 max pool, BN-ReLU-conv-avgpool transitions, closing BN without ReLU) to code the oracle's author did not
write.  What it cannot pin: TensorFlow's own kernels, and the decoder, which is specific to the reference."""
import numpy as np
import pytest
import torch

tv = pytest.importorskip("torchvision")


def _bn(mod, params):
    g, b, mu, var = params
    mod.weight.data = torch.from_numpy(np.asarray(g, np.float32))
    mod.bias.data = torch.from_numpy(np.asarray(b, np.float32))
    mod.running_mean.data = torch.from_numpy(np.asarray(mu, np.float32))
    mod.running_var.data = torch.from_numpy(np.asarray(var, np.float32))
    mod.eps = 1.001e-5                                   # densenet.py:61,67,101,118,134


def _conv(mod, k_hwio):
    w = torch.from_numpy(np.ascontiguousarray(np.asarray(k_hwio, np.float32).transpose(3, 2, 0, 1)))
    assert tuple(mod.weight.shape) == tuple(w.shape), (mod.weight.shape, w.shape)
    mod.weight.data = w


def test_encoder_matches_torchvision_densenet121():
    from digipathai_b200.models.densenet import init_densenet_weights
    from oracle import densenet_ref
    rng = np.random.default_rng(3)
    x = ((rng.integers(0, 256, (2, 128, 128, 3)).astype(np.float32)) - 128.0) / 128.0
    w = init_densenet_weights(0)
    densenet_ref.calibrate_bn(w, x)
    taps = {}
    densenet_ref.forward(w, x, taps=taps)

    net = tv.models.densenet121(weights=None).eval()
    f = net.features
    _conv(f.conv0, w["conv1/conv"])
    _bn(f.norm0, w["conv1/bn"])
    for b, n in zip((2, 3, 4, 5), densenet_ref.BLOCKS):
        blk = getattr(f, f"denseblock{b - 1}")
        for i in range(1, n + 1):
            layer, p = getattr(blk, f"denselayer{i}"), f"conv{b}_block{i}"
            _bn(layer.norm1, w[p + "_0_bn"]); _conv(layer.conv1, w[p + "_1_conv"])
            _bn(layer.norm2, w[p + "_1_bn"]); _conv(layer.conv2, w[p + "_2_conv"])
        if b < 5:
            tr = getattr(f, f"transition{b - 1}")
            _bn(tr.norm, w[f"pool{b}_bn"]); _conv(tr.conv, w[f"pool{b}_conv"])
    _bn(f.norm5, w["bn"])

    def close(name, got_nchw):
        got = got_nchw.permute(0, 2, 3, 1).numpy()
        want = taps[name]
        assert got.shape == want.shape, (name, got.shape, want.shape)
        err = np.abs(got - want).max()
        assert err <= 2e-5 * max(1.0, np.abs(want).max()), (name, err, np.abs(want).max())

    with torch.no_grad():
        t = torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()
        t = f.relu0(f.norm0(f.conv0(t)));      close("conv1", t)
        t = f.pool0(t);                        close("pool1", t)          # -inf padding == zero padding after a ReLU
        for b in (2, 3, 4, 5):
            t = getattr(f, f"denseblock{b - 1}")(t)
            if b < 5:
                close(f"conv{b}", t)
                t = getattr(f, f"transition{b - 1}")(t); close(f"pool{b}", t)
        t = f.norm5(t);                        close("conv5", t)
    assert taps["conv5"].shape == (2, 4, 4, 1024)
