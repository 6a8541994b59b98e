"""GPU parity: every mode of the tcgen05 implicit-GEMM conv kernel against a CPU evaluation of the same
packed problem (fp32 math on the fp16 operands), through the C ABI (dp_debug_run_ops)."""
import numpy as np
import pytest

import conv_cases
from gpu_bringup import emu_single

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(conv_cases.CASES))
def test_conv_case(name):
    import torch
    from digipathai_b200.engine import TileModel
    pr, x, B = conv_cases.build_case(name)
    op = pr.ops[0]
    tta_out = 5 if op.head else 0
    ref = emu_single(pr, x, tta_out)
    m = TileModel(pr, device=0, max_batch=B)
    got = {}
    for label, naive in (("naive", 1), ("tc", 0)):
        m.set_option("naive_conv", naive)
        m.write_buffer(0, x)
        m.write_buffer(1, np.zeros((B,) + pr.bufs[1], dtype=np.float16))
        probs = torch.zeros((B, pr.patch, pr.patch), dtype=torch.float32, device="cuda") if op.head else None
        m.run_ops(B, 0, 1, tta_out, probs)
        torch.cuda.synchronize()
        if op.head:
            got[label] = probs.cpu().numpy()
        else:
            full = m.read_buffer(1, B).astype(np.float32)
            got[label] = full[..., op.out_choff:op.out_choff + op.cout].copy()
            full[..., op.out_choff:op.out_choff + op.cout] = 0
            assert np.abs(full).max() == 0, "wrote outside its channel range of the concat buffer"
    # tolerance: one fp16 rounding of the output (half an ulp at |ref|max) + fp32 accumulation-order noise
    tol = 1e-5 if op.head else float(np.abs(ref).max()) * 2.0 ** -10
    assert np.abs(got["tc"] - ref).max() <= tol, (name, np.abs(got["tc"] - ref).max(), tol)
    # the naive path rounds the conv output to fp16 before its separate head kernel
    assert np.abs(got["naive"] - ref).max() <= (2e-3 if op.head else tol)
    m.close()


@pytest.mark.parametrize("name", list(conv_cases.TAP_CASES))
def test_tap_conv_case(name):
    """Generic tap tables (1x7, 7x1, 5x5, ...), stride-2 'same' convs through element-strided TMA boxes, the
    in-place residual epilogue and clipped N tiles (Inception-ResNet-v2 shapes) against the CPU evaluation."""
    import torch
    from digipathai_b200.engine import TileModel
    pr, x, out0, B = conv_cases.build_tap_case(name)
    op = pr.ops[0]
    ref = emu_single(pr, x, 0, residual=out0.astype(np.float32))
    m = TileModel(pr, device=0, max_batch=B)
    for label, naive in (("naive", 1), ("tc", 0)):
        m.set_option("naive_conv", naive)
        m.write_buffer(0, x)
        m.write_buffer(1, out0)
        m.run_ops(B, 0, 1)
        torch.cuda.synchronize()
        full = m.read_buffer(1, B).astype(np.float32)
        got = full[..., op.out_choff:op.out_choff + op.cout]
        tol = float(np.abs(ref).max()) * 2.0 ** -10
        err = np.abs(got - ref).max()
        assert err <= tol, (name, label, err, tol)
        keep = out0.astype(np.float32).copy()
        full[..., op.out_choff:op.out_choff + op.cout] = 0
        keep[..., op.out_choff:op.out_choff + op.cout] = 0
        assert np.array_equal(full, keep), (name, label, "wrote outside its channel range")
    m.close()


@pytest.mark.parametrize("name", list(conv_cases.DENSE_CASES))
def test_fused_dense_layer(name):
    """dense_layer_kernel (1x1 -> bottleneck in smem -> 3x3) against fp32 math on the same fp16 operands."""
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    pr, x, B = conv_cases.build_dense_case(name)
    op = pr.ops[0]
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    xs = f(x.astype(np.float32))[..., op.in_choff:op.in_choff + op.cin]
    a = torch.relu(xs * f(op.pro_scale[:op.cin]) + f(op.pro_shift[:op.cin])).half().float()
    t = torch.relu(a @ f(op.w[0].astype(np.float32)).T + f(op.epi_shift)).half().float()
    w2 = f(op.w2.astype(np.float32))
    ref = torch.zeros(*t.shape[:3], 32)
    for e, (dy, dx, g) in enumerate(emulator.entries(3)):
        ref += emulator._shift(t, dy, dx) @ w2[e].T
    ref = ref.numpy()
    m = TileModel(pr, device=0, max_batch=B)
    for label, naive in (("naive", 1), ("tc", 0)):
        m.set_option("naive_conv", naive)
        m.write_buffer(0, x)
        m.run_ops(B, 0, 1)
        torch.cuda.synchronize()
        full = m.read_buffer(0, B).astype(np.float32)
        got = full[..., op.out_choff:op.out_choff + 32]
        tol = float(np.abs(ref).max()) * 2.0 ** -9   # two fp16 rounding points (bottleneck, output)
        assert np.abs(got - ref).max() <= tol, (name, label, np.abs(got - ref).max(), tol)
        # everything outside the 32 new channels is untouched
        keep = x.astype(np.float32).copy()
        full[..., op.out_choff:op.out_choff + 32] = 0
        keep[..., op.out_choff:op.out_choff + 32] = 0
        assert np.array_equal(full, keep), (name, label, "wrote outside its channel range")
    m.close()
