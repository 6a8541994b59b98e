"""EXPERIMENTAL, opt-in: the fused dense layer with its 3x3 taps stacked along N = 96 (dense_layer.cuh STACK variant,
planner switch DP_DL_STACK) against the same fp32 reference as tests/test_gpu_conv.py::test_fused_dense_layer.

The variant was written after round 1's GPU budget was spent and has never run on hardware.  A mistake in an mbarrier
protocol shows up as a hang, not as a failure, so these tests only run when DP_TEST_UNVERIFIED=1 -- under a shell
`timeout` -- and the suite does not depend on them:

    timeout 600 env DP_TEST_UNVERIFIED=1 python -m pytest tests/test_gpu_stack_dense.py -m gpu -x -q
"""
import os

import numpy as np
import pytest

import conv_cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DP_TEST_UNVERIFIED") != "1",
                                 reason="unverified kernel variant: opt in with DP_TEST_UNVERIFIED=1 under a timeout")]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", list(conv_cases.DENSE_CASES))
def test_stacked_dense_layer(name, mode, monkeypatch):
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    monkeypatch.setenv("DP_DL_STACK", str(mode))         # read by the planner when the first plan is built
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
    m.set_option("naive_conv", 0)
    m.write_buffer(0, x)
    m.run_ops(B, 0, 1)
    torch.cuda.synchronize()
    full = m.read_buffer(0, B).astype(np.float32)
    got = full[..., op.out_choff:op.out_choff + 32]
    tol = float(np.abs(ref).max()) * 2.0 ** -9
    assert np.abs(got - ref).max() <= tol, (name, mode, np.abs(got - ref).max(), tol)
    keep = x.astype(np.float32).copy()
    full[..., op.out_choff:op.out_choff + 32] = 0
    keep[..., op.out_choff:op.out_choff + 32] = 0
    assert np.array_equal(full, keep), (name, mode, "wrote outside its channel range")
    m.close()
