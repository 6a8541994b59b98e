"""GPU parity of the DenseCRF mean-field kernels (dp_crf_tiles) against the exact numpy oracle of the model
post_process_crf configures (DigiPathAI/helpers/utils.py:568-603).  Tolerance: fp32 sums over <= 4096 pair weights
against the fp64 oracle -- marginals within 2e-4, labels identical away from |q - 0.5| < 1e-3."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(h, w, seed):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.float32)
    img[:] = (235, 235, 238)
    yy, xx = np.mgrid[0:h, 0:w]
    blob = ((yy - h * 0.45) / (h * 0.3)) ** 2 + ((xx - w * 0.4) / (w * 0.25)) ** 2 < 1
    img[blob] = (170, 90, 160)
    img = np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)
    p = np.clip(np.where(blob, 0.65, 0.25) + 0.3 * rng.standard_normal((h, w)), 0, 1).astype(np.float32)
    return img, p


@pytest.mark.parametrize("h,w,n_iter", [(24, 32, 10), (40, 40, 10), (17, 50, 3), (64, 64, 5)])
def test_crf_matches_exact_mean_field_oracle(h, w, n_iter):
    import torch
    from digipathai_b200 import engine
    from oracle import crf_ref
    cases = [_case(h, w, s) for s in (0, 1, 2)]
    rgb = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    p1 = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    lab, q1 = engine.dense_crf(rgb, p1, n_iter=n_iter, return_marginal=True, method="exact")
    lab, q1 = lab.cpu().numpy(), q1.cpu().numpy()
    changed = 0
    for k, (img, p) in enumerate(cases):
        want_lab, want_q = crf_ref.dense_crf(img, p, n_iter=n_iter)
        assert np.abs(q1[k] - want_q).max() < 2e-4, (k, np.abs(q1[k] - want_q).max())
        sure = np.abs(want_q - 0.5) > 1e-3
        assert np.array_equal(lab[k][sure], want_lab[sure])
        changed += int((want_lab != (p >= 0.5)).sum())
    assert changed > 0          # the CRF actually moved labels (the test has teeth)


def test_crf_batch_independence_and_zero_iterations():
    import torch
    from digipathai_b200 import engine
    (img_a, p_a), (img_b, p_b) = _case(32, 32, 5), _case(32, 32, 6)
    rgb = torch.from_numpy(np.stack([img_a, img_b])).cuda()
    p1 = torch.from_numpy(np.stack([p_a, p_b])).cuda()
    for method in ("exact", "lattice"):
        both = engine.dense_crf(rgb, p1, method=method).cpu().numpy()
        one = engine.dense_crf(rgb[1:].contiguous(), p1[1:].contiguous(), method=method).cpu().numpy()
        assert np.array_equal(both[1], one[0]), method
        again = engine.dense_crf(rgb, p1, method=method).cpu().numpy()
        assert np.array_equal(both, again), method            # bit-reproducible (the lattice splat adds in fixed point)
        unary_only = engine.dense_crf(rgb, p1, n_iter=0, method=method).cpu().numpy()
        assert np.array_equal(unary_only, (np.stack([p_a, p_b]) > 0.5).astype(np.uint8)), method


@pytest.mark.parametrize("h,w,n_iter,n", [(24, 32, 10, 3), (40, 40, 10, 2), (17, 50, 3, 2), (64, 64, 5, 11)])
def test_lattice_crf_matches_the_permutohedral_oracle(h, w, n_iter, n):
    """dp_crf_tiles_lattice against oracle/lattice_ref.py (numpy restatement of densecrf's permutohedral filter and
    inference loop): marginals within 2e-4 (float32 both sides, different summation order in the splat), labels identical
    away from |q - 0.5| < 1e-3; 11 tiles exercise the 8-tile chunking of the workspace."""
    import torch
    from digipathai_b200 import engine
    from oracle import lattice_ref
    cases = [_case(h, w, s) for s in range(n)]
    rgb = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    p1 = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    lab, q1 = engine.dense_crf(rgb, p1, n_iter=n_iter, return_marginal=True, method="lattice")
    lab, q1 = lab.cpu().numpy(), q1.cpu().numpy()
    changed = 0
    for k in (0, n - 1) if n > 3 else range(n):
        img, p = cases[k]
        want_lab, want_q = lattice_ref.dense_crf(img, p, n_iter=n_iter)
        assert np.abs(q1[k] - want_q).max() < 2e-4, (k, np.abs(q1[k] - want_q).max())
        sure = np.abs(want_q - 0.5) > 1e-3
        assert np.array_equal(lab[k][sure], want_lab[sure])
        changed += int((want_lab != (p >= 0.5)).sum())
    assert changed > 0


def test_lattice_crf_on_a_full_tile_is_close_to_the_exact_filter_and_fast():
    """256 x 256 (the path's tile): the lattice is an approximation of the exact Gaussian filters, so the two methods
    agree on the vast majority of labels; prints the time per tile of both."""
    import time
    import torch
    from digipathai_b200 import engine
    cases = [_case(256, 256, s) for s in range(4)]
    rgb = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    p1 = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    out = {}
    for method in ("lattice", "exact"):
        engine.dense_crf(rgb[:1].contiguous(), p1[:1].contiguous(), method=method)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out[method] = engine.dense_crf(rgb, p1, method=method).cpu().numpy()
        dt = (time.perf_counter() - t0) / len(cases)
        print(f"\n{method}: {dt * 1e3:.2f} ms per 256x256 tile")
    agree = (out["lattice"] == out["exact"]).mean()
    print(f"label agreement lattice vs exact: {agree:.4f}")
    assert agree > 0.97
