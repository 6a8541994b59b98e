"""GPU parity of dp_stitch / dp_finalize: BIT-EXACT against the golden vectors produced by the reference's own
get_prediction loop (tests/golden/pipeline_golden.npz), plus edge cases."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_stitch_and_finalize_reproduce_the_reference_loop_bit_exactly():
    import torch
    from digipathai_b200 import engine, tta
    from digipathai_b200.slide import synthetic_slide
    from digipathai_b200.tissue import TileGrid
    from oracle import pipeline_ref
    from standin import StandInModel
    z = np.load(os.path.join(G, "pipeline_golden.npz"))
    cfg = json.loads(str(z["config"]))
    slide = synthetic_slide(cfg["width"], cfg["height"], cfg["seed"], cfg["n_levels"])
    models = [StandInModel(*p) for p in cfg["models"]]
    grid = TileGrid(slide, cfg["patch"], cfg["stride"], cfg["batch"])
    ds = pipeline_ref.WSIStridedPatchDataset(slide, cfg["patch"], True, cfg["stride"], True)
    W, H = slide.level_dimensions[0]
    B, P = cfg["batch"], cfg["patch"]
    dev = torch.device("cuda", 0)
    mean = torch.zeros((W, H), dtype=torch.float32, device=dev)
    var = torch.zeros_like(mean)
    cnt = torch.zeros((W, H), dtype=torch.uint8, device=dev)
    passes = tta.pass_codes(cfg["tta"])
    for b in range(grid.n_batches):
        x = np.stack([ds[b * B + i][0] for i in range(B)])
        coords = grid.coords[b * B:(b + 1) * B]
        preds = []
        for (cin, cout) in passes:
            xin = np.stack([tta.apply(cin, t) for t in x])
            for m in models:
                p1 = m.predict(xin)[..., 1]
                preds.append(np.stack([tta.apply(tta.inverse(cout), t) for t in p1]))
        probs = torch.from_numpy(np.stack(preds).astype(np.float32)).to(dev)
        engine.stitch(probs, torch.from_numpy(coords).to(dev), mean, var, cnt)
    label = torch.empty((W, H), dtype=torch.uint8, device=dev)
    engine.finalize(mean, var, cnt, 0.3, label)
    torch.cuda.synchronize()
    assert sha(mean.cpu().numpy()) == str(z["mean_sha"])
    assert sha(var.cpu().numpy()) == str(z["var_sha"])
    assert sha(label.cpu().numpy().astype(np.float32)) == str(z["thr_sha"])


def test_stitch_edge_cases_against_oracle_arithmetic():
    import torch
    from digipathai_b200 import engine
    rng = np.random.default_rng(0)
    dev = torch.device("cuda", 0)
    P, W, H = 32, 100, 90
    # heavy overlap inside one batch (incl. identical origins, Q5) and a wrapping uint8 count (Q6)
    coords = np.array([[0, 0], [0, 0], [5, 7], [68, 58], [68, 58], [30, 30], [31, 29], [0, 58]], np.int32)
    N, B = 5, len(coords)
    probs = rng.random((N, B, P, P), dtype=np.float32)
    mean = np.zeros((W, H), np.float32); var = np.zeros((W, H), np.float32)
    cnt = rng.integers(250, 256, (W, H)).astype(np.uint8)
    cnt0 = cnt.copy()
    m = np.mean(probs, axis=0); v = np.var(probs, axis=0)
    for i, (x, y) in enumerate(coords):
        mean[x:x + P, y:y + P] += m[i]; var[x:x + P, y:y + P] += v[i]
        cnt[x:x + P, y:y + P] += np.ones((P, P), np.uint8)
    tm = torch.zeros((W, H), dtype=torch.float32, device=dev); tv = torch.zeros_like(tm)
    tc = torch.from_numpy(cnt0).to(dev)
    engine.stitch(torch.from_numpy(probs).to(dev), torch.from_numpy(coords).to(dev), tm, tv, tc)
    torch.cuda.synchronize()
    assert np.array_equal(tm.cpu().numpy(), mean) and np.array_equal(tv.cpu().numpy(), var)
    assert np.array_equal(tc.cpu().numpy(), cnt)
    # normalise: count==0 -> 1, mean/count, var/count**2.0 (float64 power in numpy), threshold >= 0.3
    np.place(cnt, cnt == 0, 1)
    mean /= cnt
    var /= cnt ** 2.0
    lab = torch.empty((W, H), dtype=torch.uint8, device=dev)
    engine.finalize(tm, tv, tc, 0.3, lab)
    torch.cuda.synchronize()
    assert np.array_equal(tm.cpu().numpy(), mean) and np.array_equal(tv.cpu().numpy(), var)
    want = np.where(mean >= np.float32(0.3), 255, 0).astype(np.uint8)
    assert np.array_equal(lab.cpu().numpy(), want)
    # single pass: variance is exactly zero
    tm.zero_(); tv.zero_(); tc.zero_()
    engine.stitch(torch.from_numpy(probs[:1].copy()).to(dev), torch.from_numpy(coords).to(dev), tm, tv, tc)
    assert float(tv.abs().max()) == 0.0
    # pyramid level
    lvl = engine.pyramid_down2(tm).cpu().numpy()
    a = tm.cpu().numpy()
    assert np.allclose(lvl, 0.25 * (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]), atol=1e-6)
