"""GPU parity of the drop-in entry points (get_prediction / getSegmentation) against the oracle pipeline."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(calibrated_weights):
    from digipathai_b200.slide import synthetic_slide
    from oracle import densenet_ref
    w, _ = calibrated_weights
    slide = synthetic_slide(1024, 768, seed=5)
    return w, slide, {"dense": densenet_ref.OracleModel(w)}


def test_get_prediction_matches_oracle_pipeline(env):
    from digipathai_b200.Segmentation import get_prediction, load_trained_models
    from oracle import pipeline_ref
    w, slide, omodels = env
    kw = dict(batch_size=4, patch_size=256, stride_size=128, tta_list=['FLIP_LEFT_RIGHT'])
    _, want = pipeline_ref.get_prediction(slide, models=omodels, **kw)
    model = load_trained_models('dense', w, 256, max_batch=4)
    status = {}
    _, got = get_prediction(slide, models={'dense': model}, status=status, **kw)
    assert got['mean'].shape == want['mean'].shape == (1024, 768)
    d = np.abs(got['mean'] - want['mean'])
    touched = want['count'] > 0
    print(f"\nget_prediction: max|mean-oracle| {d.max():.3e} mean {d[touched].mean():.3e}; "
          f"max|var-oracle| {np.abs(got['var'] - want['var']).max():.3e}")
    # fp16 mode, 1.5 x measured on the B200 (2.6e-2 max / 8.3e-4 mean on the stitched mean, 1.0e-2 on the variance)
    assert d.max() <= 3.9e-2 and d[touched].mean() <= 1.3e-3
    assert np.abs(got['var'] - want['var']).max() <= 1.5e-2
    # pixels no tile touched stay exactly zero on both sides; zero pattern of the planes is identical
    zero_w, zero_g = want['mean'] == 0, got['mean'] == 0
    assert np.array_equal(zero_w, zero_g)
    assert 0 <= status['progress'] < 100
    model.close()


def test_get_segmentation_drop_in(env, tmp_path):
    from digipathai_b200.Segmentation import getSegmentation
    from oracle import pipeline_ref
    w, slide, omodels = env
    want_thr, want_mean, _ = pipeline_ref.getSegmentation(slide, omodels, 256, 128, 4)
    status = {}
    mask_path = str(tmp_path / "mask.tiff")
    probs_path = str(tmp_path / "probs.tiff")
    got = getSegmentation(slide, patch_size=256, stride_size=128, batch_size=4, quick=True, tta_list=None,
                          crf=False, save_path=mask_path, status=status, probs_path=probs_path, weights=w,
                          pyramidal=False)
    assert got.dtype == np.float32 and got.shape == want_thr.shape and set(np.unique(got)) <= {0.0, 255.0}
    mism = got != want_thr
    # labels may only differ where the oracle probability is within the fp16 band of the 0.3 threshold
    assert (np.abs(want_mean - 0.3)[mism] <= 3.3e-2).all()
    print(f"\ngetSegmentation: {int(mism.sum())} / {mism.size} label mismatches, all inside the +-3.3e-2 band "
          f"({int((np.abs(want_mean - 0.3) <= 3.3e-2).sum())} pixels in band)")
    assert mism.mean() < 0.02
    assert status['status'] == "Saving Prediction Uncertanity..." and status['progress'] == 0
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    saved = np.asarray(Image.open(mask_path))
    assert saved.shape == (768, 1024) and np.array_equal(saved, got.T.astype(np.uint8))   # saved transposed (:345)


def test_errors_match_the_reference(env):
    from digipathai_b200.Segmentation import getSegmentation
    w, slide, _ = env
    with pytest.raises(ValueError, match="Unknown mode"):
        getSegmentation(slide, mode='kidney', weights=w)
    with pytest.raises(ValueError, match="Unknown model"):
        getSegmentation(slide, model='resnet', weights=w)
    with pytest.raises(ValueError, match="holds a 'dense' weight dict"):
        getSegmentation(slide, quick=False, weights=w)      # the 3-model ensemble needs all three weight sets
    with pytest.raises(FileNotFoundError):
        getSegmentation(slide, batch_size=4)          # no converted weights under ~/.DigiPathAI


def test_ensemble_dense_plus_inception_matches_oracle_pipeline(env):
    """Two-model ensemble (Segmentation.py:150-170: TTA outer, model inner, mean/var over all passes) with the
    Inception-ResNet-v2 U-Net as the second member, against the oracle loop driving both oracle graphs."""
    from digipathai_b200.Segmentation import get_prediction, load_trained_models
    from digipathai_b200.models.inception import init_inception_weights
    from oracle import inception_ref, pipeline_ref
    w, slide, omodels = env
    rng = np.random.default_rng(11)
    calib = (rng.integers(0, 256, (2, 256, 256, 3)).astype(np.float32) - 128.0) / 128.0
    wi = inception_ref.calibrate_bn(init_inception_weights(4), calib)
    both = {"dense": omodels["dense"], "inception": inception_ref.OracleModel(wi)}
    kw = dict(batch_size=4, patch_size=256, stride_size=256, tta_list=['ROTATE_90'])
    _, want = pipeline_ref.get_prediction(slide, models=both, **kw)
    models = {"dense": load_trained_models('dense', w, 256, max_batch=4),
              "inception": load_trained_models('inception', wi, 256, max_batch=4)}
    _, got = get_prediction(slide, models=models, **kw)
    d = np.abs(got['mean'] - want['mean'])
    touched = want['count'] > 0
    print(f"\nensemble get_prediction: max|mean-oracle| {d.max():.3e} mean {d[touched].mean():.3e}; "
          f"max|var-oracle| {np.abs(got['var'] - want['var']).max():.3e}")
    # fp16 mode, 1.5 x measured (1.3e-2 max / 2.6e-4 mean; variance 5.8e-3)
    assert d.max() <= 2e-2 and d[touched].mean() <= 4e-4
    assert np.abs(got['var'] - want['var']).max() <= 9e-3
    assert np.array_equal(want['mean'] == 0, got['mean'] == 0)
    for m in models.values():
        m.close()


def test_quick_false_runs_the_three_model_ensemble(env):
    """quick=False == DenseNet + Inception-ResNet-v2 + DeepLabv3+ (Segmentation.py:288-291), model inner loop order
    of the reference (dict insertion order), against the oracle loop driving the three oracle graphs."""
    from digipathai_b200.Segmentation import getSegmentation
    from digipathai_b200.models.deeplab import init_deeplab_weights
    from digipathai_b200.models.inception import init_inception_weights
    from oracle import deeplab_ref, inception_ref, pipeline_ref
    w, slide, omodels = env
    rng = np.random.default_rng(12)
    calib = (rng.integers(0, 256, (2, 256, 256, 3)).astype(np.float32) - 128.0) / 128.0
    wi = inception_ref.calibrate_bn(init_inception_weights(5), calib)
    wd = deeplab_ref.calibrate_bn(init_deeplab_weights(6), calib)
    three = {"dense": omodels["dense"], "inception": inception_ref.OracleModel(wi), "deeplabv3": deeplab_ref.OracleModel(wd)}
    want_thr, want_mean, want_var = pipeline_ref.getSegmentation(slide, three, 256, 256, 4)
    got = getSegmentation(slide, patch_size=256, stride_size=256, batch_size=4, quick=False,
                          weights={"dense": w, "inception": wi, "deeplabv3": wd})
    mism = got != want_thr
    print(f"\n3-model ensemble: {int(mism.sum())} / {mism.size} label mismatches")
    assert (np.abs(want_mean - 0.3)[mism] <= 1e-1).all() and mism.mean() < 0.05


def test_get_segmentation_fp32_mode_returns_the_reference_label_map(env):
    """precision='fp32' / 'tf32x3': the returned {0, 255} map equals the oracle's except where the oracle's own probability lies
    within 1e-3 of the threshold (BASELINE.json: 1e-3 on the mask probabilities, identical label map) -- single model
    and the quick=False three-model ensemble."""
    from digipathai_b200.Segmentation import getSegmentation
    from digipathai_b200.models.deeplab import init_deeplab_weights
    from digipathai_b200.models.inception import init_inception_weights
    from oracle import deeplab_ref, inception_ref, pipeline_ref
    w, slide, omodels = env
    want_thr, want_mean, _ = pipeline_ref.getSegmentation(slide, omodels, 256, 128, 4)
    for prec in ("fp32", "tf32x3"):               # CUDA-core FMA mode and the 3xTF32 tensor-core mode
        got = getSegmentation(slide, patch_size=256, stride_size=128, batch_size=4, quick=True, weights=w, precision=prec)
        mism = got != want_thr
        print(f"\ngetSegmentation {prec} mode: {int(mism.sum())} / {mism.size} label mismatches")
        assert (np.abs(want_mean - 0.3)[mism] <= 1e-3).all() and mism.sum() <= (np.abs(want_mean - 0.3) <= 1e-3).sum()
    rng = np.random.default_rng(12)
    calib = (rng.integers(0, 256, (2, 256, 256, 3)).astype(np.float32) - 128.0) / 128.0
    wi = inception_ref.calibrate_bn(init_inception_weights(5), calib)
    wd = deeplab_ref.calibrate_bn(init_deeplab_weights(6), calib)
    three = {"dense": omodels["dense"], "inception": inception_ref.OracleModel(wi), "deeplabv3": deeplab_ref.OracleModel(wd)}
    want_thr, want_mean, _ = pipeline_ref.getSegmentation(slide, three, 256, 256, 4)
    got = getSegmentation(slide, patch_size=256, stride_size=256, batch_size=4, quick=False,
                          weights={"dense": w, "inception": wi, "deeplabv3": wd}, precision="fp32")
    mism = got != want_thr
    print(f"3-model ensemble fp32 mode: {int(mism.sum())} / {mism.size} label mismatches")
    assert (np.abs(want_mean - 0.3)[mism] <= 1e-3).all()


def test_config5_ensemble_plus_crf_matches_the_oracle_pipeline(env):
    """BASELINE configs[4] in miniature: quick=False (the reference's three-model ensemble, Segmentation.py:288-291) +
    crf=True, precision='fp32'.  Oracle = pipeline_ref.getSegmentation on the three oracle graphs, then the lattice
    restatement of post_process_crf (oracle/lattice_ref.py, utils.py:568-603) on the same non-overlapping 256 x 256
    blocks.  The label map must be identical wherever the oracle's CRF marginal is further than 1e-3 from 0.5."""
    from digipathai_b200.Segmentation import getSegmentation
    from digipathai_b200.models.deeplab import init_deeplab_weights
    from digipathai_b200.models.inception import init_inception_weights
    from digipathai_b200.slide import level0_xy_raster
    from oracle import deeplab_ref, inception_ref, lattice_ref, pipeline_ref
    w, slide, omodels = env                       # 1024 x 768, one level: 4 x 3 CRF blocks
    rng = np.random.default_rng(12)
    calib = (rng.integers(0, 256, (2, 256, 256, 3)).astype(np.float32) - 128.0) / 128.0
    wi = inception_ref.calibrate_bn(init_inception_weights(5), calib)
    wd = deeplab_ref.calibrate_bn(init_deeplab_weights(6), calib)
    three = {"dense": omodels["dense"], "inception": inception_ref.OracleModel(wi), "deeplabv3": deeplab_ref.OracleModel(wd)}
    want_thr, want_mean, _ = pipeline_ref.getSegmentation(slide, three, 256, 256, 4)
    got = getSegmentation(slide, patch_size=256, stride_size=256, batch_size=4, quick=False, crf=True,
                          weights={"dense": w, "inception": wi, "deeplabv3": wd}, precision="fp32")
    raster = level0_xy_raster(slide)
    want = want_thr.copy()
    sure = np.ones(want.shape, bool)
    n_blocks = 0
    for x in range(0, want.shape[0], 256):
        for y in range(0, want.shape[1], 256):
            blk = want_mean[x:x + 256, y:y + 256]
            if blk.max() <= 1e-5:
                continue
            lab, q = lattice_ref.dense_crf(raster[x:x + 256, y:y + 256], blk)
            want[x:x + 256, y:y + 256] = lab * 255
            sure[x:x + 256, y:y + 256] = np.abs(q - 0.5) > 1e-3
            n_blocks += 1
    assert n_blocks > 0 and sure.mean() > 0.95
    mism = (got != want) & sure
    print(f"\nconfig-5 miniature: {n_blocks} CRF blocks, {int((got != want).sum())} label differences, "
          f"{int(mism.sum())} outside the |q - 0.5| < 1e-3 band")
    assert mism.sum() == 0
    assert (want != want_thr).sum() > 0          # the CRF changed labels


def test_get_segmentation_writes_pyramidal_tiffs(env, tmp_path):
    """Default export = what the reference leaves on disk after its ImageMagick pass (Segmentation.py:333-352): tiled
    pyramidal JPEG TIFFs; level 0 of the mask file is the returned label map up to JPEG loss."""
    from PIL import Image
    from digipathai_b200.Segmentation import getSegmentation
    w, slide, _ = env
    mask_path, unc_path = str(tmp_path / "mask.tiff"), str(tmp_path / "unc.tiff")
    got = getSegmentation(slide, patch_size=256, stride_size=256, batch_size=4, weights=w, mask_path=mask_path,
                          uncertainty_path=unc_path)
    Image.MAX_IMAGE_PIXELS = None
    im = Image.open(mask_path)
    assert im.n_frames == 3 and im.size == (1024, 768)            # 1024x768 -> 512x384 -> 256x192
    lvl0 = np.asarray(im).astype(np.float32)
    assert np.abs(lvl0 - got.T).mean() < 6.0 and ((lvl0 > 127) == (got.T > 127)).mean() > 0.995
    im.seek(2)
    assert im.size == (256, 192)
    assert Image.open(unc_path).n_frames == 3


def test_get_segmentation_crf_flag(env):
    """crf=True refines the thresholded map tile-wise with the fully connected CRF (the reference's commented-out
    intent, Segmentation.py:327-331): every touched tile equals engine.dense_crf on that tile's inputs."""
    import torch
    from digipathai_b200 import engine
    from digipathai_b200.Segmentation import getSegmentation, get_prediction, load_trained_models
    from digipathai_b200.slide import open_slide, upload_xy_raster
    w, slide, _ = env
    kw = dict(patch_size=256, stride_size=256, batch_size=4, weights=w)
    plain = getSegmentation(slide, crf=False, **kw)
    refined = getSegmentation(slide, crf=True, **kw)
    assert refined.shape == plain.shape and set(np.unique(refined)) <= {0.0, 255.0}
    model = load_trained_models('dense', w, 256, max_batch=4)
    _, pm = get_prediction(slide, models={'dense': model}, batch_size=4, patch_size=256, stride_size=256,
                           return_device=True)
    model.close()
    raster = upload_xy_raster(open_slide(slide), 0, 1024, torch.device("cuda", 0))
    x, y = 256, 256
    want = engine.dense_crf(raster[x:x + 256, y:y + 256][None].contiguous(),
                            pm['mean'][x:x + 256, y:y + 256][None].contiguous())[0].cpu().numpy() * 255.0
    assert np.array_equal(refined[x:x + 256, y:y + 256], want)
    assert (refined != plain).any()
