"""Generates the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN CODE (run in the build
container, where /root/reference is mounted; the fixtures travel, the reference does not).

The reference package cannot be imported (tensorflow / openslide / skimage / imgaug ... are not installable
here), but the functions on the hot path are plain numpy / cv2 / torch-DataLoader code.  This script lifts
their source text out of the reference files with ``ast`` -- nothing is copied into the repo -- and executes it
unmodified in a namespace where only the missing third-party *entry points* are stubbed:

    openslide.OpenSlide(path)        -> an in-memory slide whose read_region returns PIL RGBA images
    skimage rgb2hsv / threshold_otsu -> restatements of skimage's published algorithm (flagged: these two are
                                        NOT reference-executed; everything around them is)
    Keras Model.predict              -> tests/standin.py (IEEE-exact elementwise float32 arithmetic)
    home directory                   -> a scratch dir inside the repo (the reference writes memmaps to ~)

Executed reference code: apply_tta, transform_prob, BinMorphoProcessMaskOS, TissueMaskGenerationOS
(DigiPathAI/helpers/utils.py), class WSIStridedPatchDataset (DigiPathAI/loaders/dataloader.py) and
get_prediction (DigiPathAI/Segmentation.py) including torch's DataLoader(drop_last=True).

    python tests/golden/make_golden.py        # rewrites tta_golden.npz, morph_golden.npz, pipeline_golden.npz
"""
import ast
import hashlib
import json
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/DigiPathAI"


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
    missing = set(names) - set(out)
    assert not missing, missing
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    import torch
    from PIL import Image
    from torch.utils.data import DataLoader, Dataset
    from tqdm import tqdm

    from digipathai_b200.slide import synthetic_slide
    from oracle import pipeline_ref
    from standin import StandInModel

    # ------------------------------------------------------------------ lift + exec reference source
    u = lift(f"{REF}/helpers/utils.py",
             ["apply_tta", "transform_prob", "BinMorphoProcessMaskOS", "TissueMaskGenerationOS"])

    def rgb2hsv_stub(img):
        out = np.zeros(img.shape, dtype=np.float64)
        out[..., 1] = pipeline_ref._rgb2hsv_s(img)
        return out

    ns_u = {"np": np, "cv2": cv2, "rgb2hsv": rgb2hsv_stub, "threshold_otsu": pipeline_ref._otsu}
    for s in u.values():
        exec(compile(s, "reference:utils.py", "exec"), ns_u)

    class PILSlide:
        """OpenSlide-shaped view of an ArraySlide: read_region returns RGBA PIL images."""

        def __init__(self, arr_slide):
            self._s = arr_slide
            self.level_dimensions = arr_slide.level_dimensions
            self.level_downsamples = arr_slide.level_downsamples
            self.level_count = arr_slide.level_count

        def read_region(self, loc, level, size):
            rgb = self._s.read_region(loc, level, size)
            rgba = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=-1)
            return Image.fromarray(rgba, mode="RGBA")

    registry = {}
    openslide_stub = types.SimpleNamespace(OpenSlide=lambda p: registry[p])
    d = lift(f"{REF}/loaders/dataloader.py", ["WSIStridedPatchDataset"])
    ns_d = {"np": np, "os": os, "Image": Image, "openslide": openslide_stub, "Dataset": Dataset,
            "TissueMaskGenerationOS": ns_u["TissueMaskGenerationOS"],
            "BinMorphoProcessMaskOS": ns_u["BinMorphoProcessMaskOS"]}
    exec(compile(d["WSIStridedPatchDataset"], "reference:dataloader.py", "exec"), ns_d)

    scratch = os.path.join(HERE, "_scratch_home")
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(scratch)
    g = lift(f"{REF}/Segmentation.py", ["get_prediction"])
    ns_g = {"np": np, "os": os, "home": scratch, "DataLoader": DataLoader, "tqdm": tqdm,
            "WSIStridedPatchDataset": ns_d["WSIStridedPatchDataset"], "apply_tta": ns_u["apply_tta"],
            "transform_prob": ns_u["transform_prob"]}
    exec(compile(g["get_prediction"], "reference:Segmentation.py", "exec"), ns_g)

    # ------------------------------------------------------------------ 1. TTA golden
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 6, 6, 2)).astype(np.float32)
    seqs = [[], ["FLIP_LEFT_RIGHT"], ["FLIP_LEFT_RIGHT", "ROTATE_90"], ["ROTATE_90", "ROTATE_180", "ROTATE_270"],
            ["ROTATE_270", "nonsense", "FLIP_LEFT_RIGHT", "ROTATE_90"]]
    tta_out = {"x": x, "sequences": json.dumps(seqs)}
    for si, seq in enumerate(seqs):
        patches = x.copy()
        for k, nm in enumerate(["DEFAULT"] + seq):
            patches = ns_u["apply_tta"](patches, nm)
            tta_out[f"in_{si}_{k}"] = patches.copy()
            tta_out[f"out_{si}_{k}"] = ns_u["transform_prob"](patches.copy(), nm)
    np.savez_compressed(os.path.join(HERE, "tta_golden.npz"), **tta_out)

    # ------------------------------------------------------------------ 2. morphology golden
    mrng = np.random.default_rng(5)
    m = (mrng.random((220, 170)) > 0.995).astype(np.uint8)
    m[60:140, 40:120] |= (mrng.random((80, 80)) > 0.3).astype(np.uint8)
    morph = {"mask": np.packbits(m), "shape": np.array(m.shape)}
    for lvl in (0, 3, 4):
        morph[f"out_{lvl}"] = np.packbits(ns_u["BinMorphoProcessMaskOS"](m.copy(), lvl))
    try:
        ns_u["BinMorphoProcessMaskOS"](m.copy(), 5)
        morph["level5_raises"] = np.array(0)
    except ValueError:
        morph["level5_raises"] = np.array(1)
    np.savez_compressed(os.path.join(HERE, "morph_golden.npz"), **morph)

    # ------------------------------------------------------------------ 3. pipeline golden
    cfg = dict(width=640, height=512, seed=3, n_levels=2, patch=128, stride=64, batch=4,
               tta=["FLIP_LEFT_RIGHT", "ROTATE_90"], models=[[2.0, 1.0, 0.5], [1.5, -0.7, 0.9]])
    slide = synthetic_slide(cfg["width"], cfg["height"], cfg["seed"], cfg["n_levels"])
    registry["synthetic"] = PILSlide(slide)
    models = {f"m{i}": StandInModel(*p) for i, p in enumerate(cfg["models"])}
    torch.manual_seed(0)
    status = {}
    ref_slide, probs_map = ns_g["get_prediction"]("synthetic", mask_path=None, label_path=None,
                                                  batch_size=cfg["batch"], models=models, tta_list=cfg["tta"],
                                                  num_workers=0, verbose=0, patch_size=cfg["patch"],
                                                  stride_size=cfg["stride"], mask_level=-1, status=status)
    mean = np.array(probs_map["mean"])
    var = np.array(probs_map["var"])
    ds = ns_d["WSIStridedPatchDataset"]("synthetic", None, None, image_size=cfg["patch"], normalize=True,
                                        flip=None, rotate=None, sampling_stride=cfg["stride"], mask_level=-1,
                                        roi_masking=True)
    tile0 = ds[0]
    tile_last = ds[len(ds) - 1]
    # threshold exactly as getSegmentation does (Segmentation.py:336-337)
    thr = mean.copy()
    np.place(thr, thr >= 0.3, 255)
    np.place(thr, thr < 0.3, 0)
    np.savez_compressed(
        os.path.join(HERE, "pipeline_golden.npz"),
        config=json.dumps(cfg), X_idcs=ds._X_idcs, Y_idcs=ds._Y_idcs, mask=np.packbits(ds._mask > 0),
        mask_shape=np.array(ds._mask.shape), resolution=np.array(ds._resolution), level=np.array(ds._level),
        factor=np.array(ds._sampling_stride),
        tile0=tile0[0], tile0_xy=np.array(tile0[1:3]), tile_last=tile_last[0], tile_last_xy=np.array(tile_last[1:3]),
        mean_sha=sha(mean), var_sha=sha(var), thr_sha=sha(thr), mean_shape=np.array(mean.shape),
        mean_sub=mean[::8, ::8], var_sub=var[::8, ::8], mean_sum=np.float64(mean.astype(np.float64).sum()),
        progress_last=np.array(status.get("progress", -1)))
    shutil.rmtree(scratch, ignore_errors=True)
    print("tiles", len(ds), "mean range", mean.min(), mean.max(), "var max", var.max(), "progress", status)


if __name__ == "__main__":
    main()
