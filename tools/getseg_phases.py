import os, sys, time, tempfile
sys.path.insert(0, '.')
import torch
from digipathai_b200.Segmentation import getSegmentation
from digipathai_b200.models.densenet import init_densenet_weights
from digipathai_b200.slide import synthetic_slide_device
S = 40000
slide = synthetic_slide_device(S, S, torch.device("cuda", 0), seed=0, n_levels=5)
w = init_densenet_weights(0)
with tempfile.TemporaryDirectory() as d:
    for rep in range(2):
        ph = {}
        t0 = time.perf_counter()
        out = getSegmentation(slide, patch_size=256, stride_size=128, batch_size=32, quick=True,
                              tta_list=['FLIP_LEFT_RIGHT', 'ROTATE_90', 'ROTATE_180'], crf=False,
                              save_path=os.path.join(d, "mask.tiff"), probs_path=os.path.join(d, "probs.tiff"),
                              uncertainty_path=os.path.join(d, "unc.tiff"), weights=w, status={}, timings=ph)
        print(rep, round(time.perf_counter() - t0, 2), {k: round(v) for k, v in ph.items()}, flush=True)
        del out
