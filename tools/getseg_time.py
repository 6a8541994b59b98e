"""Wall time of the whole getSegmentation call (BASELINE configs[2] shape at a bounded size): synthetic S x S slide resident in
HBM, DenseNet U-Net, tta_list of 3, batch 32, the three pyramidal JPEG TIFFs written to /tmp.  Phases from the status
protocol + timers around save_pyramidal.    python tools/getseg_time.py [S=16384]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digipathai_b200 import tiffio
from digipathai_b200.Segmentation import getSegmentation
from digipathai_b200.models.densenet import init_densenet_weights
from digipathai_b200.slide import synthetic_slide_device

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
levels = 1
while S // (2 ** (levels - 1)) > 2500 and levels < 5:
    levels += 1
slide = synthetic_slide_device(S, S, torch.device("cuda", 0), seed=0, n_levels=levels)
w = init_densenet_weights(0)
save_t = []
_orig = tiffio.save_pyramidal
def timed(*a, **k):
    t = time.perf_counter(); r = _orig(*a, **k); save_t.append(time.perf_counter() - t); return r
tiffio.save_pyramidal = timed
with tempfile.TemporaryDirectory() as d:
    for rep in range(2):
        save_t.clear()
        t0 = time.perf_counter()
        out = getSegmentation(slide, patch_size=256, stride_size=128, batch_size=32, quick=True,
                              tta_list=['FLIP_LEFT_RIGHT', 'ROTATE_90', 'ROTATE_180'], crf=False,
                              save_path=os.path.join(d, "mask.tiff"), probs_path=os.path.join(d, "probs.tiff"),
                              uncertainty_path=os.path.join(d, "unc.tiff"), weights=w, status={})
        dt = time.perf_counter() - t0
        sizes = {f: os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)}
        print(f"run {rep}: getSegmentation on a {S}x{S} slide: {dt:.2f} s total, of which writing the three pyramidal TIFFs "
              f"{sum(save_t):.2f} s ({['%.2f' % t for t in save_t]}); files {sizes}; host cores {os.cpu_count()}", flush=True)
