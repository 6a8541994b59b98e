"""Runs the HBM-bound slide-plane kernels once at realistic sizes, for an `ncu` capture (tools/profile_aux.sh):
stem gather (inside one forward), stitch of one batch of 32 tiles x 4 passes on a 16384^2 plane, finalize and
pyramid_down2 on that plane, tissue mask + morphology on a 2500^2 level image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200 import engine, tissue
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
dev = torch.device("cuda", 0)
S = 16384
g = torch.Generator(device=dev); g.manual_seed(0)
slide = torch.randint(0, 256, (4096, 4096, 3), dtype=torch.uint8, device=dev, generator=g)
m = engine.TileModel(densenet121_unet_program(init_densenet_weights(0), 256), device=0, max_batch=32)
m.set_option("use_graph", 0)
xs = (torch.arange(32, dtype=torch.int32) % 8) * 128 + 1024
ys = (torch.arange(32, dtype=torch.int32) // 8) * 128 + 2048
coords = torch.stack([xs, ys], 1).contiguous().to(dev)
probs = torch.rand((4, 32, 256, 256), dtype=torch.float32, device=dev, generator=g)
mean = torch.zeros((S, S), dtype=torch.float32, device=dev); var = torch.zeros_like(mean)
count = torch.zeros((S, S), dtype=torch.uint8, device=dev)
label = torch.empty((S, S), dtype=torch.uint8, device=dev)
rgb = torch.randint(0, 256, (2500, 2500, 3), dtype=torch.uint8, device=dev, generator=g)
for rep in range(2):   # first pass warms up, the profiler is told to skip it
    m.forward_tiles(slide, coords, 0, 0, out=probs[0])
    engine.stitch(probs, coords, mean, var, count)
    engine.finalize(mean, var, count, 0.3, label)
    engine.pyramid_down2(mean)
    mk = tissue.tissue_mask_device(rgb)
    tissue.morpho_process_device(mk, 4)
    torch.cuda.synchronize()
print("done")
