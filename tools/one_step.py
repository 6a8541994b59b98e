"""Three batch-32 forwards of the DenseNet U-Net on direct launches (no CUDA graph), for `ncu` launch lists:
the last one is the profiled step (`--launch-skip 2*n --launch-count n`, n = launches per forward, printed here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digipathai_b200.engine import TileModel, kernel_launch_count
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"        # fp16 | fp32 | tf32x3
m = TileModel(densenet121_unet_program(init_densenet_weights(0), 256, precision=precision), device=0, max_batch=32)
m.set_option("use_graph", 0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
slide = torch.randint(0, 256, (4096, 4096, 3), dtype=torch.uint8, device="cuda", generator=g)
coords = torch.randint(0, 4096 - 256, (32, 2), dtype=torch.int32).cuda()
out = torch.empty((32, 256, 256), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for i in range(3):
    n0 = kernel_launch_count()
    m.forward_tiles(slide, coords, 0, 0, out=out)
    torch.cuda.synchronize()
print("launches per forward:", kernel_launch_count() - n0)
