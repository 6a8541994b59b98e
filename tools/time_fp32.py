"""Times the fp32 precision mode's forward at batch 32 (CUDA events, L2 not flushed: the working set is > L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200.engine import TileModel
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
w = init_densenet_weights(0)
m = TileModel(densenet121_unet_program(w, 256, precision="fp32"), device=0, max_batch=32)
t = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    m.forward_tile_batch(t)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    m.forward_tile_batch(t)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"fp32 mode: {ms:.2f} ms per batch of 32 = {32e3 / ms:.0f} tiles/s")
m.set_option("profile", 1)
m.forward_tile_batch(t); torch.cuda.synchronize()
tt = m.op_times_ms()
names = [o.name for o in m.program.ops]
for i in np.argsort(-tt)[:12]:
    print(f"  {names[i]:24s} {tt[i]:.3f} ms")
