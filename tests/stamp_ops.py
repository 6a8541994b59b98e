"""Debug tool (not a test): %globaltimer entry/exit of every tensor-core kernel of one forward (direct launches, PDL on)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200.engine import TileModel
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights

prog = densenet121_unet_program(init_densenet_weights(0), 256)
m = TileModel(prog, 0, 32)
m.set_option("use_graph", 0)
for a in sys.argv[1:]:
    k, v = a.split("=")
    m.set_option(k, int(v))
tiles = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    m.forward_tile_batch(tiles)
torch.cuda.synchronize()
m.set_option("stamp", 1)
m.forward_tile_batch(tiles)
torch.cuda.synchronize()
st = m.read_stamps()
t0 = st[st[:, 0] > 0, 0].min()
prev_end = None
tot_busy = 0
for i, o in enumerate(prog.ops):
    a, b = st[i]
    if a == 0:
        continue
    gap = (a - prev_end) / 1e3 if prev_end else 0.0
    print(f"{i:3d} {o.name:16s} start {(a - t0) / 1e3:9.1f} us  dur {(b - a) / 1e3:7.1f} us  start-prev_end {gap:7.1f} us")
    prev_end = b
print(f"span {(st[:, 1].max() - t0) / 1e3:.1f} us")
