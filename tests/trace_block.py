"""Debug tool (not a test): role timelines of CTA 0 of a persistent dense-block kernel (item = layer index).
Usage: python tests/trace_block.py <first op of the block> <layer> [<layer> ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200.engine import TileModel
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights

first = int(sys.argv[1]) if len(sys.argv) > 1 else 25
layers = [int(a) for a in sys.argv[2:]] or [10, 11]
prog = densenet121_unet_program(init_densenet_weights(0), 256)
m = TileModel(prog, 0, 32)
m.set_option("use_graph", 0)
tiles = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    m.forward_tile_batch(tiles)
torch.cuda.synchronize()
m.set_option("trace_block", first)
m.forward_tile_batch(tiles)
torch.cuda.synchronize()
tr = m.read_trace()
ROLE = {0: "prod", 1: "mma", 2: "epi", 3: "xform"}
EV = {(0, 1): "A_issued", (0, 2): "tail_slot", (1, 1): "A_ready", (1, 3): "ph1_issued", (1, 4): "ph2_start",
      (1, 5): "ph2_issued", (2, 0): "mid_start", (2, 2): "mid_done", (2, 3): "final_start", (2, 1): "final_done",
      (2, 4): "published", (3, 1): "xf_start", (3, 0): "xf_done"}
t0 = min(t for _, _, _, t in tr)
tr = sorted(tr, key=lambda e: (e[3] - t0) & 0xFFFFFFFF)
pub = {it: (t - t0) & 0xFFFFFFFF for r, e, it, t in tr if (r, e) == (2, 4)}
ks = sorted(pub)
print("publish period clk:", [pub[b] - pub[a] for a, b in zip(ks[:-1], ks[1:])])
for r, e, it, t in tr:
    if it in layers and r in ROLE:
        print(f"  {(t - t0) & 0xFFFFFFFF:9d} clk  {ROLE[r]:6s} {EV.get((r, e), e):12s} layer {it}")
