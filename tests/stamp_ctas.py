"""Debug tool (not a test): per-CTA %globaltimer stamps of consecutive dense-layer kernels (direct launches, PDL on):
when does every CTA of layer k+1 enter, when does its grid-dependency wait return, when does it exit -- relative to
layer k.  Usage: python tests/stamp_ctas.py 35 36 37 [key=value options]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200.engine import TileModel
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights

ops = [int(a) for a in sys.argv[1:] if "=" not in a] or [35, 36, 37, 57, 58, 59]
prog = densenet121_unet_program(init_densenet_weights(0), 256)
m = TileModel(prog, 0, 32)
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("="); m.set_option(k, int(v))
m.set_option("use_graph", 0)
tiles = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    m.forward_tile_batch(tiles)
torch.cuda.synchronize()
m.set_option("stamp_ctas", 1)
m.forward_tile_batch(tiles)
torch.cuda.synchronize()
t0 = None
for op in ops:
    st = m.read_cta_stamps(op)
    st = st[st[:, 0] > 0]
    if not len(st):
        print(op, prog.ops[op].name, "no stamps"); continue
    if t0 is None:
        t0 = st[:, 0].min()
    e, w, x = (st[:, 0] - t0) / 1e3, (st[:, 1] - t0) / 1e3, (st[:, 2] - t0) / 1e3
    q = lambda v: f"min {v.min():7.2f} p10 {np.percentile(v, 10):7.2f} med {np.median(v):7.2f} p90 {np.percentile(v, 90):7.2f} max {v.max():7.2f}"
    print(f"op {op} {prog.ops[op].name} ({len(st)} CTAs on {len(set(st[:, 3]))} SMs)  [us since first entry of op {ops[0]}]")
    print(f"   entry        {q(e)}")
    print(f"   wait return  {q(w)}")
    print(f"   exit         {q(x)}")
    print(f"   entry->exit  {q(x - e)}   wait->exit {q(x - w)}")
