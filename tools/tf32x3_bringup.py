"""Bring-up of the 3xTF32 mode: output and stage buffers against the fp32 FMA mode (same program, same weights),
against the oracle for the DenseNet, and batch-32 timing.   python tools/tf32x3_bringup.py [dense|inception|deeplabv3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from digipathai_b200.engine import TileModel

name = sys.argv[1] if len(sys.argv) > 1 else "dense"
rng = np.random.default_rng(1)
tiles = rng.integers(0, 256, (4, 256, 256, 3)).astype(np.uint8)
x = (tiles.astype(np.float32) - 128.0) / 128.0
if name == "dense":
    from digipathai_b200.models.densenet import densenet121_unet_program as build, init_densenet_weights as init
    from oracle import densenet_ref as ref
elif name == "inception":
    from digipathai_b200.models.inception import inception_resnet_v2_unet_program as build, init_inception_weights as init
    from oracle import inception_ref as ref
else:
    from digipathai_b200.models.deeplab import deeplabv3plus_xception_program as build, init_deeplab_weights as init
    from oracle import deeplab_ref as ref
w = init(0)
ref.calibrate_bn(w, x[:2])
t = torch.from_numpy(tiles).cuda()
outs = {}
for prec in ("fp32", "tf32x3"):
    prog = build(w, 256, precision=prec)
    m = TileModel(prog, device=0, max_batch=32)
    outs[prec] = m.forward_tile_batch(t).cpu().numpy()
    bufs = {}
    if prec == "fp32":
        keep = {i: m.read_buffer(i, 4) for i in range(len(prog.bufs))}
    else:
        for i in range(len(prog.bufs)):
            a, b = keep[i], m.read_buffer(i, 4)
            d = np.abs(a - b).max()
            s = max(np.abs(a).max(), 1e-9)
            flag = "" if d <= 1e-4 * max(1.0, s) else "   <-- differs"
            print(f"  buf {i:3d} {prog.bufs[i].name if hasattr(prog.bufs[i], 'name') else ''} max|fp32 - tf32x3| {d:.3e} (scale {s:.3e}){flag}")
    tb = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((32, 256, 256), dtype=torch.float32, device="cuda")
    for _ in range(2):
        m.forward_tile_batch(tb, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        m.forward_tile_batch(tb, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name} {prec}: {ms:.2f} ms per batch of 32 = {32 / ms * 1e3:.0f} tiles/s", flush=True)
    if prec == "tf32x3":
        m.set_option("profile", 1)
        m.forward_tile_batch(tb, out=out)
        tt = m.op_times_ms()
        m.set_option("profile", 0)
        names = [o.name for o in prog.ops]
        for i in np.argsort(-tt)[:12]:
            print(f"    {names[i]:24s} {tt[i]:.3f} ms")
    m.close()
want = ref.forward(w, x)[..., 1]
for prec in outs:
    d = np.abs(outs[prec] - want)
    mism = (outs[prec] >= 0.3) != (want >= 0.3)
    print(f"{name} {prec} vs oracle: max {d.max():.3e} mean {d.mean():.3e} label mismatches {int(mism.sum())} "
          f"(outside band: {int((np.abs(want - 0.3)[mism] > d.max()).sum())})")
print(f"fp32 vs tf32x3: max {np.abs(outs['fp32'] - outs['tf32x3']).max():.3e}")
