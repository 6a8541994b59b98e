"""Evidence for engine.ForwardLanes: %globaltimer entry / exit of every tensor-core kernel of three batch-32 forwards
issued together on three lanes (CUDA-graph replays captured with the stamps on), against one forward alone.  Prints how
much of the burst's wall time has kernels of 1, 2 and 3 different batches running at the same time and which ops overlap.
    python tools/lanes_timeline.py [lanes=3] > profiles/r2_lanes_timeline.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200 import engine
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights

L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prog = densenet121_unet_program(init_densenet_weights(0), 256)
slide = torch.randint(0, 256, (8192, 8192, 3), dtype=torch.uint8, device="cuda")
coords = torch.randint(0, 8192 - 256, (64, 32, 2), dtype=torch.int32).cuda()


def run(lanes, rounds):
    m = engine.TileModel(prog, 0, 32)                 # fresh model: its graphs are captured with the stamps on
    pool = engine.ForwardLanes({"m": m}, lanes)
    models = pool.models["m"]
    outs = [torch.empty((32, 256, 256), dtype=torch.float32, device="cuda") for _ in range(lanes)]
    for mm in models:
        mm.set_option("stamp", 1)
    pool.begin()
    for k in range(rounds * lanes):                   # warm-up (graph capture with the stamps on, clocks, L2)
        pool.forward("m", slide, coords[k % 64], 0, 0, out=outs[k % lanes])
    pool.join()
    torch.cuda.synchronize()
    pool.begin()
    for k in range(lanes):                            # the burst that is analysed: ONE forward per lane, all stamped
        pool.forward("m", slide, coords[(7 + k) % 64], 0, 0, out=outs[k])
    pool.join()
    torch.cuda.synchronize()
    st = [mm.read_stamps() for mm in models]
    m.close()
    return st


def report(st, label):
    ev = []
    for l, s in enumerate(st):
        for i, (a, b) in enumerate(s):
            if a > 0 and b > a:
                ev.append((int(a), int(b), l, i))
    t0 = min(e[0] for e in ev); t1 = max(e[1] for e in ev)
    pts = sorted({e[0] for e in ev} | {e[1] for e in ev})
    hist = {}
    pair_time = {}
    for x0, x1 in zip(pts[:-1], pts[1:]):
        act = [(l, i) for (a, b, l, i) in ev if a <= x0 and b >= x1]
        n = len({l for l, _ in act})
        hist[n] = hist.get(n, 0) + (x1 - x0)
        if n >= 2:
            names = tuple(sorted(prog.ops[i].name.split("_block")[0] for _, i in act))
            pair_time[names] = pair_time.get(names, 0) + (x1 - x0)
    span = (t1 - t0) / 1e3
    busy = [sum(b - a for (a, b, l, _) in ev if l == k) / 1e3 for k in range(len(st))]
    print(f"{label}: span {span:.1f} us for {len(st)} forwards = {span / len(st):.1f} us per forward; "
          f"sum of tensor-core kernel time per lane {['%.0f' % b for b in busy]} us")
    for n in sorted(hist):
        print(f"   {n} batch(es) with a tensor-core kernel running: {hist[n] / 1e3:8.1f} us ({100 * hist[n] / (t1 - t0):.1f} %)")
    for names, t in sorted(pair_time.items(), key=lambda kv: -kv[1])[:12]:
        print(f"      overlap {' + '.join(names):40s} {t / 1e3:7.1f} us")


report(run(1, 6), "1 lane: one forward alone")
report(run(L, 6), f"{L} lanes: a burst of {L} forwards issued together, one per lane (includes the ramp-up and the drain of the burst)")
