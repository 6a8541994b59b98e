"""Debug tool (not a test): per-role timeline of CTA 0 for selected conv ops of the DenseNet forward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200.engine import TileModel
from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights

ops = [int(a) for a in sys.argv[1:]] or [135, 134, 8, 65, 66, 108]
prog = densenet121_unet_program(init_densenet_weights(0), 256)
m = TileModel(prog, 0, 32)
if os.environ.get("TRACE_FIRST"):
    show_first = int(os.environ["TRACE_FIRST"])
else:
    show_first = 3
m.set_option("use_graph", 0)
tiles = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    m.forward_tile_batch(tiles)
torch.cuda.synchronize()
ROLE = {0: "prod", 1: "mma", 2: "epi", 3: "xform", 9: "setup"}
EV_DL = {(0, 1): "A_issued", (1, 1): "A_ready", (1, 2): "B_ready", (1, 3): "ph1_issued", (1, 4): "ph2_start",
         (1, 5): "ph2_issued", (2, 0): "mid_start", (2, 2): "mid_done", (2, 3): "final_start", (2, 1): "final_done",
         (3, 1): "xf_start", (3, 0): "xf_done", (3, 2): "xf_stored", (3, 3): "xf_fenced", (9, 0): "setup_done", (9, 2): "kernel_entry", (9, 1): "kernel_end"}
EV = {(0, 1): "A_issued", (0, 2): "B_issued", (1, 0): "acc_free", (1, 1): "A_ready", (1, 2): "B_ready",
      (1, 3): "item_issued", (1, 4): "tap_issued", (1, 5): "B_committed", (2, 0): "acc_full", (2, 1): "epi_done", (3, 0): "xform_done", (9, 0): "setup_done", (9, 2): "kernel_entry", (9, 1): "kernel_end"}
for op in ops:
    m.set_option("trace_op", op)
    m.forward_tile_batch(tiles)
    torch.cuda.synchronize()
    tr = m.read_trace()
    m.set_option("trace_op", -1)
    if not tr:
        print("op", op, "no trace"); continue
    t0 = min(t for _, _, _, t in tr)
    tr = sorted(tr, key=lambda e: (e[3] - t0) & 0xFFFFFFFF)
    print(f"=== op {op} {prog.ops[op].name}: {len(tr)} events")
    evmap = EV_DL if prog.ops[op].type == 6 else EV
    items = sorted({it for r, e, it, t in tr if r != 9})
    show = set(items[:show_first] + items[-2:])
    last = {}
    for r, e, it, t in tr:
        dt = (t - t0) & 0xFFFFFFFF
        if r == 9 or it in show:
            print(f"  {dt:9d} clk  {ROLE.get(r, r):6s} {evmap.get((r, e), e):12s} item {it}")
    # per-item period on the mma role
    iss = [((t - t0) & 0xFFFFFFFF) for r, e, it, t in tr if (r, e) == (1, 3)]
    if len(iss) > 2:
        d = np.diff(iss)
        print(f"  items on CTA0: {len(iss)}; item period clk: median {np.median(d):.0f} min {d.min()} max {d.max()}; total {iss[-1]}")
    eps = [((t - t0) & 0xFFFFFFFF) for r, e, it, t in tr if (r, e) == (2, 1)]
    epf = [((t - t0) & 0xFFFFFFFF) for r, e, it, t in tr if (r, e) == (2, 0)]
    if eps and epf:
        print(f"  epilogue duration clk: median {np.median(np.array(eps) - np.array(epf[:len(eps)])):.0f}")
