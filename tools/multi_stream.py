"""Experiment: K forward steps of batch 32 issued round-robin over S model replicas on S CUDA streams.

    python tools/multi_stream.py [dense|inception|deeplabv3] [steps]

Each replica is its own dp_model (own activation buffers, own captured graph), so consecutive batches are independent
and kernels of different batches may share the GPU wherever one batch's kernel leaves SMs idle.
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digipathai_b200 import engine

name = sys.argv[1] if len(sys.argv) > 1 else "dense"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B, P = 32, 256
dev = torch.device("cuda", 0)


def build():
    if name == "dense":
        from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
        return densenet121_unet_program(init_densenet_weights(0), P)
    if name == "inception":
        from digipathai_b200.models.inception import inception_resnet_v2_unet_program, init_inception_weights
        return inception_resnet_v2_unet_program(init_inception_weights(0), P)
    from digipathai_b200.models.deeplab import deeplabv3plus_xception_program, init_deeplab_weights
    return deeplabv3plus_xception_program(init_deeplab_weights(0), P)


prog = build()
SW = 8192
slide = torch.randint(0, 256, (SW, SW, 3), dtype=torch.uint8, device=dev)
coords = torch.randint(0, SW - P, (steps + 16, B, 2), dtype=torch.int32).to(dev)
res = {}
for S in (1, 2, 3, 4):
    models = [engine.TileModel(prog, device=0, max_batch=B) for _ in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    outs = [torch.empty((B, P, P), dtype=torch.float32, device=dev) for _ in range(S)]

    def run(n, off):
        for k in range(n):
            s = k % S
            with torch.cuda.stream(streams[s]):
                models[s].forward_tiles(slide, coords[off + k], 0, 0, out=outs[s])

    run(2 * S + 4, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    run(steps, 8)
    for st in streams:
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res[S] = {"ms_per_step": round(ms, 4), "tiles_per_s": round(B / ms * 1e3, 1)}
    print(name, "streams", S, res[S], flush=True)
    for m in models:
        m.close()
    del models, outs
    torch.cuda.empty_cache()
print(json.dumps({"model": name, "steps": steps, "by_streams": res}))
