"""GPU bring-up / bisect script (not a pytest file).  Each case runs in its own process so that a trapped
kernel (sticky CUDA error) cannot poison the next one.

    python tests/gpu_bringup.py            # all cases, each under `timeout`
    python tests/gpu_bringup.py --case X   # one case in this process
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def emu_single(pr, x, tta_out=0, residual=None):
    """Reference for a one-op program on CPU (fp32 math on the fp16 operands)."""
    import numpy as np
    import torch
    import emulator
    from digipathai_b200 import tta
    from digipathai_b200.program import KIND_UP2
    op = pr.ops[0]
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    xs = f(x.astype(np.float32))[..., op.in_choff:op.in_choff + op.cin]
    if op.pro:
        xs = torch.relu(xs * f(op.pro_scale[:op.cin]) + f(op.pro_shift[:op.cin])).half().float()
    y = emulator.conv_eval(op, xs)
    if op.residual:
        y = y + f(np.asarray(residual, dtype=np.float32))[..., op.out_choff:op.out_choff + op.cout]
    if op.relu:
        y = torch.relu(y)
    if op.head:
        p = torch.sigmoid(y @ f(op.head_w) + op.head_b).numpy()
        return np.stack([tta.apply(tta.inverse(tta_out), t) for t in p])
    return y.numpy()


def run_conv_case(name):
    import numpy as np
    import torch
    import conv_cases
    from digipathai_b200.engine import TileModel
    pr, x, B = conv_cases.build_case(name)
    op = pr.ops[0]
    tta_out = 5 if op.head else 0
    ref = emu_single(pr, x, tta_out)
    m = TileModel(pr, device=0, max_batch=B)
    res = {}
    combos = [("naive", 1, 0, 0), ("tc_b0", 0, 0, 0), ("tc_b1", 0, 1, 0)]
    if name.startswith("h_"):
        combos += [("tc_b0_pad8", 0, 0, 1), ("tc_b1_pad8", 0, 1, 1)]
    for label, naive, dmode, pad8 in combos:
        m.set_option("naive_conv", naive)
        m.set_option("desc_base_mode", dmode)
        m.set_option("halo_pad8", pad8)
        m.write_buffer(0, x)
        m.write_buffer(1, np.zeros((B,) + pr.bufs[1], dtype=np.float16))
        probs = torch.zeros((B, pr.patch, pr.patch), dtype=torch.float32, device="cuda") if op.head else None
        t0 = time.time()
        m.run_ops(B, 0, 1, tta_out, probs)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if op.head:
            got = probs.cpu().numpy()
            err = np.abs(got - ref).max()
        else:
            got = m.read_buffer(1, B).astype(np.float32)[..., op.out_choff:op.out_choff + op.cout]
            err = np.abs(got - ref).max()
            # channels outside the written range must stay zero
            full = m.read_buffer(1, B).astype(np.float32)
            full[..., op.out_choff:op.out_choff + op.cout] = 0
            if np.abs(full).max() != 0:
                print(f"  {label}: WROTE OUTSIDE ITS CHANNEL RANGE")
        res[label] = err
        print(f"  {name:16s} {label:9s} max_abs_err {err:.3e}  ref_max {np.abs(ref).max():.2f}  ({dt*1e3:.1f} ms)", flush=True)
    return res


def run_tap_case(name):
    import numpy as np
    import torch
    import conv_cases
    from digipathai_b200.engine import TileModel
    pr, x, out0, B = conv_cases.build_tap_case(name)
    op = pr.ops[0]
    ref = emu_single(pr, x, 0, residual=out0.astype(np.float32))
    m = TileModel(pr, device=0, max_batch=B)
    for label, naive in (("naive", 1), ("tc", 0)):
        m.set_option("naive_conv", naive)
        m.write_buffer(0, x)
        m.write_buffer(1, out0)
        t0 = time.time()
        m.run_ops(B, 0, 1)
        torch.cuda.synchronize()
        dt = time.time() - t0
        full = m.read_buffer(1, B).astype(np.float32)
        err = np.abs(full[..., op.out_choff:op.out_choff + op.cout] - ref).max()
        keep = out0.astype(np.float32).copy()
        full[..., op.out_choff:op.out_choff + op.cout] = 0
        keep[..., op.out_choff:op.out_choff + op.cout] = 0
        outside = "" if np.array_equal(full, keep) else "  WROTE OUTSIDE ITS CHANNEL RANGE"
        print(f"  {name:16s} {label:9s} max_abs_err {err:.3e}  ref_max {np.abs(ref).max():.2f}  ({dt*1e3:.1f} ms){outside}", flush=True)


def run_full_inception(batch=2, patch=256):
    import numpy as np
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.inception import inception_resnet_v2_unet_program, init_inception_weights
    from oracle import inception_ref as R
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (batch, patch, patch, 3)).astype(np.uint8)
    w = init_inception_weights(0)
    x = (tiles.astype(np.float32) - 128) / 128
    R.calibrate_bn(w, x)
    y = R.forward(w, x)[..., 1]
    prog = inception_resnet_v2_unet_program(w, patch)
    emu, ebufs = emulator.run(prog, tiles, keep=True)
    print(f"  emulator(fp16 storage) vs oracle: {np.abs(emu - y).max():.3e}", flush=True)
    m = TileModel(prog, device=0, max_batch=batch)
    t = torch.from_numpy(tiles).cuda()
    for label, naive in (("naive", 1), ("tc", 0)):
        m.set_option("naive_conv", naive)
        p = m.forward_tile_batch(t).cpu().numpy()
        print(f"  inception forward {label:5s}: vs oracle {np.abs(p - y).max():.3e}  vs emulator {np.abs(p - emu).max():.3e}", flush=True)
        for bi, nm in enumerate(prog.buf_names):
            got = m.read_buffer(bi, batch).astype(np.float32)
            want = ebufs[bi].numpy()
            d = np.abs(got - want)
            print(f"     buf {nm:12s} max_abs_diff {d.max():.3e} mean {d.mean():.3e} (ref max {np.abs(want).max():.2f})", flush=True)


def run_full_deeplab(batch=2, patch=256):
    import numpy as np
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.deeplab import deeplabv3plus_xception_program, init_deeplab_weights
    from oracle import deeplab_ref as R
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (batch, patch, patch, 3)).astype(np.uint8)
    w = init_deeplab_weights(0)
    x = (tiles.astype(np.float32) - 128) / 128
    R.calibrate_bn(w, x)
    y = R.forward(w, x)[..., 1]
    prog = deeplabv3plus_xception_program(w, patch)
    emu, ebufs = emulator.run(prog, tiles, keep=True)
    print(f"  emulator(fp16 storage) vs oracle: max {np.abs(emu - y).max():.3e} mean {np.abs(emu - y).mean():.3e}", flush=True)
    m = TileModel(prog, device=0, max_batch=batch)
    t = torch.from_numpy(tiles).cuda()
    for label, naive in (("naive", 1), ("tc", 0)):
        m.set_option("naive_conv", naive)
        p = m.forward_tile_batch(t).cpu().numpy()
        print(f"  deeplab forward {label:5s}: vs oracle max {np.abs(p - y).max():.3e} mean {np.abs(p - y).mean():.3e}  "
              f"vs emulator max {np.abs(p - emu).max():.3e} mean {np.abs(p - emu).mean():.3e}", flush=True)
        for bi, nm in enumerate(prog.buf_names):
            if nm == "LG":
                continue
            got = m.read_buffer(bi, batch).astype(np.float32)
            want = ebufs[bi].numpy()
            d = np.abs(got - want)
            print(f"     buf {nm:26s} max_abs_diff {d.max():.3e} mean {d.mean():.3e} (ref max {np.abs(want).max():.2f})", flush=True)


def run_full(batch=2, naive_too=True):
    import numpy as np
    import torch
    import emulator
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
    from oracle import densenet_ref as R
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (batch, 256, 256, 3)).astype(np.uint8)
    w = init_densenet_weights(0)
    x = (tiles.astype(np.float32) - 128) / 128
    R.calibrate_bn(w, x)
    y = R.forward(w, x)[..., 1]
    prog = densenet121_unet_program(w, 256)
    emu, ebufs = emulator.run(prog, tiles, keep=True)
    print(f"  emulator(fp16 storage) vs oracle: {np.abs(emu - y).max():.3e}", flush=True)
    m = TileModel(prog, device=0, max_batch=batch)
    t = torch.from_numpy(tiles).cuda()
    for label, naive in (("naive", 1), ("tc", 0)) if naive_too else (("tc", 0),):
        m.set_option("naive_conv", naive)
        p = m.forward_tile_batch(t).cpu().numpy()
        print(f"  full forward {label:5s}: vs oracle {np.abs(p - y).max():.3e}  vs emulator {np.abs(p - emu).max():.3e}", flush=True)
        # per-buffer drift against the emulator (where does an error first appear?)
        for bi, nm in enumerate(prog.buf_names):
            got = m.read_buffer(bi, batch).astype(np.float32)
            want = ebufs[bi].numpy()
            d = np.abs(got - want).max()
            print(f"     buf {nm:12s} max_abs_diff {d:.3e} (ref max {np.abs(want).max():.2f})", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--set", default="dense", choices=["dense", "inception", "deeplab"])
    args = ap.parse_args()
    if args.case:
        import conv_cases
        if args.case == "full":
            run_full()
        elif args.case == "full_inception":
            run_full_inception()
        elif args.case == "full_deeplab":
            run_full_deeplab()
        elif args.case == "full_deeplab64":
            run_full_deeplab(3, 64)
        elif args.case == "full_inception64":
            run_full_inception(3, 64)
        elif args.case in conv_cases.TAP_CASES:
            run_tap_case(args.case)
        else:
            run_conv_case(args.case)
        return
    import conv_cases
    names = list(conv_cases.CASES) + ["full"] if args.set == "dense" else \
        (["full_deeplab64", "full_deeplab"] if args.set == "deeplab" else
         list(conv_cases.TAP_CASES) + ["full_inception64", "full_inception"])
    for name in names:
        print(f"== {name}", flush=True)
        try:
            r = subprocess.run(["timeout", str(args.timeout), sys.executable, os.path.abspath(__file__), "--case", name],
                               capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in out if l.startswith("  ") or "rror" in l or "dp:" in l or "Traceback" in l]
            print("\n".join(keep[-40:]), flush=True)
            print(f"   exit {r.returncode}", flush=True)
        except Exception as e:  # noqa: BLE001
            print("   launcher error", e, flush=True)


if __name__ == "__main__":
    main()
