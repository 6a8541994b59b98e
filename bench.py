#!/usr/bin/env python
"""Benchmark of the getSegmentation hot path (metric of BASELINE.json: tiles/s, 256x256, batch 32).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                           # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of 32 synthetic 256x256x3 uint8 tiles:
tile crop + normalise + DenseNet-121 U-Net forward + softmax channel 1 (dp_forward_tiles), i.e. BASELINE.json
configs[1].  `value` is device-timed with the slide raster already resident in HBM; `e2e` is the same step
through the Keras-``predict``-shaped host API (pinned host uint8 tiles -> H2D -> forward -> D2H of the
probabilities) with the copies inside the timed region.  Every line also carries
  * `slide`: BASELINE configs[2]/[3] -- the whole get_prediction loop (tissue mask + tile grid, forward x 4 TTA
    passes, stitch, normalise) on ONE synthetic 40 000 x 40 000 slide, sharded by x-stripes over the N ranks with one
    halo exchange, plus the same slide on rank 0 alone (`n1_seconds`) so that `speedup_vs_n1` is measured in the run;
  * `parity`: the fp16 forward and the fp32 precision mode against the fp32 oracle on calibrated weights;
  * `fp32_mode`: tiles/s of the precision modes (`--precision fp32|tf32x3` makes one the whole line's subject);
  * `getseg` (N = 1): BASELINE configs[2] through the reference's entry point -- getSegmentation on the same 40 000 x 40 000
    slide including the three pyramidal TIFFs on disk and the float32 map returned to the host.
`--workload slide` runs only the slide part, at `--slide` pixels a side.

The oracle (oracle/) is executed here only for the `cpu_baseline` leg and for `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DTYPE = {"fp16": "f16 (fp32 accumulate)", "fp32": "f32", "tf32x3": "f32 storage, 3xTF32 tensor-core products (fp32 accumulate)"}
METRIC = "tiles_per_sec_256x256_b32"
UNIT = "tiles/s"
PATCH, BATCH = 256, 32
REF_FLOP_PER_TILE = 42316333056  # 21 158 166 528 conv MACs x 2 (reference graph)
REF_FLOP = {"dense": 42316333056, "inception": 57406652416,   # SURVEY.md 8(d); inception: 28 703 326 208 MACs x 2
            "deeplabv3": 25603100672}                          # 12 801 550 336 conv + depthwise MACs x 2
MODEL_DESC = {"dense": "DenseNet-121 U-Net", "inception": "Inception-ResNet-v2 U-Net (--model inception; not the headline config)",
              "deeplabv3": "DeepLabv3+ Xception OS16 (--model deeplabv3; not the headline config)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Starts the sampler process (nvidia-smi needs ~100 ms to come up, so this is called well before the timed
        region); rows are time-stamped on receipt and `mark_begin` / `mark_end` bracket the timed region."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.02]
        note = "sampled every 20 ms inside the timed region"
        if not inside:
            # a timed region shorter than the sampling period: fall back to the samples taken while the same loop
            # ran as warm-up / e2e legs right around it (same kernels, same load)
            inside = [r for (_, r) in self.rows]
            note = "timed region shorter than the 20 ms sampling period: samples from the surrounding warm-up and e2e loops"
        self.note = note
        sm, mx, reasons = [], [], set()
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "note": note}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def oracle_tiles_per_sec(n_tiles: int, threads: int, seed: int = 0, model: str = "dense", budget_s: float = 0.0,
                         batch: int = 4):
    """Times the CPU oracle (fp32 PyTorch-CPU restatement of the reference graph + crop/normalise) on a sample:
    ``n_tiles`` tiles, or -- with ``budget_s`` -- as many batches of ``batch`` (cycling over the sample) as fit."""
    import torch
    torch.set_num_threads(threads)
    rng = np.random.default_rng(seed)
    if model == "deeplabv3":
        from digipathai_b200.models.deeplab import init_deeplab_weights
        from oracle import deeplab_ref as densenet_ref
        w = init_deeplab_weights(0)
    elif model == "inception":
        from digipathai_b200.models.inception import init_inception_weights
        from oracle import inception_ref as densenet_ref
        w = init_inception_weights(0)
    else:
        from digipathai_b200.models.densenet import init_densenet_weights
        from oracle import densenet_ref
        w = init_densenet_weights(0)
    tiles = rng.integers(0, 256, (n_tiles, PATCH, PATCH, 3)).astype(np.uint8)
    densenet_ref.forward(w, (tiles[:1].astype(np.float32) - 128.0) / 128.0)  # warm-up (thread pools, allocs)
    t0 = time.perf_counter()
    done = 0
    s = 0
    while True:
        x = (tiles[s:s + batch].astype(np.float32) - 128.0) / 128.0
        densenet_ref.forward(w, x)
        done += len(x)
        s = (s + batch) % n_tiles
        dt = time.perf_counter() - t0
        if (dt >= budget_s) if budget_s > 0 else (done >= n_tiles):
            break
    return done / dt, dt


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    sample = BATCH                          # one step = one batch of 32 tiles, as on the GPU arm
    steps = max(1, min(args.steps, 8))      # bounded: ~1.5 s of CPU work per step on 16 cores
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_tiles_per_sec(4, cores, model=args.model)
    t_all = 0.0
    for _ in range(steps):
        v, dt = oracle_tiles_per_sec(sample, cores, model=args.model, batch=BATCH)
        vals.append(v); t_all += dt
    value = float(np.sum([sample] * steps) / t_all)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: {MODEL_DESC[args.model]} forward on synthetic 256x256x3 uint8 tiles, "
                               "batch 32 per step",
                   "note": "reference CPU path = oracle port (fp32 torch-CPU restatement of the Keras graph; the "
                           "reference itself needs TensorFlow 1.x and cannot be installed offline)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"one batch of {sample} tiles per step, {steps} steps (of {args.steps} asked: bounded "
                                   "so that the CPU arm ends within minutes)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (200 x 1.8 ms keeps ~18 clock samples in the region)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="forward", choices=["forward", "slide", "config5"])
    ap.add_argument("--model", default="dense", choices=["dense", "inception", "deeplabv3"],
                    help="graph to run (BASELINE configs[1] names the DenseNet U-Net: the default)")
    ap.add_argument("--slide", type=int, default=8192, help="--workload slide: side of the synthetic slide")
    ap.add_argument("--tta", default="", help="--workload slide: comma separated tta_list")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32", "tf32x3"],
                    help="fp16 = tensor cores (BASELINE configs[1]); fp32 / tf32x3 = the library's 1e-3 precision modes "
                         "(fp32 FMA on the CUDA cores / 3xTF32 split products on the tensor cores)")
    ap.add_argument("--no-slide", action="store_true", help="skip the 40 000^2 slide record of the default line")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity record of the default line")
    ap.add_argument("--line-slide", type=int, default=40000, help="side of the slide behind the line's `slide` record")
    ap.add_argument("--split", type=int, default=0, help="sub-batches captured as parallel graph branches (0 = library default)")
    ap.add_argument("--lanes", type=int, default=3,
                    help="forward steps in flight on the GPU (engine.ForwardLanes: one stream + one model clone per lane; "
                         "1 = one stream, L2 flushed between steps as in round 1)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--epi-direct", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-pair", action="store_true")
    ap.add_argument("--no-resident", action="store_true")
    ap.add_argument("--dump-ops", default="", help="write the per-op device-time table (instrumented pass) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    # The library reads a few DP_* environment variables (planner experiments, and DP_DBG_SKIP / DP_NAIVE_CONV which
    # switch work off or swap kernels for bring-up).  A bench line taken with a work-skipping knob set is not a
    # measurement: refuse.  Every other DP_* override in effect is written into the line's config.
    if os.environ.get("DP_DBG_SKIP", "0") not in ("", "0"):
        raise SystemExit("bench.py: DP_DBG_SKIP is set (it disables parts of the conv kernel); unset it")
    env_overrides = {k: v for k, v in sorted(os.environ.items()) if k.startswith("DP_")}

    import torch
    import torch.distributed as dist
    from digipathai_b200 import engine
    from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a collective that cannot complete (mismatched sizes, a dead rank) aborts after 5 min instead of hanging
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=5))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks, peak_src = load_peaks()
    if args.workload == "config5":
        rec = config5_record(args, rank, world, local, dev, barrier, max_over_ranks)
        if rank == 0:
            print(json.dumps({"metric": "ensemble_tiles_per_sec_256x256_b64_crf", "value": rec["tiles_per_s"], "unit": UNIT,
                              "n_gpus": world, "steps": rec["reps"], "warmup": 1, "ms_per_step": rec["seconds"] * 1e3,
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                              "dtype": DTYPE[args.precision], "data": "synthetic",
                              "config": {"workload": rec["workload"]}, "config5": rec,
                              "gpu_launches": rec["gpu_launches"]}))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.model == "deeplabv3":
        from digipathai_b200.models.deeplab import deeplabv3plus_xception_program, init_deeplab_weights
        model = engine.TileModel(deeplabv3plus_xception_program(init_deeplab_weights(0), PATCH, precision=args.precision),
                                 device=local, max_batch=BATCH)
    elif args.model == "inception":
        from digipathai_b200.models.inception import inception_resnet_v2_unet_program, init_inception_weights
        model = engine.TileModel(inception_resnet_v2_unet_program(init_inception_weights(0), PATCH,
                                                                  precision=args.precision), device=local, max_batch=BATCH)
    else:
        weights = init_densenet_weights(0)
        model = engine.TileModel(densenet121_unet_program(weights, PATCH, precision=args.precision), device=local,
                                 max_batch=BATCH)
    ref_flop_per_tile = REF_FLOP[args.model]

    if args.split:
        model.set_option("split", args.split)
    if args.no_graph:
        model.set_option("use_graph", 0)
    if args.no_pdl:
        model.set_option("use_pdl", 0)
    if args.epi_direct:
        model.set_option("epi_direct", 1)
    if args.no_overlap:
        model.set_option("use_overlap", 0)
    if args.no_pair:
        model.set_option("b_pair", 0)
    if args.no_resident:
        model.set_option("b_resident", 0)

    if args.workload == "slide":
        tta_list = [t for t in args.tta.split(",") if t] or None
        rec = slide_record(model, args.slide, tta_list, rank, world, local, dev, barrier, max_over_ranks,
                           reps=max(1, min(args.steps, 3)), with_n1=False)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": rec["tiles_per_s"], "unit": UNIT, "n_gpus": world,
                              "steps": rec["reps"], "warmup": 1, "ms_per_step": rec["seconds"] * 1e3,
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                              "dtype": DTYPE[args.precision], "data": "synthetic",
                              "config": {"workload": rec["workload"]}, "slide": rec,
                              "gpu_launches": rec["gpu_launches"]}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- synthetic input, resident in HBM
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    SW = SH = 16384                          # 805 MB raster: the tiles of a step come from far more than L2 holds
    slide = torch.randint(0, 256, (SW, SH, 3), dtype=torch.uint8, device=dev, generator=g)   # [x, y, c]
    n_coord_sets = args.steps + args.warmup + 8
    cg = torch.Generator(); cg.manual_seed(99 + rank)
    coords_all = torch.randint(0, SW - PATCH, (n_coord_sets, BATCH, 2), generator=cg, dtype=torch.int32).to(dev)
    probs = torch.empty((BATCH, PATCH, PATCH), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step(i):
        model.forward_tiles(slide, coords_all[i % n_coord_sets], 0, 0, out=probs)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                     # nvidia-smi takes ~100 ms to come up: start it before the warm-up
    for i in range(args.warmup):
        step(i)
    barrier()
    # ---- one stream, L2 flushed between steps (the round-1 / round-2 definition; kept as `single_stream`)
    n_single = args.steps if args.lanes <= 1 else min(args.steps, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_single)]
    n0 = engine.kernel_launch_count()
    barrier()
    if args.lanes <= 1:
        sampler.mark_begin()
    for k in range(n_single):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        ev[k][0].record()
        step(args.warmup + k)
        ev[k][1].record()
    barrier()
    if args.lanes <= 1:
        sampler.mark_end()
    launches = engine.kernel_launch_count() - n0
    ms_single = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / n_single
    single_stream = {"value": world * BATCH / (ms_single * 1e-3), "ms_per_step": ms_single, "steps": n_single,
                     "l2": "flushed between timed steps (256 MiB memset, untimed)"}
    ms_per_step, value = ms_single, single_stream["value"]
    if args.lanes > 1:
        # ---- the headline: the same K steps with `lanes` of them in flight (one stream + one model clone per lane).
        # Steps are independent tile batches, as in a slide run; the whole region is timed on the device from the
        # first launch to the completion of the last step.  No flush is possible between overlapping steps: every
        # step's tiles come from random origins of an 805 MB raster and every lane cycles ~1 GB of activations, both
        # far beyond the 126 MB L2 (only the 35 MB of weights stay resident, as they do on a slide).
        pool = engine.ForwardLanes({"m": model}, args.lanes)
        probs_l = [torch.empty((BATCH, PATCH, PATCH), dtype=torch.float32, device=dev) for _ in range(args.lanes)]

        def lane_step(i):
            pool.forward("m", slide, coords_all[i % n_coord_sets], 0, 0, out=probs_l[i % args.lanes])

        pool.begin()
        for i in range(max(args.warmup, 2 * args.lanes)):
            lane_step(i)
        pool.join()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = engine.kernel_launch_count()
        sampler.mark_begin()
        t0.record()
        pool.begin()
        for k in range(args.steps):
            lane_step(args.warmup + k)
        pool.join()
        t1.record()
        barrier()
        sampler.mark_end()
        launches = engine.kernel_launch_count() - n0
        ms_total = max_over_ranks(t0.elapsed_time(t1))
        ms_per_step = ms_total / args.steps
        value = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---------------------------------------------------------------- end to end through the host API
    host_in = torch.randint(0, 256, (BATCH, PATCH, PATCH, 3), dtype=torch.uint8).pin_memory()
    host_out = torch.empty((BATCH, PATCH, PATCH), dtype=torch.float32).pin_memory()
    dev_in = torch.empty((BATCH, PATCH, PATCH, 3), dtype=torch.uint8, device=dev)

    def e2e_step():
        dev_in.copy_(host_in, non_blocking=True)
        model.forward_tile_batch(dev_in, 0, 0, out=probs)
        host_out.copy_(probs, non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_serial_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_serial_value = world * BATCH * args.steps / (e2e_serial_ms * 1e-3)
    # The public host-buffer API for a stream of batches (engine.HostBatchPipeline): same per-step copies (every
    # step's inputs leave pinned host memory, every step's probabilities land in pinned host memory), but H2D of
    # batch k+1 / forward of batch k / D2H of batch k-1 overlap on three streams.  Timed from the first H2D to the
    # completion of the last D2H.
    pipe = engine.HostBatchPipeline(model, BATCH, lanes=max(1, args.lanes))
    host_out_flat = host_out
    for _ in range(3):
        pipe.submit(host_in, host_out_flat)
    pipe.drain()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(pipe.s_in)
    for _ in range(args.steps):
        pipe.submit(host_in, host_out_flat)
    p1.record(pipe.s_out)
    pipe.drain()
    barrier()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(p0.elapsed_time(p1))
    e2e_value = world * BATCH * args.steps / (e2e_ms * 1e-3)

    clocks = sampler.stop() if rank == 0 else None

    # ---------------------------------------------------------------- per-kernel roofline (instrumented pass)
    model.set_option("profile", 1)
    n_prof = min(10, args.steps)
    acc = None
    for k in range(n_prof):
        flush.zero_()
        step(k)
        t = model.op_times_ms()
        acc = t if acc is None else acc + t
    model.set_option("profile", 0)
    per_op = acc / n_prof
    infos = [model.op_info(i) for i in range(model.n_ops())]
    conv_ms = float(sum(t for t, inf in zip(per_op, infos) if inf["type"] in (3, 6)))
    all_ms = float(per_op.sum())
    aux_ms = all_ms - conv_ms               # gather, pools, BN passes: HBM-bound helpers, accurately timed one by one
    n_conv = sum(1 for inf in infos if inf["type"] in (3, 6))
    flop_step = ref_flop_per_tile * BATCH
    # The conv family's time is taken from the TIMED step, not from the instrumented pass: inside the CUDA graph the
    # kernels overlap through programmatic dependent launch, so event-bracketed launches sum to more than the step
    # they decompose.  family time = ms_per_step - (helper kernels' time)  <=  ms_per_step.
    # With several steps in flight the helper kernels of one step overlap the conv kernels of another, so nothing is
    # subtracted: achieved = algorithmic FLOPs / the whole timed step.
    family_ms = max(ms_per_step - aux_ms, 1e-6) if args.lanes <= 1 else ms_per_step
    achieved_tf = flop_step / (family_ms * 1e-3) / 1e12
    exec_macs = model.executed_macs(BATCH)
    peak_tf = peaks["bf16_tflops_sustained"]
    top = sorted(range(len(per_op)), key=lambda i: -per_op[i])[:5]
    top_desc = [{"op": i, "kind": infos[i]["kind"], "cin": infos[i]["cin"], "cout": infos[i]["cout"],
                 "hw": infos[i]["h"], "ms": round(float(per_op[i]), 4),
                 "exec_tflops": round(2 * infos[i]["macs_per_tile"] * BATCH / (per_op[i] * 1e-3) / 1e12, 1)
                 if per_op[i] > 0 and infos[i]["macs_per_tile"] else None} for i in top]

    if args.dump_ops and rank == 0:
        names = [o.name for o in model.program.ops]
        with open(args.dump_ops, "w") as f:
            f.write("op,name,type,kind,cin,cout,hw,ms,exec_tflops\n")
            for i, (t, inf) in enumerate(zip(per_op, infos)):
                tf = 2 * inf["macs_per_tile"] * BATCH / (t * 1e-3) / 1e12 if t > 0 and inf["macs_per_tile"] else 0
                f.write(f"{i},{names[i]},{inf['type']},{inf['kind']},{inf['cin']},{inf['cout']},{inf['h']},"
                        f"{t:.4f},{tf:.1f}\n")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # reported at N = 1 only (the host cores are shared)
        threads = os.cpu_count() or 1
        v, dt = oracle_tiles_per_sec(32, threads, model=args.model, budget_s=12.0)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{int(round(v * dt))} tiles of the same workload in batches of 4 "
                                  f"({dt:.1f} s of CPU work)"}

    # ---------------------------------------------------------------- whole-step DRAM traffic (from the committed ncu pass)
    traffic, traffic_note = None, "no ncu capture on file for this model / precision"
    tp = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
    if os.path.exists(tp) and args.model == "dense" and args.precision == "fp16":
        td = json.load(open(tp))
        traffic = td.get("dram_bytes_per_step")
        traffic_note = td.get("note", "")

    # ---------------------------------------------------------------- parity on calibrated weights; precision mode
    parity = fp32_mode = None
    if rank == 0 and args.model == "dense" and not args.no_parity:
        parity, fp32_mode = parity_and_fp32_records(local, model if args.precision == "fp16" else None)
    barrier()

    # ---------------------------------------------------------------- BASELINE configs[2]/[3]: one 40k slide over N ranks
    slide_rec = None
    if args.model == "dense" and not args.no_slide:
        slide = flush = None                 # release the forward leg's raster and flush buffer
        torch.cuda.empty_cache()
        slide_rec = slide_record(model, args.line_slide, ["FLIP_LEFT_RIGHT", "ROTATE_90", "ROTATE_180"], rank, world, local,
                                 dev, barrier, max_over_ranks, reps=1, with_n1=(world > 1))

    # ---------------------------------------------------------------- BASELINE configs[2] through the public entry point
    getseg = None
    if rank == 0 and world == 1 and args.model == "dense" and not args.no_slide and args.precision == "fp16":
        getseg = getseg_record(args.line_slide, dev)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE[args.precision], "data": "synthetic",
            "config": {"workload": f"configs[1]: {MODEL_DESC[args.model]} forward on synthetic 256x256x3 uint8 tiles, "
                                   "batch 32 per GPU, tiles cropped from an HBM-resident raster",
                       "l2": ("flushed between timed steps (256 MiB memset, untimed)" if args.lanes <= 1 else
                              "inputs larger than L2: tiles cropped at random origins of an 805 MB raster, ~1 GB of "
                              "activations per lane; steps overlap, so no flush between them (single_stream has the "
                              "flushed one-stream figure)"),
                       "lanes": args.lanes,
                       "env_overrides": env_overrides,
                       "weights": "random-init (He-normal), BN stats (0,1)",
                       "graph": "off" if args.no_graph else f"one CUDA graph per step, {args.split or 1} sub-batch branch(es), PDL between conv kernels",
                       "executed_gflop_per_tile": 2 * exec_macs / BATCH / 1e9,
                       "reference_gflop_per_tile": ref_flop_per_tile / 1e9},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_in.numel()),
                    "d2h_bytes_per_step": int(host_out.numel() * 4),
                    "api": f"engine.HostBatchPipeline(lanes={max(1, args.lanes)}): pinned host buffers, H2D / forward / "
                           "D2H on separate streams, one device buffer pair per step in flight; every step copies its "
                           "own inputs and results",
                    "serial_value": e2e_serial_value,
                    "serial_api": "copy in, TileModel.forward_tile_batch, copy out on one stream (no overlap)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "kernel": (f"conv_tc_kernel{' + dense_layer_kernel' if args.model == 'dense' else ''}, tcgen05 implicit-GEMM family"
                                    if args.precision == "fp16" else "conv_f32_kernel (fp32 FMA implicit GEMM)") +
                                   f" ({n_conv} launches per step); achieved = algorithmic FLOPs of the reference graph per step "
                                   + ("/ (timed ms_per_step - helper-kernel ms)" if args.lanes <= 1 else
                                      f"/ timed ms_per_step ({args.lanes} steps in flight; helper kernels not subtracted)"),
                         "peak_source": f"{peak_src} bf16_tflops_sustained",
                         "family_ms_per_step": family_ms, "aux_ms_per_step": aux_ms,
                         "frac_whole_step": flop_step / (ms_per_step * 1e-3) / 1e12 / peak_tf,
                         "instrumented_conv_ms": conv_ms, "instrumented_all_ops_ms": all_ms,
                         "traffic_note": traffic_note, "top_ops": top_desc},
            "cpu_baseline": cpu_baseline,
            "single_stream": single_stream,
        }
        if parity is not None:
            line["parity"] = parity
        if fp32_mode is not None:
            line["fp32_mode"] = fp32_mode
        if slide_rec is not None:
            line["slide"] = slide_rec
        if getseg is not None:
            line["getseg"] = getseg
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def getseg_record(S, dev):
    """BASELINE configs[2] end to end through the reference's own entry point: ``getSegmentation`` on one synthetic S x S
    slide resident in HBM (3 TTA transforms = 4 passes, batch 32, stride 128), the three pyramidal JPEG TIFFs written to a
    temporary directory, the float32 {0, 255} map returned to the host.  Host wall clock of the second of two calls (the
    first pays graph capture and lane allocation inside the fresh models it builds); N = 1 only."""
    import tempfile
    import torch
    from digipathai_b200 import tiffio
    from digipathai_b200.Segmentation import getSegmentation
    from digipathai_b200.models.densenet import init_densenet_weights
    from digipathai_b200.slide import synthetic_slide_device
    levels = 1
    while S // (2 ** (levels - 1)) > 2500 and levels < 5:
        levels += 1
    slide = synthetic_slide_device(S, S, dev, seed=0, n_levels=levels)
    w = init_densenet_weights(0)
    save_t = []
    orig = tiffio.save_pyramidal

    def timed(*a, **k):
        t = time.perf_counter()
        r = orig(*a, **k)
        save_t.append(time.perf_counter() - t)
        return r

    tiffio.save_pyramidal = timed
    runs = []
    try:
        with tempfile.TemporaryDirectory() as d:
            for _ in range(2):
                save_t.clear()
                phases = {}
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = getSegmentation(slide, timings=phases, patch_size=PATCH, stride_size=128, batch_size=BATCH, quick=True,
                                      tta_list=["FLIP_LEFT_RIGHT", "ROTATE_90", "ROTATE_180"], crf=False,
                                      save_path=os.path.join(d, "mask.tiff"), probs_path=os.path.join(d, "probs.tiff"),
                                      uncertainty_path=os.path.join(d, "unc.tiff"), weights=w, status={})
                runs.append((time.perf_counter() - t0, sum(save_t)))
                shape = tuple(out.shape)
                del out
                file_bytes = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    finally:
        tiffio.save_pyramidal = orig
    del slide
    torch.cuda.empty_cache()
    return {"workload": f"getSegmentation(img, 256, 128, 32, quick=True, tta_list of 3, crf=False, save_path / probs_path / "
                        f"uncertainty_path) on one synthetic {S}x{S} slide resident in HBM; returns float32 {shape}",
            "seconds": runs[-1][0], "write_seconds": runs[-1][1], "first_call_seconds": runs[0][0],
            "phases_ms": {k: round(v, 1) for k, v in phases.items()},
            "result_file_bytes": int(file_bytes), "host_cores": os.cpu_count(),
            "note": "JPEG tiles of the three pyramidal TIFFs are encoded on the GPU (csrc/jpeg_enc.cuh)"}


def parity_and_fp32_records(local, fp16_model):
    """(parity, fp32_mode) records of the bench line: both arithmetic modes against the fp32 oracle on CALIBRATED
    weights (BN statistics from an oracle pass, so activations are O(1) and the softmax is informative), 4 tiles; and
    the precision mode's throughput at batch 32 (device-timed, same synthetic tiles as the headline)."""
    import torch
    from digipathai_b200 import engine
    from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
    from oracle import densenet_ref
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (4, PATCH, PATCH, 3)).astype(np.uint8)
    x = (tiles.astype(np.float32) - 128.0) / 128.0
    w = init_densenet_weights(0)
    densenet_ref.calibrate_bn(w, x[:2])
    want = densenet_ref.forward(w, x)[..., 1]
    t = torch.from_numpy(tiles).cuda(local)

    def rec(got):
        d = np.abs(got - want)
        mism = (got >= 0.3) != (want >= 0.3)
        return {"max_abs": float(d.max()), "mean_abs": float(d.mean()), "label_mismatch": int(mism.sum()),
                "band_pixels": int((np.abs(want - 0.3) <= d.max()).sum()),
                "mismatch_outside_band": int((np.abs(want - 0.3)[mism] > d.max()).sum()), "pixels": int(want.size)}

    m16 = engine.TileModel(densenet121_unet_program(w, PATCH), device=local, max_batch=4)
    r16 = rec(m16.forward_tile_batch(t).cpu().numpy())
    m16.close()
    tb = torch.randint(0, 256, (BATCH, PATCH, PATCH, 3), dtype=torch.uint8, device=f"cuda:{local}")
    out = torch.empty((BATCH, PATCH, PATCH), dtype=torch.float32, device=f"cuda:{local}")
    recs, modes = {}, {}
    what = {"fp32": "precision='fp32': fp32 weights / activations / FMA accumulation on the CUDA cores (csrc/precise.cuh), batch 32",
            "tf32x3": "precision='tf32x3': fp32 weights / activations, every product as three tcgen05 kind::tf32 MMAs on "
                      "hi/lo split operands, chunked accumulation (csrc/precise_tc.cuh), batch 32"}
    for prec in ("fp32", "tf32x3"):
        m32 = engine.TileModel(densenet121_unet_program(w, PATCH, precision=prec), device=local, max_batch=BATCH)
        recs[prec] = rec(m32.forward_tile_batch(t).cpu().numpy())
        for _ in range(2):
            m32.forward_tile_batch(tb, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n = 5
        e0.record()
        for _ in range(n):
            m32.forward_tile_batch(tb, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        m32.close()
        modes[prec] = {"tiles_per_s": BATCH / (ms * 1e-3), "ms_per_step": ms, "unit": UNIT,
                       "achieved_tflops_fp32": REF_FLOP["dense"] * BATCH / (ms * 1e-3) / 1e12, "what": what[prec]}
    parity = {"against": "fp32 oracle (oracle/densenet_ref.py) on calibrated seed-0 weights, 4 uniform-noise tiles",
              "threshold": 0.3, "fp16": r16, "fp32": recs["fp32"], "tf32x3": recs["tf32x3"],
              "note": "the fp16 figure is this random-init instance's own amplification of one 10-bit-mantissa rounding "
                      "(profiles/r2_parity_conditioning.md), not a kernel error; the fp32 and tf32x3 modes carry BASELINE's 1e-3"}
    fp32_mode = modes["fp32"]
    fp32_mode["tf32x3"] = modes["tf32x3"]
    return parity, fp32_mode


def config5_record(args, rank, world, local, dev, barrier, max_over_ranks):
    """BASELINE configs[4]: the reference's ensemble (quick=False: DenseNet + Inception-ResNet-v2 + DeepLabv3+ U-Nets,
    Segmentation.py:288-291) at batch 64 on one synthetic slide, threshold 0.3, then crf=True (post_process_crf on the
    permutohedral lattice, per 256 x 256 block), sharded by x-stripes over the ranks.  One untimed run, then timed
    runs (host wall clock between device-synchronising barriers, max over ranks)."""
    import torch
    from digipathai_b200 import engine
    from digipathai_b200.dist import sharded_get_prediction
    from digipathai_b200.models.deeplab import deeplabv3plus_xception_program, init_deeplab_weights
    from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
    from digipathai_b200.models.inception import inception_resnet_v2_unet_program, init_inception_weights
    from digipathai_b200.slide import synthetic_slide_device
    B5, S = 64, args.slide
    prec = args.precision
    models = {
        "dense": engine.TileModel(densenet121_unet_program(init_densenet_weights(0), PATCH, precision=prec), local, B5),
        "inception": engine.TileModel(inception_resnet_v2_unet_program(init_inception_weights(0), PATCH, precision=prec), local, B5),
        "deeplabv3": engine.TileModel(deeplabv3plus_xception_program(init_deeplab_weights(0), PATCH, precision=prec), local, B5),
    }
    levels = 1
    while S // (2 ** (levels - 1)) > 2500 and levels < 5:
        levels += 1
    slide = synthetic_slide_device(S, S, dev, seed=0, n_levels=levels)
    reps = max(1, min(args.steps, 3))
    n0 = engine.kernel_launch_count()
    times, info, grid = [], None, None
    for it in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        grid, out, info = sharded_get_prediction(slide, models, B5, None, PATCH, 128, device=local, threshold=0.3, crf=True)
        barrier()
        times.append(max_over_ranks(time.perf_counter() - t0))
        del out
    best = min(times[1:])
    tm = info["timings_ms"]
    n_tiles = len(grid.coords)
    crf_tiles = int(max_over_ranks(float(info["crf_tiles"])))
    rec = {"workload": f"configs[4]: 3-model ensemble (DenseNet-121 / Inception-ResNet-v2 / DeepLabv3+ U-Nets) + crf=True "
                       f"on one synthetic {S}x{S} slide, patch 256 stride 128 batch 64, 1 pass, {n_tiles} tiles = "
                       f"{3 * n_tiles} tile-forwards, x-stripe sharding over {world} rank(s)",
           "n_gpus": world, "tiles": n_tiles, "seconds": best, "tiles_per_s": n_tiles / best, "reps": reps,
           "all_seconds": [round(t, 4) for t in times], "crf_blocks_max_rank": crf_tiles,
           "loop_ms": max_over_ranks(float(tm.get("loop_ms", 0.0))), "crf_ms": max_over_ranks(float(tm.get("crf_ms", 0.0))),
           "halo_ms": max_over_ranks(float(tm.get("halo_ms", 0.0))), "grid_ms": max_over_ranks(float(tm.get("grid_ms", 0.0))),
           "crf": "permutohedral lattice (dp_crf_tiles_lattice), 10 mean-field iterations, non-overlapping 256x256 blocks",
           "gpu_launches": int(engine.kernel_launch_count() - n0)}
    for m in models.values():
        m.close()
    return rec


def slide_record(model, S, tta_list, rank, world, local, dev, barrier, max_over_ranks, reps=1, with_n1=False):
    """BASELINE configs[2] (N = 1) / configs[3] (N > 1): the whole get_prediction loop on ONE synthetic S x S slide
    generated in HBM -- tissue mask + tile grid, forward x TTA passes, stitch, halo exchange, normalise -- sharded by
    x-stripes over the N ranks (dist.sharded_get_prediction; N = 1 goes through the same function).  One untimed
    warm-up run, then `reps` timed runs: host wall clock between barriers that synchronise the device, max over
    ranks.  `with_n1`: rank 0 afterwards runs the same slide alone, so the speed-up is measured inside this run."""
    import torch
    from digipathai_b200 import engine
    from digipathai_b200.dist import sharded_get_prediction
    from digipathai_b200.slide import synthetic_slide_device
    levels = 1
    while S // (2 ** (levels - 1)) > 2500 and levels < 5:
        levels += 1
    slide = synthetic_slide_device(S, S, dev, seed=0, n_levels=levels)
    n_pass = 1 + (len(tta_list) if tta_list else 0)
    n0 = engine.kernel_launch_count()
    times, info, grid = [], None, None
    for it in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        grid, out, info = sharded_get_prediction(slide, {"dense": model}, BATCH, tta_list, PATCH, 128, device=local)
        barrier()
        times.append(max_over_ranks(time.perf_counter() - t0))
        del out
    launches = engine.kernel_launch_count() - n0
    best = min(times[1:])
    tm = info["timings_ms"]
    phase = {k: max_over_ranks(float(tm.get(k, 0.0))) for k in ("grid_ms", "upload_ms", "loop_ms", "halo_ms", "normalise_ms")}
    halo_bytes = int(max_over_ranks(float(info["halo_bytes_sent"])))
    n_tiles = len(grid.coords)
    rec = {"workload": f"get_prediction on one synthetic {S}x{S} slide ({levels} virtual levels) resident in HBM, patch 256 "
                       f"stride 128 batch 32, {n_pass} passes (tta_list={tta_list}), {n_tiles} tiles after drop_last = "
                       f"{n_tiles * n_pass} tile-forwards; x-stripe sharding over {world} rank(s), one P2P halo exchange",
           "n_gpus": world, "tiles": n_tiles, "tile_forwards": n_tiles * n_pass, "seconds": best,
           "tiles_per_s": n_tiles * n_pass / best, "reps": reps, "all_seconds": [round(t, 4) for t in times],
           "timing": "host wall clock between device-synchronising barriers, max over ranks; first run untimed",
           "prologue_ms": phase["grid_ms"] + phase["upload_ms"], "grid_ms": phase["grid_ms"],
           "upload_ms": phase["upload_ms"], "loop_ms": phase["loop_ms"], "halo_ms": phase["halo_ms"],
           "normalise_ms": phase["normalise_ms"], "halo_bytes": halo_bytes, "gpu_launches": int(launches),
           "n1_seconds": None, "speedup_vs_n1": None}
    if world == 1:
        rec["n1_seconds"], rec["speedup_vs_n1"] = best, 1.0
    elif with_n1:
        barrier()
        t1 = None
        if rank == 0:
            t0 = time.perf_counter()
            _, out, _ = sharded_get_prediction(slide, {"dense": model}, BATCH, tta_list, PATCH, 128, device=local,
                                               shard=(0, 1))
            torch.cuda.synchronize()
            t1 = time.perf_counter() - t0
            del out
        barrier()
        if rank == 0:
            rec["n1_seconds"], rec["speedup_vs_n1"] = t1, t1 / best
            rec["n1_note"] = "same slide, same process, rank 0 alone (shard=(0, 1)); one run, model already warm"
    del slide
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    main()
