"""B200-native drop-in for ``DigiPathAI/Segmentation.py``: ``getSegmentation`` and ``get_prediction``.

Same call signatures, ``status`` protocol, return values and error behaviour as the reference
(DigiPathAI/Segmentation.py:65-76,192-205 and README.md:79-87 -- the union of the three signatures the
reference ships), but the batch loop runs on the GPU through ``libdigipath_b200.so``:

    reference (per batch, host)                              here (per batch, device-resident)
    ---------------------------------------------------      ------------------------------------------------
    DataLoader workers: read_region + (v-128)/128            tile crop + normalise fused into the stem gather
    apply_tta -> Model.predict -> transform_prob  (x N)      dp_forward_tiles(tta_in, tta_out)            (x N)
    np.mean / np.var over passes, `+=` into 3 memmaps        dp_stitch (deterministic, reference add order)
    count==0 -> 1, mean /= count, var /= count**2            dp_finalize
    mean >= 0.3 -> 255 else 0                                dp_finalize (label plane)

Reference quirks kept on purpose: cumulative in-place TTA (Q1), unknown TTA names are identity
passes (Q2), ``drop_last`` (Q3), [x, y] plane orientation (Q4), clamped tile origins (Q5), uint8 count that
wraps and var / count**2 (Q6), levels > 4 raise (Q7), ``quick`` meaning and ignored ``crf`` / ``mask_level``
(Q8), only softmax channel 1 is used (Q9).
"""
from __future__ import annotations

import os
from os.path import expanduser

import numpy as np

from . import tta as _tta
from .slide import open_slide, upload_xy_raster
from .tissue import TileGrid

home = expanduser("~")

_MODEL_SETS = {  # Segmentation.py:232-278
    "colon": ("digestpath_models", "digestpath"),
    "liver": ("paip_models", "paip"),
    "breast": ("camelyon_models", "camelyon"),
}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("digipathai_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def get_prediction(wsi_path, mask_path=None, label_path=None, batch_size=64, models=None, tta_list=None,
                   num_workers=8, verbose=0, patch_size=256, stride_size=256, mask_level=-1, status=None,
                   *, device=0, tile_range=None, return_device=False, finalize=True, tissue_mask=None, grid=None,
                   timings=None, lanes=None):
    """Patch based segmentor (reference: Segmentation.py:65-189).

    ``models`` maps a name to a ``TileModel`` (engine.py) -- the object that replaces the Keras model.
    Returns ``(slide, {'mean': [W,H] float32, 'var': [W,H] float32})`` like the reference (numpy arrays instead
    of memmaps).  Keyword-only extensions: ``device`` (CUDA ordinal), ``tile_range=(lo, hi)`` restricts the run
    to a slice of the post-``drop_last`` tile list (multi-GPU sharding, dist.py), ``return_device`` keeps the
    planes as torch CUDA tensors, ``finalize=False`` skips the normalisation (sharded runs normalise after the
    halo exchange), ``tissue_mask`` supplies a precomputed RAW [x, y] mask instead of the Otsu/HSV heuristic (the
    morphology of utils.py:200-219 is still applied to it), ``grid`` a ready ``TileGrid`` of this slide (sharded
    runs build it once and index it with ``tile_range``), ``timings`` a dict that receives this call's phase times in
    milliseconds (grid, raster upload, tile loop; each phase ends with a device synchronisation), ``lanes`` the number
    of forwards kept in flight on the GPU (engine.ForwardLanes: the passes / models of a batch and consecutive batches
    run on ``lanes`` streams, each with its own clone of every model; default ``DIGIPATH_LANES`` or 3; the planes do
    not depend on it, bit for bit).
    ``mask_path`` names a ``.tiff`` tissue mask read at the slide's top level instead of running the heuristic
    (dataloader.py:256-263; any other extension leaves the reference's dataset without a mask -- AttributeError --
    and does so here); ``label_path`` / ``num_workers`` / ``mask_level`` are accepted for signature compatibility:
    the live reference passes None for the first and ignores the last (dataloader.py:240-241).
    """
    from . import engine
    torch = _torch()
    if not models:
        raise ValueError("get_prediction needs at least one model")
    slide = open_slide(wsi_path)
    if grid is None and tissue_mask is None and mask_path is not None:
        from .tissue import mask_from_slide
        if not os.path.basename(mask_path).endswith('.tiff'):
            raise AttributeError("mask_path must name a .tiff file: the reference's dataset is left without a "
                                 "'_mask' attribute for any other extension (dataloader.py:256-263)")
        tissue_mask = mask_from_slide(open_slide(mask_path), len(slide.level_dimensions) - 1)
    import time
    _t0 = time.perf_counter()
    if grid is None:
        grid = TileGrid(slide, patch_size=patch_size, stride_size=stride_size, batch_size=batch_size,
                        roi_masking=True, mask=tissue_mask, device=device)
        if timings is not None:
            timings['grid_ms'] = (time.perf_counter() - _t0) * 1e3
    elif (grid.patch_size, grid.stride_size, grid.batch_size) != (int(patch_size), int(stride_size), int(batch_size)):
        raise ValueError("grid= was built for another patch / stride / batch size")
    n_batches = len(grid)
    print("Length of DataLoader: {}".format(n_batches))
    passes = _tta.pass_codes(tta_list)
    names = list(models.keys())
    for nm in names:
        if models[nm].patch != patch_size:
            raise ValueError(f"model '{nm}' was built for patch {models[nm].patch}, not {patch_size}")
        if models[nm].max_batch < batch_size:
            raise ValueError(f"model '{nm}' holds buffers for {models[nm].max_batch} tiles, batch is {batch_size}")
    W, H = slide.level_dimensions[0]
    dev = torch.device("cuda", device)
    P = int(patch_size)

    coords_all = grid.coords
    b_lo, b_hi = 0, n_batches
    if tile_range is not None:
        lo, hi = tile_range
        if lo % batch_size or hi % batch_size:
            raise ValueError("tile_range must be aligned to the batch size")
        b_lo, b_hi = lo // batch_size, hi // batch_size
    # stripe of the planes this call owns: [x_lo, x_hi)
    if tile_range is None or b_hi <= b_lo:
        x_lo, x_hi = 0, W
    else:
        sel = coords_all[b_lo * batch_size:b_hi * batch_size, 0]
        x_lo, x_hi = int(sel.min()), int(sel.max()) + P
    with torch.cuda.device(dev):
        _t0 = time.perf_counter()
        raster = upload_xy_raster(slide, x_lo, x_hi, dev)                          # uint8 [x, y, c] stripe in HBM
        if timings is not None:
            torch.cuda.synchronize(dev)
            timings['upload_ms'] = (time.perf_counter() - _t0) * 1e3
            _t0 = time.perf_counter()
        mean = torch.zeros((x_hi - x_lo, H), dtype=torch.float32, device=dev)
        var = torch.zeros_like(mean)
        count = torch.zeros((x_hi - x_lo, H), dtype=torch.uint8, device=dev)
        n_pass = len(passes) * len(names)
        if lanes is None:
            lanes = int(os.environ.get("DIGIPATH_LANES", "3"))
        if b_hi <= b_lo or not all(hasattr(m, "clone") for m in models.values()):
            lanes = 1
        pool = engine.ForwardLanes(models, lanes)
        # one probability buffer per batch in flight (+1 being stitched): [slot][pass x model][tile][P][P]
        n_slots = 1 if pool.L == 1 else max(2, -(-pool.L // n_pass) + 1)
        probs = [torch.empty((n_pass, batch_size, P, P), dtype=torch.float32, device=dev) for _ in range(n_slots)]
        stitched = [None] * n_slots
        coords_dev = torch.from_numpy(coords_all - np.array([x_lo, 0], np.int32)).to(dev) if len(coords_all) else None
        coords_abs = torch.from_numpy(coords_all).to(dev) if len(coords_all) else None
        pool.begin()                                 # the lanes start after the raster upload and the plane memsets
        for ii in range(b_lo, b_hi):
            if status is not None:
                # same arithmetic as Segmentation.py:139 (an ensemble therefore tops out below 100 %)
                status['progress'] = int(ii * 100.0 / (len(names) * n_batches))
            c_local = coords_dev[ii * batch_size:(ii + 1) * batch_size]
            slot = (ii - b_lo) % n_slots
            if stitched[slot] is not None:
                pool.wait_event(stitched[slot])      # the stitch that last read probs[slot] has finished
            k = 0
            for (t_in, t_out) in passes:            # Segmentation.py:150-160: tta outer, model inner
                for nm in names:
                    pool.forward(nm, raster, c_local, t_in, t_out, out=probs[slot][k])
                    k += 1
            pool.join()
            engine.stitch(probs[slot], coords_abs[ii * batch_size:(ii + 1) * batch_size], mean, var, count, x_lo=x_lo)
            if pool.L > 1:
                stitched[slot] = torch.cuda.Event()
                stitched[slot].record()
        pool.close()
        if timings is not None:
            torch.cuda.synchronize(dev)
            timings['loop_ms'] = (time.perf_counter() - _t0) * 1e3
        if finalize:
            engine.finalize(mean, var, count, 0.0, None)
        torch.cuda.synchronize(dev)
    out = {'mean': mean, 'var': var}
    if not finalize:
        out['count'] = count
    out['x_range'] = (x_lo, x_hi)
    if not return_device:
        out = {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in out.items()}
    return (slide, out)


def crf_block_starts(W, P):
    """First columns of the non-overlapping P-wide CRF blocks of a slide W columns wide (the last one clamped inside)."""
    return sorted({min(x, W - P) for x in range(0, W, P)})


def _crf_refine(engine, torch, slide, mean, label, P, batch, x_lo=0, block_xs=None):
    """``crf=True``: the reference's intent (Segmentation.py:327-331, commented out there) is
    ``post_process_crf(img, [1 - mean, mean], 2)`` on the whole slide, which pydensecrf cannot hold in memory at WSI
    size.  Here the fully connected CRF of utils.py:568-603 is applied per non-overlapping P x P block of the slide
    (blocks at the far edges clamped inside) wherever the probability map is not identically background; the MAP label
    replaces the thresholded label ({0, 255}).  Planes are in the reference's [x, y] orientation; the CRF is symmetric
    in the two image axes, so no transpose is needed.

    ``mean`` / ``label`` may be a stripe of the slide starting at column ``x_lo`` (sharded runs); ``block_xs`` then
    lists the first columns of the blocks this call owns (dist.py assigns every block to exactly one rank)."""
    Ws, H = mean.shape
    W = slide.level_dimensions[0][0]
    if W < P or H < P:
        return 0
    xs = crf_block_starts(W, P) if block_xs is None else sorted(block_xs)
    xs = [x for x in xs if x >= x_lo and x + P <= x_lo + Ws]         # the block must lie inside this stripe
    ys = sorted({min(y, H - P) for y in range(0, H, P)})
    if not xs:
        return 0
    raster = upload_xy_raster(slide, x_lo, x_lo + Ws, mean.device)    # uint8 [x, y, c] of the stripe
    # block maxima in one pass: blocks whose probability map is identically background are skipped
    col_max = torch.stack([mean[x - x_lo:x - x_lo + P].amax(0) for x in xs])              # [nx, H]
    blk_max = torch.stack([col_max[:, y:y + P].amax(1) for y in ys], 1).cpu().numpy()     # [nx, ny]
    todo = [(xs[i], ys[j]) for i in range(len(xs)) for j in range(len(ys)) if blk_max[i, j] > 1e-5]
    for s in range(0, len(todo), batch):
        chunk = todo[s:s + batch]
        rgb = torch.stack([raster[x - x_lo:x - x_lo + P, y:y + P] for x, y in chunk]).contiguous()
        p1 = torch.stack([mean[x - x_lo:x - x_lo + P, y:y + P] for x, y in chunk]).contiguous()
        lab = engine.dense_crf(rgb, p1)
        for (x, y), l in zip(chunk, lab):
            label[x - x_lo:x - x_lo + P, y:y + P] = l * 255
    return len(todo)


def _default_weight_path(mode, model):
    folder, prefix = _MODEL_SETS[mode]
    suffix = {'dense': 'densenet', 'inception': 'inception', 'deeplabv3': 'deeplabv3'}[model]
    return os.path.join(home, '.DigiPathAI', folder, f'{prefix}_{suffix}.npz')


def load_trained_models(model, path, patch_size=256, *, device=0, max_batch=32, precision="fp16"):
    """Counterpart of utils.py:427-448: build the graph and load its weights; returns a ``TileModel``.

    ``path`` is one of the reference's Keras ``.h5`` checkpoints (read in pure Python: h5lite.py + keras_h5.py), a
    flat ``.npz`` of Keras-named arrays (tools/h5_to_npz.py output) or an in-memory weight dict.
    ``precision``: 'fp16' (tensor cores, fp32 accumulation), 'fp32' (the reference's arithmetic on the CUDA cores:
    matches its fp32 ``Model.predict`` within 1e-3 on any weight set; ~20x slower) or 'tf32x3' (fp32 storage, every
    product as three TF32 tensor-core MMAs: same tolerance, ~2x faster than 'fp32') -- see program.py.
    """
    from .engine import TileModel
    from .models.densenet import densenet121_unet_program
    from .models.inception import inception_resnet_v2_unet_program
    if model.__contains__('dense'):
        weights = path if isinstance(path, dict) else _load_weights('dense', path)
        return TileModel(densenet121_unet_program(weights, patch_size, precision=precision), device=device,
                         max_batch=max_batch)
    if model.__contains__('inception'):
        weights = path if isinstance(path, dict) else _load_weights('inception', path)
        return TileModel(inception_resnet_v2_unet_program(weights, patch_size, precision=precision), device=device,
                         max_batch=max_batch)
    if model.__contains__('deeplabv3'):
        from .models.deeplab import deeplabv3plus_xception_program
        weights = path if isinstance(path, dict) else _load_weights('deeplabv3', path)
        return TileModel(deeplabv3plus_xception_program(weights, patch_size, precision=precision), device=device,
                         max_batch=max_batch)
    raise ValueError("Unknown model provided, allowed models ['dense', 'inception', 'deeplabv3']")


def _load_weights(model, path):
    """``.h5`` / ``.hdf5`` -> Keras checkpoint (what utils.py:447 hands to ``load_weights``); anything else -> ``.npz``.
    When the ``.npz`` the default paths name is missing but the reference's ``.h5`` sits beside it, that is read."""
    path = str(path)
    h5 = path if path.endswith(('.h5', '.hdf5')) else None
    if h5 is None and not os.path.exists(path) and os.path.exists(os.path.splitext(path)[0] + '.h5'):
        h5 = os.path.splitext(path)[0] + '.h5'
    if h5 is not None:
        if not os.path.exists(h5):
            raise FileNotFoundError(f"{h5} not found (the reference downloads it, DigiPathAI/helpers/utils.py:58-98)")
        from .keras_h5 import load_keras_h5
        return load_keras_h5(model, h5)
    return _load_npz(path)


def _load_npz(path):
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found. The reference downloads Keras .h5 weights at this point "
            "(DigiPathAI/helpers/utils.py:58-98); convert them once with tools/h5_to_npz.py, or pass "
            "weights=<dict> to getSegmentation.")
    z = np.load(path)
    out = {}
    for k in z.files:
        if k.endswith('::gamma'):
            b = k[:-7]
            out[b] = (z[b + '::gamma'], z[b + '::beta'], z[b + '::mean'], z[b + '::var'])
        elif '::' not in k:
            out[k] = z[k]
    return out


def save_npz(weights: dict, path: str):
    flat = {}
    for k, v in weights.items():
        if isinstance(v, tuple):
            for nm, a in zip(('gamma', 'beta', 'mean', 'var'), v):
                flat[f'{k}::{nm}'] = a
        else:
            flat[k] = v
    np.savez(path, **flat)


def getSegmentation(img_path,
                    patch_size=256,
                    stride_size=128,
                    batch_size=32,
                    quick=True,
                    tta_list=None,
                    crf=False,
                    save_path=None,
                    status=None,
                    *,
                    probs_path=None,
                    mask_path=None,
                    uncertainty_path=None,
                    mask_level=-1,
                    model='dense',
                    mode='colon',
                    weights=None,
                    device=0,
                    return_device=False,
                    pyramidal=True,
                    precision="fp16",
                    timings=None):
    """Whole-slide segmentation (reference: Segmentation.py:192-356, README.md:79-87).

    Returns the thresholded map -- float32 ``[W, H]`` of {0, 255}, NOT transposed -- exactly what the reference
    returns (Segmentation.py:336-337,356); writes probs / mask / uncertainty TIFFs when paths are given
    (``save_path`` is the README's name for ``mask_path``).  ``mask_level`` is accepted and ignored, as in the
    live reference (dataloader.py:240-241).  ``crf=True`` refines the label map with the fully connected CRF of
    ``post_process_crf`` (utils.py:568-603), tile-wise -- see ``_crf_refine``; the reference accepts the flag but
    its CRF call is commented out (Segmentation.py:327-331), so ``crf=False`` is the reference-identical path.
    The three files are written the way the reference leaves them after its ImageMagick pass: tiled (256x256)
    pyramidal JPEG-q90 TIFFs (tiffio.save_pyramidal; probabilities and uncertainty scaled to 8 bit);
    ``pyramidal=False`` writes lossless single-level TIFFs instead (float32 probabilities / uncertainty).
    Extensions (keyword only): ``weights`` = dict / ``.npz`` path per model name or a single dict for
    ``model``; ``device``; ``return_device`` returns the uint8 label plane as a CUDA tensor instead;
    ``precision='fp32'`` runs the networks in the reference's fp32 arithmetic (see ``load_trained_models``).
    """
    from . import engine
    import time
    torch = _torch()
    _t = [time.perf_counter()]

    def _lap(key):
        if timings is not None:
            now = time.perf_counter()
            timings[key] = timings.get(key, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now

    mode = mode.lower()
    print("==================================================")
    print(mode)
    if mode not in ['colon', 'liver', 'breast']:
        raise ValueError("Unknown mode found, allowed fields are: ['colon', 'liver', 'breast']")
    if mask_path is None:
        mask_path = save_path

    print("---------------------- {}, {} ---------------".format(model, quick))
    if not quick:
        names = ['dense', 'inception', 'deeplabv3']          # Segmentation.py:288-291
    else:
        if model not in ('dense', 'inception', 'deeplabv3'):
            raise ValueError("Unknown model provided, allowed models ['dense', 'inception', 'deeplabv3']")
        names = [model]

    def weight_source(nm):
        if isinstance(weights, dict) and nm in weights and not isinstance(weights[nm], np.ndarray):
            return weights[nm]
        if isinstance(weights, dict) and ('conv1/conv' in weights or 'conv2d_1' in weights or
                                          'entry_flow_conv1_1' in weights):
            family = 'dense' if 'conv1/conv' in weights else ('inception' if 'conv2d_1' in weights else 'deeplabv3')
            if family != nm:
                raise ValueError(f"weights= holds a '{family}' weight dict but model '{nm}' was requested; pass "
                                 "weights={'dense': ..., 'inception': ..., 'deeplabv3': ...}")
            return weights          # a single flat weight dict for `model`
        if isinstance(weights, str):
            return weights
        return _default_weight_path(mode, nm)

    missing = [nm for nm in names if isinstance(weight_source(nm), str) and not os.path.exists(weight_source(nm))]
    if status is not None:
        status['status'] = "Downloading Trained Models" if missing else "Found Trained Models, Skipping download"
    if status is not None:
        status['status'] = "Loading Trained weights"
    models = {}
    for nm in names:
        models[nm] = load_trained_models(nm, weight_source(nm), patch_size=patch_size, device=device,
                                         max_batch=batch_size, precision=precision)

    _lap('load_ms')
    threshold = 0.3
    if status is not None:
        status['status'] = "Running segmentation"
    slide, probs_map = get_prediction(img_path, mask_path=None, mask_level=mask_level, label_path=None,
                                      batch_size=batch_size, tta_list=tta_list, models=models,
                                      patch_size=patch_size, stride_size=stride_size, status=status,
                                      device=device, return_device=True, finalize=False)
    _lap('predict_ms')
    mean, var, count = probs_map['mean'], probs_map['var'], probs_map['count']
    label = torch.empty(mean.shape, dtype=torch.uint8, device=mean.device)
    with torch.cuda.device(mean.device):
        engine.finalize(mean, var, count, threshold, label)
        if crf:
            _crf_refine(engine, torch, slide, mean, label, int(patch_size), int(batch_size))
        torch.cuda.synchronize()
    for m in models.values():
        m.close()
    _lap('finalize_ms')

    from .tiffio import save_plane, save_pyramidal
    if pyramidal:
        _save = lambda path, plane, scale: save_pyramidal(path, plane if scale == 1 else plane * scale)
    else:
        _save = lambda path, plane, scale: save_plane(path, plane if scale in (1, 0) else plane * scale)
    if probs_path:
        _save(probs_path, mean.T, 255 if pyramidal else 0)               # Segmentation.py:333-334
    if status is not None:
        status['progress'] = 100
    if status is not None:
        status['status'] = "Saving Prediction Mask..."
    if mask_path:
        _save(mask_path, label.T, 1)                                     # Segmentation.py:345-346
    if status is not None:
        status['status'] = "Saving Prediction Uncertanity..."
    if uncertainty_path:
        _save(uncertainty_path, var.T, 255)                              # Segmentation.py:351-352
    if status is not None:
        status['progress'] = 0
    _lap('save_ms')
    if return_device:
        return label
    # the reference returns float32 {0, 255} [W, H] (Segmentation.py:356): 1 byte per pixel crosses PCIe, the widening runs
    # on all host cores (torch's CPU cast is threaded; numpy's astype is not: 1.3 s less on a 40 000^2 slide)
    out = label.cpu().to(torch.float32).numpy()
    _lap('return_ms')
    return out
