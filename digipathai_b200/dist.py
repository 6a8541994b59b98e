"""Multi-GPU sharding of one slide: contiguous tile ranges per rank + one halo exchange of boundary stripes.

The reference is single-GPU (``CUDA_VISIBLE_DEVICES='0'``, DigiPathAI/Segmentation.py:62); this module adds
the one strategy the path admits: tiles are independent, coupling exists only through the
``+=`` into overlapping plane windows (Segmentation.py:164-173).  Tiles are ordered x-major
(``np.where``, dataloader.py:311), so a contiguous range of the post-``drop_last`` tile list is an x-stripe.

  partition_batches   equal numbers of batches per rank (balances tissue tiles, not slide area)
  stripe_of           the [x_lo, x_hi) plane stripe a tile range touches
  halo_exchange       every pair of ranks whose stripes intersect swaps its partial sums over the
                      intersection (P2P send/recv, NCCL on GPUs / gloo in the CPU tests); each rank then adds
                      the contributions in ascending rank order, so all ranks hold bit-identical totals.
                      fp32 addition is not associative: (rank0 partial) + (rank1 partial) can differ from the
                      single-GPU sequential order in the last ulp (SURVEY.md 8(e) determinism caveat).
  gather_planes       disjoint "owned" sub-stripes to rank 0 -> full [W, H] planes
"""
from __future__ import annotations

import numpy as np


def partition_batches(n_batches: int, world_size: int):
    """[(b_lo, b_hi)] per rank; earlier ranks take the remainder."""
    base, rem = divmod(n_batches, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def stripe_of(coords: np.ndarray, t_lo: int, t_hi: int, patch: int):
    """[x_lo, x_hi) covered by tiles t_lo..t_hi-1 (empty range -> (0, 0))."""
    if t_hi <= t_lo:
        return (0, 0)
    xs = coords[t_lo:t_hi, 0]
    return (int(xs.min()), int(xs.max()) + patch)


def stripes_for(coords: np.ndarray, batch_size: int, world_size: int, patch: int):
    n_batches = len(coords) // batch_size
    parts = partition_batches(n_batches, world_size)
    return parts, [stripe_of(coords, lo * batch_size, hi * batch_size, patch) for lo, hi in parts]


def halo_exchange(planes, stripes, rank: int, group=None):
    """In-place: add every other rank's partial sums over the intersection of its stripe with mine.

    ``planes`` is a list of torch tensors shaped [x_hi - x_lo, H] (float32 mean, float32 var, uint8 count);
    ``stripes`` the [x_lo, x_hi) of every rank.  Returns the number of bytes this rank sent.
    """
    import torch
    import torch.distributed as dist
    my_lo, my_hi = stripes[rank]
    peers = []
    for s, (lo, hi) in enumerate(stripes):
        if s == rank:
            continue
        a, b = max(lo, my_lo), min(hi, my_hi)
        if b > a:
            peers.append((s, a, b))
    if not peers:
        return 0
    own = {s: [p[a - my_lo:b - my_lo].clone() for p in planes] for (s, a, b) in peers}   # my partials, pre-sum
    recv = {s: [torch.empty_like(t) for t in own[s]] for s in own}
    ops, sent = [], 0
    for (s, a, b) in peers:
        for t_send, t_recv in zip(own[s], recv[s]):
            ops.append(dist.P2POp(dist.isend, t_send, s, group))
            ops.append(dist.P2POp(dist.irecv, t_recv, s, group))
            sent += t_send.numel() * t_send.element_size()
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    # ascending-rank summation over every intersected region; regions of different peers may overlap each
    # other (stripes narrower than a patch), so the x axis is cut at every region boundary and each segment is
    # rebuilt from the original partials (a segment is read before it is written and never read again).
    order = sorted([s for (s, _, _) in peers] + [rank])
    bounds = {s: (a, b) for (s, a, b) in peers}
    xs = sorted({my_lo, my_hi, *[v for ab in bounds.values() for v in ab]})
    for x0, x1 in zip(xs[:-1], xs[1:]):
        contributors = [s for s in order if s == rank or (bounds[s][0] <= x0 and x1 <= bounds[s][1])]
        if contributors == [rank]:
            continue
        for k, p in enumerate(planes):
            acc = None
            for s in contributors:
                if s == rank:
                    part = p[x0 - my_lo:x1 - my_lo]
                else:
                    a, _ = bounds[s]
                    part = recv[s][k][x0 - a:x1 - a]
                acc = part.clone() if acc is None else acc + part   # uint8 adds wrap like the reference's
            p[x0 - my_lo:x1 - my_lo] = acc
    return sent


def owned_ranges(stripes):
    """Disjoint [lo, hi) per rank: a rank owns its stripe minus what a lower rank already owns."""
    out, covered = [], 0
    for (lo, hi) in stripes:
        a = max(lo, covered)
        out.append((a, max(a, hi)))
        covered = max(covered, hi)
    return out


def crf_blocks_of(stripes, rank: int, W: int, patch: int):
    """First columns of the CRF blocks rank ``rank`` refines: a block belongs to the LOWEST rank whose stripe contains
    it entirely (stripes overlap by patch - stride columns, and after the halo exchange both ranks hold identical
    values there); blocks no stripe contains lie outside every tile (probability 0) and are nobody's."""
    from .Segmentation import crf_block_starts
    mine = []
    for bx in crf_block_starts(W, patch):
        owner = next((r for r, (lo, hi) in enumerate(stripes) if hi > lo and lo <= bx and bx + patch <= hi), None)
        if owner == rank:
            mine.append(bx)
    return mine


def gather_planes(planes, stripes, rank: int, world_size: int, W: int, group=None):
    """Rank 0 returns full [W, H] planes (zeros where no tile touched), other ranks None."""
    import torch
    import torch.distributed as dist
    own = owned_ranges(stripes)
    my_lo = stripes[rank][0]
    if rank == 0:
        full = [torch.zeros((W,) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device) for p in planes]
        a, b = own[0]
        for f, p in zip(full, planes):
            f[a:b] = p[a - my_lo:b - my_lo]
        for s in range(1, world_size):
            a, b = own[s]
            if b <= a:
                continue
            for f in full:
                buf = torch.empty((b - a,) + tuple(f.shape[1:]), dtype=f.dtype, device=f.device)
                dist.recv(buf, s, group)
                f[a:b] = buf
        return full
    a, b = own[rank]
    if b > a:
        for p in planes:
            dist.send(p[a - my_lo:b - my_lo].contiguous(), 0, group)
    return None


def local_part(slide, models, rank: int, world: int, batch_size=32, tta_list=None, patch_size=256, stride_size=128,
               status=None, device=None, tissue_mask=None, grid=None, timings=None):
    """One rank's share before the halo exchange: ``(grid, [mean, var, count] partial-sum planes over this rank's
    stripe, stripes of all ranks, (b_lo, b_hi))``.  The grid is built once here (or handed in) and passed to
    ``get_prediction`` as is (``tissue_mask`` is a RAW mask, see TileGrid).  A rank whose batch range is empty
    (more ranks than batches) holds zero-width planes and never touches the raster."""
    import time
    import torch
    from .Segmentation import get_prediction
    from .tissue import TileGrid
    t0 = time.perf_counter()
    if grid is None:
        grid = TileGrid(slide, patch_size, stride_size, batch_size, True, tissue_mask, device=device)
    if timings is not None:
        timings['grid_ms'] = (time.perf_counter() - t0) * 1e3
    parts, stripes = stripes_for(grid.coords, batch_size, world, patch_size)
    lo, hi = parts[rank]
    if hi <= lo:
        H = slide.level_dimensions[0][1]
        dev = torch.device("cuda", device if device is not None else torch.cuda.current_device())
        planes = [torch.zeros((0, H), dtype=dt, device=dev) for dt in (torch.float32, torch.float32, torch.uint8)]
        return grid, planes, stripes, (lo, hi)
    _, out = get_prediction(slide, batch_size=batch_size, models=models, tta_list=tta_list,
                            patch_size=patch_size, stride_size=stride_size, status=status, device=device,
                            tile_range=(lo * batch_size, hi * batch_size), return_device=True, finalize=False,
                            grid=grid, timings=timings)
    if tuple(out['x_range']) != tuple(stripes[rank]):
        # mismatched geometry would turn into mismatched P2P sizes in halo_exchange, i.e. a hang: fail here instead
        raise RuntimeError(f"rank {rank}: planes cover x {out['x_range']}, stripe plan says {stripes[rank]}")
    return grid, [out['mean'], out['var'], out['count']], stripes, (lo, hi)


def sharded_get_prediction(wsi_path, models, batch_size=32, tta_list=None, patch_size=256, stride_size=128,
                           status=None, device=None, gather=False, tissue_mask=None, shard=None, threshold=None,
                           crf=False):
    """``get_prediction`` across the ranks of the default process group (one process per GPU; a process without an
    initialised group is a world of one; ``shard=(rank, world)`` overrides both -- ``shard=(0, 1)`` runs the whole
    slide on the calling rank alone, which is how bench.py measures the one-GPU time beside an N-GPU run).

    Every rank computes the same tile grid (deterministic; on the GPU when the slide's raster is resident there),
    takes its contiguous range of batches, runs the device loop on its stripe, swaps halos, then normalises its
    stripe.  Returns ``(grid, planes_dict, info)``; with ``gather=True`` rank 0's dict holds the full planes.
    ``info['timings_ms']`` splits this rank's wall time into grid / raster upload / tile loop / halo / normalise
    (each phase ends with a device synchronisation).  ``threshold`` (getSegmentation's 0.3) adds the rank's uint8
    label stripe ``res['label']``; ``crf=True`` then refines it tile-wise (Segmentation._crf_refine) on the P x P blocks
    this rank owns (``crf_blocks_of``: every block of the slide has exactly one owner).
    """
    import time
    import torch
    import torch.distributed as dist
    from . import engine
    from .slide import open_slide
    if shard is not None:
        rank, world = int(shard[0]), int(shard[1])
    elif dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    if device is None:
        device = torch.cuda.current_device()
    slide = open_slide(wsi_path)
    tm = {}
    grid, planes, stripes, (lo, hi) = local_part(slide, models, rank, world, batch_size, tta_list, patch_size,
                                                 stride_size, status, device, tissue_mask, timings=tm)
    t0 = time.perf_counter()
    sent = halo_exchange(planes, stripes, rank) if (hi > lo and world > 1) else 0
    if hi > lo:
        torch.cuda.synchronize(planes[0].device)
    tm['halo_ms'] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    label = None
    if hi > lo:
        with torch.cuda.device(planes[0].device):
            if threshold is not None:
                label = torch.empty(planes[0].shape, dtype=torch.uint8, device=planes[0].device)
            engine.finalize(planes[0], planes[1], planes[2], float(threshold) if threshold is not None else 0.0, label)
            torch.cuda.synchronize()
    tm['normalise_ms'] = (time.perf_counter() - t0) * 1e3
    n_crf = 0
    if crf and label is not None:
        from .Segmentation import _crf_refine
        t0 = time.perf_counter()
        with torch.cuda.device(planes[0].device):
            n_crf = _crf_refine(engine, torch, slide, planes[0], label, int(patch_size), int(batch_size),
                                x_lo=stripes[rank][0],
                                block_xs=crf_blocks_of(stripes, rank, slide.level_dimensions[0][0], int(patch_size)))
            torch.cuda.synchronize()
        tm['crf_ms'] = (time.perf_counter() - t0) * 1e3
    info = {'rank': rank, 'world': world, 'batches': (lo, hi), 'stripe': stripes[rank], 'halo_bytes_sent': sent,
            'timings_ms': tm, 'crf_tiles': n_crf}
    res = {'mean': planes[0], 'var': planes[1], 'x_range': stripes[rank]}
    if label is not None:
        res['label'] = label
    if gather:
        W = slide.level_dimensions[0][0]
        full = gather_planes(planes[:2], stripes, rank, world, W) if world > 1 else planes[:2]
        if rank == 0:
            if world == 1 and tuple(stripes[0]) != (0, W):      # a world of one still owns only its stripe
                a, b = stripes[0]
                full = [torch.zeros((W,) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device) for p in planes[:2]]
                for f, p in zip(full, planes[:2]):
                    f[a:b] = p
            res = {'mean': full[0], 'var': full[1], 'x_range': (0, W)}
    return grid, res, info
