"""N>1 host logic on CPU: partitioning + halo exchange over a world_size-2 (and 3) gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from digipathai_b200 import dist as dpd


def test_partition_and_stripes():
    assert dpd.partition_batches(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert dpd.partition_batches(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    coords = np.array([[x, y] for x in (0, 0, 128, 256, 384) for y in (0, 128)], np.int32)
    parts, stripes = dpd.stripes_for(coords, 2, 2, 256)
    assert parts == [(0, 3), (3, 5)]
    assert stripes == [(0, 384), (256, 640)]
    assert dpd.owned_ranges(stripes) == [(0, 384), (384, 640)]
    assert dpd.owned_ranges([(0, 300), (100, 200), (250, 400)]) == [(0, 300), (300, 300), (300, 400)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, stripes, H, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed + rank)
        lo, hi = stripes[rank]
        mean = torch.from_numpy(rng.random((hi - lo, H), dtype=np.float32))
        var = torch.from_numpy(rng.random((hi - lo, H), dtype=np.float32))
        cnt = torch.from_numpy(rng.integers(0, 256, (hi - lo, H)).astype(np.uint8))
        planes = [mean.clone(), var.clone(), cnt.clone()]
        sent = dpd.halo_exchange(planes, stripes, rank)
        W = max(h for _, h in stripes)
        full = dpd.gather_planes(planes, stripes, rank, world, W)
        q.put((rank, [mean.numpy(), var.numpy(), cnt.numpy()], [p.numpy() for p in planes],
               None if full is None else [f.numpy() for f in full], sent))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("stripes", [[(0, 40), (24, 64)], [(0, 40), (30, 50), (36, 80)], [(0, 32), (32, 64)]])
def test_halo_exchange_gloo(stripes):
    world, H = len(stripes), 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, stripes, H, 11, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, orig, summed, full, sent = q.get(timeout=120)
        res[r] = (orig, summed, full, sent)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W = max(h for _, h in stripes)
    # expected: per plane, sum contributions in ascending rank order over the global x axis
    for k, dt in enumerate((np.float32, np.float32, np.uint8)):
        acc = np.zeros((W, H), dt)
        touched = np.zeros(W, bool)
        for r in range(world):
            lo, hi = stripes[r]
            part = res[r][0][k]
            fresh = ~touched[lo:hi]
            blk = acc[lo:hi]
            blk[fresh] = part[fresh]
            blk[~fresh] = (blk[~fresh] + part[~fresh]).astype(dt)
            touched[lo:hi] = True
        for r in range(world):
            lo, hi = stripes[r]
            assert np.array_equal(res[r][1][k], acc[lo:hi]), f"plane {k} rank {r}"
        if k < 2:
            assert np.array_equal(res[0][2][k], acc) if k < 2 else True
    if stripes == [(0, 32), (32, 64)]:
        assert res[0][3] == 0 and res[1][3] == 0        # disjoint stripes exchange nothing
    else:
        assert res[0][3] > 0


def test_every_crf_block_inside_the_tiled_area_has_exactly_one_owner():
    """dist.crf_blocks_of: with x-stripes that overlap by patch - stride, each 256-wide CRF block that any stripe contains
    is refined by exactly one rank, and a single stripe owns the same set of blocks the ranks own together."""
    from digipathai_b200 import dist as dpd
    rng = np.random.default_rng(0)
    P, stride, W = 256, 128, 5000
    xs = np.sort(rng.choice(np.arange(0, W - P + 1, stride), 30, replace=False))
    coords = np.stack([np.repeat(xs, 4), np.tile(np.arange(4) * stride, len(xs))], 1).astype(np.int32)
    whole = dpd.crf_blocks_of([dpd.stripe_of(coords, 0, len(coords), P)], 0, W, P)
    for world in (2, 3, 5):
        parts, stripes = dpd.stripes_for(coords, 4, world, P)
        owned = [dpd.crf_blocks_of(stripes, r, W, P) for r in range(world)]
        flat = [b for o in owned for b in o]
        assert len(flat) == len(set(flat))                       # no block has two owners
        for r, o in enumerate(owned):
            lo, hi = stripes[r]
            assert all(lo <= b and b + P <= hi for b in o)       # an owner holds the whole block
        assert set(flat) <= set(whole)
        for b in set(whole) - set(flat):                         # a block nobody owns straddles a gap between stripes:
            lo_hi = [(lo, hi) for lo, hi in stripes if hi > lo]  # it is not inside any rank's tiled area
            assert not any(lo <= b and b + P <= hi for lo, hi in lo_hi)
