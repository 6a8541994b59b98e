"""Size-independent properties of the CUDA path at (or near) BASELINE.json's full sizes, plus edge cases:
determinism, batch-position independence, sharded == unsharded, large-plane stitch against numpy windows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model32():
    import torch
    from digipathai_b200.engine import TileModel
    from digipathai_b200.models.densenet import densenet121_unet_program, init_densenet_weights
    m = TileModel(densenet121_unet_program(init_densenet_weights(0), 256), device=0, max_batch=32)
    yield m, torch
    m.close()


def test_forward_is_deterministic_and_batch_position_independent(model32):
    m, torch = model32
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    tiles = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=g)
    a = m.forward_tile_batch(tiles).clone()
    b = m.forward_tile_batch(tiles).clone()
    assert torch.equal(a, b)                                   # graph replay is bit-reproducible
    perm = torch.randperm(32, device="cuda", generator=g)
    c = m.forward_tile_batch(tiles[perm].contiguous())
    assert torch.equal(c, a[perm])                             # a tile's result does not depend on its slot
    d = m.forward_tile_batch(tiles[:5].contiguous())           # odd batch: other plan, other tile shapes
    assert float((d - a[:5]).abs().max()) < 2e-2               # only fp32 accumulation order may differ
    assert torch.isfinite(a).all() and float(a.min()) >= 0.0 and float(a.max()) <= 1.0


def test_all_eight_d4_codes_roundtrip(model32):
    """tta_in = g, tta_out = g must equal running the transformed tile untransformed and undoing g on the host."""
    from digipathai_b200 import tta
    m, torch = model32
    rng = np.random.default_rng(11)
    tiles = rng.integers(0, 256, (2, 256, 256, 3)).astype(np.uint8)
    for g in range(8):
        want_in = np.stack([tta.apply(g, t) for t in tiles])
        plain = m.forward_tile_batch(torch.from_numpy(want_in).cuda()).cpu().numpy()
        want = np.stack([tta.apply(tta.inverse(g), p) for p in plain])
        got = m.forward_tile_batch(torch.from_numpy(tiles).cuda(), g, g).cpu().numpy()
        assert np.array_equal(got, want), g                    # gather / scatter index maps are exact


def test_graph_and_direct_paths_agree(model32):
    m, torch = model32
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    tiles = torch.randint(0, 256, (8, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=g)
    a = m.forward_tile_batch(tiles).clone()
    for key in ("use_graph", "use_pdl", "use_overlap", "b_resident", "epi_direct", "dense_block"):
        m.set_option(key, 0)
        b = m.forward_tile_batch(tiles).clone()
        m.set_option(key, 1)
        assert torch.equal(a, b), key                          # scheduling options never change arithmetic
    m.set_option("b_pair", 1)
    b = m.forward_tile_batch(tiles).clone()
    m.set_option("b_pair", 0)
    assert torch.equal(a, b)


def test_persistent_dense_block_kernel_is_bit_identical_to_the_per_layer_kernels(model32):
    """conv4 / conv5 (16x16 / 8x8 maps) run as one persistent kernel per block (csrc/dense_block.cuh: regions pinned to
    CTAs, neighbour flags instead of kernel boundaries).  Same arithmetic in the same order: every probability and the
    block outputs themselves must equal the one-launch-per-layer path bit for bit -- at batch 32 (128 / 32 regions),
    at a ragged batch, under TTA, with and without the CUDA graph, and on repeated calls (flags are re-armed)."""
    m, torch = model32
    g = torch.Generator(device="cuda"); g.manual_seed(21)
    prog = m.program
    for B in (32, 5, 1):
        tiles = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=g)
        m.set_option("dense_block", 0)
        want = m.forward_tile_batch(tiles, 5, 4).clone()
        want_d4 = m.read_buffer(prog.buf("D4"), B)
        want_d5 = m.read_buffer(prog.buf("D5"), B)
        m.set_option("dense_block", 2)           # 2 = every block the kernel applies to (1 = planner's choice)
        for graph in (1, 0, 1):
            m.set_option("use_graph", graph)
            for _ in range(2):
                got = m.forward_tile_batch(tiles, 5, 4).clone()
                assert torch.equal(got, want), (B, graph)
            assert np.array_equal(m.read_buffer(prog.buf("D4"), B), want_d4)
            assert np.array_equal(m.read_buffer(prog.buf("D5"), B), want_d5)
        m.set_option("use_graph", 1)
        m.set_option("dense_block", 1)
        assert torch.equal(m.forward_tile_batch(tiles, 5, 4), want), B


def test_sharded_tile_ranges_equal_unsharded():
    """Two 'ranks' on one GPU (tile_range halves + host-side halo sum) reproduce the single-range planes."""
    import torch
    from digipathai_b200 import dist as dpd
    from digipathai_b200.Segmentation import get_prediction, load_trained_models
    from digipathai_b200.models.densenet import init_densenet_weights
    from digipathai_b200.slide import synthetic_slide
    from digipathai_b200.tissue import TileGrid
    slide = synthetic_slide(2048, 1536, seed=2)
    model = load_trained_models('dense', init_densenet_weights(0), 256, max_batch=8)
    kw = dict(batch_size=8, patch_size=256, stride_size=128, models={'dense': model})
    _, full = get_prediction(slide, finalize=False, **kw)
    grid = TileGrid(slide, 256, 128, 8)
    parts, stripes = dpd.stripes_for(grid.coords, 8, 2, 256)
    acc = {k: np.zeros_like(full[k]) for k in ('mean', 'var', 'count')}
    for (lo, hi), (x0, x1) in zip(parts, stripes):
        _, part = get_prediction(slide, tile_range=(lo * 8, hi * 8), finalize=False, **kw)
        assert part['x_range'] == (x0, x1)
        for k in acc:
            acc[k][x0:x1] += part[k]
    assert np.array_equal(acc['count'], full['count'])
    assert np.abs(acc['mean'] - full['mean']).max() <= 4e-7 * max(1.0, float(full['mean'].max()))  # fp32 re-association
    assert np.abs(acc['var'] - full['var']).max() <= 1e-6
    model.close()


def test_local_parts_of_every_rank_sum_to_the_unsharded_planes():
    """dist.local_part (what each rank of sharded_get_prediction computes before the halo exchange) for world
    sizes 2 and 3, run one rank after the other on this GPU: the stripes it reports tile the unsharded result,
    with the default tissue heuristic and with a caller-supplied raw mask."""
    from digipathai_b200 import dist as dpd
    from digipathai_b200.Segmentation import get_prediction, load_trained_models
    from digipathai_b200.models.densenet import init_densenet_weights
    from digipathai_b200.slide import synthetic_slide
    from digipathai_b200.tissue import TileGrid
    slide = synthetic_slide(2048, 1536, seed=3, n_levels=2)
    model = load_trained_models('dense', init_densenet_weights(0), 256, max_batch=8)
    kw = dict(batch_size=8, patch_size=256, stride_size=128)
    raw = TileGrid(slide, 256, 128, 8).raw_mask
    for mask in (None, raw):
        _, full = get_prediction(slide, models={'dense': model}, finalize=False, tissue_mask=mask, **kw)
        assert int(full['count'].max()) > 0
        for world in (2, 3):
            acc = {k: np.zeros_like(full[k]) for k in ('mean', 'var', 'count')}
            n_tiles = 0
            for rank in range(world):
                grid, planes, stripes, (lo, hi) = dpd.local_part(slide, {'dense': model}, rank, world, device=0,
                                                                 tissue_mask=mask, **kw)
                x0, x1 = stripes[rank]
                n_tiles += (hi - lo) * 8
                if hi > lo:
                    for k, p in zip(('mean', 'var', 'count'), planes):
                        assert p.shape[0] == x1 - x0
                        acc[k][x0:x1] += p.cpu().numpy()
            assert n_tiles == len(grid.coords)
            assert np.array_equal(acc['count'], full['count'])
            assert np.abs(acc['mean'] - full['mean']).max() <= 4e-7 * max(1.0, float(full['mean'].max()))
            assert np.abs(acc['var'] - full['var']).max() <= 1e-6
    model.close()


def test_sharded_get_prediction_in_a_world_of_one_is_get_prediction(tmp_path):
    import torch
    import torch.distributed as dist
    from digipathai_b200 import dist as dpd
    from digipathai_b200.Segmentation import get_prediction, load_trained_models
    from digipathai_b200.models.densenet import init_densenet_weights
    from digipathai_b200.slide import synthetic_slide
    slide = synthetic_slide(1536, 1280, seed=4, n_levels=2)
    model = load_trained_models('dense', init_densenet_weights(0), 256, max_batch=8)
    kw = dict(batch_size=8, patch_size=256, stride_size=128, tta_list=['FLIP_LEFT_RIGHT'])
    _, want = get_prediction(slide, models={'dense': model}, **kw)
    assert not dist.is_initialized()
    dist.init_process_group('gloo', init_method=f'file://{tmp_path}/pg', rank=0, world_size=1)
    try:
        grid, got, info = dpd.sharded_get_prediction(slide, {'dense': model}, device=0, gather=True, **kw)
    finally:
        dist.destroy_process_group()
    assert info['halo_bytes_sent'] == 0 and info['batches'] == (0, len(grid))
    x0, x1 = info['stripe']
    assert np.array_equal(got['mean'].cpu().numpy()[x0:x1], want['mean'][x0:x1])
    assert np.array_equal(got['var'].cpu().numpy()[x0:x1], want['var'][x0:x1])
    assert not want['mean'][:x0].any() and not want['mean'][x1:].any()
    model.close()


def test_stitch_on_a_large_plane_matches_numpy_windows():
    import torch
    from digipathai_b200 import engine
    rng = np.random.default_rng(4)
    W, H, P, B, N = 6000, 5000, 256, 32, 4
    dev = torch.device("cuda", 0)
    mean = torch.zeros((W, H), dtype=torch.float32, device=dev)
    var = torch.zeros_like(mean)
    cnt = torch.zeros((W, H), dtype=torch.uint8, device=dev)
    ref_m = np.zeros((W, H), np.float32); ref_v = np.zeros((W, H), np.float32); ref_c = np.zeros((W, H), np.uint8)
    for _ in range(6):
        coords = np.stack([rng.integers(0, W - P, B) // 64 * 64, rng.integers(0, H - P, B) // 64 * 64], 1).astype(np.int32)
        probs = rng.random((N, B, P, P), dtype=np.float32)
        engine.stitch(torch.from_numpy(probs).to(dev), torch.from_numpy(coords).to(dev), mean, var, cnt)
        m_, v_ = np.mean(probs, axis=0), np.var(probs, axis=0)
        for i, (x, y) in enumerate(coords):
            ref_m[x:x + P, y:y + P] += m_[i]; ref_v[x:x + P, y:y + P] += v_[i]
            ref_c[x:x + P, y:y + P] += np.ones((P, P), np.uint8)
    torch.cuda.synchronize()
    assert np.array_equal(mean.cpu().numpy(), ref_m) and np.array_equal(var.cpu().numpy(), ref_v)
    assert np.array_equal(cnt.cpu().numpy(), ref_c)
    # idempotence of the normalise step on an already normalised plane with count == 1
    engine.finalize(mean, var, cnt, 0.3, None)
    ones = torch.ones_like(cnt)
    before = mean.clone()
    engine.finalize(mean, var, ones, 0.3, None)
    assert torch.equal(before, mean)


def test_lanes_do_not_change_a_bit(model32):
    """Several forwards in flight (engine.ForwardLanes: model clones sharing the weights, one stream each) give the
    planes of the one-stream loop bit for bit -- with TTA passes, with one pass per batch (more lanes than passes) and
    for the two-model ensemble; a clone's forward equals its parent's; the host pipeline with lanes equals predict."""
    m, torch = model32
    from digipathai_b200 import engine
    from digipathai_b200.Segmentation import get_prediction
    from digipathai_b200.slide import synthetic_slide
    g = torch.Generator(device="cuda"); g.manual_seed(21)
    tiles = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=g)
    want = m.forward_tile_batch(tiles).clone()
    c = m.clone()
    assert c.device_bytes < m.device_bytes                                  # the clone holds no second copy of the weights
    assert torch.equal(c.forward_tile_batch(tiles), want)
    c.close()
    assert torch.equal(m.forward_tile_batch(tiles), want)                   # the parent survives its clone

    slide = synthetic_slide(2048, 1536, seed=4, n_levels=2)
    for tta_list in (['FLIP_LEFT_RIGHT', 'ROTATE_90'], None):
        kw = dict(batch_size=8, patch_size=256, stride_size=128, models={'dense': m}, tta_list=tta_list)
        _, one = get_prediction(slide, lanes=1, **kw)
        for lanes in (2, 3, 4):
            _, many = get_prediction(slide, lanes=lanes, **kw)
            assert np.array_equal(one['mean'], many['mean']) and np.array_equal(one['var'], many['var']), (tta_list, lanes)

    # host pipeline: 7 different batches through 3 lanes
    pipe = engine.HostBatchPipeline(m, 32, lanes=3)
    ins = [torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8).pin_memory() for _ in range(7)]
    outs = [torch.empty((32, 256, 256), dtype=torch.float32).pin_memory() for _ in range(7)]
    for a, b in zip(ins, outs):
        pipe.submit(a, b)
    pipe.drain()
    for a, b in zip(ins, outs):
        assert torch.equal(m.forward_tile_batch(a.cuda()).cpu(), b)
    pipe.close()
