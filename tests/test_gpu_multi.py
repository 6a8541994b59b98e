"""Two real GPUs, one process each over NCCL: the x-stripe sharded ``get_prediction`` (dist.sharded_get_prediction)
against the single-GPU planes.  Skipped on a one-GPU box (run it with ``gpurun --gpus 2``); the same orchestration runs
on CPU over gloo with test doubles in tests/test_dist_pipeline.py.

Tolerance: halo columns are summed as (rank-0 partial) + (rank-1 partial), the single GPU adds tile by tile -- fp32
re-association, <= 4e-7 relative on the un-normalised sums (SURVEY.md 8(e) determinism caveat); everything outside
the halo must be bit-identical.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from digipathai_b200 import dist as dpd
        from digipathai_b200.Segmentation import load_trained_models
        from digipathai_b200.models.densenet import init_densenet_weights
        from digipathai_b200.slide import synthetic_slide
        slide = synthetic_slide(3072, 2048, seed=5, n_levels=2)
        model = load_trained_models('dense', init_densenet_weights(0), 256, device=rank, max_batch=8)
        grid, res, info = dpd.sharded_get_prediction(slide, {'dense': model}, 8, ['FLIP_LEFT_RIGHT'], 256, 128,
                                                     device=rank, gather=True)
        torch.cuda.synchronize()
        # second run: threshold + tile-wise CRF on the blocks each rank owns (config-5 shape)
        _, res2, info2 = dpd.sharded_get_prediction(slide, {'dense': model}, 8, None, 256, 128, device=rank,
                                                    threshold=0.3, crf=True)
        from digipathai_b200.dist import crf_blocks_of
        torch.cuda.synchronize()
        stripes = dpd.stripes_for(grid.coords, 8, world, 256)[1]
        mine = crf_blocks_of(stripes, rank, 3072, 256)
        lab = res2['label'].cpu().numpy()
        blocks = {bx: lab[bx - info2['stripe'][0]: bx - info2['stripe'][0] + 256] for bx in mine}
        q.put((rank, info, len(grid.coords),
               res['mean'].cpu().numpy() if rank == 0 else None, res['var'].cpu().numpy() if rank == 0 else None,
               blocks, info2['crf_tiles']))
        dist.barrier()
        model.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_slide_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        item = q.get(timeout=600)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    from digipathai_b200.Segmentation import get_prediction, load_trained_models
    from digipathai_b200.models.densenet import init_densenet_weights
    from digipathai_b200.slide import synthetic_slide
    slide = synthetic_slide(3072, 2048, seed=5, n_levels=2)
    model = load_trained_models('dense', init_densenet_weights(0), 256, max_batch=8)
    _, want = get_prediction(slide, batch_size=8, models={'dense': model}, tta_list=['FLIP_LEFT_RIGHT'],
                             patch_size=256, stride_size=128)
    model.close()
    info0, info1 = got[0][0], got[1][0]
    assert info0['batches'][1] == info1['batches'][0] and info1['batches'][1] * 8 == got[0][1]
    assert info0['halo_bytes_sent'] > 0 and info0['halo_bytes_sent'] == info1['halo_bytes_sent']
    mean, var = got[0][2], got[0][3]
    a, b = info1['stripe'][0], info0['stripe'][1]          # halo = [a, b)
    assert a < b
    outside = np.ones(mean.shape[0], bool)
    outside[a:b] = False
    assert np.array_equal(mean[outside], want['mean'][outside])
    assert np.array_equal(var[outside], want['var'][outside])
    assert np.abs(mean[a:b] - want['mean'][a:b]).max() <= 4e-7
    assert np.abs(var[a:b] - want['var'][a:b]).max() <= 1e-6
    # sharded threshold + CRF: every block is refined by exactly one rank, and (up to the <= 4e-7 halo re-association
    # feeding the CRF) its labels are those of the one-GPU run
    from digipathai_b200 import dist as dpd
    model = load_trained_models('dense', init_densenet_weights(0), 256, max_batch=8)
    _, one, info_one = dpd.sharded_get_prediction(slide, {'dense': model}, 8, None, 256, 128, device=0, threshold=0.3,
                                                  crf=True, shard=(0, 1))
    lab_one, lo_one = one['label'].cpu().numpy(), info_one['stripe'][0]
    model.close()
    blocks = {**got[0][4], **got[1][4]}
    assert len(blocks) == len(got[0][4]) + len(got[1][4]) and got[0][5] + got[1][5] == info_one['crf_tiles'] > 0
    same = total = 0
    for bx, lab in blocks.items():
        ref = lab_one[bx - lo_one: bx - lo_one + 256]
        same += int((lab == ref).sum()); total += lab.size
    assert same / total > 0.9999, same / total
