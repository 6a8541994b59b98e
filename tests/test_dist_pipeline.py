"""The multi-GPU orchestration (dist.sharded_get_prediction -> get_prediction(tile_range, grid) -> halo exchange ->
normalise -> gather) end to end on CPU: world_size 2 and 3 over gloo, with TEST DOUBLES for the three device entry
points the loop calls (TileModel.forward_tiles, engine.stitch, engine.finalize) -- numpy restatements of
Segmentation.py:162-177 living in this file only.  The doubles emit dyadic values, so every fp32 sum is exact and
the sharded planes must equal the single-process planes bit for bit.

This is host-logic coverage (which tiles a rank runs, which stripe it owns, what it swaps); the CUDA kernels
themselves are covered by the ``-m gpu`` tests.
"""
import contextlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

P, STRIDE, BATCH = 64, 32, 4


class _FakeModel:
    """forward_tiles double: 'probability' = red channel / 256, scaled by a per-pass dyadic factor."""
    patch, max_batch = P, 32

    def forward_tiles(self, raster, coords, t_in=0, t_out=0, out=None):
        for b, (x, y) in enumerate(coords.tolist()):
            out[b] = raster[x:x + P, y:y + P, 0].to(torch.float32) / 256.0 / (1 + (t_in > 0) + 2 * (t_out > 4))
        return out

    def close(self):
        pass


def _stitch(probs, coords, mean, var, count, x_lo=0):
    m = probs.mean(0)                                        # Segmentation.py:162-163
    v = ((probs - m[None]) ** 2).mean(0)
    for b, (x, y) in enumerate(coords.tolist()):
        mean[x - x_lo:x - x_lo + P, y:y + P] += m[b]          # :164-173
        var[x - x_lo:x - x_lo + P, y:y + P] += v[b]
        count[x - x_lo:x - x_lo + P, y:y + P] += 1


def _finalize(mean, var, count, threshold, label=None):
    c = count.clone()
    c[c == 0] = 1                                            # :175-177
    mean /= c.to(torch.float32)
    var /= (c * c).to(torch.float32)
    if label is not None:                                    # :336-337
        label[...] = (mean >= threshold).to(torch.uint8) * 255


class _TorchShim:
    """What Segmentation._torch() returns in this test: torch, with 'cuda' devices mapped to the CPU."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*_):
        return torch.device("cpu")


def _install_doubles():
    from digipathai_b200 import Segmentation, engine
    Segmentation._torch = lambda: _TorchShim()
    engine.stitch, engine.finalize = _stitch, _finalize
    torch.cuda.device = lambda *_: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *_: None


def _slide(seed):
    from digipathai_b200.slide import ArraySlide
    rng = np.random.default_rng(seed)
    W, H = 640, 448
    img = np.full((H, W, 3), 240, np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    blob = ((yy - 230) / 170.0) ** 2 + ((xx - 300) / 260.0) ** 2 < 1
    img[blob] = (170, 90, 160)
    img = np.clip(img.astype(np.int16) + rng.integers(-8, 9, img.shape), 0, 255).astype(np.uint8)
    return ArraySlide(img, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tta_list, use_raw_mask, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_doubles()
        from digipathai_b200 import dist as dpd
        from digipathai_b200.tissue import TileGrid
        slide = _slide(0)
        mask = TileGrid(slide, P, STRIDE, BATCH).raw_mask if use_raw_mask else None
        grid, res, info = dpd.sharded_get_prediction(slide, {"m": _FakeModel()}, BATCH, tta_list, P, STRIDE,
                                                     device=0, gather=True, tissue_mask=mask)
        q.put((rank, info, len(grid.coords), res["mean"].numpy(), res["var"].numpy(), res["x_range"]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,tta_list,use_raw_mask", [(2, None, False), (2, ["FLIP_LEFT_RIGHT"], True),
                                                         (3, ["FLIP_LEFT_RIGHT", "ROTATE_90"], False)])
def test_sharded_get_prediction_equals_single_process(world, tta_list, use_raw_mask):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tta_list, use_raw_mask, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        item = q.get(timeout=180)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # single-process result through the same doubles
    from digipathai_b200 import Segmentation, engine
    saved = (Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize)
    try:
        _install_doubles()
        _, want = Segmentation.get_prediction(_slide(0), batch_size=BATCH, models={"m": _FakeModel()},
                                              tta_list=tta_list, patch_size=P, stride_size=STRIDE)
    finally:
        Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize = saved

    infos = [got[r][0] for r in range(world)]
    n_tiles = got[0][1]
    assert sum(b - a for a, b in (i["batches"] for i in infos)) * BATCH == n_tiles > 4 * BATCH * world
    assert all(i["halo_bytes_sent"] > 0 for i in infos)          # 50 % overlap: neighbouring stripes intersect
    mean, var, x_range = got[0][2], got[0][3], got[0][4]
    assert x_range == (0, 640) and mean.shape == want["mean"].shape
    assert want["mean"].max() > 0.3 and (not tta_list or want["var"].max() > 0)
    assert np.array_equal(mean, want["mean"])
    assert np.array_equal(var, want["var"])
    # every rank's own (un-gathered) stripe agrees too
    for r in range(1, world):
        x0, x1 = got[r][4]
        assert (x0, x1) == infos[r]["stripe"]
        assert np.array_equal(got[r][2], want["mean"][x0:x1])
        assert np.array_equal(got[r][3], want["var"][x0:x1])


def test_local_part_refuses_a_grid_for_other_geometry():
    from digipathai_b200 import Segmentation, engine
    from digipathai_b200.tissue import TileGrid
    saved = (Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize)
    try:
        _install_doubles()
        slide = _slide(1)
        grid = TileGrid(slide, P, STRIDE, BATCH)
        with pytest.raises(ValueError):
            Segmentation.get_prediction(slide, batch_size=BATCH * 2, models={"m": _FakeModel()}, patch_size=P,
                                        stride_size=STRIDE, grid=grid)
        with pytest.raises(ValueError):
            Segmentation.get_prediction(slide, batch_size=BATCH, models={"m": _FakeModel()}, patch_size=P,
                                        stride_size=STRIDE, grid=grid, tile_range=(1, 5))
    finally:
        Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize = saved


def test_mask_path_tiff_replaces_the_tissue_heuristic(tmp_path):
    """get_prediction(mask_path=<.tiff>) (dataloader.py:256-263) through the same doubles: equals a run with the
    mask handed over directly, differs from the heuristic run, and a non-.tiff name fails like the reference."""
    from PIL import Image
    from digipathai_b200 import Segmentation, engine
    saved = (Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize)
    try:
        _install_doubles()
        slide = _slide(0)                                           # 640 x 448, one level
        m = np.zeros((448, 640, 3), np.uint8)                       # image layout [y, x, c]
        m[100:300, 200:520] = (0, 0, 3)                             # luma of (0, 0, 3) rounds to 0 ...
        m[150:250, 250:400] = (0, 0, 5)                             # ... of (0, 0, 5) to 1: only this box counts
        path = str(tmp_path / "mask.tiff")
        Image.fromarray(m).save(path)
        kw = dict(batch_size=BATCH, models={"m": _FakeModel()}, patch_size=P, stride_size=STRIDE)
        _, got = Segmentation.get_prediction(slide, mask_path=path, **kw)
        raw = np.zeros((640, 448), np.uint8)
        raw[250:400, 150:250] = 255
        _, want = Segmentation.get_prediction(slide, tissue_mask=raw, **kw)
        _, auto = Segmentation.get_prediction(slide, **kw)
        assert got["mean"].max() > 0 and np.array_equal(got["mean"], want["mean"])
        assert not np.array_equal(got["mean"], auto["mean"])
        with pytest.raises(AttributeError):
            Segmentation.get_prediction(slide, mask_path=str(tmp_path / "mask.png"), **kw)
    finally:
        Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize = saved


def test_get_segmentation_orchestration_status_files_and_return(tmp_path, monkeypatch):
    """getSegmentation end to end through the doubles: status protocol (the strings and the final progress the
    viewer polls, Segmentation.py:292-354), the three result files, and the returned {0, 255} map."""
    from PIL import Image
    from digipathai_b200 import Segmentation, engine
    saved = (Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize)
    try:
        _install_doubles()
        monkeypatch.setattr(Segmentation, "load_trained_models", lambda *a, **k: _FakeModel())
        slide = _slide(0)

        class Status(dict):
            log = []

            def __setitem__(self, k, v):
                if k == "status":
                    self.log.append(v)
                super().__setitem__(k, v)

        status = Status()
        paths = {k: str(tmp_path / f"{k}.tiff") for k in ("probs", "mask", "unc")}
        out = Segmentation.getSegmentation(slide, patch_size=P, stride_size=STRIDE, batch_size=BATCH, quick=True,
                                           tta_list=["FLIP_LEFT_RIGHT"], crf=False, status=status,
                                           probs_path=paths["probs"], mask_path=paths["mask"],
                                           uncertainty_path=paths["unc"], weights={"conv1/conv": np.zeros(1)})
        _, want = Segmentation.get_prediction(slide, batch_size=BATCH, models={"m": _FakeModel()},
                                              tta_list=["FLIP_LEFT_RIGHT"], patch_size=P, stride_size=STRIDE)
    finally:
        Segmentation._torch, engine.stitch, engine.finalize, torch.cuda.device, torch.cuda.synchronize = saved
    assert out.dtype == np.float32 and out.shape == (640, 448) and set(np.unique(out)) == {0.0, 255.0}
    assert np.array_equal(out, np.where(want["mean"] >= np.float32(0.3), 255, 0))
    assert Status.log == ["Found Trained Models, Skipping download", "Loading Trained weights", "Running segmentation",
                          "Saving Prediction Mask...", "Saving Prediction Uncertanity..."]
    assert status["progress"] == 0
    for k, path in paths.items():
        im = Image.open(path)
        assert im.size == (640, 448) and im.n_frames == 3          # [W, H] planes are written transposed: image layout
    mask = np.asarray(Image.open(paths["mask"]))
    assert ((mask > 127) == (out.T > 127)).mean() > 0.999
    probs = np.asarray(Image.open(paths["probs"])).astype(np.float32) / 255.0
    assert np.abs(probs - want["mean"].T).mean() < 0.01
