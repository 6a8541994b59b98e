"""Device-side tile model: the object that stands where the reference's Keras ``Model`` stood.

``TileModel.predict`` mirrors ``models[name].predict(image_patches, batch_size=...)``
(DigiPathAI/Segmentation.py:154-156): host float32 ``[B,P,P,3]`` in [-1,1] in, host float32 ``[B,P,P,2]``
softmax out.  ``forward_tiles`` is the resident path the B200 ``get_prediction`` loop uses: tiles are cropped
from the slide raster already in HBM, and the softmax channel-1 plane stays on the device for the stitch.

PyTorch is used for device memory and streams only; all arithmetic happens in libdigipath_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .program import Program, serialize


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class TileModel:
    def __init__(self, program: Program | bytes | None, device: int = 0, max_batch: int = 32, *, _clone_of=None):
        if not torch.cuda.is_available():
            raise RuntimeError("digipathai_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._h = _lib.c_model_p()
        if _clone_of is not None:
            self.program, self.device, self.max_batch = _clone_of.program, _clone_of.device, _clone_of.max_batch
            _lib.check(_lib.lib.dp_model_clone(_clone_of._h, C.byref(self._h)), "dp_model_clone")
        else:
            blob = program if isinstance(program, (bytes, bytearray)) else serialize(program)
            self.program = program if isinstance(program, Program) else None
            self.device = int(device)
            self.max_batch = int(max_batch)
            buf = (C.c_char * len(blob)).from_buffer_copy(blob)
            _lib.check(_lib.lib.dp_model_create(buf, len(blob), self.device, self.max_batch, C.byref(self._h)),
                       "dp_model_create")
        p, mb, nb = C.c_int(), C.c_int(), C.c_uint64()
        _lib.check(_lib.lib.dp_model_info(self._h, C.byref(p), C.byref(mb), C.byref(nb)))
        self.patch, self.device_bytes = p.value, nb.value
        self.precision = {0: "fp16", 1: "fp32", 2: "tf32x3"}[_lib.lib.dp_model_precision(self._h)]
        self._buf_dtype = np.float16 if self.precision == "fp16" else np.float32
        self._tile_coords = {}

    # ------------------------------------------------------------------ lifetime
    def clone(self) -> "TileModel":
        """Another execution lane of this model (dp_model_clone): shares the weights in HBM, owns its activation
        buffers and captured graphs, so it can run a different tile batch on another stream at the same time."""
        return TileModel(None, _clone_of=self)

    def lane(self, i: int) -> "TileModel":
        """Lane ``i`` of this model: the model itself for 0, otherwise a clone created on first use and kept until
        ``close`` (so repeated slide runs do not re-allocate ~1 GB of activations per lane)."""
        if i == 0:
            return self
        lanes = self.__dict__.setdefault("_lanes", [])
        while len(lanes) < i:
            lanes.append(self.clone())
        return lanes[i - 1]

    def close(self):
        for m in self.__dict__.pop("_lanes", []):
            m.close()
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib.dp_model_destroy(self._h)
            self._h = _lib.c_model_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ hot path
    def forward_tiles(self, slide: torch.Tensor, coords: torch.Tensor, tta_in: int = 0, tta_out: int = 0,
                      out: torch.Tensor | None = None) -> torch.Tensor:
        """slide: cuda uint8 [W,H,3] ([x,y,c]); coords: cuda int32 [B,2]; returns cuda float32 [B,P,P]."""
        assert slide.is_cuda and slide.dtype == torch.uint8 and slide.dim() == 3 and slide.shape[2] == 3
        assert slide.is_contiguous() and coords.is_cuda and coords.dtype == torch.int32 and coords.is_contiguous()
        B = coords.shape[0]
        if out is None:
            out = torch.empty((B, self.patch, self.patch), dtype=torch.float32, device=slide.device)
        assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == B * self.patch * self.patch
        _lib.check(
            _lib.lib.dp_forward_tiles(self._h, C.c_void_p(slide.data_ptr()), slide.shape[0], slide.shape[1],
                                      C.c_void_p(coords.data_ptr()), B, int(tta_in), int(tta_out),
                                      C.c_void_p(out.data_ptr()), _stream_ptr()),
            "dp_forward_tiles")
        return out

    def _coords_for_batch(self, B: int, device) -> torch.Tensor:
        key = (B, str(device))
        if key not in self._tile_coords:
            c = torch.zeros((B, 2), dtype=torch.int32)
            c[:, 0] = torch.arange(B, dtype=torch.int32) * self.patch
            self._tile_coords[key] = c.to(device)
        return self._tile_coords[key]

    def forward_tile_batch(self, tiles: torch.Tensor, tta_in: int = 0, tta_out: int = 0,
                           out: torch.Tensor | None = None) -> torch.Tensor:
        """tiles: cuda uint8 [B,P,P,3] already gathered -> cuda float32 [B,P,P]."""
        B, P = tiles.shape[0], self.patch
        assert tiles.shape[1:] == (P, P, 3)
        return self.forward_tiles(tiles.reshape(B * P, P, 3), self._coords_for_batch(B, tiles.device), tta_in,
                                  tta_out, out)

    def predict(self, image_patches, batch_size=None, verbose=0, steps=None) -> np.ndarray:
        """Keras-``Model.predict``-shaped entry: host array in, host softmax ``[B,P,P,2]`` out.

        Accepts the reference's float32 tiles in [-1,1] (exact multiples of 1/128, as produced by
        ``(img - 128.0)/128.0`` in dataloader.py:383-388) or the raw uint8 tiles.
        """
        x = np.asarray(image_patches)
        if x.dtype != np.uint8:
            u = np.rint(x.astype(np.float32) * 128.0 + 128.0)
            if np.abs(u - (x * 128.0 + 128.0)).max() > 1e-3 or u.min() < 0 or u.max() > 255:
                raise ValueError("predict() expects tiles normalised as (uint8 - 128)/128")
            x = u.astype(np.uint8)
        dev = torch.device("cuda", self.device)
        out = np.empty(x.shape[:3] + (2,), dtype=np.float32)
        step = self.max_batch
        for s in range(0, x.shape[0], step):
            t = torch.from_numpy(np.ascontiguousarray(x[s:s + step])).pin_memory().to(dev, non_blocking=True)
            p1 = self.forward_tile_batch(t).cpu().numpy()
            out[s:s + step, ..., 1] = p1
            out[s:s + step, ..., 0] = 1.0 - p1
        return out

    # ------------------------------------------------------------------ introspection / debugging
    def set_option(self, key: str, value: int):
        _lib.check(_lib.lib.dp_model_set_option(self._h, key.encode(), int(value)))

    def buffer_shape(self, buf: int):
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib.dp_model_buffer_shape(self._h, buf, C.byref(h), C.byref(w), C.byref(c)))
        return h.value, w.value, c.value

    def read_buffer(self, buf: int, n_tiles: int) -> np.ndarray:
        h, w, c = self.buffer_shape(buf)
        a = np.empty((n_tiles, h, w, c), dtype=self._buf_dtype)
        _lib.check(_lib.lib.dp_debug_read_buffer(self._h, buf, n_tiles, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    def write_buffer(self, buf: int, arr: np.ndarray):
        a = np.ascontiguousarray(arr, dtype=self._buf_dtype)
        _lib.check(_lib.lib.dp_debug_write_buffer(self._h, buf, a.shape[0], a.ctypes.data_as(C.c_void_p), a.nbytes))

    def run_ops(self, n_tiles: int, op_begin: int, op_end: int, tta_out: int = 0, probs: torch.Tensor | None = None):
        ptr = C.c_void_p(probs.data_ptr()) if probs is not None else C.c_void_p(None)
        _lib.check(_lib.lib.dp_debug_run_ops(self._h, n_tiles, op_begin, op_end, int(tta_out), ptr, _stream_ptr()),
                   "dp_debug_run_ops")

    def n_ops(self) -> int:
        a, b = C.c_int(), C.c_int()
        _lib.check(_lib.lib.dp_model_program_size(self._h, C.byref(a), C.byref(b)))
        return a.value

    def op_times_ms(self) -> np.ndarray:
        n = self.n_ops()
        arr = (C.c_float * n)()
        _lib.check(_lib.lib.dp_model_op_times(self._h, arr, n))
        return np.array(arr[:], dtype=np.float64)

    def op_info(self, op: int) -> dict:
        v = [C.c_int() for _ in range(6)]
        macs = C.c_uint64()
        _lib.check(_lib.lib.dp_model_op_info(self._h, op, *[C.byref(x) for x in v], C.byref(macs)))
        keys = ("type", "kind", "cin", "cout", "h", "w")
        d = {k: x.value for k, x in zip(keys, v)}
        d["macs_per_tile"] = macs.value
        return d

    def read_stamps(self) -> np.ndarray:
        n = 2 * self.n_ops()
        arr = (C.c_uint64 * n)()
        _lib.check(_lib.lib.dp_debug_read_stamps(self._h, arr, n))
        return np.array(arr[:], dtype=np.int64).reshape(-1, 2)

    def read_cta_stamps(self, op: int) -> np.ndarray:
        """[256, 4] per CTA of dense-layer op ``op``: %globaltimer at entry, when the grid-dependency wait returned, at
        exit, and the SM id (option 'stamp_ctas'; rows of CTAs that did not run are zero)."""
        arr = (C.c_uint64 * 1024)()
        _lib.check(_lib.lib.dp_debug_read_cta_stamps(self._h, int(op), arr, 1024))
        return np.array(arr[:], dtype=np.int64).reshape(256, 4)

    def read_trace(self):
        """[(role, event, item, clock32)] of CTA 0 for the op selected with set_option('trace_op', i)."""
        arr = (C.c_uint64 * 10016)()
        _lib.check(_lib.lib.dp_debug_read_trace(self._h, arr, 10016))
        out = []
        for role in range(5):
            n = min(int(arr[role]), 2000)
            base = 8 + 2000 * role
            for v in arr[base:base + n]:
                out.append((9 if role == 4 else role, (v >> 48) & 0xFF, (v >> 32) & 0xFFFF, v & 0xFFFFFFFF))
        return out

    def executed_macs(self, n_tiles: int) -> int:
        v = C.c_uint64()
        _lib.check(_lib.lib.dp_model_executed_macs(self._h, n_tiles, C.byref(v)))
        return v.value


class ForwardLanes:
    """``lanes`` tile batches in flight on one GPU.

    The batches of a slide (and the TTA passes / ensemble members of one batch, Segmentation.py:150-160) are
    independent, while a batch-32 forward has long stretches that cannot fill 148 SMs (the 16x16 / 8x8 dense blocks run
    128 / 32 CTAs, every kernel has a tail).  Lane ``s`` is a CUDA stream plus one clone of every model
    (``TileModel.clone``: shared weights, own activations); consecutive ``forward`` calls go to consecutive lanes, so
    the hardware overlaps one batch's thin kernels with another batch's wide ones.  Results are bit-identical to the
    single-stream order (same kernels on the same data).  ``lanes=1`` issues on the caller's stream, unchanged.

        lanes = ForwardLanes({'dense': model}, lanes=3)
        lanes.begin()                                  # lanes wait for what the caller's stream has queued
        lanes.forward('dense', raster, coords, t_in, t_out, out=probs[k])
        lanes.join()                                   # caller's stream waits for every lane
    """

    def __init__(self, models: dict, lanes: int = 3):
        self.L = max(1, int(lanes))
        self.models = {nm: [m.lane(i) for i in range(self.L)] if self.L > 1 else [m] for nm, m in models.items()}
        self.streams = []
        if self.L > 1:
            with torch.cuda.device(torch.device("cuda", next(iter(models.values())).device)):
                self.streams = [torch.cuda.Stream() for _ in range(self.L)]
        self.j = 0

    def begin(self):
        self.wait_event(None)

    def wait_event(self, ev=None):
        """Every lane waits for ``ev`` (default: for the work queued on the caller's stream so far)."""
        if self.L == 1:
            return
        if ev is None:
            ev = torch.cuda.Event()
            ev.record()
        for st in self.streams:
            st.wait_event(ev)

    def forward(self, name, slide, coords, tta_in=0, tta_out=0, out=None):
        if self.L == 1:
            return self.models[name][0].forward_tiles(slide, coords, tta_in, tta_out, out=out)
        s = self.j % self.L
        self.j += 1
        with torch.cuda.stream(self.streams[s]):
            return self.models[name][s].forward_tiles(slide, coords, tta_in, tta_out, out=out)

    def join(self):
        if self.L == 1:
            return
        cur = torch.cuda.current_stream()
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    def close(self):
        """Lanes are owned by their models (TileModel.lane) and live until those are closed; this only waits."""
        self.join()


class HostBatchPipeline:
    """Host-buffer front end for a stream of tile batches (what stands where a loop of ``model.predict`` calls
    stood, Segmentation.py:150-156): the H2D copy of batch k+1, the forward of batch k and the D2H copy of batch
    k-1 run on three CUDA streams with double-buffered device tensors, so the PCIe transfers (6.3 MB in, 8.4 MB
    out per batch of 32) hide behind the 1.9 ms forward instead of adding ~0.27 ms to it.

        pipe = HostBatchPipeline(model, batch=32)
        for tiles_u8, probs in batches:          # pinned uint8 [B,P,P,3] in, pinned float32 [B,P,P] out
            pipe.submit(tiles_u8, probs)
        pipe.drain()                             # results are in the host buffers after this returns
    """

    def __init__(self, model: "TileModel", batch: int | None = None, tta_in: int = 0, tta_out: int = 0,
                 lanes: int = 1):
        self.model, self.B, self.P = model, int(batch or model.max_batch), model.patch
        assert self.B <= model.max_batch
        self.tta = (int(tta_in), int(tta_out))
        dev = torch.device("cuda", model.device)
        self.dev = dev
        self.L = max(1, int(lanes))                      # forwards in flight (ForwardLanes: one model clone each)
        self.n_slots = 2 * self.L                        # device buffers: one being filled + one in flight per lane
        self.lane_models = [model.lane(i) for i in range(self.L)]
        with torch.cuda.device(dev):
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
            self.s_lanes = [torch.cuda.Stream() for _ in range(self.L)]
            self.s_comp = self.s_lanes[0]
            n = self.n_slots
            self.d_in = [torch.empty((self.B, self.P, self.P, 3), dtype=torch.uint8, device=dev) for _ in range(n)]
            self.d_out = [torch.empty((self.B, self.P, self.P), dtype=torch.float32, device=dev) for _ in range(n)]
            self.in_ready = [torch.cuda.Event() for _ in range(n)]
            self.comp_done = [torch.cuda.Event() for _ in range(n)]
            self.out_done = [torch.cuda.Event() for _ in range(n)]
        self.k = 0

    def submit(self, tiles_u8: torch.Tensor, probs_out: torch.Tensor) -> None:
        """tiles_u8: pinned host uint8 [B,P,P,3]; probs_out: pinned host float32 [B,P,P] (filled asynchronously)."""
        assert not tiles_u8.is_cuda and tiles_u8.is_pinned() and tiles_u8.dtype == torch.uint8
        assert not probs_out.is_cuda and probs_out.is_pinned() and probs_out.dtype == torch.float32
        assert tuple(tiles_u8.shape) == (self.B, self.P, self.P, 3) and probs_out.numel() == self.B * self.P * self.P
        slot, reuse = self.k % self.n_slots, self.k >= self.n_slots
        lane = self.k % self.L
        s_comp = self.s_lanes[lane]
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_in):
                if reuse:
                    self.s_in.wait_event(self.comp_done[slot])     # the forward that last used d_in[slot] has consumed it
                self.d_in[slot].copy_(tiles_u8, non_blocking=True)
                self.in_ready[slot].record(self.s_in)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(self.in_ready[slot])
                if reuse:
                    s_comp.wait_event(self.out_done[slot])         # the D2H that last used d_out[slot] has drained it
                self.lane_models[lane].forward_tile_batch(self.d_in[slot], self.tta[0], self.tta[1],
                                                          out=self.d_out[slot])
                self.comp_done[slot].record(s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.comp_done[slot])
                probs_out.view(self.B, self.P, self.P).copy_(self.d_out[slot], non_blocking=True)
                self.out_done[slot].record(self.s_out)
        self.k += 1

    def drain(self) -> None:
        for st in [self.s_in, *self.s_lanes, self.s_out]:
            st.synchronize()

    def close(self) -> None:
        self.drain()


# ---------------------------------------------------------------------- slide-plane kernels
def stitch(probs: torch.Tensor, coords: torch.Tensor, mean: torch.Tensor, var: torch.Tensor, count: torch.Tensor,
           x_lo: int = 0):
    """probs cuda f32 [N,B,P,P]; coords cuda int32 [B,2]; planes cuda [w,h] (Segmentation.py:162-173)."""
    N, B, P, _ = probs.shape
    assert probs.is_contiguous() and mean.is_contiguous() and var.is_contiguous() and count.is_contiguous()
    assert mean.dtype == torch.float32 and var.dtype == torch.float32 and count.dtype == torch.uint8
    _lib.check(
        _lib.lib.dp_stitch(C.c_void_p(probs.data_ptr()), N, B, P, C.c_void_p(coords.data_ptr()),
                           C.c_void_p(mean.data_ptr()), C.c_void_p(var.data_ptr()), C.c_void_p(count.data_ptr()),
                           mean.shape[0], mean.shape[1], int(x_lo), _stream_ptr()),
        "dp_stitch")


def finalize(mean: torch.Tensor, var: torch.Tensor, count: torch.Tensor, threshold: float,
             label: torch.Tensor | None = None):
    """In-place normalise (Segmentation.py:175-177) and optional threshold to {0,255} (:336-337)."""
    lp = C.c_void_p(label.data_ptr()) if label is not None else C.c_void_p(None)
    _lib.check(
        _lib.lib.dp_finalize(C.c_void_p(mean.data_ptr()), C.c_void_p(var.data_ptr()), C.c_void_p(count.data_ptr()),
                             mean.numel(), float(np.float32(threshold)), lp, _stream_ptr()),
        "dp_finalize")


def pyramid_down2(plane: torch.Tensor) -> torch.Tensor:
    w, h = plane.shape
    out = torch.empty((w // 2, h // 2), dtype=torch.float32, device=plane.device)
    _lib.check(_lib.lib.dp_pyramid_down2(C.c_void_p(plane.data_ptr()), w, h, C.c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def dense_crf(rgb: torch.Tensor, p1: torch.Tensor, n_iter: int = 10, sdims_gauss: float = 10.0,
              compat_gauss: float = 3.0, sdims_bilateral: float = 50.0, schan_bilateral: float = 20.0,
              compat_bilateral: float = 10.0, return_marginal: bool = False, method: str = "lattice"):
    """Fully connected CRF refinement (``post_process_crf``, DigiPathAI/helpers/utils.py:568-603; defaults = its
    parameters).  rgb cuda uint8 [n,h,w,3]; p1 cuda float32 [n,h,w]; returns cuda uint8 labels [n,h,w] in {0,1}
    (and the label-1 marginal when asked).  ``method='lattice'``: Gaussian filters on the permutohedral lattice, as
    pydensecrf evaluates them (csrc/crf_lattice.cuh); ``'exact'``: all-pairs evaluation of the same filters
    (csrc/crf.cuh, ~15x slower)."""
    assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.dim() == 4 and rgb.shape[3] == 3 and rgb.is_contiguous()
    assert p1.is_cuda and p1.dtype == torch.float32 and p1.shape == rgb.shape[:3] and p1.is_contiguous()
    n, h, w = p1.shape
    if method not in ("lattice", "exact"):
        raise ValueError("method must be 'lattice' or 'exact'")
    ws_fn, run_fn = ((_lib.lib.dp_crf_lattice_workspace_bytes, _lib.lib.dp_crf_tiles_lattice) if method == "lattice"
                     else (_lib.lib.dp_crf_workspace_bytes, _lib.lib.dp_crf_tiles))
    nbytes = int(ws_fn(n, h, w))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=p1.device)
    labels = torch.empty((n, h, w), dtype=torch.uint8, device=p1.device)
    q1 = torch.empty((n, h, w), dtype=torch.float32, device=p1.device) if return_marginal else None
    _lib.check(
        run_fn(C.c_void_p(rgb.data_ptr()), C.c_void_p(p1.data_ptr()), n, h, w, int(n_iter),
                              float(sdims_gauss), float(compat_gauss), float(sdims_bilateral), float(schan_bilateral),
                              float(compat_bilateral), C.c_void_p(ws.data_ptr()), nbytes,
                              C.c_void_p(labels.data_ptr()),
                              C.c_void_p(q1.data_ptr()) if q1 is not None else C.c_void_p(None), _stream_ptr()),
        "dp_crf_tiles" if method == "exact" else "dp_crf_tiles_lattice")
    return (labels, q1) if return_marginal else labels


def kernel_launch_count() -> int:
    return int(_lib.lib.dp_kernel_launch_count())
