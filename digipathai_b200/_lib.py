"""ctypes binding of ``libdigipath_b200.so`` (C ABI declared in ``include/digipath_b200.h``).

There is deliberately no fallback: if the shared library is missing or fails to load, importing this module
raises, and every product entry point above it fails loudly.  Build it with ``python -c "import
__graft_entry__ as g; g.build()"`` (or ``digipathai_b200/csrc/build.sh``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdigipath_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension has not been built. Run __graft_entry__.build() "
        "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

c_model_p = C.c_void_p

_SIGS = {
    "dp_abi_version": (C.c_int, []),
    "dp_last_error": (C.c_char_p, []),
    "dp_model_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(c_model_p)]),
    "dp_model_clone": (C.c_int, [c_model_p, C.POINTER(c_model_p)]),
    "dp_model_destroy": (C.c_int, [c_model_p]),
    "dp_model_info": (C.c_int, [c_model_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "dp_model_precision": (C.c_int, [c_model_p]),
    "dp_forward_tiles": (
        C.c_int,
        [c_model_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "dp_stitch": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
         C.c_int64, C.c_int64, C.c_void_p],
    ),
    "dp_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "dp_pyramid_down2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "dp_jpeg_encode_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "dp_jpeg_encode_gray_tiles": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                            C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "dp_jpeg_compact": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "dp_tissue_hist": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dp_tissue_mask": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "dp_morph_rect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dp_crf_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dp_crf_tiles": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
         C.c_float, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "dp_crf_lattice_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dp_crf_tiles_lattice": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
         C.c_float, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "dp_d4_src": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dp_kernel_launch_count": (C.c_uint64, []),
    "dp_model_set_option": (C.c_int, [c_model_p, C.c_char_p, C.c_int]),
    "dp_model_program_size": (C.c_int, [c_model_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dp_model_buffer_shape": (C.c_int, [c_model_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dp_debug_read_buffer": (C.c_int, [c_model_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "dp_debug_write_buffer": (C.c_int, [c_model_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "dp_debug_run_ops": (C.c_int, [c_model_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dp_model_op_times": (C.c_int, [c_model_p, C.POINTER(C.c_float), C.c_int]),
    "dp_model_op_info": (C.c_int, [c_model_p, C.c_int] + [C.POINTER(C.c_int)] * 6 + [C.POINTER(C.c_uint64)]),
    "dp_debug_read_stamps": (C.c_int, [c_model_p, C.POINTER(C.c_uint64), C.c_int]),
    "dp_debug_read_cta_stamps": (C.c_int, [c_model_p, C.c_int, C.POINTER(C.c_uint64), C.c_int]),
    "dp_debug_read_trace": (C.c_int, [c_model_p, C.POINTER(C.c_uint64), C.c_int]),
    "dp_model_executed_macs": (C.c_int, [c_model_p, C.c_int, C.POINTER(C.c_uint64)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


class DigiPathB200Error(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.dp_last_error().decode("utf-8", "replace")
        raise DigiPathB200Error(f"{what}: {msg}" if what else msg)
