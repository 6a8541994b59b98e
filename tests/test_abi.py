"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "digipath_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from digipathai_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/digipath_b200.h but not exported"
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding and header out of sync"


def test_ingest_library_exports_every_declared_symbol():
    from digipathai_b200 import ingest
    src = open(os.path.join(ROOT, "include", "digipath_ingest.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", src)))
    assert len(syms) == 7
    for s in syms:
        assert hasattr(ingest.lib, s), f"{s} declared in include/digipath_ingest.h but not exported"
    assert set(syms) == set(ingest.EXPORTED_SYMBOLS)
    assert ingest.lib.dp_ingest_abi_version() == 1
    # argument checks that need no GPU
    assert ingest.lib.dp_jpeg_decoder_create(0, None) != 0 and b"null" in ingest.lib.dp_ingest_last_error()
    assert ingest.lib.dp_scatter_tiles_xy(None, 1, 8, 8, None, None, 0, 8, 8, None) != 0


def test_abi_version_and_error_channel():
    from digipathai_b200 import _lib
    assert _lib.lib.dp_abi_version() == 1
    h = _lib.c_model_p()
    bad = b"NOPE" + b"\0" * 200
    rc = _lib.lib.dp_model_create(bad, len(bad), 0, 4, ctypes.byref(h))
    assert rc != 0
    assert b"DPB1" in _lib.lib.dp_last_error()
    rc = _lib.lib.dp_model_create(bad[:8], 8, 0, 4, ctypes.byref(h))
    assert rc != 0 and b"too small" in _lib.lib.dp_last_error()
    assert _lib.lib.dp_finalize(None, None, None, 10, 0.3, None, None) != 0


def test_d4_source_map_matches_host_table():
    from digipathai_b200 import _lib, tta
    P = 7
    for code in range(8):
        for i in range(P):
            for j in range(P):
                a, b = ctypes.c_int(), ctypes.c_int()
                _lib.lib.dp_d4_src(code, i, j, P, ctypes.byref(a), ctypes.byref(b))
                assert (a.value, b.value) == tta.src(code, i, j, P)


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "digipathai_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_container_roundtrip_header():
    import struct
    from conv_cases import build_case
    from digipathai_b200.program import serialize
    pr, _, _ = build_case("h_up2_16")
    blob = serialize(pr)
    magic, ver, nb, no, patch, _, doff, total = struct.unpack_from("<4s5I2Q", blob, 0)
    assert magic == b"DPB1" and ver == 1 and nb == 2 and no == 1 and total == len(blob) and doff % 256 == 0
    op = struct.unpack_from("<12if3i8q", blob, 72 + 16 * nb)
    assert op[0] == 3 and op[7] == 4          # OP_CONV, KIND_UP2
    w_off = op[16]
    w = np.frombuffer(blob, dtype=np.float16, count=16 * 64 * 64, offset=doff + w_off).reshape(16, 64, 64)
    assert np.array_equal(w, pr.ops[0].w)
