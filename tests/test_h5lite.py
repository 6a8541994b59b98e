"""digipathai_b200/h5lite.py (pure-Python HDF5 reader for Keras ``.h5`` weight files, rows a2 / N4) against files
emitted by the independent classic-layout writer in tests/h5_writer.py: groups deep and wide enough for multi-level
B-trees, contiguous / compact / chunked (+deflate, +shuffle) datasets, fixed- and variable-length string attributes,
attributes in a continuation block, big-endian data; and the Keras layout end to end:
file -> read_keras_weights -> tools/h5_to_npz.map_layers -> the DenseNet weight dict, bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from h5_writer import Writer  # noqa: E402

from digipathai_b200 import h5lite  # noqa: E402


def test_groups_datasets_and_attributes_round_trip():
    rng = np.random.default_rng(0)
    w = Writer()
    a = rng.standard_normal((3, 3, 16, 8)).astype(np.float32)
    b = rng.integers(-5, 5, (7,)).astype(np.int64)
    c = rng.standard_normal((5, 4)).astype(">f8")
    w.group("/", attrs={"layer_names": np.array([b"conv1", b"a_much_longer_layer_name", b"bn"]), "backend": "tensorflow",
                        "keras_version": np.bytes_(b"2.2.4"), "n": np.int32(7)}, continuation=True)
    w.dataset("/conv1/conv1/kernel:0", a, attrs={"scale": np.float64(0.5)})
    w.dataset("/conv1/conv1/bias:0", b)
    w.dataset("/bn/be", c, layout="compact")
    w.group("/conv1", attrs={"weight_names": ["conv1/kernel:0", "conv1/bias:0"]})
    f = h5lite.File(w.finish())
    assert sorted(f.keys()) == ["bn", "conv1"]
    at = f.attrs
    assert [x.decode() for x in at["layer_names"]] == ["conv1", "a_much_longer_layer_name", "bn"]
    assert at["backend"] == "tensorflow" and at["keras_version"] == b"2.2.4" and int(at["n"]) == 7
    assert list(f["conv1"].attrs["weight_names"]) == ["conv1/kernel:0", "conv1/bias:0"]
    k = f["conv1/conv1/kernel:0"]
    assert k.shape == a.shape and np.array_equal(np.asarray(k), a) and float(k.attrs["scale"]) == 0.5
    assert np.array_equal(np.asarray(f["/conv1/conv1/bias:0"]), b)
    got = np.asarray(f["bn/be"])
    assert got.dtype.byteorder in "=<|" and np.array_equal(got, c.astype("<f8"))
    assert "nope" not in f
    with pytest.raises(KeyError):
        f["conv1/missing"]


def test_wide_group_needs_a_multi_level_btree():
    w = Writer()
    names = [f"layer_{i:04d}" for i in range(700)]                 # 88 leaves -> 3 internal nodes -> a level-2 root
    for i, n in enumerate(names):
        w.dataset(f"/{n}/v", np.full((2,), i, np.float32))
    f = h5lite.File(w.finish())
    assert sorted(f.keys()) == names
    for i in (0, 1, 255, 256, 699):
        assert np.array_equal(np.asarray(f[f"{names[i]}/v"]), np.full((2,), i, np.float32))


@pytest.mark.parametrize("deflate,shuffle", [(False, False), (True, False), (True, True)])
def test_chunked_datasets(deflate, shuffle):
    rng = np.random.default_rng(1)
    a = rng.standard_normal((70, 33)).astype(np.float32)
    big = np.arange(40 * 40 * 40, dtype=np.int32).reshape(40, 40, 40)
    w = Writer()
    w.dataset("/a", a, chunks=(16, 8), deflate=deflate, shuffle=shuffle)       # ragged edge chunks
    w.dataset("/big", big, chunks=(8, 8, 8), deflate=deflate, shuffle=shuffle)   # 125 chunks: two B-tree levels
    f = h5lite.File(w.finish())
    assert np.array_equal(np.asarray(f["a"]), a)
    assert np.array_equal(np.asarray(f["big"]), big)


def test_rejects_what_it_cannot_read():
    with pytest.raises(h5lite.H5Error):
        h5lite.File(b"not an hdf5 file at all" * 100)
    data = bytearray(Writer().finish())
    data[8] = 9                                                     # unknown superblock version
    with pytest.raises(h5lite.H5Error):
        h5lite.File(bytes(data))


def test_keras_layout_to_densenet_weights(tmp_path):
    import h5_to_npz
    from digipathai_b200.models import densenet as DN
    from test_h5_mapping import _as_keras_layers
    w, shapes = DN.init_densenet_weights(3), DN.layer_shapes()
    uc = {n + "_conv" for n, _, _ in DN.DECODER} | {"head"}
    ub = {n + "_norm" for n, _, _ in DN.DECODER}
    layers = _as_keras_layers(w, shapes, 17, uc, ub, "_bias")
    wr = Writer()
    prefix = "/model_weights"                                        # a full model.save() file nests the weights here
    names = list(layers.keys()) + ["input_1", "relu_without_weights"]
    wr.group(prefix, attrs={"layer_names": np.array([n.encode() for n in names]), "backend": np.bytes_(b"tensorflow"),
                            "keras_version": np.bytes_(b"2.2.4")})
    for ln, wd in layers.items():
        wr.group(f"{prefix}/{ln}", attrs={"weight_names": np.array([k.encode() for k in wd])})
        for k, v in wd.items():
            wr.dataset(f"{prefix}/{ln}/{k}", np.asarray(v, np.float32))
    for ln in names[-2:]:
        wr.group(f"{prefix}/{ln}", attrs={"weight_names": np.zeros((0,), "S1")})
    path = tmp_path / "digestpath_densenet.h5"
    path.write_bytes(wr.finish())
    got = h5_to_npz.map_layers("dense", h5lite.read_keras_weights(str(path)))
    assert got.keys() == w.keys()
    for k in w:
        if isinstance(w[k], tuple):
            assert all(np.array_equal(x, y) for x, y in zip(got[k], w[k])), k
        else:
            assert np.array_equal(got[k], w[k]), k


def test_load_weights_dispatches_on_the_extension(tmp_path):
    """Segmentation._load_weights (what load_trained_models calls with a path): .h5 -> Keras reader, .npz -> flat file,
    a missing default .npz with the reference's .h5 beside it -> that .h5."""
    from digipathai_b200 import Segmentation
    from digipathai_b200.models import deeplab as DL
    from test_h5_mapping import _as_keras_layers
    w, shapes = DL.init_deeplab_weights(2), DL.layer_shapes()
    layers = _as_keras_layers(w, shapes, 0, set(), set(), "/bias")
    wr = Writer()
    wr.group("/", attrs={"layer_names": np.array([n.encode() for n in layers])})
    for ln, wd in layers.items():
        wr.group(f"/{ln}", attrs={"weight_names": np.array([k.encode() for k in wd])})
        for k, v in wd.items():
            wr.dataset(f"/{ln}/{k}", np.asarray(v, np.float32))
    (tmp_path / "paip_deeplabv3.h5").write_bytes(wr.finish())
    for name in ("paip_deeplabv3.h5", "paip_deeplabv3.npz"):
        got = Segmentation._load_weights("deeplabv3", str(tmp_path / name))
        assert got.keys() == w.keys()
        k = next(k for k in w if not isinstance(w[k], tuple))
        assert np.array_equal(got[k], w[k])
    with pytest.raises(FileNotFoundError):
        Segmentation._load_weights("deeplabv3", str(tmp_path / "other.h5"))


def test_reads_a_file_written_by_libhdf5():
    """The one libhdf5-written file this image holds: scipy's test fixture ``testhdf5_7.4_GLNX86.mat`` (a MATLAB 7.4 v7.3
    MAT-file = HDF5 1.x behind a 512-byte user block; copied to tests/golden/ from
    site-packages/scipy/io/matlab/tests/data, BSD licence).  It pins the classic-layout parsing -- superblock 0, version-1
    object header, symbol-table group (B-tree + local heap + SNOD), dataspace / datatype / layout messages, a
    fixed-length string attribute -- against bytes libhdf5 itself produced.  Known answer: MATLAB's ``0:pi/4:2*pi``."""
    path = os.path.join(ROOT, "tests", "golden", "testhdf5_7.4_GLNX86.mat")
    f = h5lite.File(path)
    assert f.keys() == ["testdouble"]
    ds = f["testdouble"]
    assert ds.attrs["MATLAB_class"] == b"double"
    a = np.asarray(ds)
    assert a.dtype == np.float64 and a.shape == (9, 1)
    assert np.array_equal(a[:, 0], np.arange(9) * (np.pi / 4))
    with pytest.raises(h5lite.H5Error):
        h5lite.File(open(path, "rb").read()[:600])            # truncated behind the user block
