"""Test infrastructure: a minimal HDF5 *writer* for the classic on-disk layout libhdf5 produces with its default
``libver='earliest'`` (superblock 0, version-1 object headers, symbol-table groups = v1 B-tree + local heap + SNOD
leaves, contiguous / chunked datasets, version-1 attribute messages, a global heap for variable-length strings).
Written from the HDF5 File Format Specification independently of digipathai_b200/h5lite.py (it imports nothing from
it) so that the reader is not merely checked against its own assumptions; libhdf5 itself is not available here.

    w = Writer()
    w.group("/", attrs={"layer_names": np.array([b"conv1", b"bn1"])})
    w.dataset("/conv1/conv1/kernel:0", array)
    open(path, "wb").write(w.finish())
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + bytes(-len(b) % 8)


def _dt_float(size, big=False):
    exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
    bits = bytes([0x20 | (1 if big else 0), 8 * size - 1, 0])
    props = struct.pack("<HHBBBBI", 0, 8 * size, man_bits, exp_bits, 0, man_bits, bias)
    return bytes([0x11]) + bits + struct.pack("<I", size) + props


def _dt_int(size, signed, big=False):
    bits = bytes([(0x08 if signed else 0) | (1 if big else 0), 0, 0])
    return bytes([0x10]) + bits + struct.pack("<I", size) + struct.pack("<HH", 0, 8 * size)


def _dt_string(size):
    return bytes([0x13, 0x01, 0, 0]) + struct.pack("<I", size)          # null-padded ASCII


def _dt_vlen_string():
    return bytes([0x19, 0x01, 0, 0]) + struct.pack("<I", 16) + _dt_string(1)


def _datatype(dtype: np.dtype) -> bytes:
    if dtype.kind == "f":
        return _dt_float(dtype.itemsize, dtype.byteorder == ">")
    if dtype.kind in "iu":
        return _dt_int(dtype.itemsize, dtype.kind == "i", dtype.byteorder == ">")
    if dtype.kind == "S":
        return _dt_string(dtype.itemsize)
    raise TypeError(dtype)


def _dataspace(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _message(mtype, body: bytes, flags=0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


class Writer:
    def __init__(self):
        self.buf = bytearray(96)                 # superblock is patched in finish()
        self.tree = {"/": {"kind": "group", "attrs": {}, "children": {}}}
        self.vlen = []                           # (placeholder position in the attribute data, bytes)

    # ---------------------------------------------------------------- building the tree (in memory)
    def _node(self, path, create=True):
        node = self.tree["/"]
        for part in [p for p in path.split("/") if p]:
            if part not in node["children"]:
                if not create:
                    raise KeyError(path)
                node["children"][part] = {"kind": "group", "attrs": {}, "children": {}}
            node = node["children"][part]
        return node

    def group(self, path, attrs=None, continuation=False):
        n = self._node(path)
        n["attrs"].update(attrs or {})
        n["continuation"] = continuation
        return n

    def dataset(self, path, array, attrs=None, chunks=None, deflate=False, shuffle=False, layout="contiguous"):
        parent, name = path.rsplit("/", 1)
        p = self._node(parent or "/")
        p["children"][name] = {"kind": "dataset", "array": np.asarray(array), "attrs": dict(attrs or {}),
                               "chunks": chunks, "deflate": deflate, "shuffle": shuffle, "layout": layout}

    # ---------------------------------------------------------------- allocation
    def _alloc(self, data: bytes) -> int:
        self.buf += bytes(-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---------------------------------------------------------------- attributes
    def _attr_message(self, name, value) -> bytes:
        nm = name.encode() + b"\0"
        if isinstance(value, str):               # scalar variable-length string (what h5py writes for str)
            gaddr = self._global_heap([value.encode()])
            dt, ds = _dt_vlen_string(), _dataspace(())
            data = struct.pack("<IQI", len(value.encode()), gaddr, 1)
        elif isinstance(value, (list, tuple)) and value and isinstance(value[0], str):
            enc = [v.encode() for v in value]
            gaddr = self._global_heap(enc)
            dt, ds = _dt_vlen_string(), _dataspace((len(enc),))
            data = b"".join(struct.pack("<IQI", len(e), gaddr, i + 1) for i, e in enumerate(enc))
        else:
            a = np.asarray(value)
            dt, ds, data = _datatype(a.dtype), _dataspace(a.shape), a.tobytes()
        head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        return _message(0x0C, head + _pad8(nm) + _pad8(dt) + _pad8(ds) + data)

    def _global_heap(self, objects) -> int:
        body = b""
        for i, o in enumerate(objects):
            body += struct.pack("<HH4xQ", i + 1, 1, len(o)) + _pad8(o)
        free = 4096 - 16 - len(body) if len(body) + 16 + 16 <= 4096 else 16
        body += struct.pack("<HH4xQ", 0, 0, free) + bytes(free - 16 if free >= 16 else 0)
        return self._alloc(b"GCOL" + struct.pack("<B3xQ", 1, 16 + len(body)) + body)

    # ---------------------------------------------------------------- object headers
    def _object_header(self, messages, continuation=False) -> int:
        """messages: list of encoded messages.  With ``continuation`` everything after the first message moves into a
        continuation block referenced by a 0x10 message."""
        if continuation and len(messages) > 1:
            tail = b"".join(messages[1:])
            taddr = self._alloc(tail)
            first = messages[0] + _message(0x10, struct.pack("<QQ", taddr, len(tail)))
            n = len(messages) + 1
        else:
            first, n = b"".join(messages), len(messages)
        return self._alloc(struct.pack("<BxHII4x", 1, n, 1, len(first)) + first)

    def _write_dataset(self, d) -> int:
        a = d["array"]
        msgs = [_message(0x01, _dataspace(a.shape)), _message(0x03, _datatype(a.dtype), flags=1)]
        if d["chunks"]:
            cd = tuple(d["chunks"])
            filters = []
            if d["shuffle"]:
                filters.append((2, [a.dtype.itemsize]))
            if d["deflate"]:
                filters.append((1, [6]))
            if filters:
                fb = struct.pack("<BB6x", 1, len(filters))
                for fid, cdv in filters:
                    fb += struct.pack("<HHHH", fid, 0, 1, len(cdv)) + b"".join(struct.pack("<I", v) for v in cdv)
                    if len(cdv) % 2:
                        fb += bytes(4)
                msgs.append(_message(0x0B, fb))
            entries = []
            grid = [range(0, s, c) for s, c in zip(a.shape, cd)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = [g[i] for g, i in zip(grid, offs)]
                chunk = np.zeros(cd, a.dtype)
                sl = tuple(slice(oo, min(oo + c, s)) for oo, c, s in zip(o, cd, a.shape))
                chunk[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
                raw = chunk.tobytes()
                if d["shuffle"]:
                    es = a.dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
                if d["deflate"]:
                    raw = zlib.compress(raw, 6)
                entries.append((o, len(raw), self._alloc(raw)))
            bt = self._chunk_btree(entries, a.ndim, a.shape, cd)
            msgs.append(_message(0x08, struct.pack("<BBBQ", 3, 2, a.ndim + 1, bt) +
                                 b"".join(struct.pack("<I", c) for c in cd) + struct.pack("<I", a.dtype.itemsize)))
        elif d["layout"] == "compact":
            raw = a.tobytes()
            msgs.append(_message(0x08, struct.pack("<BBH", 3, 0, len(raw)) + raw))
        else:
            raw = a.tobytes()
            addr = self._alloc(raw) if raw else UNDEF
            msgs.append(_message(0x08, struct.pack("<BBQQ", 3, 1, addr, len(raw))))
        msgs += [self._attr_message(k, v) for k, v in d["attrs"].items()]
        return self._object_header(msgs)

    def _chunk_btree(self, entries, rank, shape, cd) -> int:
        def key(offs, size):
            return struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

        def node(level, items):                  # items: (first offsets, size, child address)
            body = b""
            for offs, size, child in items:
                body += key(offs, size) + struct.pack("<Q", child)
            body += key([s for s in shape], 0)   # closing key
            return self._alloc(b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF) + body)

        level, items = 0, entries
        while True:
            groups = [items[i:i + 2 * INTERNAL_K] for i in range(0, len(items), 2 * INTERNAL_K)] or [[]]
            nodes = [(g[0][0] if g else [0] * rank, g[0][1] if g else 0, node(level, g)) for g in groups]
            if len(nodes) == 1:
                return nodes[0][2]
            level, items = level + 1, nodes

    def _write_group(self, g) -> int:
        names = sorted(g["children"].keys(), key=lambda s: s.encode())
        addrs = {}
        for n in names:
            c = g["children"][n]
            addrs[n] = self._write_group(c) if c["kind"] == "group" else self._write_dataset(c)
        # local heap: offset 0 holds the empty string
        seg, offsets = bytearray(8), {}
        for n in names:
            offsets[n] = len(seg)
            seg += _pad8(n.encode() + b"\0")
        seg_addr = self._alloc(bytes(seg))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), UNDEF, seg_addr))
        # leaves: at most 2 * LEAF_K symbols each
        leaves = []
        for i in range(0, len(names), 2 * LEAF_K):
            part = names[i:i + 2 * LEAF_K]
            body = b"".join(struct.pack("<QQII16x", offsets[n], addrs[n], 0, 0) for n in part)
            body += bytes(40 * (2 * LEAF_K - len(part)))
            leaves.append((offsets[part[-1]], self._alloc(b"SNOD" + struct.pack("<BxH", 1, len(part)) + body)))
        level, items = 0, leaves
        while True:
            groups = [items[i:i + 2 * INTERNAL_K] for i in range(0, len(items), 2 * INTERNAL_K)] or [[]]
            nodes = []
            for grp in groups:
                body = struct.pack("<Q", 0)
                for last_key, child in grp:
                    body += struct.pack("<QQ", child, last_key)
                nodes.append((grp[-1][0] if grp else 0,
                              self._alloc(b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), UNDEF, UNDEF) + body)))
            if len(nodes) == 1:
                btree = nodes[0][1]
                break
            level, items = level + 1, nodes
        msgs = [_message(0x11, struct.pack("<QQ", btree, heap))]
        msgs += [self._attr_message(k, v) for k, v in g["attrs"].items()]
        g["_btree"], g["_heap"] = btree, heap
        return self._object_header(msgs, continuation=g.get("continuation", False))

    # ---------------------------------------------------------------- file
    def finish(self) -> bytes:
        root = self.tree["/"]
        raddr = self._write_group(root)
        self.buf += bytes(-len(self.buf) % 8)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, raddr, 1, 0) + struct.pack("<QQ", root["_btree"], root["_heap"])
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)
