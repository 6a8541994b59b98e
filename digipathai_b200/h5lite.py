"""Minimal pure-Python reader for the HDF5 files Keras writes with ``model.save_weights('x.h5')`` -- what the
reference loads with ``model.load_weights(path)`` (DigiPathAI/helpers/utils.py:427-448; the nine ``.h5`` checkpoints it
downloads are listed at utils.py:58-98).  h5py / libhdf5 do not exist in this image, so the container format is read
directly, following the public "HDF5 File Format Specification Version 3.0":

* superblock versions 0-3; object headers version 1 (what libhdf5's default ``libver='earliest'`` writes, i.e. every
  Keras file of the TensorFlow-1.x era) and version 2 (``OHDR`` / ``OCHK`` blocks);
* groups: symbol-table groups (v1 B-tree + local heap + ``SNOD`` leaves) and compact new-style groups (link messages);
  densely stored links / attributes (fractal heaps) are refused with an explicit error;
* datasets: contiguous, compact and chunked (v1 B-tree chunk index; deflate and shuffle filters) layouts of
  fixed-point, floating-point and fixed-length string types;
* attributes (message versions 1-3) of those types and of variable-length strings (global heap), which is how Keras
  stores ``layer_names`` / ``weight_names`` / ``keras_version``.

Only reading is implemented.  The version-2 paths (superblock 2/3, ``OHDR`` headers, link messages: files written with
``libver='latest'``) are written from the specification but NOT exercised by any test -- the independent writer emits
the classic layout only, which is what Keras files of the reference's era use.  **Pinning:** (1) a file libhdf5 itself wrote -- scipy's MATLAB-7.4 v7.3 fixture (HDF5 1.x behind a 512-byte user block),
committed as tests/golden/testhdf5_7.4_GLNX86.mat: group walk, dataset values (known answer 0:pi/4:2*pi) and string
attribute are read back exactly; (2) an independent minimal writer (tests/h5_writer.py) that emits the classic layout
byte by byte from the same specification, for everything that small file does not contain (multi-level B-trees, chunked
/ deflated data, continuation blocks, variable-length strings, the Keras layout).  A real Keras checkpoint has still not
been available.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class _Buf:
    """Little-endian cursor over the file bytes with the superblock's offset / length sizes."""

    def __init__(self, data: bytes, pos: int = 0, osz: int = 8, lsz: int = 8):
        self.d, self.p, self.osz, self.lsz = data, pos, osz, lsz

    def at(self, pos):
        return _Buf(self.d, pos, self.osz, self.lsz)

    def bytes(self, n):
        if self.p + n > len(self.d):
            raise H5Error(f"read of {n} bytes at {self.p} runs past the end of the file ({len(self.d)} bytes)")
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def skip(self, n):
        self.p += n

    def u8(self):
        return self.bytes(1)[0]

    def u16(self):
        return struct.unpack("<H", self.bytes(2))[0]

    def u32(self):
        return struct.unpack("<I", self.bytes(4))[0]

    def u64(self):
        return struct.unpack("<Q", self.bytes(8))[0]

    def uint(self, n):
        return int.from_bytes(self.bytes(n), "little")

    def off(self):
        v = self.uint(self.osz)
        return UNDEF if v == (1 << (8 * self.osz)) - 1 else v

    def length(self):
        return self.uint(self.lsz)

    def align(self, base, a=8):
        self.p = base + (self.p - base + a - 1) // a * a


# ---------------------------------------------------------------------------------------------- message types
MSG_DATASPACE, MSG_LINK_INFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK = 0x1, 0x2, 0x3, 0x4, 0x5, 0x6
MSG_LAYOUT, MSG_GROUP_INFO, MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION = 0x8, 0xA, 0xB, 0xC, 0x10
MSG_SYMBOL_TABLE, MSG_ATTR_INFO = 0x11, 0x15


class _Datatype:
    """Datatype message (spec IV.A.2.d): class, size and what is needed to build a numpy dtype."""

    def __init__(self, b: _Buf):
        cv = b.u8()
        self.cls, self.version = cv & 0x0F, cv >> 4
        bits = b.bytes(3)
        self.size = b.u32()
        self.bits = bits
        self.base = None
        self.vlen_string = False
        order = ">" if bits[0] & 1 else "<"
        if self.cls == 0:                                   # fixed point: bit offset, precision
            b.skip(4)
            signed = bool(bits[0] & 0x08)
            self.np = np.dtype(f"{order}{'i' if signed else 'u'}{self.size}")
        elif self.cls == 1:                                 # floating point: 12 bytes of layout properties
            b.skip(12)
            if self.size not in (2, 4, 8):
                raise H5Error(f"unsupported floating-point size {self.size}")
            self.np = np.dtype(f"{order}f{self.size}")
        elif self.cls == 3:                                 # fixed-length string (padding in bits 0-3, charset 4-7)
            self.np = np.dtype(f"S{self.size}")
        elif self.cls == 9:                                 # variable length: base type follows
            self.vlen_string = (bits[0] & 0x0F) == 1
            self.base = _Datatype(b)
            self.np = None
        else:
            raise H5Error(f"unsupported datatype class {self.cls} (only integer / float / string / vlen are read)")


def _read_dataspace(b: _Buf):
    """Dataspace message (spec IV.A.2.b) -> shape tuple (() for a scalar, None for a null dataspace)."""
    version = b.u8()
    rank = b.u8()
    flags = b.u8()
    if version == 1:
        b.skip(5)
    elif version == 2:
        kind = b.u8()
        if kind == 2:
            return None
    else:
        raise H5Error(f"dataspace message version {version}")
    dims = tuple(b.length() for _ in range(rank))
    if flags & 1:
        b.skip(rank * b.lsz)
    return dims


class _Object:
    """One object header: its messages, parsed lazily into links / attributes / dataset description."""

    def __init__(self, f: "File", addr: int):
        self.f, self.addr = f, addr
        self.msgs = []                                      # (type, flags, _Buf positioned at the message body, size)
        self._read_header()

    # ---- header blocks
    def _read_header(self):
        b = self.f.buf.at(self.addr)
        if b.d[self.addr:self.addr + 4] == b"OHDR":
            self._read_v2(b)
            return
        version = b.u8()
        if version != 1:
            raise H5Error(f"object header version {version} at {self.addr}")
        b.skip(1)
        n_msgs = b.u16()
        b.skip(4)                                           # reference count
        size = b.u32()
        b.align(self.addr, 8)                               # the first block starts 8-aligned after the 12-byte prefix
        blocks = [(b.p, size)]
        while blocks and len(self.msgs) < n_msgs:
            start, length = blocks.pop(0)
            c = self.f.buf.at(start)
            while c.p + 8 <= start + length and len(self.msgs) < n_msgs:
                mtype, msize, mflags = c.u16(), c.u16(), c.u8()
                c.skip(3)
                body = c.p
                if mtype == MSG_CONTINUATION:
                    cb = self.f.buf.at(body)
                    blocks.append((cb.off(), cb.length()))
                self.msgs.append((mtype, mflags, body, msize))
                c.p = body + msize

    def _read_v2(self, b: _Buf):
        b.skip(4)
        version = b.u8()
        if version != 2:
            raise H5Error(f"OHDR version {version}")
        flags = b.u8()
        if flags & 0x20:
            b.skip(16)                                      # access / modification / change / birth times
        if flags & 0x10:
            b.skip(4)                                       # max compact / min dense attribute counts
        size = b.uint(1 << (flags & 3))
        tracked = bool(flags & 0x04)
        blocks = [(b.p, size)]
        while blocks:
            start, length = blocks.pop(0)
            c = self.f.buf.at(start)
            while c.p + 4 <= start + length:
                mtype, msize, mflags = c.u8(), c.u16(), c.u8()
                if tracked:
                    c.skip(2)
                body = c.p
                if body + msize > start + length:
                    break                                   # gap / checksum region
                if mtype == MSG_CONTINUATION:
                    cb = self.f.buf.at(body)
                    o, l = cb.off(), cb.length()
                    if self.f.buf.d[o:o + 4] != b"OCHK":
                        raise H5Error("continuation block without OCHK signature")
                    blocks.append((o + 4, l - 8))           # minus signature and trailing checksum
                self.msgs.append((mtype, mflags, body, msize))
                c.p = body + msize

    def _bodies(self, mtype):
        return [(self.f.buf.at(body), size, flags) for (t, flags, body, size) in self.msgs if t == mtype]

    # ---- attributes
    def attrs(self) -> dict:
        if any(t == MSG_ATTR_INFO and self._attr_info_is_dense(body) for (t, _, body, _) in self.msgs):
            raise H5Error("attributes are stored densely (fractal heap): not supported by this reader")
        out = {}
        for b, size, flags in self._bodies(MSG_ATTRIBUTE):
            if flags & 0x02:
                raise H5Error("shared attribute messages are not supported")
            start = b.p
            version = b.u8()
            aflags = b.u8()
            nsz, tsz, ssz = b.u16(), b.u16(), b.u16()
            if version == 3:
                b.skip(1)                                   # name character set
            elif version not in (1, 2):
                raise H5Error(f"attribute message version {version}")
            pad = (lambda n: (n + 7) // 8 * 8) if version == 1 else (lambda n: n)
            if version >= 2 and aflags & 3:
                raise H5Error("attributes with shared datatype / dataspace are not supported")
            name = b.bytes(nsz).split(b"\0", 1)[0].decode("utf8")
            b.p = start + 8 + (1 if version == 3 else 0) + pad(nsz)
            t0 = b.p
            dt = _Datatype(b.at(t0))
            s0 = t0 + pad(tsz)
            shape = _read_dataspace(b.at(s0))
            d0 = s0 + pad(ssz)
            out[name] = self.f._decode(dt, shape, b.at(d0))
        return out

    def _attr_info_is_dense(self, body):
        b = self.f.buf.at(body)
        b.skip(1)
        flags = b.u8()
        if flags & 1:
            b.skip(2)
        return b.off() != UNDEF                             # fractal heap address defined = dense storage in use

    # ---- group links
    def links(self) -> dict:
        """name -> object header address."""
        out = {}
        for b, _, _ in self._bodies(MSG_SYMBOL_TABLE):
            btree, heap = b.off(), b.off()
            self.f._walk_group_btree(btree, self.f._local_heap(heap), out)
        for b, _, _ in self._bodies(MSG_LINK_INFO):
            b.skip(1)
            flags = b.u8()
            if flags & 1:
                b.skip(8)
            if b.off() != UNDEF:
                raise H5Error("group with densely stored links (fractal heap): not supported by this reader")
        for b, _, _ in self._bodies(MSG_LINK):
            version, flags = b.u8(), b.u8()
            ltype = b.u8() if flags & 0x08 else 0
            if flags & 0x04:
                b.skip(8)
            if flags & 0x10:
                b.skip(1)
            nlen = b.uint(1 << (flags & 3))
            name = b.bytes(nlen).decode("utf8")
            if ltype == 0:
                out[name] = b.off()
        return out

    def is_group(self):
        return any(t in (MSG_SYMBOL_TABLE, MSG_LINK_INFO, MSG_LINK, MSG_GROUP_INFO) for (t, _, _, _) in self.msgs)

    # ---- dataset
    def read_dataset(self) -> np.ndarray:
        dts, sps, lay = self._bodies(MSG_DATATYPE), self._bodies(MSG_DATASPACE), self._bodies(MSG_LAYOUT)
        if not (dts and sps and lay):
            raise H5Error("object is not a dataset")
        dt = _Datatype(dts[0][0])
        shape = _read_dataspace(sps[0][0])
        if dt.np is None:
            raise H5Error("variable-length datasets are not supported")
        if shape is None:
            return np.zeros((0,), dt.np)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        b = lay[0][0]
        version = b.u8()
        if version in (1, 2):
            rank = b.u8()
            cls = b.u8()
            b.skip(5)
            addr = b.off() if cls != 0 else None
            dims = [b.u32() for _ in range(rank)]
            if cls == 0:
                raw = b.bytes(b.u32())
            elif cls == 1:
                raw = self.f.buf.at(addr).bytes(n * dt.size) if addr != UNDEF else bytes(n * dt.size)
            else:
                raw = self._read_chunked(addr, dims[:-1], shape, dt)
        elif version == 3:
            cls = b.u8()
            if cls == 0:
                raw = b.bytes(b.u16())
            elif cls == 1:
                addr, size = b.off(), b.length()
                raw = self.f.buf.at(addr).bytes(n * dt.size) if addr != UNDEF else bytes(n * dt.size)
            elif cls == 2:
                rank = b.u8()
                addr = b.off()
                dims = [b.u32() for _ in range(rank)]
                raw = self._read_chunked(addr, dims[:-1], shape, dt)
            else:
                raise H5Error(f"data layout class {cls}")
        else:
            raise H5Error(f"data layout message version {version} (chunk indexes of the 1.10 format are not supported)")
        return np.frombuffer(raw, dtype=dt.np, count=n).reshape(shape).astype(dt.np.newbyteorder("="))

    def _filters(self):
        out = []
        for b, _, _ in self._bodies(MSG_FILTERS):
            version = b.u8()
            nf = b.u8()
            if version == 1:
                b.skip(6)
            for _ in range(nf):
                fid = b.u16()
                nlen = b.u16() if (version == 1 or fid >= 256) else 0
                b.skip(2)                                   # flags
                ncd = b.u16()
                if nlen:
                    b.skip((nlen + 7) // 8 * 8 if version == 1 else nlen)
                cd = [b.u32() for _ in range(ncd)]
                if version == 1 and ncd % 2:
                    b.skip(4)
                out.append((fid, cd))
        return out

    def _read_chunked(self, btree, chunk_dims, shape, dt) -> bytes:
        filters = self._filters()
        out = np.zeros(shape, dtype=dt.np)
        if btree == UNDEF:
            return out.tobytes()
        rank = len(shape)
        for offs, size, mask, addr in self.f._walk_chunk_btree(btree, rank):
            raw = self.f.buf.at(addr).bytes(size)
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                              # shuffle: bytes of each element were de-interleaved
                    es = cd[0] if cd else dt.size
                    a = np.frombuffer(raw, np.uint8)
                    ne = len(a) // es
                    raw = a[:ne * es].reshape(es, ne).T.tobytes() + a[ne * es:].tobytes()
                elif fid == 3:                              # fletcher32: checksum appended, not verified
                    raw = raw[:-4]
                else:
                    raise H5Error(f"unsupported filter id {fid}")
            chunk = np.frombuffer(raw, dt.np, count=int(np.prod(chunk_dims))).reshape(chunk_dims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk_dims, shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out.tobytes()


class Group:
    def __init__(self, f: "File", obj: _Object, name: str):
        self._f, self._o, self.name = f, obj, name
        self._links = None

    @property
    def attrs(self) -> dict:
        return self._o.attrs()

    def keys(self):
        if self._links is None:
            self._links = self._o.links()
        return list(self._links.keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            if node._links is None:
                node._links = node._o.links()
            if part not in node._links:
                raise KeyError(f"{path!r}: no member {part!r} in {node.name!r}")
            obj = _Object(node._f, node._links[part])
            child = (node.name.rstrip("/") + "/" + part)
            node = Group(node._f, obj, child) if obj.is_group() else Dataset(node._f, obj, child)
        return node


class Dataset:
    def __init__(self, f: "File", obj: _Object, name: str):
        self._f, self._o, self.name = f, obj, name

    @property
    def attrs(self) -> dict:
        return self._o.attrs()

    def __array__(self, dtype=None, copy=None):
        a = self._o.read_dataset()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, key):
        return self._o.read_dataset()[key]

    @property
    def shape(self):
        return _read_dataspace(self._o._bodies(MSG_DATASPACE)[0][0])


class File(Group):
    """``File(path)`` / ``File(bytes)``: the root group.  ``f['a/b']`` -> Group or Dataset; ``np.asarray(dataset)``;
    ``.attrs`` -> dict of numpy arrays / scalars / str."""

    def __init__(self, src):
        data = src if isinstance(src, (bytes, bytearray, memoryview)) else open(src, "rb").read()
        data = bytes(data)
        base = 0
        while data[base:base + 8] != SIGNATURE:             # the superblock may sit at 0, 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base + 8 > len(data):
                raise H5Error("not an HDF5 file (signature not found)")
        b = _Buf(data, base + 8)
        version = b.u8()
        if version in (0, 1):
            b.skip(4)                                       # free-space, root-entry, reserved, shared-header versions
            osz, lsz = b.u8(), b.u8()
            b.skip(1)
            b.skip(4)                                       # group leaf / internal node K
            b.skip(4)                                       # file consistency flags
            if version == 1:
                b.skip(4)
            b.osz, b.lsz = osz, lsz
            self.base = b.off()
            b.off(); b.off(); b.off()                       # free-space info, end of file, driver info
            b.off()                                         # root entry: link name offset
            root_addr = b.off()
        elif version in (2, 3):
            osz, lsz = b.u8(), b.u8()
            b.skip(1)
            b.osz, b.lsz = osz, lsz
            self.base = b.off()
            b.off(); b.off()                                # superblock extension, end of file
            root_addr = b.off()
        else:
            raise H5Error(f"superblock version {version}")
        if self.base == UNDEF:
            self.base = 0
        if self.base != base:
            raise H5Error(f"superblock at {base} but base address {self.base}: not supported")
        if base:                                            # user block (e.g. MATLAB v7.3 files): addresses are relative
            data = data[base:]                              # to the base address = the superblock's position
        self.buf = _Buf(data, 0, osz, lsz)
        self._heaps = {}
        super().__init__(self, _Object(self, root_addr), "/")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- group internals
    def _local_heap(self, addr):
        if addr not in self._heaps:
            b = self.buf.at(addr)
            if b.bytes(4) != b"HEAP":
                raise H5Error(f"no local heap at {addr}")
            b.skip(4)
            size = b.length()
            b.length()
            self._heaps[addr] = (b.off(), size)
        return self._heaps[addr]

    def _heap_string(self, heap, off):
        seg, size = heap
        d = self.buf.d
        end = d.index(b"\0", seg + off, seg + size)
        return d[seg + off:end].decode("utf8")

    def _walk_group_btree(self, addr, heap, out):
        b = self.buf.at(addr)
        sig = b.bytes(4)
        if sig == b"SNOD":
            b.skip(2)
            n = b.u16()
            for _ in range(n):
                name_off, obj = b.off(), b.off()
                b.skip(4 + 4 + 16)
                out[self._heap_string(heap, name_off)] = obj
            return
        if sig != b"TREE":
            raise H5Error(f"no B-tree node at {addr}")
        ntype, level = b.u8(), b.u8()
        if ntype != 0:
            raise H5Error("group B-tree expected")
        n = b.u16()
        b.off(); b.off()
        b.length()                                          # key 0
        for _ in range(n):
            child = b.off()
            b.length()
            self._walk_group_btree(child, heap, out)

    def _walk_chunk_btree(self, addr, rank):
        b = self.buf.at(addr)
        if b.bytes(4) != b"TREE":
            raise H5Error(f"no B-tree node at {addr}")
        ntype, level = b.u8(), b.u8()
        if ntype != 1:
            raise H5Error("chunk B-tree expected")
        n = b.u16()
        b.off(); b.off()
        for _ in range(n):
            size, mask = b.u32(), b.u32()
            offs = [b.u64() for _ in range(rank + 1)][:rank]
            child = b.off()
            if level == 0:
                yield offs, size, mask, child
            else:
                yield from self._walk_chunk_btree(child, rank)

    # ---- values
    def _global_heap_object(self, addr, index) -> bytes:
        b = self.buf.at(addr)
        if b.bytes(4) != b"GCOL":
            raise H5Error(f"no global heap collection at {addr}")
        b.skip(4)
        end = addr + b.length()
        while b.p + 8 + b.lsz <= end:
            idx = b.u16()
            b.skip(6)
            size = b.length()
            if idx == 0:
                break
            if idx == index:
                return b.bytes(size)
            b.skip((size + 7) // 8 * 8)
        raise H5Error(f"global heap object {index} not found in the collection at {addr}")

    def _decode(self, dt: _Datatype, shape, b: _Buf):
        if shape is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if dt.cls == 9:
            if not (dt.vlen_string or dt.base.cls == 3 or (dt.base.cls == 0 and dt.base.size == 1)):
                raise H5Error("variable-length sequences of non-character types are not supported")
            vals = []
            for _ in range(n):
                ln, addr, idx = b.u32(), b.off(), b.u32()
                vals.append(self._global_heap_object(addr, idx)[:ln] if ln and addr not in (0, UNDEF) else b"")
            if not shape:
                return vals[0].decode("utf8")
            a = np.empty(n, dtype=object)
            a[:] = [v.decode("utf8") for v in vals]
            return a.reshape(shape)
        raw = b.bytes(n * dt.size)
        a = np.frombuffer(raw, dtype=dt.np, count=n).reshape(shape).astype(dt.np.newbyteorder("=") if dt.cls != 3 else dt.np)
        if dt.cls == 3:
            a = np.asarray(np.char.rstrip(a, b"\0")) if a.size else a
        return a[()] if not shape else a


# ---------------------------------------------------------------------------------------------- Keras weight files
def _text(v):
    return v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def _chunked_attr(attrs: dict, name: str):
    """Keras splits attributes above 64 KB into ``name0``, ``name1`` ... (saving.py: save_attributes_to_hdf5_group)."""
    if name in attrs:
        return [_text(x) for x in np.asarray(attrs[name]).reshape(-1)]
    out, i = [], 0
    while f"{name}{i}" in attrs:
        out += [_text(x) for x in np.asarray(attrs[f"{name}{i}"]).reshape(-1)]
        i += 1
    if not out and i == 0:
        raise H5Error(f"attribute {name!r} not found (is this a Keras weight file?)")
    return out


def read_keras_weights(path) -> dict:
    """Keras ``.h5`` (``save_weights`` layout, or a full ``model.save`` file whose weights sit under
    ``model_weights``) -> {layer name: {weight name: float array}} -- the structure tools/h5_to_npz.map_layers takes."""
    f = File(path)
    g = f["model_weights"] if "model_weights" in f.keys() else f
    layers = {}
    for ln in _chunked_attr(g.attrs, "layer_names"):
        lg = g[ln]
        wn = _chunked_attr(lg.attrs, "weight_names") if lg.attrs else []
        if wn:
            layers[ln] = {n: np.asarray(lg[n]) for n in wn}
    return layers
