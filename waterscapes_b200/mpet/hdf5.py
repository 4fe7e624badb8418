"""Mesh / marker ingestion and result output in DOLFIN's HDF5 layout, without h5py or DOLFIN.

The reference's brain runs read their meshes with
``HDF5File(MPI.comm_world, "colin27_coarse_boundaries.h5", "r").read(mesh, "/mesh", False)`` and
``.read(boundaries, "/boundaries")`` (sandbox/brain-simulations/MPET-4networks-colin27.py:28-45) and store fields
with ``HDF5File(...).write(u, "/u", t)`` (:197-199, :257-259).  DOLFIN 2017-2019 writes classic HDF5 files
(superblock version 0, symbol-table groups, contiguous or unfiltered chunked datasets); this module reads that
subset with numpy -- checked against the reference's own ``src/mpet/test/donut2D.h5`` in tests/test_hdf5.py -- and
writes the same subset, so files round-trip through ``HDF5File`` here.

Layout (DOLFIN ``HDF5File``): ``<name>/coordinates`` f8[Nv, gdim], ``<name>/topology`` i8[Nc, d+1] (attribute
``celltype``) for a mesh; ``<name>/topology`` i8[n, k] + ``<name>/values`` [n] for a MeshFunction over entities given
by their vertices; ``<name>/vector_<k>`` (attribute ``timestamp``) for the k-th stored vector of a Function.
Host-side I/O only: nothing here is on the timestep path.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# ============================================================================================== reader
class H5Reader:
    def __init__(self, path):
        self.raw = open(path, "rb").read()
        r = self.raw
        if r[:8] != _SIG:
            raise ValueError("%s is not an HDF5 file" % path)
        if r[8] != 0:
            raise NotImplementedError("HDF5 superblock version %d (DOLFIN writes version 0)" % r[8])
        self.O, self.L = r[13], r[14]
        if (self.O, self.L) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths")
        self.base = self._u(24, 8)
        root = 24 + 4 * self.O
        self.root = self._entry(root)

    # ---- primitives
    def _u(self, off, n):
        return int.from_bytes(self.raw[off:off + n], "little")

    def _entry(self, e):
        d = dict(name_off=self._u(e, 8), ohdr=self._u(e + 8, 8), cache=self._u(e + 16, 4))
        if d["cache"] == 1:
            d["btree"], d["heap"] = self._u(e + 24, 8), self._u(e + 32, 8)
        return d

    def _messages(self, addr):
        """[(type, flags, bytes)] of a version-1 object header, following continuation blocks."""
        r = self.raw
        if r[addr] != 1:
            raise NotImplementedError("object header version %d" % r[addr])
        nmsg, size = self._u(addr + 2, 2), self._u(addr + 8, 4)
        blocks, out = [(addr + 16, addr + 16 + size)], []
        while blocks and len(out) < nmsg:
            p, end = blocks.pop(0)
            while p + 8 <= end and len(out) < nmsg:
                t, s, fl = self._u(p, 2), self._u(p + 2, 2), r[p + 4]
                data = r[p + 8:p + 8 + s]
                if t == 0x10:
                    a, ln = int.from_bytes(data[:8], "little"), int.from_bytes(data[8:16], "little")
                    blocks.append((a, a + ln))
                out.append((t, fl, data))
                p += 8 + s
        return out

    def _group_entries(self, btree, heap):
        r = self.raw
        assert r[heap:heap + 4] == b"HEAP"
        hdata = self._u(heap + 8 + 2 * 8, 8)
        out = {}

        def walk(node):
            assert r[node:node + 4] == b"TREE" and r[node + 4] == 0
            level, used = r[node + 5], self._u(node + 6, 2)
            p = node + 8 + 16
            for _ in range(used):
                child = self._u(p + 8, 8)
                p += 16
                if level > 0:
                    walk(child)
                    continue
                assert r[child:child + 4] == b"SNOD"
                for j in range(self._u(child + 6, 2)):
                    e = self._entry(child + 8 + 40 * j)
                    s0 = hdata + e["name_off"]
                    out[r[s0:r.index(b"\0", s0)].decode()] = e

        walk(btree)
        return out

    def _children(self, entry):
        if entry.get("cache") == 1:
            return self._group_entries(entry["btree"], entry["heap"])
        for t, _, data in self._messages(entry["ohdr"]):
            if t == 0x11:
                return self._group_entries(int.from_bytes(data[:8], "little"), int.from_bytes(data[8:16], "little"))
        return None

    def _find(self, path):
        e = self.root
        for part in [p for p in path.split("/") if p]:
            ch = self._children(e)
            if ch is None or part not in ch:
                raise KeyError(path)
            e = ch[part]
        return e

    # ---- public
    def has(self, path):
        try:
            self._find(path)
            return True
        except KeyError:
            return False

    def keys(self, path="/"):
        ch = self._children(self._find(path))
        return sorted(ch) if ch else []

    @staticmethod
    def _dtype(data):
        cls, ver = data[0] & 0x0F, data[0] >> 4
        size = int.from_bytes(data[4:8], "little")
        if data[1] & 1:
            raise NotImplementedError("big-endian data")
        if cls == 0:
            return np.dtype("<%s%d" % ("i" if data[1] & 0x08 else "u", size))
        if cls == 1:
            return np.dtype("<f%d" % size)
        if cls == 3:
            return np.dtype("S%d" % size)
        raise NotImplementedError("HDF5 datatype class %d" % cls)

    @staticmethod
    def _shape(data):
        ver, rank, flags = data[0], data[1], data[2]
        p = 8 if ver == 1 else 4
        return tuple(int.from_bytes(data[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    def read(self, path):
        msgs = self._messages(self._find(path)["ohdr"])
        shape = dtype = layout = None
        for t, _, data in msgs:
            if t == 0x01:
                shape = self._shape(data)
            elif t == 0x03:
                dtype = self._dtype(data)
            elif t == 0x08:
                layout = data
            elif t == 0x0B:
                raise NotImplementedError("filtered (compressed) datasets")
        if shape is None or dtype is None or layout is None:
            raise KeyError("%s is not a dataset" % path)
        if layout[0] != 3:
            raise NotImplementedError("data layout message version %d" % layout[0])
        count = int(np.prod(shape)) if shape else 1
        if layout[1] == 1:        # contiguous
            addr = int.from_bytes(layout[2:10], "little")
            if addr == _UNDEF:
                return np.zeros(shape, dtype=dtype)
            return np.frombuffer(self.raw, dtype=dtype, count=count, offset=self.base + addr).reshape(shape).copy()
        if layout[1] == 0:        # compact
            n = int.from_bytes(layout[2:4], "little")
            return np.frombuffer(layout[4:4 + n], dtype=dtype, count=count).reshape(shape).copy()
        if layout[1] == 2:        # chunked, no filters: B-tree of chunks
            nd = layout[2]
            bt = int.from_bytes(layout[3:11], "little")
            cdims = [int.from_bytes(layout[11 + 4 * i:15 + 4 * i], "little") for i in range(nd)][:-1]
            out = np.zeros(shape, dtype=dtype)
            self._read_chunks(bt, nd, cdims, out)
            return out
        raise NotImplementedError("layout class %d" % layout[1])

    def _read_chunks(self, node, nd, cdims, out):
        r = self.raw
        assert r[node:node + 4] == b"TREE" and r[node + 4] == 1
        level, used = r[node + 5], self._u(node + 6, 2)
        ksize = 8 + 8 * nd
        p = node + 8 + 16
        for _ in range(used):
            nbytes, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = [self._u(p + 8 + 8 * i, 8) for i in range(nd - 1)]
            child = self._u(p + ksize, 8)
            p += ksize + 8
            if level > 0:
                self._read_chunks(child, nd, cdims, out)
                continue
            if mask:
                raise NotImplementedError("filtered chunk")
            chunk = np.frombuffer(r, dtype=out.dtype, count=int(np.prod(cdims)), offset=self.base + child).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

    def attrs(self, path):
        """Version-1 attribute messages of an object: {name: scalar / array / str}."""
        out = {}
        for t, _, data in self._messages(self._find(path)["ohdr"]):
            if t != 0x0C or data[0] != 1:
                continue
            nsz, tsz, ssz = (int.from_bytes(data[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
            pad = lambda n: (n + 7) & ~7
            p = 8
            name = data[p:p + nsz].split(b"\0")[0].decode()
            p += pad(nsz)
            dt = data[p:p + tsz]
            p += pad(tsz)
            sp = data[p:p + ssz]
            p += pad(ssz)
            try:
                dtype = self._dtype(dt)
            except NotImplementedError:
                continue
            shape = self._shape(sp) if sp[1] else ()
            cnt = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(data[p:p + cnt * dtype.itemsize], dtype=dtype, count=cnt)
            if dtype.kind == "S":
                out[name] = val[0].split(b"\0")[0].decode()
            else:
                out[name] = val.reshape(shape) if shape else val[0]
        return out


# ============================================================================================== writer
class _Node:
    def __init__(self):
        self.children = {}      # name -> _Node (group) or ("data", array, attrs)
        self.attrs = {}


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0x00
        return bytes([0x10, bits, 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize == 8:
        return bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "S":
        return bytes([0x13, 0, 0, 0]) + struct.pack("<I", dt.itemsize)
    raise NotImplementedError(str(dt))


def _space_msg(shape):
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(t, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", t, len(data), flags) + data


def _attr_msg(name, value):
    if isinstance(value, str):
        raw = value.encode() + b"\0"
        dt, sp, payload = _dtype_msg(np.dtype("S%d" % len(raw))), _space_msg(()), raw
    else:
        arr = np.atleast_1d(np.asarray(value))
        arr = arr.astype("<f8" if arr.dtype.kind == "f" else "<i8")
        shape = () if np.ndim(value) == 0 else arr.shape
        dt, sp, payload = _dtype_msg(arr.dtype), _space_msg(shape), arr.tobytes()
    nm = name.encode() + b"\0"
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + payload
    return _msg(0x0C, body)


def _ohdr(msgs):
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def write_h5(path, tree):
    """Write ``tree`` = {"/a/b": array | (array, {attr: value})} as a classic (superblock 0) HDF5 file with
    symbol-table groups and contiguous datasets."""
    root = _Node()
    for name, val in tree.items():
        arr, attrs = val if isinstance(val, tuple) else (val, {})
        parts = [p for p in name.split("/") if p]
        g = root
        for p in parts[:-1]:
            g = g.children.setdefault(p, _Node())
        g.children[parts[-1]] = ("data", np.ascontiguousarray(arr), dict(attrs))
    K = 64                                   # group leaf node K: up to 2K entries per group (one symbol node)
    buf = bytearray(96)                      # superblock + root symbol-table entry, filled in at the end

    def alloc(data):
        off = len(buf)
        buf.extend(data)
        buf.extend(b"\0" * (-len(buf) % 8))
        return off

    def emit_dataset(arr, attrs):
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8")
        elif arr.dtype.kind in "iub":
            arr = arr.astype("<i8" if arr.dtype.kind != "u" else "<u8")
        addr = alloc(arr.tobytes()) if arr.size else _UNDEF
        msgs = [_msg(0x01, _space_msg(arr.shape)), _msg(0x03, _dtype_msg(arr.dtype), 1),
                _msg(0x05, bytes([2, 2, 2, 1, 0, 0, 0, 0]), 1),
                _msg(0x08, bytes([3, 1]) + struct.pack("<QQ", addr, arr.nbytes), 1)]
        msgs += [_attr_msg(k, v) for k, v in attrs.items()]
        return alloc(_ohdr(msgs))

    def emit_group(node):
        names = sorted(node.children)
        if len(names) > 2 * K:
            raise NotImplementedError("more than %d entries in one group" % (2 * K))
        entries = []
        for nm in names:
            ch = node.children[nm]
            if isinstance(ch, _Node):
                oh, bt, hp = emit_group(ch)
                entries.append((nm, oh, 1, bt, hp))
            else:
                entries.append((nm, emit_dataset(ch[1], ch[2]), 0, 0, 0))
        heap_data = bytearray(8)             # offset 0: the empty string
        offs = []
        for nm, *_ in entries:
            offs.append(len(heap_data))
            heap_data.extend(_pad8(nm.encode() + b"\0"))
        hd = alloc(bytes(heap_data))
        heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), _UNDEF, hd))
        snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(entries)))
        for (nm, oh, cache, bt, hp), no in zip(entries, offs):
            snod += struct.pack("<QQI4xQQ", no, oh, cache, bt, hp)
        snod += b"\0" * (8 + 2 * K * 40 - len(snod))
        sn = alloc(bytes(snod))
        node_size = 8 + 16 + (2 * 16 + 1) * 8 + 2 * 16 * 8        # internal K = 16
        tree_ = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, _UNDEF, _UNDEF))
        tree_ += struct.pack("<QQQ", 0, sn, offs[-1] if offs else 0)
        tree_ += b"\0" * (node_size - len(tree_))
        bt = alloc(bytes(tree_))
        oh = alloc(_ohdr([_msg(0x11, struct.pack("<QQ", bt, heap))]))
        return oh, bt, heap

    oh, bt, hp = emit_group(root)
    head = _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", K, 16, 0)
    head += struct.pack("<QQQQ", 0, _UNDEF, len(buf), _UNDEF)
    head += struct.pack("<QQI4xQQ", 0, oh, 1, bt, hp)
    buf[:len(head)] = head
    with open(path, "wb") as f:
        f.write(bytes(buf))


# ============================================================================================== DOLFIN-style facade
class HDF5File:
    """``HDF5File(comm, filename, mode)`` with the calls the reference makes: ``read(mesh, "/mesh", False)``,
    ``read(meshfunction, "/boundaries")``, ``write(mesh | meshfunction | function, name[, t])``, ``close()``."""

    def __init__(self, comm, filename, mode="r"):
        self.filename, self.mode = filename, mode
        self._r = H5Reader(filename) if mode == "r" else None
        self._tree = {}
        self._counts = {}

    # ---- reading
    def read(self, obj, name, use_partition_from_file=False):
        from .dolfin_shim import Mesh, MeshFunction
        if isinstance(obj, Mesh):
            coords = self._r.read(name + "/coordinates")
            topo = self._r.read(name + "/topology").astype(np.int64)
            if coords.shape[1] != 3 or topo.shape[1] != 4:
                raise ValueError("the B200 path needs a tetrahedral mesh (got %s cells in %d-D)"
                                 % (self._r.attrs(name + "/topology").get("celltype", "?"), coords.shape[1]))
            Mesh.__init__(obj, coords, topo)
            return obj
        if isinstance(obj, MeshFunction):
            topo = np.sort(self._r.read(name + "/topology").astype(np.int64), axis=1)
            vals = self._r.read(name + "/values").astype(np.int64).ravel()
            F = obj.mesh.exterior_facets()["vertices"]
            if topo.shape[1] != 3:
                raise ValueError("only facet functions are used by the mpet API")
            # match stored facets to the mesh's exterior facets by their (sorted) vertex triples
            both = np.vstack([topo, F.astype(np.int64)])
            _, inv = np.unique(both, axis=0, return_inverse=True)
            inv = inv.ravel()
            value_of = np.full(inv.max() + 1, -1, dtype=np.int64)
            value_of[inv[:topo.shape[0]]] = np.arange(topo.shape[0])
            src = value_of[inv[topo.shape[0]:]]
            found = src >= 0
            obj.array_[found] = vals[src[found]]
            return obj
        raise TypeError("cannot read a %s" % type(obj).__name__)

    def read_mesh(self, name="/mesh"):
        from .dolfin_shim import Mesh
        m = Mesh.__new__(Mesh)
        return self.read(m, name, False)

    # ---- writing
    def write(self, obj, name, t=None):
        from .dolfin_shim import Mesh, MeshFunction, Function, SubFunction
        assert self.mode == "w"
        if isinstance(obj, Mesh):
            self._tree[name + "/coordinates"] = obj.coordinates
            self._tree[name + "/topology"] = (obj.cells.astype(np.int64), {"celltype": "tetrahedron"})
        elif isinstance(obj, MeshFunction):
            F = obj.mesh.exterior_facets()["vertices"]
            self._tree[name + "/topology"] = (F.astype(np.int64), {"celltype": "triangle"})
            self._tree[name + "/values"] = obj.array().astype(np.int64)
        elif isinstance(obj, (Function, SubFunction)):
            vec = obj.x.detach().cpu().numpy() if isinstance(obj, Function) else np.asarray(obj.values, dtype=float).T.ravel()
            k = self._counts.get(name, 0)
            self._counts[name] = k + 1
            self._tree["%s/vector_%d" % (name, k)] = (vec, {} if t is None else {"timestamp": float(t)})
        else:
            raise TypeError("cannot write a %s" % type(obj).__name__)

    def close(self):
        if self.mode == "w" and self._tree:
            write_h5(self.filename, self._tree)
            self._tree = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
