"""TEST INFRASTRUCTURE: a strict, hand-written reader of the subset of HDF5 the product's writer emits (no h5py in this image).

It walks the file the way libhdf5 does — superblock v0 -> root symbol table entry -> local heap + B-tree v1 -> symbol table nodes ->
object headers v1 -> Dataspace / Datatype / Fill Value / Data Layout messages -> raw data — and checks every signature, version and
size on the way ("HDF5 File Format Specification Version 1.1").  `read_xmf` restates how pymcac reads the light data
(/root/reference/pymcac/reader/xdmf_reader.py:58-112): metadata from <Information>, per-<Grid> time, Geometry / Attribute -> dataset."""
from __future__ import annotations

import struct
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(AssertionError):
    pass


def _need(cond, msg):
    if not cond:
        raise H5Error(msg)


class H5Min:
    def __init__(self, path):
        self.b = Path(path).read_bytes()
        b = self.b
        _need(b[:8] == b"\x89HDF\r\n\x1a\n", "signature")
        ver_sb, ver_fs, ver_root, _, ver_sh, so, sl, _ = struct.unpack_from("<8B", b, 8)
        _need((ver_sb, ver_fs, ver_root, ver_sh) == (0, 0, 0, 0), "superblock / free-space / root / shared-header versions must be 0")
        _need((so, sl) == (8, 8), "8-byte offsets and lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        _need(self.leaf_k >= 1 and self.internal_k >= 1, "node ranks")
        (flags,) = struct.unpack_from("<I", b, 20)
        _need(flags == 0, "file consistency flags")
        base, free, eof, driver = struct.unpack_from("<4Q", b, 24)
        _need(base == 0 and free == UNDEF and driver == UNDEF, "base / free-space / driver addresses")
        _need(eof == len(b), f"end-of-file address {eof} != file size {len(b)}")
        name_off, hdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
        _need(name_off == 0 and cache == 1, "root symbol table entry: cached symbol table")
        self.root_header = hdr
        # the root object header must carry the same B-tree / heap addresses in its Symbol Table message
        msgs = self._object_header(hdr)
        _need(len(msgs) == 1 and msgs[0][0] == 0x0011, "root header: exactly one Symbol Table message")
        _need(struct.unpack_from("<QQ", msgs[0][1]) == (btree, heap), "root header vs cached addresses")
        self.heap_data = self._heap(heap)
        self.datasets = {}
        for name, addr in self._btree(btree):
            self.datasets[name] = addr

    def _object_header(self, addr):
        b = self.b
        _need(addr % 8 == 0 and addr + 16 <= len(b), "object header address")
        ver, _, nmsg, refcount, size = struct.unpack_from("<BBHII", b, addr)
        _need(ver == 1 and refcount == 1, "object header version 1, one link")
        p, end, out = addr + 16, addr + 16 + size, []
        _need(end <= len(b), "object header size")
        while p < end:
            mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
            _need(msize % 8 == 0, "message data is 8-byte padded")
            out.append((mtype, b[p + 8:p + 8 + msize]))
            p += 8 + msize
        _need(p == end and len(out) == nmsg, "message count / total size")
        return out

    def _heap(self, addr):
        b = self.b
        _need(b[addr:addr + 4] == b"HEAP" and b[addr + 4] == 0, "local heap signature / version")
        size, free, data = struct.unpack_from("<QQQ", b, addr + 8)
        _need(free == 1, "no free block (H5HL_FREE_NULL)")
        _need(data + size <= len(b) and size % 8 == 0, "heap data segment")
        return b[data:data + size]

    def _name(self, off):
        end = self.heap_data.index(b"\0", off)
        return self.heap_data[off:end].decode()

    def _btree(self, addr):
        b = self.b
        _need(b[addr:addr + 4] == b"TREE", "B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        _need(ntype == 0 and level == 0, "group B-tree leaf level")
        left, right = struct.unpack_from("<QQ", b, addr + 8)
        _need(left == UNDEF and right == UNDEF, "no siblings")
        _need(used <= 2 * self.internal_k, "entries used vs internal node rank")
        _need(addr + 24 + (4 * self.internal_k + 1) * 8 <= len(b), "B-tree node is allocated in full")
        p = addr + 24
        (key,) = struct.unpack_from("<Q", b, p)
        _need(self._name(key) == "", "first key is the empty name")
        prev = ""
        for _ in range(used):
            child, key = struct.unpack_from("<QQ", b, p + 8)
            p += 16
            names = list(self._snod(child))
            _need(names and names[-1][0] == self._name(key), "right key = greatest name of the child")
            _need(names[0][0] > prev, "children ordered")
            prev = names[-1][0]
            yield from names

    def _snod(self, addr):
        b = self.b
        _need(b[addr:addr + 4] == b"SNOD" and b[addr + 4] == 1, "symbol table node signature / version")
        (n,) = struct.unpack_from("<H", b, addr + 6)
        _need(1 <= n <= 2 * self.leaf_k, "symbols vs leaf node rank")
        _need(addr + 8 + 2 * self.leaf_k * 40 <= len(b), "symbol table node is allocated in full")
        prev = None
        for e in range(n):
            name_off, hdr, cache = struct.unpack_from("<QQI", b, addr + 8 + 40 * e)
            _need(cache == 0, "dataset entries cache nothing")
            name = self._name(name_off)
            _need(prev is None or name > prev, "entries sorted by name (strcmp)")
            prev = name
            yield name, hdr

    def dataset(self, name) -> np.ndarray:
        msgs = dict()
        for t, d in self._object_header(self.datasets[name]):
            _need(t not in msgs, "duplicate message")
            msgs[t] = d
        _need(set(msgs) == {0x0001, 0x0003, 0x0005, 0x0008}, f"dataset messages {sorted(msgs)}")
        sp = msgs[0x0001]
        _need(sp[0] == 1 and sp[1] == 1 and sp[2] == 0, "dataspace v1, rank 1, no max dims")
        (count,) = struct.unpack_from("<Q", sp, 8)
        dt = msgs[0x0003]
        cls, ver = dt[0] & 0x0F, dt[0] >> 4
        (size,) = struct.unpack_from("<I", dt, 4)
        _need(ver == 1, "datatype v1")
        if cls == 1:
            _need(dt[1:4] == bytes([0x20, 0x3F, 0x00]) and size == 8, "IEEE little-endian double: bit field")
            _need(struct.unpack_from("<HHBBBBI", dt, 8) == (0, 64, 52, 11, 0, 52, 1023), "IEEE double properties")
            dtype = np.dtype("<f8")
        else:
            _need(cls == 0 and dt[1:4] == bytes([0x08, 0, 0]) and size in (4, 8), "signed little-endian integer")
            _need(struct.unpack_from("<HH", dt, 8) == (0, 8 * size), "integer precision")
            dtype = np.dtype("<i4" if size == 4 else "<i8")
        fv = msgs[0x0005]
        _need(fv[0] == 2 and fv[3] == 1 and struct.unpack_from("<I", fv, 4) == (0,), "fill value v2, default")
        lay = msgs[0x0008]
        _need(lay[0] == 3 and lay[1] == 1, "data layout v3, contiguous")
        addr, nbytes = struct.unpack_from("<QQ", lay, 2)
        _need(nbytes == count * dtype.itemsize, "layout size = elements x element size")
        _need(addr % 8 == 0 and addr + nbytes <= len(self.b), "raw data inside the file")
        return np.frombuffer(self.b, dtype=dtype, count=count, offset=addr).copy()


def read_xmf(path):
    """(metadata, [(time, {name: (h5 file, dataset, dimensions)})]) the way pymcac's XdmfReader extracts them."""
    root = ET.parse(path).getroot()
    _need(root.tag == "Xdmf" and root.find("Domain") is not None, "Xdmf / Domain")
    metadata = {}
    for el in root.iter("Information"):
        if el.get("Name") in {"Copyright", "Physics"}:
            continue
        metadata[el.get("Name")] = float(el.get("Value"))
    steps = []
    for grid in root.iter("Grid"):
        if grid.get("Name") == "Collection":
            _need(grid.get("GridType") == "Collection" and grid.get("CollectionType") == "Temporal", "temporal collection")
            continue
        t = float(grid.find("Time").get("Value"))
        items = {}
        geo = list(grid.find("Geometry"))[0]
        _need(grid.find("Geometry").get("Type") == "XYZ" and grid.find("Topology").get("Type") == "Polyvertex", "XYZ polyvertex")
        items["Positions"] = (*geo.text.split(":"), int(geo.get("Dimensions")))
        for at in grid.findall("Attribute"):
            di = list(at)[0]
            _need(di.get("Format") == "HDF", "heavy data in HDF5")
            items[at.get("Name")] = (*di.text.split(":"), int(di.get("Dimensions")))
        steps.append((t, items))
    return metadata, steps
