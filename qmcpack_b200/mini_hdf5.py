"""Minimal read-only HDF5 parser (no libhdf5 / h5py in this image): version-0 superblock, version-1 object headers,
symbol-table groups (v1 B-trees + local heaps), contiguous / compact / unfiltered chunked datasets of fixed-point, floating
point and fixed-length string types.  Enough for the ESHDF orbital files under the reference's tests/solids/ (written by
pw2qmcpack with the 1.8 file format) and for the spline-coefficient dumps QMCPACK writes with save_coefs="yes"
(qmcpack_b200/spline_dump.py).  `write_h5` emits the same subset (one flat group of contiguous datasets) so that the dump
reader can be exercised without libhdf5."""
import struct

import numpy as np


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "only superblock version 0 is supported"
        assert b[13] == 8 and b[14] == 8, "8-byte offsets and lengths expected"
        # 24: base address, free-space address, end-of-file address, driver info address, then the root symbol table entry
        root_entry = 24 + 32
        self.root = self._entry(root_entry)

    # symbol table entry: name offset, object header address, cache type, reserved, scratch
    def _entry(self, off):
        name_off, ohdr, cache = struct.unpack_from("<QQI", self.b, off)
        btree = heap = None
        if cache == 1:
            btree, heap = struct.unpack_from("<QQ", self.b, off + 24)
        return dict(name_off=name_off, ohdr=ohdr, btree=btree, heap=heap)

    def _messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1, "only version-1 object headers are supported"
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", b, p)
                data = p + 8
                if mtype == 0x10:  # continuation
                    o, l = struct.unpack_from("<QQ", b, data)
                    blocks.append((o, l))
                out.append((mtype, data, msize))
                p = data + ((msize + 7) & ~7)
        return out

    def _group_tables(self, ent):
        if ent["btree"] is not None:
            return ent["btree"], ent["heap"]
        for mtype, data, _ in self._messages(ent["ohdr"]):
            if mtype == 0x11:
                return struct.unpack_from("<QQ", self.b, data)
        return None

    def _heap_name(self, heap, off):
        assert self.b[heap:heap + 4] == b"HEAP"
        seg = struct.unpack_from("<Q", self.b, heap + 24)[0]
        e = self.b.index(b"\0", seg + off)
        return self.b[seg + off:e].decode()

    def _walk_btree(self, node, heap, out):
        b = self.b
        assert b[node:node + 4] == b"TREE", "bad B-tree node"
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        assert ntype == 0
        p = node + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, p + 8)[0]  # key (8) then child (8)
            p += 16
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                assert b[child:child + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                for s in range(nsym):
                    e = self._entry(child + 8 + 40 * s)
                    out[self._heap_name(heap, e["name_off"])] = e

    def listdir(self, path="/"):
        ent = self._resolve(path)
        t = self._group_tables(ent)
        if t is None:
            raise KeyError(path + " is not a group")
        out = {}
        self._walk_btree(t[0], t[1], out)
        return out

    def _resolve(self, path):
        ent = self.root
        for part in [p for p in path.split("/") if p]:
            t = self._group_tables(ent)
            if t is None:
                raise KeyError(path)
            out = {}
            self._walk_btree(t[0], t[1], out)
            if part not in out:
                raise KeyError(path)
            ent = out[part]
        return ent

    def exists(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def read(self, path):
        b = self.b
        ent = self._resolve(path)
        dims, dtype, layout = (), None, None
        for mtype, data, msize in self._messages(ent["ohdr"]):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", b, data)
                base = data + (8 if ver == 1 else 4)
                dims = struct.unpack_from("<%dQ" % rank, b, base) if rank else ()
            elif mtype == 0x03:  # datatype
                cv, bf0, _, _, size = struct.unpack_from("<BBBBI", b, data)
                cls = cv & 15
                if cls == 0:
                    dtype = np.dtype(("<" if not bf0 & 1 else ">") + ("i" if bf0 & 8 else "u") + str(size))
                elif cls == 1:
                    dtype = np.dtype(("<" if not bf0 & 1 else ">") + "f" + str(size))
                elif cls == 3:
                    dtype = np.dtype("S%d" % size)
                else:
                    raise NotImplementedError("datatype class %d at %s" % (cls, path))
            elif mtype == 0x08:  # data layout
                ver = b[data]
                if ver == 3:
                    cls = b[data + 1]
                    if cls == 0:
                        n = struct.unpack_from("<H", b, data + 2)[0]
                        layout = ("compact", data + 4, n)
                    elif cls == 1:
                        a, n = struct.unpack_from("<QQ", b, data + 2)
                        layout = ("contiguous", a, n)
                    else:
                        nd = b[data + 2]
                        bt = struct.unpack_from("<Q", b, data + 3)[0]
                        cd = struct.unpack_from("<%dI" % nd, b, data + 11)
                        layout = ("chunked", bt, cd)
                else:
                    nd, cls = b[data + 1], b[data + 2]
                    p = data + 8
                    a = None
                    if cls != 0:
                        a = struct.unpack_from("<Q", b, p)[0]
                        p += 8
                    d = struct.unpack_from("<%dI" % nd, b, p)
                    p += 4 * nd
                    if cls == 1:
                        layout = ("contiguous", a, None)
                    elif cls == 2:
                        layout = ("chunked", a, d)
                    else:
                        n = struct.unpack_from("<I", b, p)[0]
                        layout = ("compact", p + 4, n)
        if dtype is None or layout is None:
            raise KeyError(path + " is not a dataset")
        count = int(np.prod(dims)) if dims else 1
        nbytes = count * dtype.itemsize
        if layout[0] in ("contiguous", "compact"):
            if layout[1] == 0xFFFFFFFFFFFFFFFF:
                return np.zeros(dims, dtype)
            arr = np.frombuffer(b, dtype, count, layout[1])
        else:
            arr = np.zeros(count, dtype).reshape(dims)
            self._read_chunks(layout[1], layout[2], arr, dtype)
            return arr
        return arr.reshape(dims).copy()

    def _read_chunks(self, node, cdims, arr, dtype):
        b = self.b
        assert b[node:node + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        assert ntype == 1
        nd = len(cdims)  # includes the trailing element-size dimension
        keysize = 8 + 8 * nd
        p = node + 24
        for i in range(used):
            csize, fmask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % nd, b, p + 8)
            child = struct.unpack_from("<Q", b, p + keysize)[0]
            p += keysize + 8
            if level > 0:
                self._read_chunks(child, cdims, arr, dtype)
                continue
            assert fmask == 0, "filtered chunks are not supported"
            shape = tuple(cdims[:-1])
            chunk = np.frombuffer(b, dtype, int(np.prod(shape)), child).reshape(shape)
            sl = tuple(slice(o, min(o + s, a)) for o, s, a in zip(offs[:-1], shape, arr.shape))
            arr[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]


def write_h5(path, datasets):
    """Write {name: ndarray | bytes | int} as contiguous datasets of the root group in the format subset H5File reads
    (version-0 superblock, version-1 object headers, one symbol-table node: at most 8 datasets)."""
    names = sorted(datasets)
    assert 0 < len(names) <= 8, "one leaf node of the default group B-tree holds 8 symbols"
    UNDEF = 0xFFFFFFFFFFFFFFFF
    out = bytearray(96)  # superblock (56) + root symbol table entry (40)

    def align():
        out.extend(b"\0" * (-len(out) % 8))

    def message(mtype, body):
        body = body + b"\0" * (-len(body) % 8)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def object_header(msgs):
        align()
        addr = len(out)
        blob = b"".join(msgs)
        out.extend(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(blob)) + blob)
        return addr

    # local heap: offset 0 = the empty name of the root group
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data.extend(n.encode() + b"\0")
        heap_data.extend(b"\0" * (-len(heap_data) % 8))
    entries = []
    for n in names:
        v = datasets[n]
        if isinstance(v, (bytes, str)):
            v = np.array(v.encode() if isinstance(v, str) else v, dtype="S%d" % max(1, len(v)))
        a = np.ascontiguousarray(v)
        if a.dtype.kind == "S":
            dt = struct.pack("<BBBBI", 0x13, 0, 0, 0, a.dtype.itemsize)
        elif a.dtype.kind == "f":
            sz = a.dtype.itemsize
            prop = (struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127) if sz == 4 else
                    struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023))
            dt = struct.pack("<BBBBI", 0x11, 0x20, 8 * sz - 1, 0, sz) + prop
        elif a.dtype.kind in "iu":
            sz = a.dtype.itemsize
            dt = struct.pack("<BBBBI", 0x10, 0x08 if a.dtype.kind == "i" else 0, 0, 0, sz) + struct.pack("<HH", 0, 8 * sz)
        else:
            raise NotImplementedError(str(a.dtype))
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + struct.pack("<%dQ" % a.ndim, *a.shape)
        align()
        data_addr = len(out)
        out.extend(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
        layout = struct.pack("<BBQQ", 3, 1, data_addr, a.nbytes)
        ohdr = object_header([message(0x01, space), message(0x03, dt), message(0x08, layout)])
        entries.append(struct.pack("<QQII16x", name_off[n], ohdr, 0, 0))
    align()
    heap_seg = len(out)
    out.extend(heap_data)
    heap = len(out)
    out.extend(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_seg))
    snod = len(out)
    out.extend(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)) + b"".join(entries) + b"\0" * (40 * (8 - len(entries))))
    btree = len(out)
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, name_off[names[-1]])
    out.extend(node + b"\0" * (24 + 33 * 8 + 32 * 8 - len(node)))
    root = object_header([message(0x11, struct.pack("<QQ", btree, heap))])
    out[0:56] = (b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) +
                 struct.pack("<QQQQ", 0, UNDEF, len(out), UNDEF))
    out[56:96] = struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
    with open(path, "wb") as f:
        f.write(bytes(out))
