"""Minimal read-only HDF5 parser (no libhdf5 / h5py in this image): version-0 superblock, version-1 object headers,
symbol-table groups (v1 B-trees + local heaps), contiguous / compact / unfiltered chunked datasets of fixed-point, floating
point and fixed-length string types.  Enough for the ESHDF orbital files under the reference's tests/solids/ (written by
pw2qmcpack with the 1.8 file format).  Test infrastructure only: used by scripts/gen_diamondC_golden.py to turn the
reference's own DFT file into a committed fixture."""
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
