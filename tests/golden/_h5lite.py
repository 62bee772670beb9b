"""Minimal read-only HDF5 parser (test infrastructure, fixture generation only).

h5py is not installed in this image, but the reference ships its meshes
(``smash/dataset/*/mesh_*.hdf5``) and its golden file (``smash/tests/baseline.hdf5``)
as classic-format HDF5: superblock v0, symbol-table groups (TREE/HEAP/SNOD),
object headers v1, contiguous / compact / gzip-chunked datasets.  This reader
understands exactly that subset.  It is used only by ``make_golden.py``.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = f.read()
        if self.b[:8] != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        ver = self.b[8]
        if ver != 0:
            raise ValueError(f"superblock version {ver} not supported")
        self.so = self.b[13]
        self.sl = self.b[14]
        if self.so != 8 or self.sl != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        # 8 sig + 8 versions/sizes + 2+2 K + 4 flags = 24 ; then base, freespace, eof, driver
        p = 24 + 4 * 8
        # root symbol table entry
        _, root_hdr, _cache, _ = struct.unpack_from("<QQII", self.b, p)
        self.root = self._read_object(root_hdr)

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr: int):
        b = self.b
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError(f"object header v{ver} not supported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                data = p + 8
                if mtype == 0x10:  # continuation
                    coff, clen = struct.unpack_from("<QQ", b, data)
                    blocks.append((coff, clen))
                out.append((mtype, data, msize, mflags))
                p = data + msize
        return out

    def _read_object(self, addr: int):
        msgs = self._messages(addr)
        obj = {"attrs": {}, "addr": addr}
        for mtype, p, size, _ in msgs:
            if mtype == 0x11:  # symbol table => group
                btree, heap = struct.unpack_from("<QQ", self.b, p)
                obj["group"] = (btree, heap)
            elif mtype == 0x01:
                obj["shape"] = self._dataspace(p)
            elif mtype == 0x03:
                obj["dtype"] = self._datatype(p)
            elif mtype == 0x08:
                obj["layout"] = self._layout(p)
            elif mtype == 0x0B:
                obj["filters"] = self._filters(p)
            elif mtype == 0x0C:
                name, val = self._attribute(p)
                obj["attrs"][name] = val
        return obj

    def _dataspace(self, p):
        b = self.b
        ver, rank, flags = struct.unpack_from("<BBB", b, p)
        q = p + (8 if ver == 1 else 4)
        return tuple(struct.unpack_from("<" + "Q" * rank, b, q)) if rank else ()

    def _datatype(self, p):
        b = self.b
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, p)
        cls = cv & 0x0F
        if cls == 0:  # fixed point
            signed = (b0 >> 3) & 1
            order = ">" if (b0 & 1) else "<"
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:
            order = ">" if (b0 & 1) else "<"
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9:  # variable length (only used by the _save_func attribute)
            return "vlen"
        raise ValueError(f"datatype class {cls} not supported")

    def _layout(self, p):
        b = self.b
        ver, cls = struct.unpack_from("<BB", b, p)
        if ver != 3:
            raise ValueError(f"layout v{ver} not supported")
        if cls == 0:
            (size,) = struct.unpack_from("<H", b, p + 2)
            return ("compact", p + 4, size)
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", b, p + 2)
            return ("contiguous", addr, size)
        ndim = b[p + 2]
        (btree,) = struct.unpack_from("<Q", b, p + 3)
        dims = struct.unpack_from("<" + "I" * ndim, b, p + 11)
        return ("chunked", btree, dims)

    def _filters(self, p):
        b = self.b
        ver, nf = struct.unpack_from("<BB", b, p)
        if ver != 1:
            raise ValueError("filter pipeline v2 not supported")
        q = p + 8
        out = []
        for _ in range(nf):
            fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", b, q)
            q += 8 + ((nlen + 7) // 8) * 8
            cd = struct.unpack_from("<" + "I" * ncd, b, q)
            q += 4 * ncd + (4 if ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def _attribute(self, p):
        b = self.b
        ver = b[p]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            q = p + 8
            pad = lambda n: ((n + 7) // 8) * 8
            name = b[q : q + nsz].split(b"\0")[0].decode()
            q += pad(nsz)
            dt = self._datatype(q)
            q += pad(tsz)
            shape = self._dataspace(q)
            q += pad(ssz)
        else:
            raise ValueError(f"attribute v{ver} not supported")
        if isinstance(dt, str):
            return name, None
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(b, dtype=dt, count=n, offset=q).reshape(shape)
        if dt.kind == "S":
            val = arr.astype("U")
            return name, (val.item() if not shape else val)
        return name, (arr.item() if not shape else arr.copy())

    # ------------------------------------------------------------------ groups
    def _heap_str(self, heap_addr, off):
        b = self.b
        assert b[heap_addr : heap_addr + 4] == b"HEAP"
        (dseg,) = struct.unpack_from("<Q", b, heap_addr + 24)
        s = dseg + off
        e = b.index(b"\0", s)
        return b[s:e].decode()

    def _group_entries(self, btree, heap):
        b = self.b
        out = {}

        def walk(addr):
            assert b[addr : addr + 4] == b"TREE", "bad group b-tree node"
            ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
            assert ntype == 0
            p = addr + 24  # after siblings
            p += 8  # key 0
            for _ in range(used):
                (child,) = struct.unpack_from("<Q", b, p)
                p += 16  # child + next key
                if level > 0:
                    walk(child)
                else:
                    assert b[child : child + 4] == b"SNOD"
                    (nsym,) = struct.unpack_from("<H", b, child + 6)
                    q = child + 8
                    for _ in range(nsym):
                        noff, ohdr = struct.unpack_from("<QQ", b, q)
                        out[self._heap_str(heap, noff)] = ohdr
                        q += 40

        walk(btree)
        return out

    def keys(self, obj=None):
        obj = obj or self.root
        return sorted(self._group_entries(*obj["group"]).keys())

    @property
    def attrs(self):
        return self.root["attrs"]

    def __getitem__(self, name: str) -> np.ndarray:
        obj = self.root
        for part in name.strip("/").split("/"):
            entries = self._group_entries(*obj["group"])
            obj = self._read_object(entries[part])
        return self._read_dataset(obj)

    # ------------------------------------------------------------------ datasets
    def _read_dataset(self, obj) -> np.ndarray:
        b = self.b
        shape, dt, layout = obj["shape"], obj["dtype"], obj["layout"]
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            return np.frombuffer(b, dt, n, layout[1]).reshape(shape).copy()
        if layout[0] == "contiguous":
            if layout[1] == _UNDEF:
                return np.zeros(shape, dt)
            return np.frombuffer(b, dt, n, layout[1]).reshape(shape).copy()
        _, btree, cdims = layout
        cshape = tuple(cdims[:-1])
        out = np.zeros(shape, dt)
        filters = obj.get("filters", [])
        rank = len(shape)

        def walk(addr):
            assert b[addr : addr + 4] == b"TREE", "bad chunk b-tree node"
            ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
            assert ntype == 1
            p = addr + 24
            ksz = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, _mask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<" + "Q" * rank, b, p + 8)
                (child,) = struct.unpack_from("<Q", b, p + ksz)
                p += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                raw = b[child : child + csize]
                for fid, _cd in reversed(filters):
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:  # shuffle
                        a = np.frombuffer(raw, np.uint8).reshape(dt.itemsize, -1)
                        raw = a.T.tobytes()
                    else:
                        raise ValueError(f"filter {fid} not supported")
                chunk = np.frombuffer(raw, dt, int(np.prod(cshape))).reshape(cshape)
                sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                out[sl_out] = chunk[sl_in]

        if btree != _UNDEF:
            walk(btree)
        return out
