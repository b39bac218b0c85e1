"""Minimal reader for classic-format HDF5 files holding contiguous little-endian numeric datasets
(TEST INFRASTRUCTURE).  h5py / libhdf5 are not in this image; the reference's golden
vectors (src/tests/data/*.h5, read there through cv::hdf, include/orcvio/utils/
se3_ops.hpp:462-474) use superblock v0, v1 object headers, a symbol-table root group and
contiguous little-endian IEEE doubles, which is all this reader understands.
"""
import struct
import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class _File:
    def __init__(self, data):
        self.d = data
        assert data[:8] == _SIG, "not an HDF5 file"
        ver = data[8]
        assert ver == 0, f"superblock version {ver} unsupported"
        self.O = data[13]
        self.L = data[14]
        p = 24
        p += 4 * self.O      # base, free-space, eof, driver
        self.root_entry = p

    def off(self, p):
        return int.from_bytes(self.d[p:p + self.O], "little")

    def length(self, p):
        return int.from_bytes(self.d[p:p + self.L], "little")

    def entry(self, p):
        name_off = self.off(p)
        hdr = self.off(p + self.O)
        cache = struct.unpack_from("<I", self.d, p + 2 * self.O)[0]
        scratch = p + 2 * self.O + 8
        return name_off, hdr, cache, scratch

    def heap_data(self, heap_addr):
        assert self.d[heap_addr:heap_addr + 4] == b"HEAP"
        return self.off(heap_addr + 8 + 2 * self.L)

    def cstr(self, p):
        e = self.d.index(b"\0", p)
        return self.d[p:e].decode()

    def walk_btree(self, addr, heap_data, out):
        assert self.d[addr:addr + 4] == b"TREE"
        level = self.d[addr + 5]
        n = struct.unpack_from("<H", self.d, addr + 6)[0]
        p = addr + 8 + 2 * self.O
        for _ in range(n):
            p += self.L          # key
            child = self.off(p)
            p += self.O
            if level > 0:
                self.walk_btree(child, heap_data, out)
            else:
                self.read_snod(child, heap_data, out)

    def read_snod(self, addr, heap_data, out):
        assert self.d[addr:addr + 4] == b"SNOD"
        n = struct.unpack_from("<H", self.d, addr + 6)[0]
        p = addr + 8
        for _ in range(n):
            name_off, hdr, _, _ = self.entry(p)
            out[self.cstr(heap_data + name_off)] = hdr
            p += 2 * self.O + 24

    def messages(self, hdr):
        ver = self.d[hdr]
        assert ver == 1, f"object header version {ver} unsupported"
        nmsg = struct.unpack_from("<H", self.d, hdr + 2)[0]
        size = struct.unpack_from("<I", self.d, hdr + 8)[0]
        blocks = [(hdr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize = struct.unpack_from("<HH", self.d, p)
                body = p + 8
                if mtype == 0x10:
                    blocks.append((self.off(body), self.length(body + self.O)))
                msgs.append((mtype, body, msize))
                p = body + msize
        return msgs

    def dataset(self, hdr):
        dims = None
        addr = None
        dtype = "<f8"
        for mtype, p, _ in self.messages(hdr):
            if mtype == 0x1:
                ver, rank, flags = self.d[p], self.d[p + 1], self.d[p + 2]
                q = p + (8 if ver == 1 else 4)
                dims = [self.length(q + i * self.L) for i in range(rank)]
            elif mtype == 0x3:
                cls = self.d[p] & 0x0F
                sz = struct.unpack_from("<I", self.d, p + 4)[0]
                assert cls in (0, 1) and (self.d[p + 1] & 1) == 0, "only LE int/float"
                if cls == 1:
                    dtype = {4: "<f4", 8: "<f8"}[sz]
                else:
                    signed = (self.d[p + 1] >> 3) & 1
                    dtype = ("<i%d" if signed else "<u%d") % sz
            elif mtype == 0x8:
                ver = self.d[p]
                if ver == 3:
                    assert self.d[p + 1] == 1, "only contiguous layout"
                    addr = self.off(p + 2)
                else:
                    rank, cls = self.d[p + 1], self.d[p + 2]
                    assert cls == 1, "only contiguous layout"
                    addr = self.off(p + 8)
        n = int(np.prod(dims)) if dims else 1
        arr = np.frombuffer(self.d, dtype=dtype, count=n, offset=addr).copy()
        return arr.reshape(dims) if dims else arr


def read_h5(path):
    """Return {dataset name: float64 ndarray} for every dataset in the root group."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    _, _, cache, scratch = f.entry(f.root_entry)
    assert cache == 1, "root group without cached symbol table"
    btree = f.off(scratch)
    heap = f.off(scratch + f.O)
    names = {}
    f.walk_btree(btree, f.heap_data(heap), names)
    return {k: f.dataset(h) for k, h in names.items()}
