"""Read-only HDF5 subset in pure Python / numpy, for the heavy data of XDMF meshes when h5py is absent.

The reference's meshes are XDMF + HDF5 pairs written by msh2xdmf / meshio and by dolfinx's XDMFFile
(`femo/fea/utils_dolfinx.py:69-123` reads them back with dolfinx); both writers use the HDF5 library's default ("earliest")
file format.  This module implements exactly what such files contain, from the published HDF5 File Format Specification
(version 1 / 2 structures; nothing is derived from the HDF5 library's sources):

  * superblock versions 0 - 3 (a user block in front is skipped);
  * groups as symbol tables (version-1 B-tree + symbol-table nodes + local heap) and as compact link messages of version-2
    object headers (dense groups in fractal heaps are not supported);
  * object headers version 1 and version 2, with continuation blocks;
  * dataspace versions 1 - 2, fixed-point and IEEE floating-point datatypes of 1 - 8 bytes in either byte order;
  * data layout versions 1 - 3: compact, contiguous and chunked storage (version-1 chunk B-tree), with the deflate, shuffle and
    fletcher32 filters (meshio compresses its datasets with gzip).

Anything else raises NotImplementedError naming the feature.  `File(path)[name]` returns the dataset as a numpy array.
"""
import struct
import zlib

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = {4: 0xFFFFFFFF, 8: 0xFFFFFFFFFFFFFFFF}


class File:
    def __init__(self, path):
        with open(path, 'rb') as f:
            self.buf = f.read()
        off = 0
        while self.buf[off:off + 8] != _SIG:                 # the superblock sits at 0, 512, 1024, 2048, ...
            off = 512 if off == 0 else 2 * off
            if off + 8 > len(self.buf):
                raise ValueError('%s: not an HDF5 file' % path)
        self._superblock(off)

    # -- primitives ----------------------------------------------------------------------------------------------------
    def _u(self, pos, n):
        return int.from_bytes(self.buf[pos:pos + n], 'little')

    def _addr(self, pos):
        a = self._u(pos, self.so)
        return None if a == _UNDEF.get(self.so) else a + self.base

    def _superblock(self, off):
        v = self.buf[off + 8]
        if v in (0, 1):
            self.so, self.sl = self.buf[off + 13], self.buf[off + 14]
            p = off + 24 + (4 if v == 1 else 0)
            self.base = self._u(p, self.so)
            if self.base == 0 and off:
                self.base = off                                # writers that leave the field at 0 behind a user block
            p += 4 * self.so                                   # base, free-space info, end of file, driver info
            self.root = self._u(p + self.so, self.so) + self.base          # symbol-table entry: name offset, header address
        elif v in (2, 3):
            self.so, self.sl = self.buf[off + 9], self.buf[off + 10]
            p = off + 12
            self.base = self._u(p, self.so)
            if self.base == 0 and off:
                self.base = off
            self.root = self._u(p + 3 * self.so, self.so) + self.base
        else:
            raise NotImplementedError('HDF5 superblock version %d' % v)

    # -- object headers ------------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """[(type, flags, position, size)] of the object header at `addr`, following continuation blocks."""
        out = []
        if self.buf[addr:addr + 4] == b'OHDR':
            if self.buf[addr + 4] != 2:
                raise NotImplementedError('object header version %d' % self.buf[addr + 4])
            flags = self.buf[addr + 5]
            p = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            n = 1 << (flags & 3)
            size = self._u(p, n)
            blocks = [(p + n, size)]
            track = bool(flags & 0x04)
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype, msize, mflags = self.buf[p], self._u(p + 1, 2), self.buf[p + 3]
                    p += 4 + (2 if track else 0)
                    if mtype == 0x10:
                        ca, cl = self._addr(p), self._u(p + self.so, self.sl)
                        blocks.append((ca + 4, cl - 8))        # "OCHK" ... checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        if self.buf[addr] != 1:
            raise NotImplementedError('object header version %d' % self.buf[addr])
        nmsg, size = self._u(addr + 2, 2), self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and nmsg > 0:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                p += 8
                nmsg -= 1
                if mtype == 0x10:
                    blocks.append((self._addr(p), self._u(p + self.so, self.sl)))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    # -- groups --------------------------------------------------------------------------------------------------------
    def _children(self, addr):
        """name -> object header address of a group."""
        kids = {}
        for mtype, _, p, size in self._messages(addr):
            if mtype == 0x11:                                   # symbol table: B-tree + local heap
                btree, heap = self._addr(p), self._addr(p + self.so)
                if self.buf[heap:heap + 4] != b'HEAP':
                    raise ValueError('bad local heap')
                data = self._addr(heap + 8 + 2 * self.sl)
                self._group_btree(btree, data, kids)
            elif mtype == 0x06:                                 # link message (compact new-style group)
                if self.buf[p] != 1:
                    raise NotImplementedError('link message version %d' % self.buf[p])
                fl = self.buf[p + 1]
                q = p + 2
                ltype = 0
                if fl & 0x08:
                    ltype = self.buf[q]
                    q += 1
                if fl & 0x04:
                    q += 8
                if fl & 0x10:
                    q += 1
                n = 1 << (fl & 3)
                ln = self._u(q, n)
                q += n
                name = self.buf[q:q + ln].decode('utf-8')
                q += ln
                if ltype == 0:
                    kids[name] = self._addr(q)
            elif mtype == 0x02:                                 # link info: dense storage if a fractal heap is named
                q = p + 2 + (8 if self.buf[p + 1] & 1 else 0)
                if self._addr(q) is not None:
                    raise NotImplementedError('dense link storage (fractal heap) in a group')
        return kids

    def _group_btree(self, node, heap_data, kids):
        if self.buf[node:node + 4] != b'TREE' or self.buf[node + 4] != 0:
            raise ValueError('bad group B-tree node')
        level, used = self.buf[node + 5], self._u(node + 6, 2)
        p = node + 8 + 2 * self.so
        for k in range(used):
            child = self._addr(p + self.sl + k * (self.sl + self.so))
            if level:
                self._group_btree(child, heap_data, kids)
                continue
            if self.buf[child:child + 4] != b'SNOD':
                raise ValueError('bad symbol-table node')
            q = child + 8
            for _ in range(self._u(child + 6, 2)):
                noff, oaddr = self._u(q, self.so), self._addr(q + self.so)
                end = self.buf.index(b'\0', heap_data + noff)
                kids[self.buf[heap_data + noff:end].decode('utf-8')] = oaddr
                q += 2 * self.so + 24

    def _resolve(self, name):
        addr = self.root
        for part in [s for s in name.split('/') if s]:
            kids = self._children(addr)
            if part not in kids:
                raise KeyError('%s: no object %r (have %s)' % (name, part, sorted(kids)))
            addr = kids[part]
        return addr

    def keys(self, group='/'):
        return sorted(self._children(self._resolve(group)))

    # -- datasets ------------------------------------------------------------------------------------------------------
    def __getitem__(self, name):
        msgs = self._messages(self._resolve(name))
        shape = dtype = layout = None
        filters = []
        for mtype, _, p, size in msgs:
            if mtype == 0x01:
                shape = self._dataspace(p)
            elif mtype == 0x03:
                dtype = self._datatype(p)
            elif mtype == 0x08:
                layout = p
            elif mtype == 0x0B:
                filters = self._filters(p)
        if shape is None or dtype is None or layout is None:
            raise KeyError('%s is not a dataset' % name)
        return self._read(layout, shape, dtype, filters)

    def _dataspace(self, p):
        v, rank = self.buf[p], self.buf[p + 1]
        if v == 1:
            q = p + 8
        elif v == 2:
            if self.buf[p + 3] == 2:
                return (0,)                                     # null dataspace
            q = p + 4
        else:
            raise NotImplementedError('dataspace version %d' % v)
        return tuple(self._u(q + k * self.sl, self.sl) for k in range(rank))

    def _datatype(self, p):
        cls, bits0, size = self.buf[p] & 0x0F, self.buf[p + 1], self._u(p + 4, 4)
        order = '>' if bits0 & 1 else '<'
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(order + ('i' if bits0 & 0x08 else 'u') + str(size))
        if cls == 1 and size in (2, 4, 8):
            return np.dtype(order + 'f' + str(size))
        raise NotImplementedError('HDF5 datatype class %d of %d bytes' % (cls, size))

    def _filters(self, p):
        v, n = self.buf[p], self.buf[p + 1]
        q = p + (8 if v == 1 else 2)
        out = []
        for _ in range(n):
            fid = self._u(q, 2)
            q += 2
            nlen = 0
            if v == 1 or fid >= 256:
                nlen = self._u(q, 2)
                q += 2
            q += 2                                              # flags
            ncd = self._u(q, 2)
            q += 2
            q += (nlen + 7) // 8 * 8 if v == 1 else nlen
            cd = [self._u(q + 4 * k, 4) for k in range(ncd)]
            q += 4 * ncd + (4 if v == 1 and ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def _read(self, p, shape, dtype, filters):
        v = self.buf[p]
        n = int(np.prod(shape, dtype=np.int64))
        if v == 3:
            cls = self.buf[p + 1]
            if cls == 0:
                size = self._u(p + 2, 2)
                return np.frombuffer(self.buf, dtype, n, p + 4).reshape(shape).copy() if size else np.zeros(shape, dtype)
            if cls == 1:
                return self._contiguous(self._addr(p + 2), shape, dtype, n)
            if cls == 2:
                nd = self.buf[p + 2]
                btree = self._addr(p + 3)
                cdims = [self._u(p + 3 + self.so + 4 * k, 4) for k in range(nd)]
                return self._chunked(btree, shape, dtype, cdims[:-1], filters)
            raise NotImplementedError('data layout class %d' % cls)
        if v in (1, 2):
            nd, cls = self.buf[p + 1], self.buf[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._addr(q)
                q += self.so
            dims = [self._u(q + 4 * k, 4) for k in range(nd)]
            q += 4 * nd
            if cls == 1:
                return self._contiguous(addr, shape, dtype, n)
            if cls == 2:
                return self._chunked(addr, shape, dtype, dims[:-1] if len(dims) > len(shape) else dims, filters)
            size = self._u(q, 4)
            return np.frombuffer(self.buf, dtype, n, q + 4).reshape(shape).copy() if size else np.zeros(shape, dtype)
        raise NotImplementedError('data layout message version %d (files written with libver="latest"; rewrite them with the '
                                  'default format)' % v)

    def _contiguous(self, addr, shape, dtype, n):
        if addr is None or n == 0:
            return np.zeros(shape, dtype)                        # never written: fill value
        return np.frombuffer(self.buf, dtype, n, addr).reshape(shape).copy()

    def _chunked(self, btree, shape, dtype, cdims, filters):
        out = np.zeros(shape, dtype)
        if btree is None or out.size == 0:
            return out
        self._chunk_btree(btree, out, cdims, filters)
        return out

    def _chunk_btree(self, node, out, cdims, filters):
        if self.buf[node:node + 4] != b'TREE' or self.buf[node + 4] != 1:
            raise ValueError('bad chunk B-tree node')
        level, used = self.buf[node + 5], self._u(node + 6, 2)
        nd = len(cdims)
        ksize = 8 + 8 * (nd + 1)
        p = node + 8 + 2 * self.so
        for _ in range(used):
            nbytes, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = [self._u(p + 8 + 8 * k, 8) for k in range(nd)]
            child = self._addr(p + ksize)
            p += ksize + self.so
            if level:
                self._chunk_btree(child, out, cdims, filters)
                continue
            raw = self.buf[child:child + nbytes]
            for k in reversed(range(len(filters))):               # undo the pipeline, last filter first
                if mask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else out.dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    m = a.size // es
                    raw = np.concatenate([a[:m * es].reshape(es, m).T.ravel(), a[m * es:]]).tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise NotImplementedError('HDF5 filter %d' % fid)
            chunk = np.frombuffer(raw, out.dtype, int(np.prod(cdims, dtype=np.int64))).reshape(cdims)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]


def read_dataset(path, name):
    return File(path)[name]
