"""Minimal HDF5 WRITER for the tests of femo_b200/fea/hdf5_lite.py (test infrastructure; h5py is absent offline).

Writes the classic ("earliest") on-disk format of the HDF5 File Format Specification, the one meshio / dolfinx files use:
superblock version 0, version-1 object headers, groups as symbol tables (B-tree node + symbol-table node + local heap),
dataspace version 1, data layout version 3 (contiguous or chunked with a version-1 chunk B-tree, optionally two levels),
filter pipeline version 1 (shuffle + deflate, as h5py's compression="gzip", shuffle=True writes them)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Writer:
    def __init__(self, userblock=0):
        self.b = bytearray(userblock)
        self.base = userblock
        self.b += b'\0' * 96                              # superblock, filled in by close()

    def _put(self, blob):
        while (len(self.b) - self.base) % 8:
            self.b += b'\0'
        addr = len(self.b) - self.base
        self.b += blob
        return addr

    @staticmethod
    def _msg(mtype, data):
        data = data + b'\0' * (-len(data) % 8)
        return struct.pack('<HHB3x', mtype, len(data), 0) + data

    def _header(self, msgs):
        body = b''.join(msgs)
        return self._put(struct.pack('<BxHII4x', 1, len(msgs), 1, len(body)) + body)

    def dataset(self, a, chunks=None, gzip=None, shuffle=False, two_level=False):
        a = np.ascontiguousarray(a)
        dt = a.dtype
        space = struct.pack('<BBBx4x', 1, a.ndim, 0) + b''.join(struct.pack('<Q', d) for d in a.shape)
        order = 1 if dt.byteorder == '>' else 0
        if dt.kind in 'iu':
            dtype = struct.pack('<BBBBI', 0x10, order | (0x08 if dt.kind == 'i' else 0), 0, 0, dt.itemsize) + \
                struct.pack('<HH', 0, 8 * dt.itemsize)
        else:
            exp, man = {4: (8, 23), 8: (11, 52)}[dt.itemsize]
            dtype = struct.pack('<BBBBI', 0x11, order | 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + \
                struct.pack('<HHBBBBI', 0, 8 * dt.itemsize, man, exp, 0, man, (1 << (exp - 1)) - 1)
        msgs = [self._msg(1, space), self._msg(3, dtype)]
        if chunks is None:
            addr = self._put(a.tobytes())
            msgs.append(self._msg(8, struct.pack('<BBQQ', 3, 1, addr, a.nbytes)))
            return self._header(msgs)
        filt = []
        if shuffle:
            filt.append((2, [dt.itemsize]))
        if gzip is not None:
            filt.append((1, [gzip]))
        keys = []
        grid = [range(0, s, c) for s, c in zip(a.shape, chunks)]
        for offs in np.stack(np.meshgrid(*grid, indexing='ij'), -1).reshape(-1, a.ndim):
            blk = np.zeros(chunks, dt)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, a.shape))
            blk[tuple(slice(0, s.stop - s.start) for s in sel)] = a[sel]
            raw = blk.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, dt.itemsize).T.tobytes()
            if gzip is not None:
                raw = zlib.compress(raw, gzip)
            keys.append((len(raw), [int(o) for o in offs], self._put(raw)))

        def node(level, entries):
            out = struct.pack('<4sBBHQQ', b'TREE', 1, level, len(entries), UNDEF, UNDEF)
            for size, offs, child in entries:
                out += struct.pack('<II', size, 0) + b''.join(struct.pack('<Q', o) for o in offs + [0]) + struct.pack('<Q', child)
            out += struct.pack('<II', 0, 0) + b''.join(struct.pack('<Q', s) for s in list(a.shape) + [0])
            return self._put(out)
        if two_level and len(keys) > 1:
            h = len(keys) // 2
            leaves = [keys[:h], keys[h:]]
            root = node(1, [(l[0][0], l[0][1], node(0, l)) for l in leaves])
        else:
            root = node(0, keys)
        lay = struct.pack('<BBBQ', 3, 2, a.ndim + 1, root) + b''.join(struct.pack('<I', c) for c in list(chunks) + [dt.itemsize])
        msgs.append(self._msg(8, lay))
        if filt:
            fp = struct.pack('<BB6x', 1, len(filt))
            for fid, cd in filt:
                fp += struct.pack('<HHHH', fid, 0, 0, len(cd)) + b''.join(struct.pack('<I', v) for v in cd)
                if len(cd) % 2:
                    fp += b'\0' * 4
            msgs.append(self._msg(0x0B, fp))
        return self._header(msgs)

    def group(self, children):
        """children: name -> object header address (of datasets / groups written before)."""
        names = sorted(children)
        heap = bytearray(8)                               # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap)
            e = n.encode() + b'\0'
            heap += e + b'\0' * (-len(e) % 8)
        heap_data = self._put(bytes(heap))
        heap_addr = self._put(struct.pack('<4sB3xQQQ', b'HEAP', 0, len(heap), UNDEF, heap_data))
        snod = struct.pack('<4sBxH', b'SNOD', 1, len(names))
        for n in names:
            snod += struct.pack('<QQII16x', offs[n], children[n], 0, 0)
        snod_addr = self._put(snod)
        tree = struct.pack('<4sBBHQQ', b'TREE', 0, 0, 1, UNDEF, UNDEF) + struct.pack('<QQQ', 0, snod_addr, offs[names[-1]] if names else 0)
        tree_addr = self._put(tree)
        return self._header([self._msg(0x11, struct.pack('<QQ', tree_addr, heap_addr))])

    def close(self, root, path):
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack('<HHI', 4, 16, 0)
        sb += struct.pack('<QQQQ', self.base, UNDEF, len(self.b) - self.base, UNDEF)
        sb += struct.pack('<QQII16x', 0, root, 0, 0)
        assert len(sb) == 96
        self.b[self.base:self.base + 96] = sb
        with open(path, 'wb') as f:
            f.write(bytes(self.b))


def write(path, tree, userblock=0, **kw):
    """tree: nested dict name -> array | dict; kw: chunks / gzip / shuffle / two_level for every array (chunks may be a
    callable of the array)."""
    w = Writer(userblock)

    def emit(node):
        if isinstance(node, dict):
            return w.group({k: emit(v) for k, v in node.items()})
        k = dict(kw)
        if callable(k.get('chunks')):
            k['chunks'] = k['chunks'](node)
        return w.dataset(node, **k)
    w.close(emit(tree), path)
