"""Minimal HDF5 writer: the counterpart of hdf5_min.py.

Writes the classic on-disk format that h5py produces with ``libver='earliest'`` -- and that
multi-read FAST5 files use: superblock version 0, version-1 object headers, "old style" groups
(symbol-table message -> v1 B-tree -> SNOD nodes + local heap), version-1 attribute messages,
contiguous datasets and chunked datasets (layout version 3, v1 chunk B-tree) with the shuffle
and deflate filters.  Used to put FAST5 files on disk for the ingest tests (there is no h5py /
libhdf5 in this image) and as the base for HDF5 outputs.

    root = Group(attrs={'file_version': b'2.0'})
    raw = root.group('read_x').group('Raw', attrs={'duration': np.int64(4000)})
    raw.dataset('Signal', np.zeros(4000, np.int16), chunks=1024, gzip=1, shuffle=True)
    write_file(path, root)
"""
import struct
import zlib

import numpy as np

__all__ = ['Group', 'Dataset', 'Enum', 'RawAttr', 'write_file', 'FILTER_DEFLATE', 'FILTER_SHUFFLE', 'FILTER_VBZ']

_SIGNATURE = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF
FILTER_DEFLATE, FILTER_SHUFFLE, FILTER_VBZ = 1, 2, 32020
_LEAF_K, _INTERNAL_K, _CHUNK_K = 4, 16, 32


def _pad8(n):
    return (n + 7) & ~7


class Dataset:
    """``data``: numpy array (any rank, fixed-size dtype) or ``bytes`` (scalar fixed-length
    string).  ``chunks``: chunk length along axis 0 of a 1-D array (None = contiguous).
    ``encoder``: optional ``(filter_id, cd_values, encode(bytes) -> bytes)`` for a custom filter
    in place of shuffle/deflate."""

    def __init__(self, data, attrs=None, chunks=None, gzip=None, shuffle=False, encoder=None):
        self.data = data
        self.attrs = dict(attrs or {})
        self.chunks = chunks
        self.gzip = gzip
        self.shuffle = shuffle
        self.encoder = encoder


class Group:
    def __init__(self, attrs=None):
        self.attrs = dict(attrs or {})
        self.children = {}

    def group(self, name, attrs=None):
        g = self.children[name] = Group(attrs)
        return g

    def dataset(self, name, data, **kw):
        d = self.children[name] = Dataset(data, **kw)
        return d


# ---- message bodies ---------------------------------------------------------------------
def _datatype(dt):
    """Datatype message for a numpy dtype (fixed point, IEEE float, fixed string, compound)."""
    dt = np.dtype(dt)
    if dt.kind in 'iu':
        bits = 0x08 if dt.kind == 'i' else 0x00
        return struct.pack('<BBBBIHH', 0x10, bits, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == 'f':
        if dt.itemsize == 8:
            sign, prop = 63, struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
        elif dt.itemsize == 4:
            sign, prop = 31, struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
        else:
            raise ValueError('unsupported float size')
        return struct.pack('<BBBBI', 0x11, 0x20, sign, 0, dt.itemsize) + prop
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x00, 0, 0, dt.itemsize)      # null-terminated ASCII
    if dt.kind == 'V' and dt.names:
        body = b''
        for name in dt.names:
            sub, off = dt.fields[name][0], dt.fields[name][1]
            nm = name.encode() + b'\0'
            body += nm.ljust(_pad8(len(nm)), b'\0')
            body += struct.pack('<I', off) + bytes(1 + 3 + 4 + 4 + 16)   # offset, rank 0, no dims
            body += _datatype(sub)
        n = len(dt.names)
        return struct.pack('<BBBBI', 0x16, n & 0xFF, (n >> 8) & 0xFF, 0, dt.itemsize) + body
    raise ValueError('unsupported dtype %r' % dt)


class Enum:
    """Attribute value of an enumeration type: ``Enum({'unknown': 0, 'signal_positive': 2}, 2)``."""

    def __init__(self, members, value, base=np.uint8):
        self.members, self.value, self.base = dict(members), value, np.dtype(base)


class RawAttr:
    """Attribute given as ready-made datatype / dataspace messages and payload bytes (tests use it
    to plant types the readers do not know)."""

    def __init__(self, datatype, dataspace, payload):
        self.datatype, self.dataspace, self.payload = datatype, dataspace, payload


def _enum_datatype(e):
    body = _datatype(e.base)
    for name in e.members:
        nm = name.encode() + b'\0'
        body += nm.ljust(_pad8(len(nm)), b'\0')
    body += np.array(list(e.members.values()), e.base).tobytes()
    n = len(e.members)
    return struct.pack('<BBBBI', 0x18, n & 0xFF, (n >> 8) & 0xFF, 0, e.base.itemsize) + body


def _dataspace(shape):
    body = struct.pack('<BBBBI', 1, len(shape), 0, 0, 0)
    for d in shape:
        body += struct.pack('<Q', d)
    return body


def _as_array(value):
    """Attribute / dataset payload -> (numpy array, shape)."""
    if isinstance(value, (bytes, np.bytes_)):
        b = bytes(value)
        return np.array(b, dtype='S%d' % max(len(b) + 1, 1)), ()
    if isinstance(value, bool):
        return np.array(value, np.int8), ()
    if isinstance(value, int):
        return np.array(value, np.int64), ()
    if isinstance(value, float):
        return np.array(value, np.float64), ()
    a = np.asarray(value)
    return a, a.shape


def _attribute(name, value):
    a, shape = _as_array(value)
    nm = name.encode() + b'\0'
    dt, ds = _datatype(a.dtype), _dataspace(shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds))
    body += nm.ljust(_pad8(len(nm)), b'\0') + dt.ljust(_pad8(len(dt)), b'\0') + \
        ds.ljust(_pad8(len(ds)), b'\0')
    return body + np.ascontiguousarray(a).tobytes()


_VLEN_STR = struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + struct.pack('<BBBBI', 0x13, 0x00, 0, 0, 1)


def _attribute_raw(name, dt, ds, payload):
    nm = name.encode() + b'\0'
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds))
    body += nm.ljust(_pad8(len(nm)), b'\0') + dt.ljust(_pad8(len(dt)), b'\0') + \
        ds.ljust(_pad8(len(ds)), b'\0')
    return body + payload


def _shuffle(raw, elsize):
    a = np.frombuffer(raw, np.uint8)
    n = len(a) // elsize
    return a[:n * elsize].reshape(n, elsize).T.tobytes() + a[n * elsize:].tobytes()


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)                      # superblock written last

    def alloc(self, data, align=8):
        pad = (-len(self.buf)) % align
        self.buf += b'\0' * pad
        addr = len(self.buf)
        self.buf += data
        return addr

    def attribute(self, name, value):
        """``str`` values become variable-length UTF-8 strings in a global heap collection (what
        h5py writes for ``attrs[k] = 'text'``); everything else is stored inline."""
        if isinstance(value, RawAttr):
            return _attribute_raw(name, value.datatype, value.dataspace, value.payload)
        if isinstance(value, Enum):
            return _attribute_raw(name, _enum_datatype(value), _dataspace(()),
                                  np.array(value.value, value.base).tobytes())
        if not isinstance(value, str):
            return _attribute(name, value)
        return _attribute_raw(name, _VLEN_STR, _dataspace(()), self._heap_string(value))

    # ---- B-trees --------------------------------------------------------------------
    def _btree(self, node_type, leaves, key_bytes, k):
        """``leaves``: list of (first_key, child_addr, last_key).  Builds levels of up to 2k
        children; returns the root node address."""
        level = 0
        nodes = leaves
        while True:
            out = []
            for i in range(0, len(nodes), 2 * k):
                grp = nodes[i:i + 2 * k]
                body = b'TREE' + struct.pack('<BBH', node_type, level, len(grp)) + \
                    struct.pack('<QQ', _UNDEF, _UNDEF)
                body += key_bytes(grp[0][0])
                for (_first, child, last) in grp:
                    body += struct.pack('<Q', child) + key_bytes(last)
                klen = len(key_bytes(grp[0][0]))
                body = body.ljust(24 + (2 * k + 1) * klen + 2 * k * 8, b'\0')
                out.append((grp[0][0], self.alloc(body), grp[-1][2]))
            if len(out) == 1:
                return out[0][1]
            nodes = out
            level += 1

    # ---- objects ----------------------------------------------------------------------
    def _header(self, msgs):
        data = b''
        for mtype, body in msgs:
            body = body.ljust(_pad8(len(body)), b'\0')
            if len(body) > 0xFFFF:
                raise ValueError('object header message too large')
            data += struct.pack('<HHBBBB', mtype, len(body), 0, 0, 0, 0) + body
        hdr = struct.pack('<BBHII', 1, 0, len(msgs), 1, len(data)) + b'\0' * 4
        return self.alloc(hdr + data)

    def group(self, g):
        """Returns (object header address, btree address, heap address)."""
        entries = []
        heap = bytearray(8)                           # offset 0: the empty name
        for name in sorted(g.children, key=lambda s: s.encode()):
            child = g.children[name]
            off = len(heap)
            nm = name.encode() + b'\0'
            heap += nm.ljust(_pad8(len(nm)), b'\0')
            if isinstance(child, Group):
                ohdr, bt, hp = self.group(child)
                entries.append((off, ohdr, 1, struct.pack('<QQ', bt, hp)))
            else:
                entries.append((off, self.dataset(child), 0, b'\0' * 16))
        leaves = []
        for i in range(0, max(len(entries), 1), 2 * _LEAF_K):
            grp = entries[i:i + 2 * _LEAF_K]
            body = b'SNOD' + struct.pack('<BBH', 1, 0, len(grp))
            for off, ohdr, ctype, scratch in grp:
                body += struct.pack('<QQII', off, ohdr, ctype, 0) + scratch
            body = body.ljust(8 + 2 * _LEAF_K * 40, b'\0')
            first = entries[i - 1][0] if i else 0
            leaves.append((first, self.alloc(body), grp[-1][0] if grp else 0))
        bt = self._btree(0, leaves, lambda off: struct.pack('<Q', off), _INTERNAL_K)
        heap_data = self.alloc(bytes(heap))
        hp = self.alloc(b'HEAP' + struct.pack('<BBBBQQQ', 0, 0, 0, 0, len(heap), 1, heap_data))
        msgs = [(0x0011, struct.pack('<QQ', bt, hp))]
        msgs += [(0x000C, self.attribute(k, v)) for k, v in g.attrs.items()]
        return self._header(msgs), bt, hp

    def _heap_string(self, text):
        """One global heap collection holding ``text``; returns the 16-byte vlen element."""
        data = text.encode()
        obj = struct.pack('<HHIQ', 1, 1, 0, len(data)) + data.ljust(_pad8(len(data)), b'\0')
        size = 16 + len(obj) + 16                       # header, the object, the free-space object
        coll = self.alloc(b'GCOL' + struct.pack('<BBBBQ', 1, 0, 0, 0, size) + obj + bytes(16))
        return struct.pack('<IQI', len(data), coll, 1)

    def dataset(self, d):
        if isinstance(d.data, str):                     # scalar variable-length string (h5py: str data)
            elem = self._heap_string(d.data)
            msgs = [(0x0001, _dataspace(())), (0x0003, _VLEN_STR),
                    (0x0008, struct.pack('<BBQQ', 3, 1, self.alloc(elem), len(elem)))]
            msgs += [(0x000C, self.attribute(k, v)) for k, v in d.attrs.items()]
            return self._header(msgs)
        a, shape = _as_array(d.data)
        a = np.ascontiguousarray(a)
        msgs = [(0x0001, _dataspace(shape)), (0x0003, _datatype(a.dtype))]
        if d.chunks is None:
            raw = a.tobytes()
            addr = self.alloc(raw) if raw else _UNDEF
            msgs.append((0x0008, struct.pack('<BBQQ', 3, 1, addr, len(raw))))
        else:
            if a.ndim != 1:
                raise ValueError('chunked datasets: 1-D only')
            clen, es = int(d.chunks), a.dtype.itemsize
            filters = []                               # (id, cd_values, encode)
            if d.encoder is not None:
                filters.append(d.encoder)
            else:
                if d.shuffle:
                    filters.append((FILTER_SHUFFLE, [es], lambda b: _shuffle(b, es)))
                if d.gzip is not None:
                    lvl = int(d.gzip)
                    filters.append((FILTER_DEFLATE, [lvl], lambda b: zlib.compress(b, lvl)))
            leaves = []
            for start in range(0, len(a), clen):
                chunk = np.zeros(clen, a.dtype)        # edge chunks are stored full size
                part = a[start:start + clen]
                chunk[:len(part)] = part
                raw = chunk.tobytes()
                for _fid, _cd, enc in filters:
                    raw = enc(raw)
                leaves.append(((len(raw), start), self.alloc(raw), (0, start + clen)))
            if leaves:
                # key of entry i describes chunk i; the final key is past the last chunk
                fixed = [(leaves[i][0], leaves[i][1],
                          leaves[i + 1][0] if i + 1 < len(leaves) else leaves[i][2])
                         for i in range(len(leaves))]
                bt = self._btree(1, fixed,
                                 lambda k: struct.pack('<IIQQ', k[0], 0, k[1], 0), _CHUNK_K)
            else:
                bt = _UNDEF
            msgs.append((0x0008, struct.pack('<BBBQII', 3, 2, 2, bt, clen, es)))
            if filters:
                body = struct.pack('<BBHI', 1, len(filters), 0, 0)
                for fid, cd, _enc in filters:
                    body += struct.pack('<HHHH', fid, 0, 0, len(cd))
                    body += b''.join(struct.pack('<I', v) for v in cd)
                    if len(cd) & 1:
                        body += b'\0' * 4
                msgs.append((0x000B, body))
        msgs += [(0x000C, self.attribute(k, v)) for k, v in d.attrs.items()]
        return self._header(msgs)


def write_file(path, root):
    w = _Writer()
    ohdr, bt, hp = w.group(root)
    eof = len(w.buf)
    sb = _SIGNATURE + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0)
    sb += struct.pack('<HHI', _LEAF_K, _INTERNAL_K, 0)
    sb += struct.pack('<QQQQ', 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack('<QQII', 0, ohdr, 1, 0) + struct.pack('<QQ', bt, hp)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, 'wb') as f:
        f.write(w.buf)
    return eof
