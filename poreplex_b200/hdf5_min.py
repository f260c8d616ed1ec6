"""Minimal read-only HDF5 walker for Keras 2.2.4 weight files.

The reference opens its two model files with h5py (``signal_loader.py:49-75``,
``barcoding.py:51-70``).  h5py/libhdf5 do not exist in this image, so this module
understands exactly the subset of the HDF5 file format those two files use:

* superblock version 0, 8-byte offsets/lengths
* version-1 object headers (with continuation blocks)
* "old style" groups: symbol-table message -> v1 B-tree -> SNOD nodes + local heap
* contiguous (and compact) unfiltered datasets of fixed-point / IEEE-float / string / compound type
* chunked datasets (layout class 2, v1 chunk B-tree) with the deflate, shuffle, fletcher32 and
  VBZ (32020, needs libzstd) filters -- the ``Signal`` / ``Move`` datasets of FAST5 files
* version 1..3 attribute messages holding fixed- or variable-length strings and
  numeric scalars/arrays (vlen data lives in global heap collections)

Anything else raises ``Hdf5FormatError`` instead of guessing.
"""
import mmap
import os
import struct
import zlib

import numpy as np

__all__ = ['Hdf5File', 'Hdf5FormatError']

_SIGNATURE = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF
_READ_WHOLE_BELOW = 8 << 20


class Hdf5FormatError(Exception):
    pass


def _pad8(n):
    return (n + 7) & ~7


class _Datatype:
    """Decoded datatype message."""

    def __init__(self, cls, size, dtype=None, vlen_string=False, consumed=0):
        self.cls = cls
        self.size = size
        self.dtype = dtype            # numpy dtype for fixed-size types
        self.vlen_string = vlen_string
        self.consumed = consumed      # bytes of the message that were parsed


def _parse_datatype(buf, pos):
    b0 = buf[pos]
    version, cls = b0 >> 4, b0 & 0x0F
    bits = buf[pos + 1] | (buf[pos + 2] << 8) | (buf[pos + 3] << 16)
    size = struct.unpack_from('<I', buf, pos + 4)[0]
    p = pos + 8
    if cls == 0:      # fixed point
        endian = '>' if bits & 1 else '<'
        signed = bool(bits & 8)
        dt = np.dtype('%s%s%d' % (endian, 'i' if signed else 'u', size))
        return _Datatype(cls, size, dt, consumed=p + 4 - pos)
    if cls == 1:      # IEEE float
        endian = '>' if bits & 1 else '<'
        dt = np.dtype('%sf%d' % (endian, size))
        return _Datatype(cls, size, dt, consumed=p + 12 - pos)
    if cls == 3:      # fixed-length string
        return _Datatype(cls, size, np.dtype('S%d' % size), consumed=p - pos)
    if cls == 6:      # compound
        nmemb = bits & 0xFFFF
        names, formats, offsets = [], [], []
        for _ in range(nmemb):
            end = buf.index(b'\0', p)
            name = buf[p:end].decode()
            if version < 3:
                p += _pad8(end - p + 1)
            else:
                p = end + 1
            if version == 1:
                off = struct.unpack_from('<I', buf, p)[0]
                p += 4 + 1 + 3 + 4 + 4 + 16
            elif version == 2:
                off = struct.unpack_from('<I', buf, p)[0]
                p += 4
            else:
                nb = 1
                while (1 << (8 * nb)) <= size:
                    nb += 1
                off = int.from_bytes(buf[p:p + nb], 'little')
                p += nb
            sub = _parse_datatype(buf, p)
            if sub.dtype is None:
                raise Hdf5FormatError('unsupported compound member type')
            p += sub.consumed
            names.append(name)
            formats.append(sub.dtype)
            offsets.append(off)
        dt = np.dtype({'names': names, 'formats': formats, 'offsets': offsets,
                       'itemsize': size})
        return _Datatype(cls, size, dt, consumed=p - pos)
    if cls == 8:      # enumeration (e.g. Raw/end_reason): values are read as the base integer
        nmemb = bits & 0xFFFF
        base = _parse_datatype(buf, p)
        if base.dtype is None:
            raise Hdf5FormatError('unsupported enumeration base type')
        q = p + base.consumed
        for _ in range(nmemb):
            end = buf.index(b'\0', q)
            q = q + _pad8(end - q + 1) if version < 3 else end + 1
        q += nmemb * base.size
        return _Datatype(cls, size, base.dtype, consumed=q - pos)
    if cls == 9:      # variable length
        is_string = (bits & 0x0F) == 1
        base = _parse_datatype(buf, p)
        return _Datatype(cls, size, None, vlen_string=is_string,
                         consumed=p + base.consumed - pos)
    raise Hdf5FormatError('unsupported datatype class %d' % cls)


def _parse_dataspace(buf, pos):
    version = buf[pos]
    rank = buf[pos + 1]
    flags = buf[pos + 2]
    if version == 1:
        p = pos + 8
    elif version == 2:
        p = pos + 4
    else:
        raise Hdf5FormatError('unsupported dataspace version %d' % version)
    dims = struct.unpack_from('<%dQ' % rank, buf, p) if rank else ()
    p += 8 * rank
    if flags & 1:
        p += 8 * rank
    return tuple(int(d) for d in dims), p - pos


# ---- filters -------------------------------------------------------------------------------
_zstd = None


def _zstd_decompress(src, expected):
    """One zstd frame -> bytes, through the system libzstd (no Python binding in this image)."""
    global _zstd
    import ctypes as C
    if _zstd is None:
        for name in ('libzstd.so.1', 'libzstd.so'):
            try:
                _zstd = C.CDLL(name)
                break
            except OSError:
                continue
        else:
            raise Hdf5FormatError('VBZ-compressed dataset: libzstd is not available')
        _zstd.ZSTD_decompress.restype = C.c_size_t
        _zstd.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _zstd.ZSTD_isError.argtypes = [C.c_size_t]
        _zstd.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        _zstd.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
    src = bytes(src)
    size = _zstd.ZSTD_getFrameContentSize(src, len(src))
    if size >= (1 << 62):                          # unknown / error: fall back on the bound
        size = expected
    dst = C.create_string_buffer(max(int(size), 1))
    n = _zstd.ZSTD_decompress(dst, int(size), src, len(src))
    if _zstd.ZSTD_isError(n):
        raise Hdf5FormatError('zstd decompression failed')
    return dst.raw[:n]


def _svb_decode(buf, count, key_bits):
    """streamvbyte: ``count`` little-endian codes; key_bits = 1 (svb16: 1 or 2 bytes per value,
    one key bit each) or 2 (classic: 1..4 bytes, two key bits each).  Keys first, data after."""
    per = 8 // key_bits
    nkeys = (count + per - 1) // per
    keys = np.frombuffer(buf, np.uint8, nkeys)
    shifts = (np.arange(per, dtype=np.uint8) * key_bits)
    codes = ((keys[:, None] >> shifts[None, :]) & ((1 << key_bits) - 1)).reshape(-1)[:count]
    lens = codes.astype(np.int64) + 1
    ends = np.cumsum(lens)
    starts = ends - lens
    data = np.frombuffer(buf, np.uint8, int(ends[-1]) if count else 0, nkeys)
    out = np.zeros(count, np.uint32)
    for b in range(1 << key_bits):
        m = lens > b
        out[m] |= data[starts[m] + b].astype(np.uint32) << np.uint32(8 * b)
    return out


def _vbz_decode(chunk, cd_values, expected):
    """ONT VBZ filter (id 32020).  cd_values = (version, integer size, delta+zigzag, zstd level).
    Layout as published in ONT's vbz_compression: uint32 uncompressed size, then an optional zstd
    frame around the streamvbyte stream of the (delta, zigzag) coded samples; version 1 codes
    2-byte integers with svb16 (1-bit keys), everything else with the classic 2-bit keys.
    NOT checked against a file written by ONT's plugin (none available here)."""
    version, isize, zigzag, level = (list(cd_values) + [0, 0, 0, 0])[:4]
    if version > 1:
        raise Hdf5FormatError('unsupported VBZ version %d' % version)
    size = struct.unpack_from('<I', chunk, 0)[0]
    body = bytes(chunk[4:])
    if level != 0:
        body = _zstd_decompress(body, max(expected, size) * 2 + 64)
    if isize in (0, 1):
        return body[:size]
    count = size // isize
    svb16 = version == 1 and isize == 2
    vals = _svb_decode(body, count, 1 if svb16 else 2)
    if zigzag:
        if svb16:
            v = vals.astype(np.uint16)
            d = ((v >> np.uint16(1)) ^ (np.uint16(0) - (v & np.uint16(1)))).astype(np.uint16)
            vals = np.cumsum(d, dtype=np.uint16)
        else:
            d = (vals >> np.uint32(1)) ^ (np.uint32(0) - (vals & np.uint32(1)))
            vals = np.cumsum(d, dtype=np.uint32)
    return vals.astype({1: np.uint8, 2: np.uint16, 4: np.uint32}[isize]).tobytes()


def _unshuffle(raw, elsize):
    a = np.frombuffer(raw, np.uint8)
    n = len(a) // elsize
    return a[:n * elsize].reshape(elsize, n).T.tobytes() + a[n * elsize:].tobytes()


def _parse_pipeline(mbuf):
    """Filter pipeline message -> [(filter id, client data values)] in application order."""
    version, nfilters = mbuf[0], mbuf[1]
    p = 8 if version == 1 else 2
    filters = []
    for _ in range(nfilters):
        fid = struct.unpack_from('<H', mbuf, p)[0]
        p += 2
        if version == 1 or fid >= 256:
            nlen = struct.unpack_from('<H', mbuf, p)[0]
            p += 2
        else:
            nlen = 0
        _flags, ncd = struct.unpack_from('<HH', mbuf, p)
        p += 4
        p += _pad8(nlen) if version == 1 else nlen
        cd = struct.unpack_from('<%dI' % ncd, mbuf, p)
        p += 4 * ncd
        if version == 1 and ncd & 1:
            p += 4
        filters.append((fid, cd))
    return filters


def _defilter(raw, filters, mask, expected):
    for i in reversed(range(len(filters))):
        if mask & (1 << i):
            continue
        fid, cd = filters[i]
        if fid == 1:
            raw = zlib.decompress(raw)
        elif fid == 2:
            raw = _unshuffle(raw, cd[0])
        elif fid == 3:
            raw = raw[:-4]                          # fletcher32 checksum, not verified
        elif fid == 32020:
            raw = _vbz_decode(raw, cd, expected)
        else:
            raise Hdf5FormatError('unsupported filter %d' % fid)
    return raw


class _Attrs(dict):
    pass


class _Node:
    def __init__(self, h5, addr, name):
        self._h5 = h5
        self._addr = addr
        self.name = name
        self._msgs = h5._read_object_header(addr)
        self.attrs = _Attrs()
        for mtype, mbuf in self._msgs:
            if mtype == 0x000C:
                try:
                    k, v = h5._parse_attribute(mbuf)
                except (Hdf5FormatError, struct.error, ValueError, IndexError):
                    continue        # an attribute of a type outside the subset hides only itself
                self.attrs[k] = v


class Hdf5Dataset(_Node):
    def __init__(self, h5, addr, name):
        super().__init__(h5, addr, name)
        self.shape = self._dtype = self._data_addr = self._compact = None
        self._chunk_btree = self._chunk_shape = None
        self._filters = []
        self._vlen = False
        for mtype, mbuf in self._msgs:
            if mtype == 0x0001:
                self.shape, _ = _parse_dataspace(mbuf, 0)
            elif mtype == 0x0003:
                dt = _parse_datatype(mbuf, 0)
                if dt.vlen_string:
                    # elements are (length u4, global heap collection u8, object index u4)
                    self._vlen = True
                    self._dtype = np.dtype([('len', '<u4'), ('coll', '<u8'), ('idx', '<u4')])
                elif dt.dtype is None:
                    raise Hdf5FormatError('unsupported dataset datatype')
                else:
                    self._dtype = dt.dtype
            elif mtype == 0x0008:
                self._parse_layout(mbuf)
            elif mtype == 0x000B:
                self._filters = _parse_pipeline(mbuf)
        if self.shape is None or self._dtype is None:
            raise Hdf5FormatError('not a dataset: ' + name)

    @property
    def dtype(self):
        return self._dtype

    @property
    def file_offset(self):
        """Absolute byte offset of the contiguous data (None if compact)."""
        return self._data_addr

    def _parse_layout(self, mbuf):
        version = mbuf[0]
        if version == 3:
            lclass = mbuf[1]
            if lclass == 1:
                self._data_addr, _size = struct.unpack_from('<QQ', mbuf, 2)
            elif lclass == 0:
                size = struct.unpack_from('<H', mbuf, 2)[0]
                self._compact = bytes(mbuf[4:4 + size])
            elif lclass == 2:
                ndim = mbuf[2]                       # rank + 1 (the last one is the element size)
                self._chunk_btree = struct.unpack_from('<Q', mbuf, 3)[0]
                self._chunk_shape = struct.unpack_from('<%dI' % ndim, mbuf, 11)[:-1]
            else:
                raise Hdf5FormatError('unsupported layout class %d' % lclass)
        elif version in (1, 2):
            rank, lclass = mbuf[1], mbuf[2]
            if lclass == 2:
                self._chunk_btree = struct.unpack_from('<Q', mbuf, 8)[0]
                self._chunk_shape = struct.unpack_from('<%dI' % rank, mbuf, 16)[:-1]
            elif lclass == 1:
                self._data_addr = struct.unpack_from('<Q', mbuf, 8)[0]
            else:
                raise Hdf5FormatError('unsupported v1/v2 layout class %d' % lclass)
        else:
            raise Hdf5FormatError('unsupported layout version %d' % version)

    def chunks(self):
        """Yield (element offsets, stored size, filter mask, file address) of every chunk."""
        buf, rank = self._h5._buf, len(self.shape)
        klen = 8 + 8 * (rank + 1)

        def walk(node, want_level=None):
            if buf[node:node + 4] != b'TREE' or buf[node + 4] != 1:
                raise Hdf5FormatError('bad chunk B-tree node')
            level = buf[node + 5]
            # levels fall by one per step: a node that names itself (or an ancestor) as its child
            # ends the walk instead of recursing without end
            if want_level is not None and level != want_level:
                raise Hdf5FormatError('chunk B-tree levels are inconsistent')
            used = struct.unpack_from('<H', buf, node + 6)[0]
            p = node + 24
            for _ in range(used):
                size, mask = struct.unpack_from('<II', buf, p)
                offs = struct.unpack_from('<%dQ' % rank, buf, p + 8)
                child = struct.unpack_from('<Q', buf, p + klen)[0]
                p += klen + 8
                if level:
                    yield from walk(child, level - 1)
                else:
                    yield offs, size, mask, child

        if self._chunk_btree is not None and self._chunk_btree != _UNDEF:
            yield from walk(self._chunk_btree)

    def _read_chunked(self):
        out = np.zeros(self.shape, self._dtype)
        cshape = tuple(self._chunk_shape)
        cbytes = int(np.prod(cshape, dtype=np.int64)) * self._dtype.itemsize
        for offs, size, mask, addr in self.chunks():
            raw = self._h5._buf[addr:addr + size]
            if self._filters:
                raw = _defilter(raw, self._filters, mask, cbytes)
            if len(raw) < cbytes:
                raise Hdf5FormatError('short chunk in ' + self.name)
            chunk = np.frombuffer(raw, self._dtype, count=cbytes // self._dtype.itemsize).reshape(cshape)
            dst = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cshape, self.shape))
            src = tuple(slice(0, s.stop - s.start) for s in dst)
            out[dst] = chunk[src]
        return out

    def read(self):
        arr = self._read_fixed()
        if not self._vlen:
            return arr
        # variable-length strings (what h5py writes for str data): bytes objects, like h5py 3
        flat = [bytes(self._h5._global_heap_object(int(e['coll']), int(e['idx'])))[:int(e['len'])]
                for e in arr.reshape(-1)]
        if not self.shape:
            return np.array(flat[0], dtype=object)
        out = np.empty(len(flat), dtype=object)
        out[:] = flat
        return out.reshape(self.shape)

    def _read_fixed(self):
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        nbytes = n * self._dtype.itemsize
        if self._chunk_shape is not None:
            return self._read_chunked()
        if self._compact is not None:
            raw = self._compact[:nbytes]
        else:
            if self._data_addr == _UNDEF:
                return np.zeros(self.shape, self._dtype)
            raw = self._h5._buf[self._data_addr:self._data_addr + nbytes]
        return np.frombuffer(raw, self._dtype, count=n).reshape(self.shape).copy()

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        arr = self.read()
        if isinstance(key, tuple) and key == ():
            return arr if arr.shape else arr[()]
        return arr[key]


class Hdf5Group(_Node):
    def __init__(self, h5, addr, name):
        super().__init__(h5, addr, name)
        self._children = None
        self._stab = None
        for mtype, mbuf in self._msgs:
            if mtype == 0x0011:
                self._stab = struct.unpack_from('<QQ', mbuf, 0)
        if self._stab is None:
            raise Hdf5FormatError('not an old-style group: ' + name)

    def _load(self):
        if self._children is None:
            self._children = dict(self._h5._walk_group(*self._stab))
        return self._children

    def keys(self):
        return list(self._load().keys())

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def _find(self, name):
        """Object header address of child ``name`` (None if absent) by descending the B-tree
        through its keys: O(log n), no listing of a 4000-read root group per lookup."""
        if self._children is not None:
            return self._children.get(name)
        return self._h5._find_in_group(self._stab[0], self._stab[1], name.encode())

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Hdf5Group):
                raise KeyError(path)
            addr = node._find(part)
            if addr is None:
                raise KeyError(path)
            node = node._h5._open(addr, (node.name.rstrip('/') + '/' + part))
        return node

    def visit_datasets(self, prefix=''):
        """Yield (path, Hdf5Dataset) for every dataset below this group."""
        for k in self.keys():
            child = self[k]
            p = prefix + k
            if isinstance(child, Hdf5Group):
                yield from child.visit_datasets(p + '/')
            else:
                yield p, child


class Hdf5File(Hdf5Group):
    def __init__(self, path):
        # Large files are mapped, not read: a multi-read FAST5 is hundreds of MB and one read is a
        # few KB of it.  Small ones (single-read FAST5, model files) are read whole, so that a
        # batch of thousands of open single-read files holds no descriptors at all.
        self._fh = open(path, 'rb')
        if os.fstat(self._fh.fileno()).st_size <= _READ_WHOLE_BELOW:
            self._buf = self._fh.read()
            self._fh.close()
            self._fh = None
        else:
            self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        self._cache = {}
        try:
            self._open_root(path)
        except Exception:
            self.close()
            raise

    def _open_root(self, path):
        buf = self._buf
        if len(buf) < 96 or buf[:8] != _SIGNATURE:
            raise Hdf5FormatError('not an HDF5 file: ' + path)
        if buf[8] != 0:
            raise Hdf5FormatError('unsupported superblock version %d' % buf[8])
        if buf[13] != 8 or buf[14] != 8:
            raise Hdf5FormatError('only 8-byte offsets/lengths are supported')
        base = struct.unpack_from('<Q', buf, 24)[0]
        if base != 0:
            raise Hdf5FormatError('non-zero base address')
        # root symbol table entry follows the four addresses at byte 24
        root_entry = 24 + 4 * 8
        _name_off, ohdr = struct.unpack_from('<QQ', buf, root_entry)
        Hdf5Group.__init__(self, self, ohdr, '/')

    def close(self):
        self._cache = {}
        if self._fh is not None:
            try:
                self._buf.close()
            except BufferError:                     # a numpy view of the map is still alive
                pass
            self._fh.close()
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- low level ---------------------------------------------------------
    def _open(self, addr, name):
        if addr not in self._cache:
            if len(self._cache) > 8192:             # a shared handle visits every read of the file
                self._cache.clear()
            msgs = self._read_object_header(addr)
            types = {t for t, _ in msgs}
            cls = Hdf5Group if 0x0011 in types else Hdf5Dataset
            self._cache[addr] = cls(self, addr, name)
        return self._cache[addr]

    def _read_object_header(self, addr):
        buf = self._buf
        if buf[addr] != 1:
            raise Hdf5FormatError('only version-1 object headers are supported')
        nmsgs = struct.unpack_from('<H', buf, addr + 2)[0]
        hsize = struct.unpack_from('<I', buf, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsgs:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(msgs) < nmsgs:
                mtype, msize, _flags = struct.unpack_from('<HHB', buf, p)
                body = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from('<QQ', body, 0)
                    blocks.append((coff, clen))
                msgs.append((mtype, body))
        return msgs

    def _walk_group(self, btree_addr, heap_addr):
        buf = self._buf
        if buf[heap_addr:heap_addr + 4] != b'HEAP':
            raise Hdf5FormatError('bad local heap signature')
        heap_data = struct.unpack_from('<Q', buf, heap_addr + 24)[0]

        def heap_string(off):
            start = heap_data + off
            return buf[start:buf.find(b'\0', start)].decode()

        def walk(node, want_level=None):
            if buf[node:node + 4] == b'TREE':
                level = buf[node + 5]
                if want_level is not None and level != want_level:      # no cycles: see chunks()
                    raise Hdf5FormatError('group B-tree levels are inconsistent')
                used = struct.unpack_from('<H', buf, node + 6)[0]
                p = node + 8 + 16
                for i in range(used):
                    child = struct.unpack_from('<Q', buf, p + 8)[0]
                    p += 16
                    yield from walk(child, level - 1)
            elif buf[node:node + 4] == b'SNOD':
                if want_level is not None and want_level >= 0:
                    raise Hdf5FormatError('group B-tree levels are inconsistent')
                nsym = struct.unpack_from('<H', buf, node + 6)[0]
                p = node + 8
                for _ in range(nsym):
                    name_off, ohdr = struct.unpack_from('<QQ', buf, p)
                    yield heap_string(name_off), ohdr
                    p += 40
            else:
                raise Hdf5FormatError('bad group node signature')

        yield from walk(btree_addr)

    def _find_in_group(self, node, heap_addr, name):
        buf = self._buf
        if buf[heap_addr:heap_addr + 4] != b'HEAP':
            raise Hdf5FormatError('bad local heap signature')
        heap_data = struct.unpack_from('<Q', buf, heap_addr + 24)[0]

        def heap_bytes(off):
            start = heap_data + off
            return buf[start:buf.find(b'\0', start)]

        for _depth in range(64):
            sig = buf[node:node + 4]
            if sig == b'TREE':
                used = struct.unpack_from('<H', buf, node + 6)[0]
                p = node + 24                       # key0, child0, key1, child1, ...
                for i in range(used):
                    hi = heap_bytes(struct.unpack_from('<Q', buf, p + 16 * i + 16)[0])
                    if name <= hi:
                        node = struct.unpack_from('<Q', buf, p + 16 * i + 8)[0]
                        break
                else:
                    return None
            elif sig == b'SNOD':
                nsym = struct.unpack_from('<H', buf, node + 6)[0]
                for i in range(nsym):
                    name_off, ohdr = struct.unpack_from('<QQ', buf, node + 8 + 40 * i)
                    if heap_bytes(name_off) == name:
                        return ohdr
                return None
            else:
                raise Hdf5FormatError('bad group node signature')
        raise Hdf5FormatError('group B-tree too deep')

    def _global_heap_object(self, coll_addr, index):
        buf = self._buf
        if buf[coll_addr:coll_addr + 4] != b'GCOL':
            raise Hdf5FormatError('bad global heap signature')
        csize = struct.unpack_from('<Q', buf, coll_addr + 8)[0]
        p = coll_addr + 16
        end = coll_addr + csize
        while p + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from('<HHIQ', buf, p)
            if idx == 0:
                break
            if idx == index:
                return buf[p + 16:p + 16 + osize]
            p += 16 + _pad8(osize)
        raise Hdf5FormatError('global heap object not found')

    def _parse_attribute(self, mbuf):
        version = mbuf[0]
        if version == 1:
            nsize, tsize, ssize = struct.unpack_from('<HHH', mbuf, 2)
            p = 8
            name = mbuf[p:p + nsize].split(b'\0')[0].decode()
            p += _pad8(nsize)
            dt = _parse_datatype(mbuf, p)
            p += _pad8(tsize)
            shape, _ = _parse_dataspace(mbuf, p)
            p += _pad8(ssize)
        elif version in (2, 3):
            nsize, tsize, ssize = struct.unpack_from('<HHH', mbuf, 2)
            p = 8 if version == 2 else 9
            name = mbuf[p:p + nsize].split(b'\0')[0].decode()
            p += nsize
            dt = _parse_datatype(mbuf, p)
            p += tsize
            shape, _ = _parse_dataspace(mbuf, p)
            p += ssize
        else:
            raise Hdf5FormatError('unsupported attribute version %d' % version)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if dt.vlen_string:
            vals = []
            for i in range(n):
                _length, coll, idx = struct.unpack_from('<IQI', mbuf, p + 16 * i)
                vals.append(bytes(self._global_heap_object(coll, idx)))
            value = vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        elif dt.dtype is not None:
            arr = np.frombuffer(mbuf[p:p + n * dt.dtype.itemsize], dt.dtype, count=n)
            if dt.cls == 3:
                arr = np.array([s.rstrip(b'\0') for s in arr], dtype=object)
            value = arr[0] if not shape else arr.reshape(shape).copy()
        else:
            raise Hdf5FormatError('unsupported attribute datatype')
        return name, value
