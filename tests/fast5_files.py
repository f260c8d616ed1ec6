"""Test helper: put the in-memory FAST5 trees of oracle/fake_fast5.py on disk as real HDF5 files
(poreplex_b200/hdf5_write.py), and a VBZ encoder for the self-consistency test of the VBZ filter."""
import ctypes as C
import struct

import numpy as np

from poreplex_b200 import hdf5_write as W


def tree_to_hdf5(node, signal_kw=None, move_kw=None, fastq_vlen=False):
    """oracle.refshim FakeGroup -> hdf5_write.Group.  ``signal_kw`` / ``move_kw``: storage of the
    Signal / Move datasets (chunks, gzip, shuffle, encoder); others are contiguous.
    ``fastq_vlen``: store Fastq as a variable-length string (h5py's form for str data) instead of
    a fixed-length one."""
    g = W.Group(attrs=dict(node.attrs))
    for name, child in node._children.items():
        if hasattr(child, '_children'):
            g.children[name] = tree_to_hdf5(child, signal_kw, move_kw, fastq_vlen)
        else:
            kw = {}
            if name == 'Signal' and signal_kw:
                kw = signal_kw
            elif name == 'Move' and move_kw:
                kw = move_kw
            data = child._data
            if name == 'Fastq' and fastq_vlen:
                data = bytes(data).decode()
            g.children[name] = W.Dataset(data, attrs=dict(child.attrs), **kw)
    return g


def write_fast5(path, tree, signal_kw=None, move_kw=None, fastq_vlen=False):
    return W.write_file(path, tree_to_hdf5(tree, signal_kw, move_kw, fastq_vlen))


def to_single_read(multi_tree, read_id):
    """The single-read layout of fast5_file.py:76-82 from one read group of a multi-read tree."""
    from oracle import refshim
    src = multi_tree['read_' + read_id]
    f5 = refshim.FakeFile()
    ugk = f5.add_group('UniqueGlobalKey')
    ugk._children['channel_id'] = src['channel_id']
    ugk._children['tracking_id'] = src['tracking_id']
    f5.add_group('Raw').add_group('Reads')._children['Read_17'] = src['Raw']
    if 'Analyses' in src:
        f5._children['Analyses'] = src['Analyses']
    return f5


# ---- VBZ (ONT filter 32020) encoder, written from the same published description the decoders
# follow: the test it serves is a SELF-CONSISTENCY test, not a check against ONT's plugin ----
def _zstd():
    lib = C.CDLL('libzstd.so.1')
    lib.ZSTD_compress.restype = C.c_size_t
    lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    lib.ZSTD_compressBound.restype = C.c_size_t
    lib.ZSTD_compressBound.argtypes = [C.c_size_t]
    return lib


def _svb_encode(vals, key_bits):
    vals = np.asarray(vals, np.uint32)
    per = 8 // key_bits
    if key_bits == 1:
        codes = (vals > 0xFF).astype(np.uint8)
    else:
        codes = ((vals > 0xFF).astype(np.uint8) + (vals > 0xFFFF) + (vals > 0xFFFFFF)).astype(np.uint8)
    nkeys = (len(vals) + per - 1) // per
    padded = np.zeros(nkeys * per, np.uint8)
    padded[:len(vals)] = codes
    keys = np.zeros(nkeys, np.uint8)
    for j in range(per):
        keys |= padded[j::per] << np.uint8(j * key_bits)
    data = bytearray()
    for v, c in zip(vals.tolist(), codes.tolist()):
        data += int(v).to_bytes(4, 'little')[:c + 1]
    return keys.tobytes() + bytes(data)


def vbz_encoder(dtype, version=1, zigzag=True, level=1):
    """``encoder`` triple for hdf5_write.Dataset: (filter id, cd_values, encode)."""
    dtype = np.dtype(dtype)
    isize = dtype.itemsize

    def encode(raw):
        a = np.frombuffer(raw, dtype)
        if isize == 2 and version == 1:
            u = a.view(np.uint16)
            if zigzag:
                d = np.diff(u, prepend=np.uint16(0)).astype(np.uint16).view(np.int16)
                u = ((d.astype(np.int32) << 1) ^ (d.astype(np.int32) >> 15)).astype(np.uint16)
            body = _svb_encode(u, 1)
        else:
            u = a.astype(np.int32 if dtype.kind == 'i' else np.uint32).view(np.uint32)
            if zigzag:
                d = np.diff(u, prepend=np.uint32(0)).astype(np.uint32).view(np.int32)
                u = ((d.astype(np.int64) << 1) ^ (d.astype(np.int64) >> 31)).astype(np.uint32)
            body = _svb_encode(u, 2)
        if level:
            z = _zstd()
            cap = z.ZSTD_compressBound(len(body))
            dst = C.create_string_buffer(cap)
            n = z.ZSTD_compress(dst, cap, body, len(body), level)
            body = dst.raw[:n]
        return struct.pack('<I', len(raw)) + body

    return (W.FILTER_VBZ, [version, isize, 1 if zigzag else 0, level], encode)
