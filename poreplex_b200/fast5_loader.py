"""FAST5 ingest without h5py: ctypes over libpb_fast5.so (include/poreplex_b200_fast5.h).

``load_batch(reads)`` turns a list of ``(path, read_id)`` pairs -- what ``process_batch`` receives
(signal_analyzer.py:46-58) -- into the packed arrays ``SignalEngine.analyze_host`` takes: one
ragged int16 buffer (every read on a 16-byte boundary), offsets, lengths and the calibration
triple, decoded by a pool of threads.  It stands where the reference opens one h5py handle per
read (``NanoporeRead.load``, signal_loader.py:200-210; ``Fast5Reader.__init__`` / ``load_metadata``
/ the dataset read of ``get_raw_data``, fast5_file.py:65-131).  Host-side I/O only; the int16 -> pA
conversion stays on the GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc_host', 'fast5_loader.cpp')
SRC_DEPS = [os.path.join(HERE, 'csrc_host', 'inflate_fast.h')]
LIB_PATH = os.path.join(HERE, 'libpb_fast5.so')
HEADER = os.path.join(HERE, '..', 'include', 'poreplex_b200_fast5.h')

READ_OK, READ_DISAPPEARED, READ_IRREGULAR = 0, 1, 2
STATUS_NAMES = {READ_OK: 'okay', READ_DISAPPEARED: 'disappeared', READ_IRREGULAR: 'irregular_fast5'}

EXPORTS = ['pb2f_abi_version', 'pb2f_last_error', 'pb2f_open', 'pb2f_close', 'pb2f_is_multiread',
           'pb2f_num_reads', 'pb2f_read_name', 'pb2f_read_meta_get', 'pb2f_read_signal',
           'pb2f_batch_open', 'pb2f_batch_meta', 'pb2f_batch_meta_full', 'pb2f_batch_plan',
           'pb2f_batch_read', 'pb2f_batch_close', 'pb2f_inflate', 'pb2f_svb16_plan',
           'pb2f_svb16_encode']


class Fast5Error(Exception):
    pass


class ReadMeta(C.Structure):
    _fields_ = [('signal_length', C.c_int64), ('duration', C.c_int64), ('start_time', C.c_int64),
                ('digitisation', C.c_double), ('offset', C.c_double), ('range', C.c_double),
                ('sampling_rate', C.c_double), ('read_id', C.c_char * 64),
                ('channel_number', C.c_char * 16), ('run_id', C.c_char * 64),
                ('sample_id', C.c_char * 64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        for k in ('read_id', 'channel_number', 'run_id', 'sample_id'):
            d[k] = d[k].decode()
        return d


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in [SRC, HEADER] + SRC_DEPS)


def build(force=False):
    """g++ -> poreplex_b200/libpb_fast5.so (host code only: pthreads, dlopen of libzstd)."""
    if force or needs_build():
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-Wall', '-fPIC', '-shared', '-pthread',
                               '-o', LIB_PATH, SRC, '-ldl'])
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('poreplex_b200: %s is not built (run __graft_entry__.build())' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64p, dp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)
        L.pb2f_last_error.restype = C.c_char_p
        L.pb2f_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.pb2f_close.argtypes = [vp]
        L.pb2f_close.restype = None
        L.pb2f_is_multiread.argtypes = [vp]
        L.pb2f_num_reads.argtypes = [vp]
        L.pb2f_num_reads.restype = C.c_int64
        L.pb2f_read_name.argtypes = [vp, C.c_int64]
        L.pb2f_read_name.restype = C.c_char_p
        L.pb2f_read_meta_get.argtypes = [vp, C.c_char_p, C.POINTER(ReadMeta)]
        L.pb2f_read_signal.argtypes = [vp, C.c_char_p, vp, C.c_int64]
        L.pb2f_read_signal.restype = C.c_int64
        L.pb2f_batch_open.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int64,
                                      C.c_int, C.POINTER(vp)]
        L.pb2f_batch_meta.argtypes = [vp, C.POINTER(C.c_int32), i64p, dp, dp, dp, dp, i64p, i64p]
        L.pb2f_batch_meta_full.argtypes = [vp, C.c_int64, C.POINTER(ReadMeta)]
        L.pb2f_batch_plan.argtypes = [vp, i64p, i64p]
        L.pb2f_batch_plan.restype = C.c_int64
        L.pb2f_batch_read.argtypes = [vp, vp, C.c_int64, i64p, C.c_int]
        L.pb2f_batch_close.argtypes = [vp]
        L.pb2f_batch_close.restype = None
        L.pb2f_inflate.argtypes = [vp, C.c_int64, vp, C.c_int64]
        L.pb2f_inflate.restype = C.c_int64
        L.pb2f_svb16_plan.argtypes = [vp, i64p, i64p, C.c_int64, C.c_int, i64p]
        L.pb2f_svb16_plan.restype = C.c_int64
        L.pb2f_svb16_encode.argtypes = [vp, i64p, i64p, C.c_int64, C.c_int, i64p, vp]
        if L.pb2f_abi_version() != 1:
            raise RuntimeError('libpb_fast5.so: ABI version mismatch')
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise Fast5Error('%s (rc=%d)' % (load().pb2f_last_error().decode(), rc))
    return rc


class Fast5File:
    """One FAST5 file (single- or multi-read)."""

    def __init__(self, path):
        self.lib = load()
        self.handle = C.c_void_p()
        _check(self.lib.pb2f_open(os.fsencode(path), C.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.pb2f_close(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def is_multiread(self):
        return bool(self.lib.pb2f_is_multiread(self.handle))

    def read_names(self):
        n = _check(self.lib.pb2f_num_reads(self.handle))
        return [self.lib.pb2f_read_name(self.handle, i).decode() for i in range(n)]

    def meta(self, read_id=None):
        m = ReadMeta()
        _check(self.lib.pb2f_read_meta_get(self.handle, None if read_id is None else read_id.encode(),
                                           C.byref(m)))
        return m.as_dict()

    def signal(self, read_id=None):
        rid = None if read_id is None else read_id.encode()
        n = self.meta(read_id)['signal_length']
        out = np.empty(n, np.int16)
        _check(self.lib.pb2f_read_signal(self.handle, rid, out.ctypes.data_as(C.c_void_p), n))
        return out


def get_read_ids(filename, basedir=None):
    """``fast5_file.get_read_ids`` (fast5_file.py:37-58): the ``(filename, read_id)`` pairs of
    one file -- what pipeline.py:324,363 queues for ``process_batch``."""
    path = filename if basedir is None else os.path.join(basedir, filename)
    with Fast5File(path) as f:
        if f.is_multiread:
            return [(filename, rid) for rid in f.read_names()]
        try:
            return [(filename, f.meta()['read_id'])]
        except Fast5Error:                                   # the reference's KeyError branch
            return []


def _ptr(a, typ):
    return a.ctypes.data_as(C.POINTER(typ))


def load_batch(reads, inputdir=None, threads=None, pinned=False, full_meta=False, packed=False):
    """``reads``: list of ``(path, read_id)`` (``read_id`` None = first read of a single-read
    file); ``inputdir`` is joined in front of relative paths.  Returns a dict with ``raw``
    (int16, packed), ``offsets``, ``lengths`` (0 for unreadable reads), ``range``,
    ``digitisation``, ``offset``, ``sampling_rate``, ``duration``, ``start_time``, ``status``
    (READ_OK / READ_DISAPPEARED / READ_IRREGULAR) and, with ``full_meta``, ``meta`` (list of
    dicts).  ``pinned=True`` allocates ``raw`` in page-locked memory (needs torch + CUDA).
    ``packed=True`` adds ``packed`` = (uint8 buffer, offsets): the batch as streamvbyte-16
    streams for the compressed upload (``svb16_encode``)."""
    lib = load()
    n = len(reads)
    threads = int(threads or min(32, os.cpu_count() or 1))
    paths = [os.fsencode(p if inputdir is None else os.path.join(inputdir, p)) for p, _ in reads]
    ids = [None if r is None else r.encode() for _, r in reads]
    c_paths = (C.c_char_p * max(n, 1))(*paths)
    c_ids = (C.c_char_p * max(n, 1))(*ids)
    handle = C.c_void_p()
    _check(lib.pb2f_batch_open(c_paths, c_ids, n, threads, C.byref(handle)))
    try:
        out = {'status': np.zeros(n, np.int32), 'lengths': np.zeros(n, np.int64),
               'offsets': np.zeros(n, np.int64)}
        for k in ('range', 'digitisation', 'offset', 'sampling_rate'):
            out[k] = np.zeros(n, np.float64)
        for k in ('duration', 'start_time'):
            out[k] = np.zeros(n, np.int64)
        total = _check(lib.pb2f_batch_plan(handle, _ptr(out['offsets'], C.c_int64),
                                           _ptr(out['lengths'], C.c_int64)))
        if pinned:
            import torch
            raw = torch.zeros(max(total, 8), dtype=torch.int16, pin_memory=True).numpy()
        else:
            raw = np.zeros(max(total, 8), np.int16)
        _check(lib.pb2f_batch_read(handle, raw.ctypes.data_as(C.c_void_p), raw.size,
                                   _ptr(out['offsets'], C.c_int64), threads))
        sig_len = np.zeros(n, np.int64)
        _check(lib.pb2f_batch_meta(handle, _ptr(out['status'], C.c_int32), _ptr(sig_len, C.c_int64),
                                   _ptr(out['range'], C.c_double), _ptr(out['digitisation'], C.c_double),
                                   _ptr(out['offset'], C.c_double), _ptr(out['sampling_rate'], C.c_double),
                                   _ptr(out['duration'], C.c_int64), _ptr(out['start_time'], C.c_int64)))
        out['lengths'][out['status'] != READ_OK] = 0       # reads that failed while decoding
        out['raw'] = raw
        if full_meta:
            metas = []
            for i in range(n):
                m = ReadMeta()
                lib.pb2f_batch_meta_full(handle, i, C.byref(m))
                metas.append(m.as_dict())
            out['meta'] = metas
        if packed:
            out['packed'] = svb16_encode(raw, out['offsets'], out['lengths'], threads=threads,
                                         pinned=pinned)
        return out
    finally:
        lib.pb2f_batch_close(handle)


def svb16_encode(raw, offsets, lengths, threads=None, pinned=False):
    """int16 batch -> (packed uint8 buffer, packed_offsets[n + 1]): one streamvbyte-16 stream of
    zigzag deltas per read (the body of an ONT VBZ chunk without its zstd stage), the compressed
    upload form ``SignalEngine.analyze_host(..., packed=...)`` / ``pb2_batch.packed`` take."""
    import numpy as np
    L = load()
    raw = np.ascontiguousarray(raw, np.int16)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lengths = np.ascontiguousarray(lengths, np.int64)
    n = len(lengths)
    threads = int(threads or min(32, os.cpu_count() or 1))
    i64p = C.POINTER(C.c_int64)
    poff = np.zeros(n + 1, np.int64)
    total = _check(L.pb2f_svb16_plan(raw.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(i64p),
                                     lengths.ctypes.data_as(i64p), n, threads, poff.ctypes.data_as(i64p)))
    if pinned:
        import torch
        packed = torch.empty(total + 16, dtype=torch.uint8, pin_memory=True).numpy()
    else:
        packed = np.empty(total + 16, np.uint8)
    _check(L.pb2f_svb16_encode(raw.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(i64p),
                               lengths.ctypes.data_as(i64p), n, threads, poff.ctypes.data_as(i64p),
                               packed.ctypes.data_as(C.c_void_p)))
    return packed, poff
