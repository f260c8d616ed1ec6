"""FAST5 ingest without h5py: the HDF5 writer, the Python reader
(hdf5_min) and the native batch loader (libpb_fast5.so) against each other and against the
in-memory trees the oracle runs on.  There is no libhdf5 in this image, so "real" files are the
ones hdf5_write produces; the reader's group / attribute / contiguous-dataset code is also
exercised on the reference's own h5py-written model files (tests/test_host_cpu.py)."""
import os
import shutil

import numpy as np
import pytest

from fast5_files import write_fast5, to_single_read, vbz_encoder
from poreplex_b200 import fast5_loader as FL
from poreplex_b200 import hdf5_min as R
from poreplex_b200 import hdf5_write as W


@pytest.fixture(scope='module')
def lib():
    FL.build()
    return FL.load()


def _tree(n, seed=0, lengths=None, basecalls=True):
    from oracle import fake_fast5, refshim
    rng = np.random.default_rng(seed)
    f5 = refshim.FakeFile()
    sigs, ids = [], []
    for i in range(n):
        L = int(lengths[i]) if lengths is not None else int(rng.integers(900, 9000))
        raw = np.clip(rng.normal(500, 80, L), -32768, 32767).astype(np.int16)
        rid = '%08x-%04d-read' % (int(rng.integers(0, 2 ** 31)), i)
        bc = fake_fast5.synth_basecall(L, rng) if basecalls else None
        fake_fast5.add_read(f5, rid, raw, 8192.0, 1400.0 + i, 3.0 + (i % 7), 3012.0,
                            channel=str(1 + i % 512), start_time=1000 * i, run_id='run%d' % (i % 3),
                            sample_id='sample', basecall=bc)
        sigs.append(raw)
        ids.append(rid)
    return f5, ids, sigs


STORAGE = [None, dict(chunks=1000), dict(chunks=4096, gzip=1), dict(chunks=777, gzip=6, shuffle=True)]


def test_header_symbols_all_exported(lib):
    import re
    hdr = open(FL.HEADER).read()
    declared = sorted(set(re.findall(r'\b(pb2f_[a-z0-9_]+)\s*\(', hdr)))
    assert declared and sorted(FL.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pb2f_abi_version() == 1


@pytest.mark.parametrize('storage', STORAGE)
def test_python_reader_round_trip(tmp_path, storage):
    f5, ids, sigs = _tree(12, seed=1)
    path = str(tmp_path / 'multi.fast5')
    write_fast5(path, f5, signal_kw=storage, move_kw=dict(chunks=100, gzip=4) if storage else None)
    with R.Hdf5File(path) as h:
        assert 'UniqueGlobalKey' not in h
        assert sorted(h.keys()) == sorted('read_' + r for r in ids)
        for rid, sig in zip(ids, sigs):
            g = h['read_' + rid]
            node = g['Raw/Signal']
            assert len(node) == len(sig) and np.array_equal(node[0:len(node)], sig)
            assert g['Raw'].attrs['read_id'].decode() == rid
            assert int(g['Raw'].attrs['duration']) == len(sig)
            want = f5['read_' + rid]
            for grp in ('channel_id', 'tracking_id'):
                for k, v in want[grp].attrs.items():
                    got = g[grp].attrs[k]
                    assert (got.decode() == v.decode()) if isinstance(v, bytes) else (got == v), (grp, k)
            t = 'Analyses/Basecall_1D_000/BaseCalled_template/'
            assert g[t + 'Fastq'][()] == want[t + 'Fastq'][()]
            assert np.array_equal(g[t + 'Move'][()], want[t + 'Move'][()])
            assert g['Analyses/Basecall_1D_000'].name.rsplit('_', 1)[-1] == '000'


def test_python_reader_mapped_and_whole_file_paths_agree(tmp_path, monkeypatch):
    """hdf5_min reads small files whole and maps large ones; both paths must serve the same data."""
    f5, ids, sigs = _tree(8, seed=13)
    path = str(tmp_path / 'multi.fast5')
    write_fast5(path, f5, signal_kw=dict(chunks=1000, gzip=1, shuffle=True))
    for limit in (0, 1 << 30):
        monkeypatch.setattr(R, '_READ_WHOLE_BELOW', limit)
        with R.Hdf5File(path) as h:
            assert (h._fh is None) == (limit > 0)
            for rid, sig in zip(ids, sigs):
                assert np.array_equal(h['read_' + rid + '/Raw/Signal'][()], sig)
                assert h['read_' + rid + '/Raw'].attrs['read_id'].decode() == rid
            assert sorted(h.keys()) == sorted('read_' + r for r in ids)


def test_many_children_compound_and_empty(tmp_path):
    """> 256 children (two B-tree levels), a compound table, an empty group, empty datasets."""
    root = W.Group(attrs={'file_version': b'2.0', 'count': 700, 'ratio': 0.25,
                          'vec': np.arange(4, dtype=np.int32)})
    for i in range(700):
        root.group('read_%05d' % i, attrs={'i': i})
    ev = np.zeros(9, dtype=[('start', '<u8'), ('move', 'u1'), ('mean', '<f4'), ('kmer', 'S5')])
    ev['start'] = np.arange(9) * 15
    ev['mean'] = np.linspace(80, 120, 9)
    ev['kmer'] = b'ACGUA'
    root.group('tables').dataset('Events', ev)
    root.group('empty')
    root.dataset('nothing', np.zeros(0, np.int16))
    root.dataset('nothing_chunked', np.zeros(0, np.int16), chunks=64, gzip=1)
    path = str(tmp_path / 'wide.h5')
    W.write_file(path, root)
    with R.Hdf5File(path) as h:
        keys = h.keys()
        assert len(keys) == 704 and keys == sorted(keys)
        assert h.attrs['count'] == 700 and h.attrs['ratio'] == 0.25
        assert np.array_equal(h.attrs['vec'], np.arange(4))
        assert all(h['read_%05d' % i].attrs['i'] == i for i in (0, 1, 255, 256, 257, 699))
        got = h['tables/Events'][()]
        assert got.dtype.names == ev.dtype.names and all(np.array_equal(got[n], ev[n]) for n in ev.dtype.names)
        assert h['empty'].keys() == [] and len(h['nothing'][()]) == 0 and len(h['nothing_chunked'][()]) == 0


@pytest.mark.parametrize('storage', STORAGE)
def test_native_file_api(lib, tmp_path, storage):
    f5, ids, sigs = _tree(40, seed=2)
    path = str(tmp_path / 'multi.fast5')
    write_fast5(path, f5, signal_kw=storage)
    with FL.Fast5File(path) as f:
        assert f.is_multiread
        assert f.read_names() == sorted(ids)
        for rid, sig in zip(ids, sigs):
            m = f.meta(rid)
            want = f5['read_' + rid]
            assert m['signal_length'] == len(sig) == m['duration']
            assert m['read_id'] == rid and m['run_id'] == want['tracking_id'].attrs['run_id'].decode()
            assert m['channel_number'] == want['channel_id'].attrs['channel_number'].decode()
            for k in ('digitisation', 'offset', 'range', 'sampling_rate'):
                assert m[k] == want['channel_id'].attrs[k]
            assert m['start_time'] == want['Raw'].attrs['start_time']
            assert np.array_equal(f.signal(rid), sig)
        with pytest.raises(FL.Fast5Error):
            f.meta('no-such-read')
    assert FL.get_read_ids('multi.fast5', str(tmp_path)) == [('multi.fast5', r) for r in sorted(ids)]
    # single-read layout (fast5_file.py:76-82): read_id None = the first read
    spath = str(tmp_path / 'single.fast5')
    write_fast5(spath, to_single_read(f5, ids[3]), signal_kw=storage)
    assert FL.get_read_ids(spath) == [(spath, ids[3])]
    with FL.Fast5File(spath) as f:
        assert not f.is_multiread and f.read_names() == ['Read_17']
        assert f.meta()['read_id'] == ids[3]
        assert np.array_equal(f.signal(), sigs[3])
        assert np.array_equal(f.signal(ids[3]), sigs[3])
        with pytest.raises(FL.Fast5Error):          # fast5_file.py:105-108
            f.meta(ids[4])


def test_variable_length_string_attributes(lib, tmp_path):
    """h5py stores ``attrs[k] = 'text'`` as variable-length UTF-8 strings in a global heap (what
    ont_fast5_api-written files carry); MinKNOW-style fixed-length strings are the default case of
    the other tests.  Both readers must serve both."""
    from poreplex_b200 import fast5_source as FS
    import sys
    f5, ids, sigs = _tree(6, seed=7, basecalls=False)
    for rid in ids:
        g = f5['read_' + rid]
        g['Raw'].attrs['read_id'] = rid                                  # str -> vlen
        for k in ('run_id', 'sample_id'):
            g['tracking_id'].attrs[k] = g['tracking_id'].attrs[k].decode() + '-\u00b5'
        g['channel_id'].attrs['channel_number'] = g['channel_id'].attrs['channel_number'].decode()
        # newer MinKNOW files carry an enumeration attribute next to the ones the path reads, and
        # there may be types outside the subset altogether:
        # neither may hide the rest of the group
        g['Raw'].attrs['end_reason'] = W.Enum({'unknown': 0, 'signal_positive': 2, 'signal_negative': 3}, 2)
        import struct
        opaque = struct.pack('<BBBBI', 0x15, 0, 0, 0, 4) + b'tag\0'.ljust(8, b'\0')   # class 5: opaque
        g['Raw'].attrs['aaa_unknown_type'] = W.RawAttr(opaque, struct.pack('<BBBBI', 1, 0, 0, 0, 0), b'\1\2\3\4')
        g['channel_id'].attrs['zzz_unknown_type'] = W.RawAttr(opaque, struct.pack('<BBBBI', 1, 0, 0, 0, 0), b'\1\2\3\4')
    path = str(tmp_path / 'vlen.fast5')
    write_fast5(path, f5, signal_kw=dict(chunks=2048, gzip=1))
    out = FL.load_batch([(path, r) for r in ids], full_meta=True)
    assert (out['status'] == FL.READ_OK).all()
    with R.Hdf5File(path) as h:
        for i, rid in enumerate(ids):
            want = f5['read_' + rid]
            m = out['meta'][i]
            assert m['read_id'] == rid and m['run_id'] == want['tracking_id'].attrs['run_id']
            assert m['sample_id'] == 'sample-\u00b5' and m['channel_number'] == want['channel_id'].attrs['channel_number']
            assert h['read_' + rid + '/tracking_id'].attrs['run_id'].decode() == want['tracking_id'].attrs['run_id']
            assert h['read_' + rid + '/Raw'].attrs['read_id'].decode() == rid
            assert h['read_' + rid + '/Raw'].attrs['end_reason'] == 2
            assert 'aaa_unknown_type' not in h['read_' + rid + '/Raw'].attrs
            assert h['read_' + rid + '/channel_id'].attrs['range'] == want['channel_id'].attrs['range']
            assert int(h['read_' + rid + '/Raw'].attrs['duration']) == len(sigs[i])
            assert np.array_equal(out['raw'][out['offsets'][i]:out['offsets'][i] + out['lengths'][i]], sigs[i])


def test_native_batch_loader(lib, tmp_path):
    """Several files, every storage form, missing / foreign / truncated files and unknown reads:
    packed layout equals SignalEngine.pack_reads', statuses follow signal_analyzer.py:90-92 and
    signal_loader.py:200-207."""
    from poreplex_b200.engine import SignalEngine
    reads, want = [], []
    for k, storage in enumerate(STORAGE):
        f5, ids, sigs = _tree(30, seed=10 + k)
        path = str(tmp_path / ('f%d.fast5' % k))
        write_fast5(path, f5, signal_kw=storage)
        reads += [(path, r) for r in ids]
        want += sigs
    order = np.random.default_rng(0).permutation(len(reads))
    reads = [reads[i] for i in order]
    want = [want[i] for i in order]
    garbage = str(tmp_path / 'garbage.fast5')
    open(garbage, 'wb').write(b'this is not HDF5' * 100)
    empty = str(tmp_path / 'empty.fast5')
    open(empty, 'wb').close()
    bad = [(str(tmp_path / 'gone.fast5'), 'x'), (garbage, 'x'), (empty, 'x'), (reads[0][0], 'unknown-read')]
    allreads = reads[:50] + bad + reads[50:]
    for threads in (1, 4):
        out = FL.load_batch(allreads, threads=threads, full_meta=True)
        st = out['status']
        assert list(st[50:54]) == [FL.READ_DISAPPEARED, FL.READ_IRREGULAR, FL.READ_IRREGULAR, FL.READ_IRREGULAR]
        ok = np.ones(len(allreads), bool)
        ok[50:54] = False
        assert (st[ok] == FL.READ_OK).all() and (out['lengths'][~ok] == 0).all()
        assert (out['offsets'] % 8 == 0).all()
        got = [out['raw'][o:o + n] for o, n in zip(out['offsets'][ok], out['lengths'][ok])]
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        # the same layout SignalEngine.pack_reads builds from in-memory arrays
        praw, poff, plen = SignalEngine.pack_reads(want)
        assert np.array_equal(poff, out['offsets'][ok])       # unreadable reads take no space
        assert np.array_equal(plen, out['lengths'][ok])
        assert [m['read_id'] for m, k in zip(out['meta'], ok) if k] == [r for _, r in reads]
        assert (out['range'][ok] >= 1400.0).all() and (out['digitisation'][ok] == 8192.0).all()
    assert FL.load_batch([])['raw'].size >= 0


def test_batch_of_single_read_files_under_a_low_descriptor_limit(lib, tmp_path):
    """Older runs are one file per read: a batch then opens thousands of files.  The loader keeps
    mappings, not descriptors, so RLIMIT_NOFILE does not bound the batch size."""
    import resource
    f5, ids, sigs = _tree(600, seed=9, lengths=np.full(600, 950), basecalls=False)
    reads = []
    for i, rid in enumerate(ids):
        path = str(tmp_path / ('single_%04d.fast5' % i))
        write_fast5(path, to_single_read(f5, rid), signal_kw=dict(chunks=1024, gzip=1))
        reads.append((path, rid if i % 2 else None))
    soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
    resource.setrlimit(resource.RLIMIT_NOFILE, (128, hard))
    try:
        out = FL.load_batch(reads, threads=4)
    finally:
        resource.setrlimit(resource.RLIMIT_NOFILE, (soft, hard))
    assert (out['status'] == FL.READ_OK).all()
    for i in (0, 1, 299, 599):
        assert np.array_equal(out['raw'][out['offsets'][i]:out['offsets'][i] + out['lengths'][i]], sigs[i])


def test_many_open_sources_under_a_low_descriptor_limit(tmp_path, monkeypatch):
    """The drop-in keeps one Fast5Source per loaded read until the batch ends.  Without h5py the
    sources of one multi-read file share a single mapping and single-read files are read whole, so
    a large batch does not run into RLIMIT_NOFILE."""
    import resource
    import sys
    from poreplex_b200 import fast5_source as FS
    monkeypatch.setitem(sys.modules, 'h5py', None)
    f5, ids, sigs = _tree(400, seed=12, lengths=np.full(400, 950), basecalls=False)
    mpath = str(tmp_path / 'multi.fast5')
    write_fast5(mpath, f5, signal_kw=dict(chunks=1024, gzip=1))
    singles = []
    for i, rid in enumerate(ids[:300]):
        path = str(tmp_path / ('s%03d.fast5' % i))
        write_fast5(path, to_single_read(f5, rid))
        singles.append((path, rid))
    soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
    resource.setrlimit(resource.RLIMIT_NOFILE, (64, hard))
    try:
        srcs = [FS.Fast5Source(mpath, r) for r in ids] + [FS.Fast5Source(p, r) for p, r in singles]
        assert len(FS._SharedFile._open) == 1 + len(singles)
        for k in (0, 399, 400, 699):
            assert np.array_equal(srcs[k].raw_int16(), sigs[k if k < 400 else k - 400])
        for s in srcs:
            s.close()
        assert not FS._SharedFile._open
    finally:
        resource.setrlimit(resource.RLIMIT_NOFILE, (soft, hard))


def test_truncated_files_never_crash(lib, tmp_path):
    f5, ids, sigs = _tree(6, seed=3)
    path = str(tmp_path / 'whole.fast5')
    size = write_fast5(path, f5, signal_kw=dict(chunks=512, gzip=1))
    blob = open(path, 'rb').read()
    for cut in (8, 95, 96, 200, size // 4, size // 2, size - 9000, size - 100, size - 1):
        tpath = str(tmp_path / ('cut_%d.fast5' % cut))
        open(tpath, 'wb').write(blob[:cut])
        out = FL.load_batch([(tpath, r) for r in ids], threads=2)
        for i, (s, n) in enumerate(zip(out['status'], out['lengths'])):
            if s == FL.READ_OK:                     # whatever still decodes must be right
                assert np.array_equal(out['raw'][out['offsets'][i]:out['offsets'][i] + n], sigs[i])
            else:
                assert s == FL.READ_IRREGULAR and n == 0
    # flipped bytes inside the compressed chunks: zlib notices, the read becomes irregular
    dam = bytearray(blob)
    for p in range(size // 2, size // 2 + 4000, 37):
        dam[p] ^= 0x5A
    dpath = str(tmp_path / 'damaged.fast5')
    open(dpath, 'wb').write(bytes(dam))
    out = FL.load_batch([(dpath, r) for r in ids])
    assert set(out['status']) <= {FL.READ_OK, FL.READ_IRREGULAR}


def test_mutated_files_never_crash(lib, tmp_path):
    """Random byte damage anywhere in the file (headers, B-trees, heaps, chunk data): every read
    comes back okay or irregular_fast5; sizes taken from the damaged file never drive a crash, an
    endless walk or a giant allocation (bounds-checked map + sanity limits in fast5_loader.cpp)."""
    f5, ids, sigs = _tree(5, seed=6)
    path = str(tmp_path / 'whole.fast5')
    rng = np.random.default_rng(123)
    for storage in (dict(chunks=512, gzip=1, shuffle=True),
                    dict(chunks=1000, encoder=vbz_encoder(np.int16, 1, True, 1)), None):
        size = write_fast5(path, f5, signal_kw=storage)
        blob = np.frombuffer(open(path, 'rb').read(), np.uint8)
        mpath = str(tmp_path / 'mutant.fast5')
        for trial in range(150):
            m = blob.copy()
            if trial % 3 == 0:                      # a burst in the metadata-heavy tail of the file
                p0 = int(rng.integers(size // 2, size - 64))
                m[p0:p0 + 64] = rng.integers(0, 256, 64)
            else:                                   # scattered single bytes, biased to 0x00 / 0xFF
                pos = rng.integers(8, size, int(rng.integers(1, 40)))
                m[pos] = rng.choice([0, 255, 1, 128, int(rng.integers(0, 256))], len(pos))
            m.tofile(mpath)
            out = FL.load_batch([(mpath, r) for r in ids] + [(mpath, None)], threads=2)
            assert set(out['status']) <= {FL.READ_OK, FL.READ_IRREGULAR}
            assert (out['lengths'] >= 0).all() and out['lengths'].sum() < 1 << 24
            try:
                with R.Hdf5File(mpath) as h:        # the Python reader: exceptions only
                    for rid in ids:
                        node = h['read_' + rid + '/Raw/Signal']
                        if node.shape and node.shape[0] < 1 << 20:
                            node[0:len(node)]
            except Exception:
                pass


@pytest.mark.parametrize('version,zigzag,level', [(1, True, 1), (1, True, 0), (0, True, 3), (1, False, 1)])
def test_vbz_self_consistency(lib, tmp_path, version, zigzag, level):
    """VBZ (filter 32020) decoders of both readers against an encoder written from the same
    description -- NOT against ONT's plugin, which does not exist in this image."""
    f5, ids, sigs = _tree(8, seed=4)
    path = str(tmp_path / 'vbz.fast5')
    write_fast5(path, f5, signal_kw=dict(chunks=2000, encoder=vbz_encoder(np.int16, version, zigzag, level)))
    with R.Hdf5File(path) as h:
        for rid, sig in zip(ids, sigs):
            assert np.array_equal(h['read_' + rid + '/Raw/Signal'][()], sig)
    out = FL.load_batch([(path, r) for r in ids])
    assert (out['status'] == FL.READ_OK).all()
    for i, sig in enumerate(sigs):
        assert np.array_equal(out['raw'][out['offsets'][i]:out['offsets'][i] + out['lengths'][i]], sig)


def test_native_reader_walks_h5py_written_files(lib):
    """The reference's model files were written by h5py / libhdf5: the native group walker must
    get through their symbol tables (no reads in them, so the answer is an empty list)."""
    root = '/root/reference/poreplex/presets/MIN106-RNA001'
    if not os.path.isdir(root):
        pytest.skip('reference tree not present')
    for name in ('scaler-r3.hdf5', 'demux-tetra-r4.hdf5'):
        with FL.Fast5File(os.path.join(root, name)) as f:
            assert f.is_multiread and f.read_names() == []
        with R.Hdf5File(os.path.join(root, name)) as h:
            assert 'model_weights' in h


@pytest.mark.parametrize('fastq_vlen', [False, True])
def test_fast5_source_over_hdf5_min(tmp_path, monkeypatch, fastq_vlen):
    """Fast5Source (the drop-in's FAST5 access) over real files through hdf5_min gives what it
    gives over the in-memory h5py the golden runs use."""
    import sys
    from oracle import refshim
    from poreplex_b200 import fast5_source as FS
    f5, ids, sigs = _tree(5, seed=5)
    path = str(tmp_path / 'reads.fast5')
    write_fast5(path, f5, signal_kw=dict(chunks=1024, gzip=1, shuffle=True), move_kw=dict(chunks=64, gzip=1),
                fastq_vlen=fastq_vlen)
    refshim.install_fake_h5py()
    refshim.clear_fast5()
    refshim.register_fast5(path, f5)
    via_fake = [FS.Fast5Source(path, r) for r in ids]
    monkeypatch.setitem(sys.modules, 'h5py', None)          # "import h5py" now fails
    via_min = [FS.Fast5Source(path, r) for r in ids]
    for a, b in zip(via_fake, via_min):
        for k in ('duration', 'start_time', 'read_id', 'channel_number', 'digitization', 'offset',
                  'range', 'sampling_rate', 'run_id', 'sample_id', 'is_multiread'):
            assert getattr(a, k) == getattr(b, k), k
        assert np.array_equal(a.raw_int16(), b.raw_int16())
        sa, sb = a.get_basecall(want_events=True), b.get_basecall(want_events=True)
        assert set(sa) == set(sb)
        for k in sa:
            if k == 'events':
                assert set(sa[k]) == set(sb[k])
                assert all(np.array_equal(sa[k][c], sb[k][c]) for c in sa[k])
            else:
                assert sa[k] == sb[k], k
        b.close()


# ---- the inflater behind the deflate filter (csrc_host/inflate_fast.h) against zlib -----------
def _inflate(lib, src, cap):
    import ctypes as C
    dst = C.create_string_buffer(max(cap, 1))
    n = lib.pb2f_inflate(src, len(src), dst, cap)
    return n, dst.raw[:max(n, 0)]


def _inflate_cases(rng):
    yield b''
    yield b'a'
    for n in (1, 2, 7, 8, 9, 15, 16, 17, 100, 255, 256, 257, 258, 259, 270, 271, 272, 273, 300, 511,
              512, 600, 1000, 4096, 8192, 65535, 65536, 65537):
        yield bytes(rng.integers(0, 256, n, dtype=np.uint8))                        # incompressible
        yield bytes(rng.integers(0, 4, n, dtype=np.uint8))                          # short codes
        yield (b'the quick brown fox jumps over the lazy dog. ' * (n // 40 + 1))[:n]  # long matches
        yield bytes(n)                                                              # distance-1 runs
        yield (np.repeat(rng.normal(500, 60, n // 40 + 1), 20)[:n // 2] +
               rng.normal(0, 12, n // 2)).astype(np.int16).tobytes()                # signal-like
        yield bytes(np.tile(np.arange(7, dtype=np.uint8), n // 7 + 1)[:n])          # overlapping copies
        yield bytes(np.tile(np.arange(3, dtype=np.uint8), n // 3 + 1)[:n])


def test_inflater_matches_zlib(lib):
    """Every block type (stored, fixed, dynamic), compression level and strategy zlib offers,
    multi-block streams, small windows, sizes around the fast-loop margins (16 input / 272 output
    bytes); exact-size and oversized destinations; too-small destinations and truncated streams are
    errors."""
    import zlib
    rng = np.random.default_rng(0)
    n_variants = 0
    for data in _inflate_cases(rng):
        variants = [zlib.compress(data, lvl) for lvl in (0, 1, 6, 9)]
        for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strat)
            variants.append(co.compress(data) + co.flush())
        co, parts, step = zlib.compressobj(1), b'', max(len(data) // 5, 1)
        for i in range(0, len(data), step):
            parts += co.compress(data[i:i + step]) + co.flush(zlib.Z_FULL_FLUSH)
        variants.append(parts + co.flush())
        co = zlib.compressobj(9, zlib.DEFLATED, 9, 1)                                # 512-byte window
        variants.append(co.compress(data) + co.flush())
        for v in variants:
            for cap in (len(data), len(data) + 1000):
                n, out = _inflate(lib, v, cap)
                assert n == len(data) and out == data, (len(data), n)
            if data:
                assert _inflate(lib, v, len(data) - 1)[0] == -5                      # PB2F_ENOSPC
            for cut in (1, len(v) // 2, len(v) - 5, len(v) - 1):
                if 0 <= cut < len(v):
                    assert _inflate(lib, v[:cut], len(data) + 10)[0] == -3, cut      # PB2F_EFORMAT
            n_variants += 1
    assert n_variants > 1500


def test_inflater_survives_corruption(lib):
    """Bit flips anywhere in a stream: an error or -- never observed -- the right bytes; no crash,
    no write past the destination (the buffer is sized exactly + a guard)."""
    import zlib
    rng = np.random.default_rng(1)
    data = bytes(rng.integers(0, 64, 5000, dtype=np.uint8)) + b'xyz' * 700
    for lvl in (1, 6):
        v = zlib.compress(data, lvl)
        for _ in range(3000):
            m = bytearray(v)
            for _k in range(int(rng.integers(1, 4))):
                m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            n, out = _inflate(lib, bytes(m), len(data) + 300)
            assert n < 0 or out == data


def test_svb16_encoder_writes_vbz_chunk_bodies():
    """fast5_loader.svb16_encode (the compressed upload form, pb2_batch.packed) produces exactly
    the streamvbyte body of a VBZ version-1 chunk without the zstd stage: the same bytes as the
    test suite's own VBZ encoder at level 0 after its 4-byte size header."""
    from fast5_files import vbz_encoder
    from poreplex_b200 import fast5_loader
    fast5_loader.build()
    rng = np.random.default_rng(11)
    sigs = [rng.integers(-32768, 32768, size=n).astype(np.int16) for n in (1, 8, 9, 4000, 4099)]
    sigs.append((np.cumsum(rng.integers(-30, 31, size=6000)) + 400).astype(np.int16))
    lens = np.array([len(s) for s in sigs], np.int64)
    pad = (lens + 7) // 8 * 8
    off = np.zeros(len(sigs), np.int64)
    off[1:] = np.cumsum(pad)[:-1]
    raw = np.zeros(int(pad.sum()) + 8, np.int16)
    for s, o in zip(sigs, off):
        raw[o:o + len(s)] = s
    pk, po = fast5_loader.svb16_encode(raw, off, lens, threads=3)
    assert np.all(po % 16 == 0) and len(po) == len(sigs) + 1
    encode = vbz_encoder(np.int16, version=1, zigzag=True, level=0)[2]
    for i, s in enumerate(sigs):
        chunk = encode(s.tobytes())
        assert chunk[:4] == np.uint32(2 * len(s)).tobytes()
        body = chunk[4:]
        assert pk[po[i]:po[i] + len(body)].tobytes() == body
        assert po[i + 1] - po[i] == (len(body) + 15) // 16 * 16


def test_load_batch_can_hand_over_the_compressed_form(tmp_path):
    """load_batch(packed=True): the same reads as streamvbyte-16 streams beside the int16 batch."""
    from oracle import fake_fast5, refshim
    from fast5_files import write_fast5
    from poreplex_b200 import fast5_loader
    fast5_loader.build()
    rng = np.random.default_rng(4)
    tree = refshim.FakeFile()
    ids = ['%08x-0000-4000-8000-%012x' % (7, i) for i in range(5)]
    sigs = [(np.cumsum(rng.integers(-20, 21, size=n)) + 450).astype(np.int16) for n in (4000, 4001, 16, 9000, 123)]
    for rid, sgn in zip(ids, sigs):
        fake_fast5.add_read(tree, rid, sgn, 8192.0, 1400.0, 5.0, 3012.0)
    write_fast5(str(tmp_path / 'r.fast5'), tree, signal_kw=dict(chunks=2048, gzip=1, shuffle=True))
    b = fast5_loader.load_batch([('r.fast5', r) for r in ids], inputdir=str(tmp_path), threads=2, packed=True)
    pk, po = b['packed']
    assert len(po) == 6 and po[-1] < 0.65 * 2 * sum(len(x) for x in sigs) + 16 * 6
    again, off = fast5_loader.svb16_encode(b['raw'], b['offsets'], b['lengths'])
    assert np.array_equal(off, po) and np.array_equal(again[:po[-1]], pk[:po[-1]])


def test_self_referencing_btree_nodes_end_the_walk(lib, tmp_path):
    """A crafted group / chunk B-tree node that names itself as its child must not send the loader
    into a walk of used ** depth nodes: node levels have to fall by one per step, so the walk
    stops at the first inconsistency (and has a node budget besides)."""
    import struct
    import time
    f5, ids, sigs = _tree(700, seed=9, lengths=[64] * 700, basecalls=False)   # enough children for a 2-level group tree
    path = str(tmp_path / 'whole.fast5')
    write_fast5(path, f5, signal_kw=dict(chunks=16, gzip=1))
    blob = bytearray(open(path, 'rb').read())
    pos, patched = 0, {0: 0, 1: 0}
    while True:
        pos = blob.find(b'TREE', pos)
        if pos < 0:
            break
        ntype, used = blob[pos + 4], struct.unpack_from('<H', blob, pos + 6)[0]
        if ntype in (0, 1) and used > 0:
            # every entry's child pointer -> the node itself, entries used = the maximum
            klen = 8 if ntype == 0 else 24
            struct.pack_into('<H', blob, pos + 6, max(used, 16))
            for i in range(min(max(used, 16), 16)):
                at = pos + 24 + i * (klen + 8) + klen
                if at + 8 <= len(blob):
                    struct.pack_into('<Q', blob, at, pos)
            blob[pos + 5] = 3                                   # claims to be an inner node
            patched[ntype] += 1
        pos += 4
    assert patched[0] >= 1 and patched[1] >= 1
    mpath = str(tmp_path / 'loop.fast5')
    open(mpath, 'wb').write(bytes(blob))
    t0 = time.perf_counter()
    out = FL.load_batch([(mpath, r) for r in ids[:40]], threads=2)
    assert set(out['status']) <= {FL.READ_OK, FL.READ_IRREGULAR}
    try:
        with FL.Fast5File(mpath) as f:
            f.read_names()
    except FL.Fast5Error:
        pass
    assert time.perf_counter() - t0 < 20.0
