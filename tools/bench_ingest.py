"""FAST5 ingest throughput on the host cores (no GPU): multi-read files written by
poreplex_b200.hdf5_write -> packed int16 batch, through
  native : poreplex_b200.fast5_loader.load_batch (libpb_fast5.so, thread pool)
  python : poreplex_b200.hdf5_min, one read at a time (what Fast5Source does without h5py)

    python tools/bench_ingest.py [--reads 16000] [--length 4000] [--files 4] [--storage gzip1]
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from poreplex_b200 import fast5_loader as FL      # noqa: E402
from poreplex_b200 import hdf5_min as R           # noqa: E402
from poreplex_b200 import hdf5_write as W         # noqa: E402

STORAGE = {'contiguous': {}, 'chunked': dict(chunks=4096), 'gzip1': dict(chunks=4096, gzip=1),
           'gzip1-shuffle': dict(chunks=4096, gzip=1, shuffle=True)}


def vbz_storage():
    """VBZ (svb16 + zstd level 1), through the test suite's encoder: the decoders are checked only
    against that encoder, not against ONT's plugin."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
    from fast5_files import vbz_encoder
    return dict(chunks=4096, encoder=vbz_encoder(np.int16, 1, True, 1))


def make_files(tmp, n_files, per_file, length, storage, seed):
    rng = np.random.default_rng(seed)
    reads = []
    for k in range(n_files):
        root = W.Group(attrs={'file_version': b'2.0'})
        # random-walk levels + noise: compresses roughly like nanopore signal (~1.6x with gzip)
        for i in range(per_file):
            rid = '%08x-%06d' % (int(rng.integers(0, 2 ** 31)), i)
            lv = np.repeat(rng.normal(500, 60, length // 20 + 1), 20)[:length]
            sig = np.clip(lv + rng.normal(0, 12, length), 0, 2047).astype(np.int16)
            g = root.group('read_' + rid)
            g.group('Raw', attrs={'duration': np.int64(length), 'start_time': np.int64(i),
                                  'read_id': rid.encode()}).dataset('Signal', sig, **storage)
            g.group('channel_id', attrs={'channel_number': b'7', 'digitisation': 8192.0,
                                         'offset': 4.0, 'range': 1443.03, 'sampling_rate': 3012.0})
            g.group('tracking_id', attrs={'run_id': b'run', 'sample_id': b'sample'})
            reads.append((os.path.join(tmp, 'f%d.fast5' % k), rid))
        W.write_file(os.path.join(tmp, 'f%d.fast5' % k), root)
    return reads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=16000)
    ap.add_argument('--length', type=int, default=4000)
    ap.add_argument('--files', type=int, default=4)
    ap.add_argument('--storage', default='gzip1', choices=sorted(STORAGE) + ['vbz'])
    ap.add_argument('--threads', type=int, default=os.cpu_count() or 1)
    ap.add_argument('--python-reads', type=int, default=1000)
    args = ap.parse_args()
    FL.build()
    tmp = tempfile.mkdtemp(prefix='ingest_')
    storage = vbz_storage() if args.storage == 'vbz' else STORAGE[args.storage]
    reads = make_files(tmp, args.files, args.reads // args.files, args.length, storage, 1)
    file_bytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp))
    FL.load_batch(reads[:64], threads=args.threads)                       # warm-up (page cache, zlib)
    best = {}
    for threads in sorted({1, args.threads}):
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = FL.load_batch(reads, threads=threads)
            t.append(time.perf_counter() - t0)
        assert (out['status'] == 0).all()
        best[threads] = min(t)
    sample = reads[:args.python_reads]
    t0 = time.perf_counter()
    handles = {}
    for path, rid in sample:
        h = handles.get(path) or handles.setdefault(path, R.Hdf5File(path))
        node = h['read_%s/Raw/Signal' % rid]
        sig = node[0:len(node)]
        h['read_%s/channel_id' % rid].attrs['range']
    tp = time.perf_counter() - t0
    i = len(sample) - 1
    assert np.array_equal(sig, out['raw'][out['offsets'][i]:out['offsets'][i] + out['lengths'][i]])
    n = len(reads)
    line = {'metric': 'FAST5 ingest, reads/s (host)', 'reads': n, 'read_length': args.length,
            'files': args.files, 'storage': args.storage, 'file_bytes': file_bytes,
            'raw_bytes': int(2 * n * args.length), 'host_cores': os.cpu_count(),
            'native_reads_per_s': {str(k): n / v for k, v in best.items()},
            'native_raw_GBps': {str(k): 2e-9 * n * args.length / v for k, v in best.items()},
            'python_hdf5_min_reads_per_s': len(sample) / tp}
    print(json.dumps(line))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == '__main__':
    main()
