#!/usr/bin/env python3
"""Time the REFERENCE's own Python (process_batch and everything below it, run verbatim from
/root/reference) over the shims of oracle/refshim.py, under a ProcessPoolExecutor exactly as
pipeline.py:96,204 drives it -- the CPU baseline BASELINE.md section 4 describes.  The missing
third-party kernels (TensorFlow LSTM, pomegranate Viterbi) are the oracle's C restatement behind
the shims, so this flatters the reference.  Build-container only (needs /root/reference; the GPU
box has none): the result is committed under profiles/.

    python tools/cpu_reference_python.py [--reads 1024] [--length 4000] [--workers N]
"""
import argparse
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

_state = {}


def _init(tmp, length, n, seed, switches):
    """Per worker: the shims, the reads as an in-memory FAST5, the config."""
    from oracle import fake_fast5, refshim
    from poreplex_b200 import params, synth
    import make_golden as MG
    sa, sl, _, _, _ = refshim.reference_modules()
    preset = params.load_preset()
    short = length < 10500
    if short:
        preset = params.bench_short_preset(preset)
        orig = sl.SignalLoader.__init__

        def loader_init(self, config, fast5prefix):
            orig(self, config, fast5prefix)
            self.scaler_cfg['min_length'] = config['scaler_min_length_override']
        sl.SignalLoader.__init__ = loader_init
    rd = synth.to_numpy(synth.generate_reads(n, synth.SynthSpec.for_length(length), preset, seed=seed))
    rng = np.random.default_rng(seed)
    ids = ['%08x-%04x-4000-8000-%012x' % (seed, i, i) for i in range(n)]
    bcs = [fake_fast5.synth_basecall(length, rng) for _ in range(n)]
    fake_fast5.build_fast5(tmp, 'reads.fast5', rd, ids, bcs)
    cfg = MG.base_config(preset, tmp, **switches)
    _state.update(sa=sa, cfg=cfg, ids=ids)


def _batch(args):
    batchid, lo, hi = args
    res = _state['sa'].process_batch(batchid, [('reads.fast5', r) for r in _state['ids'][lo:hi]], _state['cfg'])
    if isinstance(res, tuple):
        raise RuntimeError(res[1])
    return [r['status'] for r in res]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=1024)
    ap.add_argument('--length', type=int, default=4000)
    ap.add_argument('--workers', type=int, default=os.cpu_count())
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--full', action='store_true', help='all four switches (configs[3])')
    a = ap.parse_args()
    sw = dict(trim_adapter=True, barcoding=True)
    if a.full:
        sw.update(measure_polya=True, filter_unsplit_reads=True)
    tmp = tempfile.mkdtemp(prefix='refpy_')
    jobs = [(b, lo, min(lo + a.batch, a.reads)) for b, lo in enumerate(range(0, a.reads, a.batch))]
    with ProcessPoolExecutor(a.workers, initializer=_init,
                             initargs=(tmp, a.length, a.reads, 20261017, sw)) as ex:
        list(ex.map(_batch, jobs[:a.workers]))                   # warm-up: models loaded per worker
        t0 = time.perf_counter()
        statuses = [s for part in ex.map(_batch, jobs) for s in part]
        dt = time.perf_counter() - t0
    from collections import Counter
    print(json.dumps({
        'kind': 'reference-python-over-shims', 'reads': a.reads, 'read_length': a.length,
        'switches': sorted(sw), 'workers': a.workers, 'batch_size': a.batch, 'seconds': dt,
        'reads_per_s': a.reads / dt, 'reads_per_s_per_core': a.reads / dt / a.workers,
        'status_mix': dict(Counter(statuses)),
        'note': 'unmodified poreplex.signal_analyzer.process_batch from /root/reference under '
                'ProcessPoolExecutor (pipeline.py:96,204); h5py = in-memory trees, TensorFlow / '
                'pomegranate = the oracle C restatement behind shims (faster than the real '
                'packages), csupport = the reference scrappie C; measured in the build container'}))


if __name__ == '__main__':
    main()
