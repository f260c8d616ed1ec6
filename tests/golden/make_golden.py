#!/usr/bin/env python3
"""Generate the golden fixtures by running the REFERENCE's own Python verbatim.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

For each fixture set, seeded synthetic reads are served to the unmodified reference
modules (poreplex.signal_analyzer.process_batch and everything below it) as in-memory
FAST5 trees through oracle/refshim.py.  The reference's missing third-party kernels
(pomegranate Viterbi, TensorFlow LSTM) are provided by the oracle's C restatement; its
event detector is the reference's own C (oracle/_ref).  Everything else -- pooling,
scaling, segment grouping, barcode window rule, decision rule, status/label logic,
result dicts and their order -- is the reference's code.

Written per set:
  <set>.npz   inputs (int16 signals, calibration, synthetic basecalls) + the float
              intermediates captured from inside the reference (scaling params,
              barcode windows) as raw bit patterns
  <set>.json  the reference's result dicts (list order preserved) + captured segments
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fake_fast5, refshim                     # noqa: E402
from poreplex_b200 import params, synth                    # noqa: E402

SETS = {
    # name: (preset variant, read length, reads, seed)
    'stock16k': ('stock', 16000, 40, 101),
    'short4k': ('bench-short', 4000, 88, 202),
    # long reads, half of them with a second leader+adapter+poly(A) planted in the transcript
    'chimera40k': ('stock', 40000, 18, 303),
}


def base_config(preset, inputdir, **switches):
    cfg = dict(preset)
    cfg.update({'inputdir': inputdir, 'outputdir': inputdir, 'barcoding': False,
                'measure_polya': False, 'trim_adapter': False, 'filter_unsplit_reads': False,
                'minimum_sequence_length': 10, 'dump_adapter_signals': False,
                'dump_basecalls': False, 'albacore_onthefly': False,
                'barcoding_quality_filter': 18})
    cfg.update(switches)
    km = os.path.join(inputdir, 'kmer.model')
    with open(km, 'w') as f:
        f.write('kmer\tlevel_mean\nAAAAA\t1.0\n')       # only len(index[0]) == 5 is used
    cfg['kmer_model'] = km
    return cfg


def build_inputs(variant, L, n, seed):
    preset = params.load_preset()
    if variant == 'bench-short':
        preset = params.bench_short_preset(preset)
    spec = synth.SynthSpec.for_length(L, frac_no_adapter=0.07, frac_qc_fail=0.07)
    rd = synth.to_numpy(synth.generate_reads(n, spec, preset, seed=seed))
    # every sixth read gets an adapter whose length straddles the demultiplexer's minimum
    # (260 pooled samples, 100 under bench-short): the length gate of barcoding.py:84-98
    spec2 = synth.SynthSpec.for_length(L, frac_no_adapter=0.0, frac_qc_fail=0.0)
    spec2.adapter_pooled = (245, 275) if variant == 'stock' else (90, 110)
    rd2 = synth.to_numpy(synth.generate_reads(n, spec2, preset, seed=seed + 7))
    for i in range(4, n, 6):
        for k in ('raw', 'range', 'digitisation', 'offset', 'gain', 'sampling_rate'):
            rd[k][i] = rd2[k][i]
        for k in rd['planted']:
            rd['planted'][k][i] = rd2['planted'][k][i]
    rng = np.random.default_rng(seed)
    if L >= 30000:
        plant_chimeras(rd, np.random.default_rng(seed + 1))
    lengths = np.full(n, L, np.int64)
    lengths[0] = 5000 if variant == 'stock' else 700          # scaler_signal_too_short
    lengths[1] = L - 7                                        # L % 15 != 0
    rd['length'] = lengths
    read_ids = ['%08x-%04x-4000-8000-%012x' % (seed, i, i) for i in range(n)]
    basecalls = [fake_fast5.synth_basecall(int(lengths[i]), rng) for i in range(n)]
    basecalls[2] = None                                       # not_basecalled
    # a read whose basecall is shorter than minimum_sequence_length
    short = fake_fast5.synth_basecall(int(lengths[3]), rng, p_move=0.0)
    basecalls[3] = short
    return preset, rd, read_ids, basecalls


def plant_chimeras(rd, rng, frac=0.6):
    """Copy a read's own leader+adapter+poly(A) block into its transcript (every other
    planted block is cut short so that it stays below the duration cut-offs)."""
    n, L = rd['raw'].shape
    for i in range(4, n):
        if rng.random() >= frac:
            continue
        b = rd['planted']['bounds'][i]
        blk0, blk1 = int(b[0]) * 15, int(b[4]) * 15
        blk = rd['raw'][i, blk0:blk1].copy()
        if rng.random() < 0.35:
            blk = blk[:len(blk) // 3]
        pos = int(rng.integers(blk1 + 1500, L - len(blk) - 3000))
        rd['raw'][i, pos:pos + len(blk)] = blk


def run_reference(preset, variant, rd, read_ids, basecalls, **switches):
    sa, sl, bcmod, polya, f5mod = refshim.reference_modules()
    refshim.reset_reference_persistence()
    refshim.clear_fast5()
    tmp = tempfile.mkdtemp(prefix='golden_')
    fake_fast5.build_fast5(tmp, 'reads.fast5', rd, read_ids, basecalls)
    cfg = base_config(preset, tmp, **switches)

    captured = {'segments': {}, 'scaling': {}, 'windows': {}}
    orig_detect = sa.SignalAnalysis.detect_segments
    orig_push = bcmod.BarcodeDemultiplexer.push
    orig_loader_init = sl.SignalLoader.__init__

    def detect(self, signal, elspan):
        seg = orig_detect(self, signal, elspan)
        captured['segments'][self.npread.read_id] = {k: [int(a), int(b)] for k, (a, b) in seg.items()}
        captured['scaling'][self.npread.read_id] = np.array(self.npread.scaling_params, np.float32)
        return seg

    def push(self, npread, signal):
        before = len(self.signals)
        orig_push(self, npread, signal)
        if len(self.signals) > before:
            captured['windows'][npread.read_id] = np.array(self.signals[-1], np.float32)

    def loader_init(self, config, fast5prefix):
        orig_loader_init(self, config, fast5prefix)
        if 'scaler_min_length_override' in config:            # bench-short: a data value
            self.scaler_cfg['min_length'] = config['scaler_min_length_override']

    sa.SignalAnalysis.detect_segments = detect
    bcmod.BarcodeDemultiplexer.push = push
    sl.SignalLoader.__init__ = loader_init
    try:
        reads = [('reads.fast5', rid) for rid in read_ids]
        reads.insert(5, ('missing.fast5', 'ffffffff-0000-4000-8000-000000000000'))   # disappeared
        reads.insert(9, ('reads.fast5', 'eeeeeeee-0000-4000-8000-000000000000'))     # unknown read
        results = sa.process_batch(0, reads, cfg)
    finally:
        sa.SignalAnalysis.detect_segments = orig_detect
        bcmod.BarcodeDemultiplexer.push = orig_push
        sl.SignalLoader.__init__ = orig_loader_init
    if isinstance(results, tuple):
        raise RuntimeError(results[1] + '\n' + results[2])
    return reads, results, captured


def jsonable(results):
    out = []
    for r in results:
        d = {}
        for k, v in r.items():
            if k == 'error_message':
                v = v.split('\n')[0]
            if k == 'sequence':
                v = [v[0], v[1], int(v[2])]
            if k == 'polya':
                v = {'begin': int(v['begin']), 'end': int(v['end']),
                     'dwell_time': float(v['dwell_time']),
                     'spikes': [[float(x) for x in s] for s in v['spikes']]}
            if isinstance(v, (np.integer,)):
                v = int(v)
            if isinstance(v, (np.floating,)):
                v = float(v)
            d[k] = v
        out.append(d)
    return out


def main():
    for name, (variant, L, n, seed) in SETS.items():
        preset, rd, read_ids, basecalls = build_inputs(variant, L, n, seed)
        reads, res_bc, cap = run_reference(preset, variant, rd, read_ids, basecalls,
                                           trim_adapter=True, barcoding=True)
        _, res_trim, _ = run_reference(preset, variant, rd, read_ids, basecalls,
                                       trim_adapter=True)
        _, res_polya, _ = run_reference(preset, variant, rd, read_ids, basecalls,
                                        trim_adapter=True, barcoding=True, measure_polya=True)
        full = None
        if variant == 'stock':
            _, full, _ = run_reference(preset, variant, rd, read_ids, basecalls,
                                       trim_adapter=True, barcoding=True, measure_polya=True,
                                       filter_unsplit_reads=True)
        win_ids = sorted(cap['windows'])
        np.savez_compressed(
            os.path.join(HERE, name + '.npz'),
            raw=rd['raw'], length=rd['length'], range=rd['range'],
            digitisation=rd['digitisation'], offset=rd['offset'],
            sampling_rate=rd['sampling_rate'], read_ids=np.array(read_ids),
            bc_present=np.array([b is not None for b in basecalls]),
            bc_moves=np.concatenate([b['moves'] if b else np.zeros(0, np.uint8) for b in basecalls]),
            bc_nmoves=np.array([len(b['moves']) if b else 0 for b in basecalls]),
            bc_seq=np.array([b['sequence'] if b else '' for b in basecalls]),
            bc_qual=np.array([b['qstring'] if b else '' for b in basecalls]),
            bc_meanq=np.array([b['mean_qscore'] if b else 0.0 for b in basecalls]),
            scaling_ids=np.array(sorted(cap['scaling'])),
            scaling_bits=np.array([cap['scaling'][k].view(np.uint32) for k in sorted(cap['scaling'])]),
            window_ids=np.array(win_ids),
            window_bits=np.array([cap['windows'][k].view(np.uint32) for k in win_ids]
                                 ).reshape(len(win_ids), -1))
        doc = {'set': name, 'preset': variant, 'read_length': L, 'seed': seed,
               'reads': [list(r) for r in reads],
               'results_trim_barcoding': jsonable(res_bc),
               'results_trim_only': jsonable(res_trim),
               'results_trim_barcoding_polya': jsonable(res_polya),
               'segments': cap['segments']}
        if full is not None:
            doc['results_all_switches'] = jsonable(full)
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(doc, f, indent=0, sort_keys=True)
        from collections import Counter
        print(name, Counter((r['status'], r.get('label')) for r in res_bc),
              'windows', len(win_ids),
              'barcodes', Counter(r.get('barcode', 'none') for r in res_bc))


if __name__ == '__main__':
    main()
