#!/usr/bin/env python3
"""Golden outputs of the reference's summary writers (poreplex/io.py SequencingSummaryWriter,
FinalSummaryTracker) for a seeded list of result dicts.  Needs /root/reference; run from the
repo root:  python tests/golden/make_summary_golden.py   -> tests/golden/summary_*.{json,txt}"""
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def make_results(seed=5, n=600):
    """Result dicts with the shape NanoporeRead.report gives them (signal_loader.py:165-198)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    fail_status = ['scaler_signal_too_short', 'sequence_too_short', 'irregular_fast5',
                   'basecall_table_incomplete', 'adapter_not_detected', 'not_basecalled',
                   'scaling_qc_fail', 'unknown_error']
    weights = np.array([40, 3, 1, 2, 25, 9, 14, 6], float)
    res = []
    for i in range(n):
        u = rng.random()
        e = {'filename': 'batch%d/read%05d.fast5' % (i // 100, i), 'read_id': 'r%08x' % int(rng.integers(1 << 31)),
             'run_id': 'run0', 'channel': str(int(rng.integers(1, 513))), 'start_time': round(float(rng.random() * 1e4), 3),
             'duration': int(rng.integers(4000, 60000)), 'num_events': int(rng.integers(100, 4000)),
             'sequence_length': int(rng.integers(0, 3000)), 'mean_qscore': round(float(rng.random() * 12), 2),
             'sample_id': 'sample'}
        if u < 0.02:
            res.append({'filename': e['filename'], 'status': 'disappeared'})      # no label, no read_id
            continue
        if u < 0.70:
            e['status'], e['label'] = 'okay', 'pass'
        elif u < 0.76:
            e['status'], e['label'] = 'unsplit_read', 'artifact'
        elif u < 0.80:
            e['status'] = 'scaler_signal_too_short'                               # stopped before stage C: no label
        else:
            e['status'], e['label'] = str(rng.choice(fail_status, p=weights / weights.sum())), 'fail'
        if 'label' in e and rng.random() < 0.55:
            e['barcode'] = int(rng.integers(0, 4))
            e['barcode_guess'] = e['barcode']
            e['barcode_score'] = int(rng.integers(19, 30))
        if e.get('label') == 'pass' and rng.random() < 0.8:
            e['polya'] = {'begin': 100, 'end': 900, 'dwell_time': float(rng.random() * 2), 'spikes': []}
        res.append(e)
    return res


def reference_outputs(results, config, label_names, barcode_names):
    from oracle import refshim
    refshim.install()
    import types
    ps = sys.modules['pysam']
    if not hasattr(ps, 'faidx'):
        ps.faidx = None
    from poreplex import io as rio
    tmp = tempfile.mkdtemp()
    w = rio.SequencingSummaryWriter(config, tmp, label_names, barcode_names)
    w.write_results(results)
    w.close()
    t = rio.FinalSummaryTracker(label_names, barcode_names)
    t.feed_results(results)
    buf = io.StringIO()
    t.print_results(buf)
    return open(os.path.join(tmp, 'sequencing_summary.txt')).read(), buf.getvalue()


CONFIGS = {
    'barcoding_polya': ({'barcoding': True, 'measure_polya': True, 'fast5_output': True},
                        {'fail': 'fail', 'pass': 'pass', 'artifact': 'artifact'},
                        {None: 'undetermined', 0: 'BC1', 1: 'BC2', 2: 'BC3', 3: 'BC4'}),
    'plain': ({'barcoding': False, 'measure_polya': False, 'fast5_output': False},
              {'fail': 'fail', 'pass': 'pass', 'artifact': 'artifact'}, {None: '-'}),
}


def main():
    results = make_results()
    for name, (config, labels, barcodes) in CONFIGS.items():
        res = results if config['barcoding'] else [
            {k: v for k, v in e.items() if not k.startswith('barcode')} for e in results]
        seq, final = reference_outputs(res, config, labels, barcodes)
        open(os.path.join(HERE, 'summary_%s_sequencing_summary.txt' % name), 'w').write(seq)
        open(os.path.join(HERE, 'summary_%s_final.txt' % name), 'w').write(final)
        print(name, len(seq.splitlines()), 'summary rows;', len(final.splitlines()), 'table lines')


if __name__ == '__main__':
    main()
