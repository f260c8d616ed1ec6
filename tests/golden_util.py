"""Helpers shared by the tests that use the committed golden fixtures."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    with open(os.path.join(GOLDEN_DIR, name + '.json')) as f:
        doc = json.load(f)
    return z, doc


def golden_reads(z):
    """numpy dict in the layout of synth.to_numpy (+ ragged lengths)."""
    return {k: z[k] for k in ('raw', 'length', 'range', 'digitisation', 'offset', 'sampling_rate')}


def golden_basecalls(z):
    out, pos = [], 0
    for i in range(len(z['read_ids'])):
        nm = int(z['bc_nmoves'][i])
        if not z['bc_present'][i]:
            out.append(None)
        else:
            out.append({'moves': z['bc_moves'][pos:pos + nm], 'sequence': str(z['bc_seq'][i]),
                        'qstring': str(z['bc_qual'][i]), 'mean_qscore': float(z['bc_meanq'][i]),
                        'first_sample': 0, 'block_stride': 15, 'num_events': nm})
        pos += nm
    return out


def pack_golden(z):
    """(raw, offsets, lengths) with 16-byte aligned reads, like SignalEngine.pack_reads."""
    lengths = z['length'].astype(np.int64)
    padded = (lengths + 7) // 8 * 8
    offsets = np.zeros(len(lengths), np.int64)
    offsets[1:] = np.cumsum(padded)[:-1]
    raw = np.zeros(int(padded.sum()) + 8, np.int16)
    for i, (o, n) in enumerate(zip(offsets, lengths)):
        raw[o:o + n] = z['raw'][i][:n]
    return raw, offsets, lengths


def normalise_result(r):
    """Result dict -> comparable form (error text keeps its first line only)."""
    d = dict(r)
    if 'error_message' in d:
        d['error_message'] = True
    if 'sequence' in d:
        d['sequence'] = [d['sequence'][0], d['sequence'][1], int(d['sequence'][2])]
    if 'polya' in d:
        p = d['polya']
        d['polya'] = {'begin': int(p['begin']), 'end': int(p['end']),
                      'dwell_time': float(p['dwell_time']),
                      'spikes': [tuple(float(x) for x in s) for s in p['spikes']]}
    return d
