"""CPU restatement of the guppy event-table derivation (TEST INFRASTRUCTURE).

Follows, line for line, with the same numpy / scipy calls the reference makes:

  Fast5Reader.construct_events_from_moves   fast5_file.py:183-207
  Fast5Reader.convert_events_guppy          fast5_file.py:209-230
  SignalAnalysis.load_events (derived cols) signal_analyzer.py:311-326

numpy IS the reference's implementation of this arithmetic, so the restatement is the
reference minus pandas and h5py; tests/test_oracle_cpu.py pins it against the reference's own
``Fast5Reader.get_basecall`` + ``SignalAnalysis.load_events`` running over oracle/refshim.py.
"""
import numpy as np
from scipy.signal import medfilt


class EventTableError(Exception):
    pass


def derive_event_table(raw, rng, digitisation, offset, moves, sequence, qstring, first_sample,
                       block_stride=15, scaling_params=None):
    """-> dict of columns (numpy arrays) for one read."""
    moves = np.asarray(moves)
    # construct_events_from_moves, fast5_file.py:183-207
    pos = moves.cumsum() - 1
    kmer_size = len(sequence) - int(moves.sum()) + 1
    revseq = sequence[::-1].replace('U', 'T')
    qual = 1 - 10 ** -((np.frombuffer(qstring.encode(), 'B') - 33) / 10)
    if kmer_size == 5:
        posshift = 2
    elif kmer_size == 1:
        revseq = '__' + revseq + '__'
        posshift = 0
    else:
        raise EventTableError('Move table is encoded with an unknown kmer-size.')
    ev = {
        'model_state': np.array([revseq[int(x):int(x) + 5] for x in pos], dtype='S5'),
        'p_model_state': np.array([qual[int(x) + posshift] for x in pos], np.float64),
        'move': moves,
    }
    # convert_events_guppy, fast5_file.py:209-230
    n = len(moves)
    last_sample = first_sample + block_stride * n
    ev['start'] = np.arange(first_sample, last_sample, block_stride)
    rawsig = np.asarray(raw)[first_sample:min(last_sample, len(raw))]
    rawdata = np.array(rng / digitisation * (rawsig + offset), dtype=np.float32)   # :130-131
    rawdata = medfilt(rawdata, 5)
    if len(rawdata) % block_stride > 0:
        rawdata = np.pad(rawdata, [0, block_stride - len(rawdata) % block_stride], 'constant',
                         constant_values=[np.nan, np.nan])
    if len(rawdata) // block_stride != n:
        raise EventTableError('Numbers of events and raw data strides does not match.')
    sigbyevents = rawdata.reshape([n, block_stride])
    ev['mean'] = sigbyevents.mean(axis=1)
    ev['stdv'] = sigbyevents.std(axis=1)
    ev['length'] = np.full(n, block_stride)
    # load_events, signal_analyzer.py:311-326
    if scaling_params is not None:
        ev['scaled_mean'] = np.poly1d(np.asarray(scaling_params, np.float32))(ev['mean'])
    ev['pos'] = np.cumsum(ev['move'])
    duration = np.hstack((np.diff(ev['start']), [1])).astype(np.int64)
    ev['end'] = ev['start'] + duration
    return ev
