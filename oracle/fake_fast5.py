"""In-memory multi-read FAST5 trees for the oracle runs (TEST INFRASTRUCTURE).

Builds the HDF5 layout that ``poreplex/fast5_file.py`` reads on
top of oracle.refshim's fake h5py, from the tensors of poreplex_b200.synth plus a
synthetic guppy flip-flop basecall (``Move`` table, one row per 15 samples).
"""
import os

import numpy as np

from . import refshim


def synth_basecall(n_samples, rng, first_sample=0, block_stride=15, p_move=0.3):
    """A flip-flop style basecall: moves in {0,1}, moves[0] == 1, len(seq) == sum(moves)
    (fast5_file.py:183-197 requires kmer_size == 1 or 5)."""
    n_events = max((n_samples - first_sample) // block_stride, 1)
    moves = (rng.random(n_events) < p_move).astype(np.uint8)
    moves[0] = 1
    seqlen = int(moves.sum())
    seq = ''.join(rng.choice(list('ACGU'), size=seqlen))
    qual_vals = rng.integers(3, 31, size=seqlen)
    qstring = ''.join(chr(33 + int(q)) for q in qual_vals)
    mean_q = float(-10 * np.log10(np.mean(10 ** (-qual_vals / 10))))
    return {'moves': moves, 'sequence': seq, 'qstring': qstring, 'mean_qscore': mean_q,
            'first_sample': first_sample, 'block_stride': block_stride,
            'num_events': int(n_events)}


def add_read(f5, read_id, raw, digitisation, rng_range, offset, sampling_rate,
             channel='1', start_time=0, run_id='run0', sample_id='sample0', basecall=None,
             duration=None):
    """Append one read group (multi-read layout, fast5_file.py:70-75)."""
    g = f5.add_group('read_' + read_id)
    rawg = g.add_group('Raw', attrs={
        'duration': np.int64(len(raw) if duration is None else duration),
        'start_time': np.int64(start_time),
        'read_id': read_id.encode()})
    rawg.add_dataset('Signal', np.asarray(raw, np.int16))
    g.add_group('channel_id', attrs={
        'channel_number': str(channel).encode(), 'digitisation': float(digitisation),
        'offset': float(offset), 'range': float(rng_range),
        'sampling_rate': float(sampling_rate)})
    g.add_group('tracking_id', attrs={'run_id': run_id.encode(),
                                      'sample_id': sample_id.encode()})
    if basecall is not None:
        an = g.add_group('Analyses')
        bc = an.add_group('Basecall_1D_000')
        tmpl = bc.add_group('BaseCalled_template')
        fq = '@{}\n{}\n+\n{}\n'.format(read_id, basecall['sequence'], basecall['qstring'])
        tmpl.add_dataset('Fastq', fq.encode())
        tmpl.add_dataset('Move', np.asarray(basecall['moves'], np.uint8))
        summ = bc.add_group('Summary')
        summ.add_group('basecall_1d_template', attrs={
            'sequence_length': len(basecall['sequence']),
            'mean_qscore': basecall['mean_qscore'],
            'block_stride': basecall['block_stride']})
        seg = an.add_group('Segmentation_000')
        seg.add_group('Summary').add_group('segmentation', attrs={
            'num_events_template': basecall['num_events'],
            'first_sample_template': basecall['first_sample']})
    return g


def build_fast5(inputdir, filename, reads, read_ids, basecalls=None):
    """Register an in-memory FAST5 at ``inputdir/filename`` holding ``reads`` (numpy dict
    from synth.to_numpy, or a list of per-read dicts for ragged input) and create the
    empty on-disk file the reference's os.path.exists check needs
    (signal_analyzer.py:90)."""
    f5 = refshim.FakeFile()
    for i, rid in enumerate(read_ids):
        raw = reads['raw'][i]
        if 'length' in reads:
            raw = raw[:int(reads['length'][i])]
        add_read(f5, rid, raw, reads['digitisation'][i], reads['range'][i], reads['offset'][i],
                 reads['sampling_rate'][i], channel=str(1 + i % 512), start_time=1000 * i,
                 basecall=None if basecalls is None else basecalls[i])
    path = os.path.join(inputdir, filename)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, 'ab').close()
    refshim.register_fast5(path, f5)
    return f5
