"""SURVEY.md section 8f rank 3 on the GPU: guppy Move table -> event-table columns
(pb2_derive_event_tables_host; fast5_file.py:183-230, signal_analyzer.py:311-326) against the
numpy restatement that tests/test_oracle_cpu.py pins to the reference's own Fast5Reader /
SignalAnalysis.load_events.  Floats as raw bit patterns."""
import numpy as np
import pytest

from golden_util import load_golden, golden_basecalls, pack_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['stock16k', 'chimera40k', 'short4k'])
def test_event_columns_match_restatement(name, eng_stock):
    from oracle import events_restated as ER
    z, _ = load_golden(name)
    bcs = golden_basecalls(z)
    raw, off, ln = pack_golden(z)
    idx = [i for i, b in enumerate(bcs) if b is not None]
    n = len(idx)
    sub_raw, sub_off, sub_ln = eng_stock.pack_reads([raw[off[i]:off[i] + ln[i]] for i in idx])
    batch = (sub_raw, sub_off, sub_ln, z['range'][idx], z['digitisation'][idx], z['offset'][idx])
    ss = np.stack([0.9 + 0.002 * np.arange(n), 3.0 + 0.05 * np.arange(n)], axis=1).astype(np.float32)
    # flip-flop (k-mer size 1) basecalls as they are, and the same reads re-labelled as 5-mer
    # models (sequence 4 longer than the sum of the moves)
    for kmer_pad in ('', 'ACGU'):
        seqs = [bcs[i]['sequence'] + kmer_pad for i in idx]
        quals = [bcs[i]['qstring'] + '5' * len(kmer_pad) for i in idx]
        tables, err = eng_stock.derive_event_tables_host(
            batch, [bcs[i]['moves'] for i in idx], [bcs[i]['first_sample'] for i in idx], 15,
            sequences=seqs, qstrings=quals, scale_shift=ss)
        assert not err.any()
        for k, i in enumerate(idx):
            want = ER.derive_event_table(raw[off[i]:off[i] + ln[i]], z['range'][i], z['digitisation'][i],
                                         z['offset'][i], bcs[i]['moves'], seqs[k], quals[k],
                                         bcs[i]['first_sample'], 15, ss[k])
            got = tables[k]
            for col in ('mean', 'stdv', 'scaled_mean'):
                w = np.asarray(want[col], np.float32)
                assert np.array_equal(got[col].view(np.uint32), w.view(np.uint32)), (col, i)
            for col in ('start', 'end', 'length', 'pos'):
                assert np.array_equal(got[col], np.asarray(want[col], np.int64)), (col, i)
            assert np.array_equal(got['p_model_state'].view(np.uint64), want['p_model_state'].view(np.uint64))
            assert np.array_equal(got['model_state'], want['model_state']), i
    assert n >= 10


def test_event_table_error_codes(eng_stock):
    """fast5_file.py:197 (unknown k-mer size) and :221 (events vs raw strides) as per-read codes."""
    z, _ = load_golden('stock16k')
    bcs = golden_basecalls(z)
    raw, off, ln = pack_golden(z)
    i = next(k for k, b in enumerate(bcs) if b is not None and k > 4)
    sig = raw[off[i]:off[i] + ln[i]]
    r3, o3, l3 = eng_stock.pack_reads([sig, sig, sig[:len(sig) - 40]])
    batch = (r3, o3, l3, z['range'][[i] * 3], z['digitisation'][[i] * 3], z['offset'][[i] * 3])
    b = bcs[i]
    seqs = [b['sequence'], b['sequence'][:-2], b['sequence']]
    quals = [b['qstring'], b['qstring'][:-2], b['qstring']]
    _, err = eng_stock.derive_event_tables_host(batch, [b['moves']] * 3, [0, 0, 0], 15,
                                                sequences=seqs, qstrings=quals)
    assert err.tolist() == [0, 1, 2]
