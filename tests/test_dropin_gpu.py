"""Drop-in boundary on the GPU: poreplex_b200.signal_analyzer.process_batch, fed the golden
reads as in-memory FAST5 trees, must return the SAME result dicts in the SAME order as
the reference's own process_batch did when the fixtures were made
(tests/golden/make_golden.py): BASELINE config 1 (--trim-adapter only) and with
--barcoding."""
import os
import tempfile

import numpy as np
import pytest

from golden_util import load_golden, golden_reads, golden_basecalls, normalise_result

pytestmark = pytest.mark.gpu


def _config(preset, inputdir, **sw):
    cfg = dict(preset)
    cfg.update({'inputdir': inputdir, 'outputdir': inputdir, 'barcoding': False,
                'measure_polya': False, 'trim_adapter': False, 'filter_unsplit_reads': False,
                'minimum_sequence_length': 10, 'dump_adapter_signals': False,
                'dump_basecalls': False, 'albacore_onthefly': False,
                'barcoding_quality_filter': 18})
    cfg.update(sw)
    return cfg


def _serve(name):
    from oracle import fake_fast5, refshim
    refshim.install_fake_h5py()
    refshim.clear_fast5()
    z, doc = load_golden(name)
    tmp = tempfile.mkdtemp(prefix='dropin_')
    fake_fast5.build_fast5(tmp, 'reads.fast5', golden_reads(z), [str(s) for s in z['read_ids']],
                           golden_basecalls(z))
    return z, doc, tmp


@pytest.mark.parametrize('name', ['stock16k', 'short4k', 'chimera40k'])
def test_process_batch_matches_reference_dicts(name, preset, preset_short, eng_stock, eng_short):
    import torch
    assert torch.cuda.is_available()
    from poreplex_b200 import signal_analyzer as sa
    z, doc, tmp = _serve(name)
    p = preset_short if doc['preset'] == 'bench-short' else preset
    reads = [tuple(r) for r in doc['reads']]
    for key, sw in (('results_trim_only', dict(trim_adapter=True)),
                    ('results_trim_barcoding', dict(trim_adapter=True, barcoding=True)),
                    ('results_trim_barcoding_polya', dict(trim_adapter=True, barcoding=True,
                                                          measure_polya=True)),
                    ('results_all_switches', dict(trim_adapter=True, barcoding=True,
                                                  measure_polya=True, filter_unsplit_reads=True))):
        if key not in doc:
            continue
        got = sa.process_batch(0, reads, _config(p, tmp, **sw))
        assert not isinstance(got, tuple), got
        want = doc[key]
        assert len(got) == len(want)
        for g, w in zip(got, want):                 # list ORDER is part of the contract
            g = normalise_result(g)
            w = normalise_result(w)
            assert g == w
    import pickle
    pickle.loads(pickle.dumps(got))                 # crosses a process pool in pipeline.py


@pytest.mark.parametrize('name', ['short4k', 'chimera40k'])
def test_process_batch_over_real_fast5_files(name, preset, preset_short, eng_stock, eng_short,
                                             monkeypatch, tmp_path):
    """The same golden comparison with the reads in REAL FAST5 files on disk (gzip + shuffle
    chunked Signal / Move datasets) and no h5py at all: raw signals come through the native batch
    loader (libpb_fast5.so), metadata and basecall tables through poreplex_b200.hdf5_min."""
    import sys
    from oracle import fake_fast5, refshim
    from fast5_files import write_fast5
    from poreplex_b200 import fast5_loader, fast5_source, signal_analyzer as sa
    z, doc = load_golden(name)
    tree = refshim.FakeFile()
    ids = [str(s) for s in z['read_ids']]
    reads_np, basecalls = golden_reads(z), golden_basecalls(z)
    for i, rid in enumerate(ids):
        raw = reads_np['raw'][i]
        if 'length' in reads_np:
            raw = raw[:int(reads_np['length'][i])]
        fake_fast5.add_read(tree, rid, raw, reads_np['digitisation'][i], reads_np['range'][i],
                            reads_np['offset'][i], reads_np['sampling_rate'][i],
                            channel=str(1 + i % 512), start_time=1000 * i,
                            basecall=None if basecalls is None else basecalls[i])
    write_fast5(str(tmp_path / 'reads.fast5'), tree, signal_kw=dict(chunks=4096, gzip=1, shuffle=True),
                move_kw=dict(chunks=512, gzip=1))
    fast5_loader.build()
    monkeypatch.setitem(sys.modules, 'h5py', None)              # "import h5py" fails from here on
    assert fast5_source._h5py() is fast5_source._MinimalH5py
    p = preset_short if doc['preset'] == 'bench-short' else preset
    reads = [tuple(r) for r in doc['reads']]
    calls = []
    real = fast5_loader.load_batch
    monkeypatch.setattr(fast5_loader, 'load_batch', lambda *a, **k: calls.append(1) or real(*a, **k))
    checked = 0
    for key, sw in (('results_trim_barcoding', dict(trim_adapter=True, barcoding=True)),
                    ('results_all_switches', dict(trim_adapter=True, barcoding=True,
                                                  measure_polya=True, filter_unsplit_reads=True))):
        if key not in doc:
            continue
        got = sa.process_batch(0, reads, _config(p, str(tmp_path), **sw))
        assert not isinstance(got, tuple), got
        want = doc[key]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert normalise_result(g) == normalise_result(w)
        checked += 1
    assert checked and len(calls) == checked                    # the native loader did serve the batches


def test_unbuilt_switches_fail_loudly(preset, eng_stock):
    from poreplex_b200 import signal_analyzer as sa
    z, doc, tmp = _serve('stock16k')
    for sw in ('dump_basecalls', 'dump_adapter_signals'):
        res = sa.process_batch(0, [tuple(r) for r in doc['reads']], _config(preset, tmp, **{sw: True}))
        assert isinstance(res, tuple) and res[0] == -1 and 'NotImplementedError' in res[1]


def test_quality_filter_out_of_range_is_batch_error(preset, eng_stock):
    """App. G12: --barcoding-quality-filter > 28 -> ValueError -> (-1, msg, tb)."""
    from poreplex_b200 import signal_analyzer as sa
    z, doc, tmp = _serve('stock16k')
    cfg = _config(preset, tmp, barcoding=True)
    cfg['barcoding_quality_filter'] = 29
    res = sa.process_batch(0, [tuple(r) for r in doc['reads']][:4], cfg)
    assert isinstance(res, tuple) and res[0] == -1 and 'ValueError' in res[1]
