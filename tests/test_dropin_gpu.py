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
    for sw in ('albacore_onthefly',):
        res = sa.process_batch(0, [tuple(r) for r in doc['reads']], _config(preset, tmp, **{sw: True}))
        assert isinstance(res, tuple) and res[0] == -1 and 'NotImplementedError' in res[1]


def test_dump_switches_write_the_reference_layout(preset, eng_stock):
    """--dump-adapter-signals / --dump-basecalled-events
    (signal_analyzer.py:155-211,450-466).  The part files are read back the way the inventory
    builders (io.py:334-376) and the barcode training loader
    (training/barcodes/scripts/prepare_training_data.py:63-83) read them: the adapter signals,
    normalised and padded by the loader's own numpy recipe, must be the windows captured from
    inside the reference's BarcodeDemultiplexer.push; the event tables must be the reference's
    (oracle/events_restated.py, pinned to Fast5Reader / load_events by the CPU suite)."""
    import glob
    from oracle import events_restated as ER
    from poreplex_b200 import signal_analyzer as sa
    from poreplex_b200.hdf5_min import Hdf5File
    from golden_util import golden_basecalls
    z, doc, tmp = _serve('stock16k')
    reads = [tuple(r) for r in doc['reads']]
    cfg = _config(preset, tmp, trim_adapter=True, barcoding=True, measure_polya=True,
                  dump_adapter_signals=True, dump_basecalls=True)
    got = sa.process_batch(7, reads, cfg)
    assert not isinstance(got, tuple), got
    want = doc['results_trim_barcoding_polya']
    assert [normalise_result(g) for g in got] == [normalise_result(w) for w in want]
    by_id = {r['read_id']: r for r in want if 'read_id' in r}
    ids = [str(s) for s in z['read_ids']]

    # ---- adapter dumps: adapter/<batchid>/<read_id> + catalog/adapter/<batchid>
    parts = glob.glob(os.path.join(tmp, 'adapter-dumps', 'part-*.h5'))
    assert len(parts) == 1
    with Hdf5File(parts[0]) as h5:
        grp = h5['adapter/00000007']
        catalog = h5['catalog/adapter/00000007'][()]
        assert catalog.dtype.names == ('read_id', 'start', 'end')
        dumped = {rid: np.asarray(grp[rid][:]) for rid in grp.keys()}
    assert sorted(dumped) == sorted(c.decode() for c in catalog['read_id'])
    segs = doc['segments']
    for c in catalog:
        a0, a1 = segs[c['read_id'].decode()]['adapter']
        assert (int(c['start']), int(c['end'])) == (a0 * 15, (a1 + 1) * 15)
    assert set(dumped) == {rid for rid, sg in segs.items() if 'adapter' in sg}

    def normalize_signal(sig):                   # prepare_training_data.py:63-66
        med = np.median(sig)
        mad = np.median(np.abs(sig - med))
        return (sig - med) / max(0.01, (mad * 1.4826))
    checked = 0
    for k, rid in enumerate(str(s) for s in z['window_ids']):
        signal = dumped[rid]
        assert signal.dtype == np.float32
        if len(signal) < 300:                    # prepare_training_data.py:74-82
            w = np.pad(normalize_signal(signal), (300 - len(signal), 0), 'constant', constant_values=-1000.)
        elif len(signal) > 300:
            w = normalize_signal(signal[-300:])
        else:
            w = normalize_signal(signal)
        assert np.array_equal(w.astype(np.float32).view(np.uint32), z['window_bits'][k]), rid
        checked += 1
    assert checked >= 10

    # ---- basecalled events: basecalled_events/<batchid>/<read_id> (+ attributes)
    parts = glob.glob(os.path.join(tmp, 'events', 'part-*.h5'))
    assert len(parts) == 1
    bcs = golden_basecalls(z)
    scaling = {str(i): b.view(np.float32) for i, b in zip(z['scaling_ids'], z['scaling_bits'])}
    n_ev = 0
    with Hdf5File(parts[0]) as h5:
        grp = h5['basecalled_events/00000007']
        names = list(grp.keys())
        # every read that got through load_events: everything but the early exits, the reads
        # without an adapter and the read without a basecall
        expect = {rid for rid, r in by_id.items()
                  if r['status'] in ('okay', 'sequence_too_short', 'unsplit_read')}
        assert set(names) == expect
        for rid in names[::3]:
            i = ids.index(rid)
            tbl = grp[rid][()]
            assert tbl.dtype.names == ('mean', 'start', 'stdv', 'length', 'model_state', 'move',
                                       'pos', 'end', 'scaled_mean')
            assert [tbl.dtype[n].str for n in tbl.dtype.names] == \
                ['<f4', '<u8', '<f4', '<u8', '|S5', '<i4', '<u8', '<u8', '<f8']
            ss = scaling[rid]
            raw = z['raw'][i][:int(z['length'][i])]
            ref = ER.derive_event_table(raw, z['range'][i], z['digitisation'][i], z['offset'][i],
                                        bcs[i]['moves'], bcs[i]['sequence'], bcs[i]['qstring'], 0, 15, ss)
            for col in ('mean', 'stdv'):
                assert np.array_equal(tbl[col].view(np.uint32), np.asarray(ref[col], np.float32).view(np.uint32))
            assert np.array_equal(tbl['scaled_mean'], np.asarray(ref['scaled_mean'], np.float64))
            for col in ('start', 'length', 'move', 'pos', 'end'):
                assert np.array_equal(tbl[col].astype(np.int64), np.asarray(ref[col], np.int64)), col
            assert np.array_equal(tbl['model_state'], ref['model_state'])
            attrs = grp[rid].attrs
            assert np.float32(attrs['signal_scale']) == ss[0] and np.float32(attrs['signal_shift']) == ss[1]
            a0, a1 = segs[rid]['adapter']
            assert (int(attrs['adapter_begin']), int(attrs['adapter_end'])) == (a0 * 15, (a1 + 1) * 15)
            if 'polya' in by_id[rid]:
                assert int(attrs['polya_begin']) == by_id[rid]['polya']['begin']
                assert int(attrs['polya_end']) == by_id[rid]['polya']['end']
            n_ev += 1
    assert n_ev >= 8


def test_engine_survives_repickled_configs(preset, eng_stock):
    """pipeline.py:204 pickles `config` anew for every batch: equal content must reuse ONE engine
    (worker_persistence.py:46-58 keeps one set of models per process), a changed parameter must
    not, and the superseded engine is closed."""
    import pickle
    from poreplex_b200 import engine, signal_analyzer as sa
    z, doc, tmp = _serve('stock16k')
    reads = [tuple(r) for r in doc['reads']][:12]
    cfg = _config(preset, tmp, trim_adapter=True, barcoding=True)
    created = []
    real = engine.SignalEngine.__init__

    def counting(self, *a, **k):
        created.append(1)
        return real(self, *a, **k)
    engine.SignalEngine.__init__ = counting
    try:
        engine.close_engines()
        first = sa.process_batch(0, reads, pickle.loads(pickle.dumps(cfg)))
        for b in range(1, 4):
            again = sa.process_batch(b, reads, pickle.loads(pickle.dumps(cfg)))
            assert [normalise_result(r) for r in again] == [normalise_result(r) for r in first]
        assert len(created) == 1
        held = engine._engines[0][1]
        cfg2 = pickle.loads(pickle.dumps(cfg))
        cfg2['barcoding_quality_filter'] = 20
        sa.process_batch(4, reads, cfg2)
        assert len(created) == 2 and not held.handle.value          # rebuilt, old one closed
    finally:
        engine.SignalEngine.__init__ = real


def test_quality_filter_out_of_range_is_batch_error(preset, eng_stock):
    """App. G12: --barcoding-quality-filter > 28 -> ValueError -> (-1, msg, tb)."""
    from poreplex_b200 import signal_analyzer as sa
    z, doc, tmp = _serve('stock16k')
    cfg = _config(preset, tmp, barcoding=True)
    cfg['barcoding_quality_filter'] = 29
    res = sa.process_batch(0, [tuple(r) for r in doc['reads']][:4], cfg)
    assert isinstance(res, tuple) and res[0] == -1 and 'ValueError' in res[1]
