"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs.  Integer outputs (status, segment indices, barcode
calls, phred bins, counts) must be bit-exact; float outputs are compared bit-exact too
where the arithmetic contract promises it, and within 1e-5 relative otherwise (the
tolerance BASELINE.json's north_star states for the normalised signal)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
SCALE_ATOL, SHIFT_ATOL = 3.4e-5, 1.5e-2     # 0.13296 x scaler_margin_z0, 9.8256 x scaler_margin_z1 (+ f32 rounding)
PROB_ATOL = 2e-3     # approximate (tensor-core) class probabilities vs the exact f32 chain


def _reads(preset, n, L, seed, **kw):
    from poreplex_b200 import synth
    spec = synth.SynthSpec.for_length(L, **kw)
    return synth.to_numpy(synth.generate_reads(n, spec, preset, seed=seed))


def _dense_batch(rd):
    n, L = rd['raw'].shape
    return rd['raw'].reshape(-1), np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int64)


def _oracle_batch(orc, raw, offsets, lengths, rd, barcoding=True):
    return orc.process_batch(raw, offsets, lengths, rd['range'] / rd['digitisation'],
                             rd['offset'], barcoding=barcoding)


def _compare(out, ref, n_states=6, check_probs=True):
    assert np.array_equal(out['status'], ref['status'])
    okay_like = np.isin(ref['status'], [0, 5])       # scale/shift defined
    ss_ref = np.stack([ref['scale'], ref['shift']], axis=1)
    has_ss = ref['status'] != 3
    if check_probs == 'exact':
        assert np.array_equal(out['scale_shift'][has_ss].view(np.uint32),
                              ss_ref[has_ss].view(np.uint32)), 'scale/shift not bit-exact'
    else:
        # default path: (scale, shift) of reads that passed every margin test come from the
        # tensor-core scaler; they must lie inside the uncertainty box the margin tests assume
        # (scaler_margin_z0 / z1 on the raw outputs); typical errors are 100x smaller: the
        # normalised signal (~100 pA) stays within north_star's 1e-5 relative tolerance
        d = np.abs(out['scale_shift'][has_ss].astype(np.float64) - ss_ref[has_ss])
        assert d[:, 0].max(initial=0) <= SCALE_ATOL and d[:, 1].max(initial=0) <= SHIFT_ATOL, d.max(0)
        if len(d) >= 20:
            y_err = d[:, 0] * 100.0 + d[:, 1]              # error of a 100 pA sample after scaling
            assert np.median(y_err) <= REL_TOL * 100.0, np.median(y_err)
    seg_ok = okay_like
    assert np.array_equal(out['segments'][seg_ok][:, :n_states], ref['seg'][seg_ok][:, :n_states])
    pushed = ref['pushed'] == 1
    assert np.array_equal(out['barcode'][pushed], ref['barcode'][pushed])
    assert np.array_equal(out['barcode_guess'][pushed], ref['guess'][pushed])
    assert np.array_equal(out['barcode_score'][pushed], ref['phred'][pushed])
    assert np.all(out['barcode'][~pushed] == -1)
    assert np.all(out['barcode_score'][~pushed] == -1)
    if check_probs == 'exact':
        assert np.array_equal(out['class_probs'][pushed][:, :5].view(np.uint32),
                              ref['probs'][pushed][:, :5].view(np.uint32)), 'softmax not bit-exact'
    elif check_probs:
        # default path: class probabilities of reads whose call passed the margin test come
        # from the tensor-core kernels (diagnostic output, not part of the result dict)
        assert np.allclose(out['class_probs'][pushed][:, :5], ref['probs'][pushed][:, :5],
                           rtol=0, atol=PROB_ATOL), 'softmax outside tolerance'


def _assert_barcodes_exercised(ref, min_accepted, classes):
    """The accept branch (barcoding.py:108-118) must really be on the path of the test."""
    pushed = ref['pushed'] == 1
    acc = ref['barcode'][pushed]
    assert (acc >= 0).sum() >= min_accepted, np.bincount(acc + 1)
    assert len(set(acc[acc >= 0].tolist())) >= classes, np.bincount(acc + 1)
    # ... and so must the reject-below-threshold branch (a barcode guess, no barcode)
    assert ((ref['guess'][pushed] >= 0) & (acc < 0)).sum() >= 1


@pytest.mark.parametrize('mode', ['exact', 'strict'])
@pytest.mark.parametrize('which', ['short', 'stock'])
def test_whole_path_bit_exact_modes(which, mode, eng_short, eng_stock, orc_short, orc_stock,
                                    preset_short, preset):
    """The exact kernels end to end against the oracle: status, (scale, shift), segments, barcode
    calls AND class probabilities as raw bit patterns.  `strict` (exact scaler, segmentation and
    windows; tensor-core classifier + guard + exact re-run) must give the same bits for everything
    but the class probabilities of guard-passing windows."""
    eng, orc, pr = (eng_short, orc_short, preset_short) if which == 'short' else (eng_stock, orc_stock, preset)
    L, n = (4000, 256) if which == 'short' else (16000, 96)
    rd = _reads(pr, n, L, seed=21, frac_no_adapter=0.04, frac_qc_fail=0.04)
    raw, off, ln = _dense_batch(rd)
    eng.set_fast_lstm(mode)
    try:
        out = eng.analyze_host(raw, off, ln, rd['range'], rd['digitisation'], rd['offset'])
    finally:
        eng.set_fast_lstm('fast')
    ref = _oracle_batch(orc, raw, off, ln, rd)
    if mode == 'exact':
        _compare(out, ref, check_probs='exact')
    else:
        has_ss = ref['status'] != 3
        ss_ref = np.stack([ref['scale'], ref['shift']], axis=1)
        assert np.array_equal(out['scale_shift'][has_ss].view(np.uint32), ss_ref[has_ss].view(np.uint32))
        _compare(out, ref, check_probs=True)
    _assert_barcodes_exercised(ref, 20 if which == 'short' else 30, 2 if which == 'short' else 4)


@pytest.mark.parametrize('name', ['stock16k', 'short4k', 'chimera40k'])
def test_barcode_windows_match_reference_capture(name, eng_short, eng_stock):
    """A6 directly: the normalised, padded windows k_windows builds (barcoding.py:77-101) against
    the windows captured from inside the REFERENCE's own BarcodeDemultiplexer.push when the
    golden fixtures were made -- raw bit patterns."""
    import torch
    from golden_util import load_golden, pack_golden
    z, doc = load_golden(name)
    eng = eng_short if doc['preset'] == 'bench-short' else eng_stock
    raw, off, ln = pack_golden(z)
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (t(raw), t(off), t(ln), t(z['range']), t(z['digitisation']), t(z['offset']))
    pooled = eng.pool_signal(*args, max_raw_length=int(ln.max()))
    status, ss, _ = eng.fit_scalers(*args, pooled)
    seg, _ = eng.detect_segments(*args, pooled, ss, status, max_raw_length=int(ln.max()))
    win, pushed = eng.barcode_windows(*args, pooled, ss, status, seg)
    torch.cuda.synchronize()
    win, pushed = win.cpu().numpy(), pushed.cpu().numpy()
    ids = [str(s) for s in z['read_ids']]
    want_ids = [str(s) for s in z['window_ids']]
    assert sorted(ids[i] for i in np.nonzero(pushed)[0]) == sorted(want_ids)
    for k, rid in enumerate(want_ids):
        i = ids.index(rid)
        assert np.array_equal(win[i].view(np.uint32), z['window_bits'][k]), rid
    assert len(want_ids) >= 10


def test_whole_path_short_reads(eng_short, orc_short, preset_short):
    rd = _reads(preset_short, 200, 4000, seed=11, frac_no_adapter=0.05, frac_qc_fail=0.05)
    raw, off, ln = _dense_batch(rd)
    out = eng_short.analyze_host(raw, off, ln, rd['range'], rd['digitisation'], rd['offset'])
    ref = _oracle_batch(orc_short, raw, off, ln, rd)
    _compare(out, ref)
    assert (ref['status'] == 0).sum() > 150 and (ref['pushed'] == 1).sum() > 100
    assert set(np.unique(ref['status'])) >= {0, 4, 5}
    _assert_barcodes_exercised(ref, 15, 2)


def test_whole_path_stock_16k(eng_stock, orc_stock, preset):
    rd = _reads(preset, 64, 16000, seed=12)
    raw, off, ln = _dense_batch(rd)
    out = eng_stock.analyze_host(raw, off, ln, rd['range'], rd['digitisation'], rd['offset'])
    ref = _oracle_batch(orc_stock, raw, off, ln, rd)
    _compare(out, ref)
    assert (ref['pushed'] == 1).sum() > 40 and (ref['status'] == 0).sum() > 55
    _assert_barcodes_exercised(ref, 20, 4)


def test_ragged_lengths_and_exit_paths(eng_stock, orc_stock, preset):
    """G1-G4: too-short reads, heads shorter/longer than 30000, L % 15 != 0, a read past
    the 100000-sample scan limit; ragged packing with 16-byte aligned reads."""
    lengths = [5000, 8999, 9000, 9001, 9014, 12345, 16000, 29999, 30000, 30001, 45007,
               100000, 100010, 120000, 0, 14, 15]
    sigs, rngs, digs, offs = [], [], [], []
    for i, L in enumerate(lengths):
        Lgen = max(L, 9000)
        rd = _reads(preset, 1, Lgen, seed=100 + i)
        sigs.append(rd['raw'][0][:L])
        rngs.append(rd['range'][0]); digs.append(rd['digitisation'][0]); offs.append(rd['offset'][0])
    raw, off, ln = eng_stock.pack_reads(sigs)
    rd = {'range': np.array(rngs), 'digitisation': np.array(digs), 'offset': np.array(offs)}
    out = eng_stock.analyze_host(raw, off, ln, rd['range'], rd['digitisation'], rd['offset'],
                                 keep_pooled=True)
    ref = _oracle_batch(orc_stock, raw, off, ln, rd)
    _compare(out, ref)
    assert (ref['status'] == 3).sum() == 5
    # normalised pooled signal of an okay read, against the oracle's element kernels
    i = lengths.index(45007)
    pa = orc_stock.dac_to_pa(sigs[i], rngs[i] / digs[i], offs[i])
    T = len(pa) // 15
    want = orc_stock.scale(orc_stock.pool_mean(pa[:T * 15]), ref['scale'][i], ref['shift'][i])
    po = eng_stock.pooled_offsets(off)[i]
    got = out['pooled'][po:po + T]
    np.testing.assert_allclose(got, want, rtol=REL_TOL, atol=0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_stage_pool_and_scaler(eng_short, orc_short, preset_short):
    import torch
    rd = _reads(preset_short, 40, 4000, seed=13)
    raw, off, ln = _dense_batch(rd)
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (t(raw), t(off), t(ln), t(rd['range']), t(rd['digitisation']), t(rd['offset']))
    pooled = eng_short.pool_signal(*args, max_raw_length=4000)
    status, ss, z = eng_short.fit_scalers(*args, pooled)
    torch.cuda.synchronize()
    pooled = pooled.cpu().numpy(); z = z.cpu().numpy()
    po = eng_short.pooled_offsets(off)
    heads = np.zeros((40, 2000), np.float32)
    for i in range(40):
        pa = orc_short.dac_to_pa(rd['raw'][i], rd['range'][i] / rd['digitisation'][i], rd['offset'][i])
        want = orc_short.pool_mean(pa[:266 * 15])
        assert np.array_equal(pooled[po[i]:po[i] + 266].view(np.uint32), want.view(np.uint32))
        heads[i, 2000 - 266:] = want
    zref = orc_short.scaler_predict(heads)
    assert np.array_equal(z.view(np.uint32), zref.view(np.uint32))
    # explicit heads (no zero-prefix skipping) must give the same bits
    z2 = eng_short.scaler_predict(t(heads)).cpu().numpy()
    assert np.array_equal(z2.view(np.uint32), zref.view(np.uint32))


def test_stage_viterbi_paths(eng_stock, orc_stock):
    import torch
    rng = np.random.default_rng(5)
    n, ld = 64, 700
    x = np.empty((n, ld), np.float32)
    lengths = rng.integers(1, ld + 1, n).astype(np.int32)
    lengths[:3] = [1, 2, ld]
    levels = np.array([71.5, 102.0, 112.0, 80.5, 109.0, 95.0])
    for i in range(n):
        segs = np.sort(rng.integers(0, ld, 5))
        st = np.searchsorted(segs, np.arange(ld), side='right')
        x[i] = levels[st] + rng.normal(0, 4.0, ld)
    dev = torch.device('cuda', 0)
    path, logp = eng_stock.viterbi_paths(torch.from_numpy(x).to(dev),
                                         torch.from_numpy(lengths).to(dev))
    torch.cuda.synchronize()
    path = path.cpu().numpy(); logp = logp.cpu().numpy()
    for i in range(n):
        lp, p = orc_stock.viterbi(x[i, :lengths[i]])
        assert lp == logp[i]
        assert np.array_equal(p, path[i, :lengths[i]])


def test_stage_windows_and_demux(eng_stock, orc_stock):
    """G8-G11: window lengths around the 260/300/3000 limits, flat windows (MAD = 0),
    even/odd medians, decisions on arbitrary windows."""
    import torch
    rng = np.random.default_rng(7)
    dev = torch.device('cuda', 0)
    n = 96
    win = rng.normal(0, 1.2, (n, 300)).astype(np.float32)
    npad = rng.integers(0, 41, n)
    for i in range(n):
        win[i, :npad[i]] = -1000.0
    win[0] = 0.0
    eng_stock.set_fast_lstm(False)
    try:
        probs, bc, guess, score = eng_stock.demux_predict(torch.from_numpy(win).to(dev))
        torch.cuda.synchronize()
    finally:
        eng_stock.set_fast_lstm(True)
    pref = orc_stock.demux_predict(win)
    assert np.array_equal(probs.cpu().numpy()[:, :5].view(np.uint32), pref.view(np.uint32))
    for i in range(n):
        b, g, p = orc_stock.barcode_decide(pref[i])
        assert (bc[i].item(), guess[i].item(), score[i].item()) == (-1 if b is None else b, g, p)


def test_counts(eng_short):
    import torch
    rng = np.random.default_rng(9)
    n = 100000
    status = rng.integers(0, 11, n).astype(np.int32)
    label = rng.integers(0, 4, n).astype(np.int32)
    barcode = rng.integers(-1, 4, n).astype(np.int32)
    dev = torch.device('cuda', 0)
    c = eng_short.count_results(torch.from_numpy(status).to(dev), torch.from_numpy(label).to(dev),
                                torch.from_numpy(barcode).to(dev)).cpu().numpy()
    want = np.zeros((4, 5, 11), np.int64)
    np.add.at(want, (label, barcode + 1, status), 1)
    assert np.array_equal(c, want)
    assert c.sum() == n


def test_fast_division_equals_ieee_division(eng_short, preset_short):
    """The LSTM kernels' branch-free Newton division vs the verification mode that uses
    IEEE __fdiv_rn everywhere: every output identical (incl. long -1000 pad regions where
    cell states decay towards zero)."""
    rd = _reads(preset_short, 160, 4000, seed=21)
    raw, off, ln = _dense_batch(rd)
    args = (raw, off, ln, rd['range'], rd['digitisation'], rd['offset'])
    eng_short.set_fast_lstm(False)
    try:
        fast = eng_short.analyze_host(*args)
        eng_short.set_exact_division(True)
        exact = eng_short.analyze_host(*args)
    finally:
        eng_short.set_exact_division(False)
        eng_short.set_fast_lstm(True)
    for k in ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'counts'):
        assert np.array_equal(fast[k], exact[k]), k
    assert np.array_equal(fast['scale_shift'].view(np.uint32), exact['scale_shift'].view(np.uint32))
    assert np.array_equal(fast['class_probs'].view(np.uint32), exact['class_probs'].view(np.uint32))
    assert (fast['barcode_score'] >= 0).sum() > 100


def test_polya_kernel_matches_host_core(eng_stock, orc_stock, preset):
    """k_polya (thread per read) against the same core compiled for the host, which the CPU
    suite ties to the reference's polya.py: found flag, begin/end, dwell, spikes identical."""
    import ctypes as C
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostcheck_util as H
    from poreplex_b200.engine import polya_to_dict
    hc = H.load()
    Pc = H.polya_params(preset['polya_dwell'])
    found = overflowed = 0
    for L, polya_len, seed in ((16000, (20, 60), 31), (24000, (5, 300), 32), (20000, (100, 500), 33)):
        from poreplex_b200 import synth
        spec = synth.SynthSpec.for_length(L, frac_no_adapter=0.05, frac_qc_fail=0.05)
        spec.polya_pooled = polya_len
        rd = synth.to_numpy(synth.generate_reads(96, spec, preset, seed=seed))
        raw, off, ln = _dense_batch(rd)
        out = eng_stock.analyze_host(raw, off, ln, rd['range'], rd['digitisation'], rd['offset'],
                                     polya=True)
        names = eng_stock.state_names
        ia, ip = names.index('adapter'), names.index('polya-tail')
        for i in range(len(ln)):
            if out['status'][i] != 0:
                assert polya_to_dict(out['polya'][i], 3012.0) is None
                continue
            seg = out['segments'][i]
            rb, re = (int(seg[ip, 0]), int(seg[ip, 1])) if seg[ip, 0] >= 0 else (int(seg[ia, 1]) + 1, -1)
            R = H.PolyaResultC()
            sig = np.ascontiguousarray(rd['raw'][i])
            hc.hc_polya(C.byref(Pc), sig.ctypes.data_as(C.c_void_p), C.c_int64(L),
                        C.c_double(rd['range'][i] / rd['digitisation'][i]), C.c_double(rd['offset'][i]),
                        C.c_float(out['scale_shift'][i, 0]), C.c_float(out['scale_shift'][i, 1]),
                        C.c_int32(rb), C.c_int32(re), C.byref(R))
            rec = out['polya'][i]
            # the fixed-size buffers of the kernel (48 spikes, 64 recalibration anchors) are never
            # silent: a record that exceeds them raises, and the host core agrees that it does
            over = bool(R.flags & 1) or (R.found and R.n_spikes > H.MAX_SPIKES)
            if over:
                with pytest.raises(OverflowError):
                    polya_to_dict(rec, 3012.0)
                assert (int(rec['flags']), int(rec['n_spikes'])) == (R.flags, R.n_spikes)
                overflowed += 1
                continue
            got = polya_to_dict(rec, 3012.0)
            want = H.result_to_dict(R, 3012.0)
            assert want == got, (L, i, want, got)
            found += got is not None
    assert found > 150
    assert overflowed >= 1          # the 100..500-sample tails do reach the spike capacity


def test_unsplit_kernels_match_restatement(eng_stock, orc_stock, preset):
    """k_unsplit_windows / k_unsplit_decide against oracle/unsplit_restated.py (which the CPU
    suite ties to the reference's detect_unsplit_read) on the chimera fixture's event tables."""
    import tempfile
    from golden_util import load_golden, golden_reads, golden_basecalls, pack_golden
    from oracle import refshim, fake_fast5, unsplit_restated as UR
    from poreplex_b200.fast5_source import Fast5Source
    refshim.install_fake_h5py()
    refshim.clear_fast5()
    z, doc = load_golden('chimera40k')
    ids = [str(s) for s in z['read_ids']]
    tmp = tempfile.mkdtemp()
    fake_fast5.build_fast5(tmp, 'reads.fast5', golden_reads(z), ids, golden_basecalls(z))
    raw, off, ln = pack_golden(z)
    out = eng_stock.analyze_host(raw, off, ln, z['range'], z['digitisation'], z['offset'],
                                 exact_scaler=True)      # as signal_analyzer.py does for the chimera filter
    tables, rates = [], []
    for i, rid in enumerate(ids):
        src = Fast5Source(tmp + '/reads.fast5', rid)
        bc = src.get_basecall(want_events=True) if out['status'][i] == 0 else None
        tables.append(None if bc is None else bc['events'])
        rates.append(src.sampling_rate)
    status = np.where([t is None for t in tables], 10, out['status']).astype(np.int32)
    flags = eng_stock.detect_unsplit_host(tables, np.array(rates), out['scale_shift'], status,
                                          out['segments'],
                                          batch=(raw, off, ln, z['range'], z['digitisation'], z['offset']))
    from oracle import events_restated as ER
    bcs = golden_basecalls(z)
    ia = eng_stock.adapter_state
    n_true = 0
    for i, t in enumerate(tables):
        if t is None:
            assert flags[i] == 0
            continue
        # the event table as the reference's own numpy recipe builds it (fast5_file.py:183-230)
        ev = ER.derive_event_table(z['raw'][i][:ln[i]], z['range'][i], z['digitisation'][i],
                                   z['offset'][i], bcs[i]['moves'], bcs[i]['sequence'],
                                   bcs[i]['qstring'], 0, 15)
        scaled, pos, end = UR.derive_event_columns(ev['start'], ev['mean'], ev['move'],
                                                   out['scale_shift'][i, 0], out['scale_shift'][i, 1])
        want = UR.detect_unsplit_read(
            preset['unsplit_read_detection'], lambda x: orc_stock.viterbi(x, 'unsplit')[1],
            orc_stock.unsplit_names, np.asarray(ev['start'], np.int64), end, scaled, pos,
            np.asarray(ev['p_model_state'], np.float64), int(out['segments'][i, ia, 1]), rates[i])
        assert int(flags[i]) == int(want), (i, flags[i], want)
        n_true += want
    assert n_true >= 4


@pytest.mark.parametrize('pipeline', ['streamed', 'arena'])
def test_pipelined_host_path_equals_single_pass(eng_short, orc_short, preset_short, monkeypatch, pipeline):
    """pb2_analyze_host cuts big batches into chunks whose uploads overlap the kernels -- whole
    batch resident (`streamed`, the default) or two chunk-sized arenas (`arena`); results must
    equal the single-pass path read for read, including poly(A) records and the summed counts."""
    rd = _reads(preset_short, 6000, 4000, seed=41, frac_no_adapter=0.03, frac_qc_fail=0.03)
    raw, off, ln = _dense_batch(rd)
    args = (raw, off, ln, rd['range'], rd['digitisation'], rd['offset'])
    single = eng_short.analyze_host(*args, polya=True)
    monkeypatch.setenv('POREPLEX_B200_HOST_CHUNK_ELEMS', str(3_000_000))     # -> 8 chunks
    monkeypatch.setenv('POREPLEX_B200_HOST_PIPELINE', pipeline)
    piped = eng_short.analyze_host(*args, polya=True)
    monkeypatch.delenv('POREPLEX_B200_HOST_CHUNK_ELEMS')
    monkeypatch.delenv('POREPLEX_B200_HOST_PIPELINE')
    for k in single:
        a, b = single[k], piped[k]
        if a.dtype.fields:              # poly(A) records: spikes beyond n_spikes are unset
            for f in ('found', 'n_spikes', 'begin', 'end', 'dwell_samples', 'extensions'):
                assert np.array_equal(a[f], b[f]), (k, f)
            for i in np.nonzero(a['found'])[0]:
                m = min(int(a['n_spikes'][i]), 48)
                assert np.array_equal(a['spikes'][i, :m], b['spikes'][i, :m], equal_nan=True)
        elif k == 'class_probs':
            # tensor-core probabilities depend (within the approximation error) on which
            # reads share a 128-read tile, i.e. on the chunking; the calls do not
            assert np.allclose(a, b, rtol=0, atol=PROB_ATOL), k
        elif a.dtype.kind == 'f':
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), k
        else:
            assert np.array_equal(a, b), k
    assert piped['counts'].sum() == 6000
    ref = _oracle_batch(orc_short, raw[:400 * 4000], off[:400], ln[:400],
                        {k: rd[k][:400] for k in ('range', 'digitisation', 'offset')})
    _compare({k: v[:400] for k, v in piped.items() if k not in ('counts', 'polya')}, ref)


def test_degenerate_batches(eng_stock, orc_stock, preset):
    """Empty batch, a batch in which every read is too short, and one very long read
    (> scan limit, poly(A) on) through the host API."""
    e = np.zeros(0)
    out = eng_stock.analyze_host(np.zeros(8, np.int16), e.astype(np.int64), e.astype(np.int64), e, e, e)
    assert out['status'].shape == (0,) and out['counts'].sum() == 0
    sigs = [np.full(n, 500, np.int16) for n in (0, 10, 8999)]
    raw, off, ln = eng_stock.pack_reads(sigs)
    c = np.array([1400.0] * 3), np.array([8192.0] * 3), np.array([5.0] * 3)
    out = eng_stock.analyze_host(raw, off, ln, *c, polya=True)
    assert list(out['status']) == [3, 3, 3] and list(out['label']) == [3, 3, 3]
    assert out['counts'][3, 0, 3] == 3 and not out['polya']['found'].any()
    rd = _reads(preset, 2, 150000, seed=77)
    raw, off, ln = _dense_batch(rd)
    out = eng_stock.analyze_host(raw, off, ln, rd['range'], rd['digitisation'], rd['offset'], polya=True)
    ref = _oracle_batch(orc_stock, raw, off, ln, rd)
    _compare(out, ref)
    assert out['segments'][:, :6].max() < 6666          # scan limit (signal_analyzer.py:347-349)
