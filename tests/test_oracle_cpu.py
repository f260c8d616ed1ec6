"""CPU tests of the oracle itself: against numpy (which IS the reference's implementation
of the element ops), against the reference's own event-detector C (oracle/_ref), against
independent float64 re-derivations of the third-party kernels, and against the golden
fixtures captured from the reference's Python running verbatim."""
import math
import os

import numpy as np
import pytest

from golden_util import load_golden, pack_golden


def test_activation_and_exp_accuracy(oracle_mod):
    L = oracle_mod.lib()
    xs = np.linspace(-30, 30, 24001).astype(np.float32)
    t = np.array([L.orc_tanhf(float(x)) for x in xs])
    s = np.array([L.orc_sigmoidf(float(x)) for x in xs])
    assert np.abs(t - np.tanh(xs.astype(np.float64))).max() < 5e-7
    assert np.abs(s - 1 / (1 + np.exp(-xs.astype(np.float64)))).max() < 3e-7
    assert L.orc_sigmoidf(-1000.0) == 0.0 and L.orc_sigmoidf(1000.0) == 1.0   # App. E-16
    xe = np.linspace(-87, 5, 9001).astype(np.float32)
    e = np.array([L.orc_expf(float(x)) for x in xe])
    ref = np.exp(xe.astype(np.float64))
    assert (np.abs(e - ref) / ref).max() < 2e-7
    rng = np.random.default_rng(0)
    xd = -rng.uniform(0, 40, 5000)
    ed = np.array([L.orc_exp_neg(float(x)) for x in xd])
    assert (np.abs(ed - np.exp(xd)) / np.exp(xd)).max() < 3 * 2.0 ** -52
    assert L.orc_exp_neg(-41.0) == 0.0 and L.orc_exp_neg(0.0) == 1.0
    wd = rng.uniform(1, 2, 5000)
    ld = np.array([L.orc_log_1to2(float(x)) for x in wd])
    assert np.abs(ld - np.log(wd)).max() < 3e-16
    for a, b in ((-3.0, -4.0), (-700.0, -3.0), (-1e3, -1e3), (0.5, -20.0)):
        assert abs(L.orc_pair_lse(a, b) - np.logaddexp(a, b)) < 1e-14
    assert L.orc_pair_lse(-math.inf, -2.5) == -2.5 and L.orc_pair_lse(-2.5, -math.inf) == -2.5


def test_dac_pool_scale_match_numpy(orc_stock):
    """fast5_file.py:130-131, signal_loader.py:224-225/246-247/262 evaluated by numpy."""
    rng = np.random.default_rng(1)
    for n in (15, 29, 30, 4000, 4001, 16007):
        raw = rng.integers(-500, 2000, n).astype(np.int16)
        rng_pa, dig, off = 1443.03, 8192.0, 7.0
        want_pa = np.array(rng_pa / dig * (raw + off), dtype=np.float32)
        got_pa = orc_stock.dac_to_pa(raw, rng_pa / dig, off)
        assert np.array_equal(got_pa.view(np.uint32), want_pa.view(np.uint32))
        cut = n - n % 15
        want_pool = want_pa[:cut].reshape([n // 15, 15]).mean(axis=1, dtype=np.float32)
        got_pool = orc_stock.pool_mean(got_pa[:cut])
        assert np.array_equal(got_pool.view(np.uint32), want_pool.view(np.uint32))
        params = np.array([0.9731, 4.25], np.float32)
        want = np.poly1d(params)(want_pool)
        got = orc_stock.scale(got_pool, params[0], params[1])
        assert want.dtype == np.float32
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _ref_normalize(sig):           # barcoding.py:77-81, verbatim
    med = np.median(sig)
    mad = np.median(np.abs(sig - med))
    return (sig - med) / max(0.01, (mad * 1.4826))


def _ref_push(signal, minlen=260, maxlen=3000, trimlength=300, pad=-1000.):   # barcoding.py:83-98
    if not minlen <= len(signal) <= maxlen:
        return None
    if len(signal) > trimlength:
        return _ref_normalize(signal[-trimlength:])
    elif len(signal) < trimlength:
        return np.pad(_ref_normalize(signal), (trimlength - len(signal), 0), 'constant',
                      constant_values=pad)
    return _ref_normalize(signal)


def test_barcode_window_matches_numpy(orc_stock):
    rng = np.random.default_rng(2)
    for n in (259, 260, 261, 299, 300, 301, 1000, 3000, 3001):          # App. G8
        sig = (80 + 7 * rng.standard_normal(n)).astype(np.float32)
        want = _ref_push(sig)
        got = orc_stock.barcode_window(sig)
        if want is None:
            assert got is None
        else:
            assert want.dtype == np.float32
            assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
    flat = np.full(280, 81.5, np.float32)                                 # G9: MAD = 0
    assert np.array_equal(orc_stock.barcode_window(flat), _ref_push(flat).astype(np.float32))
    ties = np.repeat(np.array([70., 80., 90.], np.float32), 100)        # ties + even length
    assert np.array_equal(orc_stock.barcode_window(ties), _ref_push(ties).astype(np.float32))


def test_barcode_decision_rule(orc_stock):
    # barcoding.py:60: list(calibtable['pred_score']) -> np.float64 scalars, so both the
    # threshold compare and bisect_right run in float64 (a Python float would be cast
    # DOWN to float32 under NEP 50)
    calib = [np.float64(v) for v in list(orc_stock.model.calib)[:orc_stock.model.n_calib]]
    thr = calib[18]
    from bisect import bisect_right
    cases = [np.array([0.99, 0.0025, 0.0025, 0.0025, 0.0025], np.float32),     # decoy wins
             np.array([0.01, 0.97, 0.01, 0.005, 0.005], np.float32),           # below threshold
             np.array([0.0, np.float32(thr), 0.0, 0.0, 0.0], np.float32),
             np.array([0.001, 0.001, 0.001, 0.9985, 0.0005], np.float32),
             np.array([0.2, 0.2, 0.2, 0.2, 0.2], np.float32)]                  # first max wins
    for p in cases:
        label = int(np.argmax(p)) - 1
        score = np.amax(p)
        want_bc = label if (label >= 0 and score >= thr) else None
        want_q = 0 if score <= 0. else bisect_right(calib, score)
        assert orc_stock.barcode_decide(p) == (want_bc, label, want_q)


def test_event_detector_restatement_matches_reference_c(oracle_mod):
    if not oracle_mod.have_ref_scrappie():
        pytest.skip('oracle/_ref/libscrappie_ref.so not built (needs /root/reference)')
    rng = np.random.default_rng(3)
    signals = []
    for n in (5, 13, 14, 39, 40, 41, 500, 3000, 12000):
        levels = np.repeat(rng.normal(100, 12, n // 25 + 1), 25)[:n]
        signals.append((levels + rng.normal(0, 2.0, n)).astype(np.float32))
    signals.append(np.full(300, 108.0, np.float32))                       # G14 constant
    signals.append(np.zeros(50, np.float32))
    for sig in signals:
        for kw in ({}, dict(window_length1=7, window_length2=20, threshold1=3, threshold2=8,
                            peak_height=4)):
            a = oracle_mod.detect_events_ref(sig, **kw)
            b = oracle_mod.detect_events_restated(sig, **kw)
            assert len(a) == len(b)
            for f in ('start', 'length', 'mean', 'stdv', 'pos', 'state'):
                assert np.array_equal(a[f], b[f], equal_nan=(f in ('mean', 'stdv'))), f


def _viterbi_f64(hmm_def, x):
    """Independent float64 Viterbi with libm exp/log (pomegranate semantics)."""
    names = sorted(s['name'] for s in hmm_def)
    idx = {n: i for i, n in enumerate(names)}
    S = len(names)
    logT = np.full((S, S), -np.inf)
    start = np.full(S, -np.inf)
    em = [None] * S
    for s in hmm_def:
        i = idx[s['name']]
        em[i] = s['emission']
        if s.get('start_prob', 0) > 0:
            start[i] = math.log(s['start_prob'])
        for nxt, p in s['transition']:
            logT[i, idx[nxt]] = math.log(p)

    def emis(i, v):
        comps = em[i]
        w = np.array([c[2] if len(c) > 2 else 1.0 for c in comps])
        w = w / w.sum()
        lps = [-math.log(c[1] * 2.50662827463) - (v - c[0]) ** 2 / (2 * c[1] ** 2) + math.log(wi)
               for c, wi in zip(comps, w)]
        return lps[0] if len(lps) == 1 else np.logaddexp(lps[0], lps[1])
    T = len(x)
    v = np.array([start[i] + emis(i, float(x[0])) for i in range(S)])
    bp = np.zeros((T, S), int)
    for t in range(1, T):
        nv = np.full(S, -np.inf)
        for l in range(S):
            e = emis(l, float(x[t]))
            for k in range(S):
                c = v[k] + logT[k, l] + e
                if c > nv[l]:
                    nv[l] = c
                    bp[t, l] = k
        v = nv
    cur = int(np.argmax(v))
    path = [cur]
    for t in range(T - 1, 0, -1):
        cur = bp[t, cur]
        path.append(cur)
    return float(v.max()), np.array(path[::-1])


def test_viterbi_against_independent_float64(orc_stock, preset):
    rng = np.random.default_rng(4)
    levels = np.array([71.5, 102.0, 112.0, 80.5, 109.0, 95.0])
    for which, key in (('seg', 'segmentation_model'), ('unsplit', 'unsplit_read_detection_model')):
        for trial in range(3):
            T = 400
            cuts = np.sort(rng.integers(0, T, 5))
            st = np.searchsorted(cuts, np.arange(T), side='right')
            x = (levels[st] + rng.normal(0, 4.0, T)).astype(np.float32)
            lp, path = orc_stock.viterbi(x, which)
            lp2, path2 = _viterbi_f64(preset[key], x)
            assert np.array_equal(path, path2)
            assert abs(lp - lp2) < 1e-8 * abs(lp2)


def test_viterbi_against_exhaustive_search(orc_stock, preset):
    """Not a second dynamic programme but NO dynamic programme: for short signals every one of
    the S^T state paths is scored directly (start + emissions + transitions, float64) and the best
    one must be the path the oracle's Viterbi returns, with the same log-probability."""
    import itertools
    rng = np.random.default_rng(8)
    levels = np.array([71.5, 102.0, 112.0, 80.5, 109.0, 95.0])
    for which, key in (('seg', 'segmentation_model'), ('unsplit', 'unsplit_read_detection_model')):
        hmm_def = preset[key]
        names = sorted(s['name'] for s in hmm_def)
        idx = {n: i for i, n in enumerate(names)}
        S = len(names)
        logT = np.full((S, S), -np.inf)
        start = np.full(S, -np.inf)
        em = [None] * S
        for st in hmm_def:
            i = idx[st['name']]
            em[i] = st['emission']
            if st.get('start_prob', 0) > 0:
                start[i] = math.log(st['start_prob'])
            for nxt, pr in st['transition']:
                logT[i, idx[nxt]] = math.log(pr)

        def emis(i, v):
            comps = em[i]
            w = np.array([c[2] if len(c) > 2 else 1.0 for c in comps])
            w = w / w.sum()
            lps = [-math.log(c[1] * math.sqrt(2 * math.pi)) - (v - c[0]) ** 2 / (2 * c[1] ** 2) + math.log(wi)
                   for c, wi in zip(comps, w)]
            return lps[0] if len(lps) == 1 else float(np.logaddexp(lps[0], lps[1]))

        for T in (1, 2, 3, 5, 6):
            for trial in range(4):
                x = (levels[np.sort(rng.integers(0, S, T))] + rng.normal(0, 5.0, T)).astype(np.float32)
                E = np.array([[emis(i, float(v)) for i in range(S)] for v in x])       # [T][S]
                best, best_path = -np.inf, None
                for path in itertools.product(range(S), repeat=T):
                    lp = start[path[0]] + E[0, path[0]]
                    for t in range(1, T):
                        lp += logT[path[t - 1], path[t]] + E[t, path[t]]
                    if lp > best:
                        best, best_path = lp, path
                lp_o, path_o = orc_stock.viterbi(x, which)
                assert tuple(int(v) for v in path_o) == best_path, (which, T, trial)
                assert abs(lp_o - best) < 1e-9 * max(1.0, abs(best))


def _lstm_f64(layer, xs, reverse=False):
    H = layer.units
    W, U, b = (a.astype(np.float64) for a in (layer.kernel, layer.recurrent, layer.bias))
    h = np.zeros(H); c = np.zeros(H)
    out = np.zeros((len(xs), H))
    order = range(len(xs) - 1, -1, -1) if reverse else range(len(xs))
    sig = lambda v: 1 / (1 + np.exp(-v))
    for t in order:
        z = xs[t] @ W + h @ U + b
        i, f, g, o = sig(z[:H]), sig(z[H:2 * H]), np.tanh(z[2 * H:3 * H]), sig(z[3 * H:])
        c = f * c + i * g
        h = o * np.tanh(c)
        out[t] = h
    return out, h


def test_lstm_networks_against_float64(orc_stock):
    from poreplex_b200 import params
    p = orc_stock.preset
    sc = params.load_scaler_model(p['signal_processing']['scaler_model'])
    dm = params.load_demux_model(p['demultiplexing']['demux_model'])
    rng = np.random.default_rng(5)
    head = np.zeros(2000, np.float32)
    head[1400:] = (90 + 12 * rng.standard_normal(600)).astype(np.float32)
    z = orc_stock.scaler_predict(head[None])[0]
    s1, _ = _lstm_f64(sc.l1, head[:, None].astype(np.float64))
    _, h2 = _lstm_f64(sc.l2, s1)
    z64 = h2 @ sc.dense_kernel.astype(np.float64) + sc.dense_bias
    assert np.abs(z - z64).max() < 2e-3
    win = rng.normal(0, 1.2, 300).astype(np.float32)
    win[:30] = -1000.0
    pr = orc_stock.demux_predict(win[None])[0]
    f, _ = _lstm_f64(dm.fwd, win[:, None].astype(np.float64))
    b, _ = _lstm_f64(dm.bwd, win[:, None].astype(np.float64), reverse=True)
    _, hl = _lstm_f64(dm.l2, np.concatenate([f, b], axis=1))
    logit = hl @ dm.dense_kernel.astype(np.float64) + dm.dense_bias
    p64 = np.exp(logit - logit.max()); p64 /= p64.sum()
    assert np.abs(pr - p64).max() < 2e-3
    assert abs(pr.sum() - 1) < 1e-5


def _torch_lstm(layer, xs, reverse=False):
    """The same layer through torch.nn.LSTM (float64): a third-party implementation of the LSTM
    equations with the Keras weights mapped onto it (Keras gate blocks i|f|c|o = torch i|f|g|o;
    kernel [in, 4H] -> weight_ih [4H, in]; Keras has one bias, so bias_hh = 0)."""
    import torch
    H, I = layer.units, layer.kernel.shape[0]
    m = torch.nn.LSTM(I, H, batch_first=True).double()
    with torch.no_grad():
        m.weight_ih_l0.copy_(torch.from_numpy(np.ascontiguousarray(layer.kernel.T, np.float64)))
        m.weight_hh_l0.copy_(torch.from_numpy(np.ascontiguousarray(layer.recurrent.T, np.float64)))
        m.bias_ih_l0.copy_(torch.from_numpy(np.asarray(layer.bias, np.float64)))
        m.bias_hh_l0.zero_()
        x = torch.from_numpy(np.ascontiguousarray(xs[::-1] if reverse else xs, np.float64))[None]
        out, (h, _c) = m(x)
    seq = out[0].numpy()
    return (seq[::-1].copy() if reverse else seq), h[0, 0].numpy()


def test_lstm_networks_against_torch_lstm(orc_stock):
    """TensorFlow cannot be installed here, so the Keras LSTM restatement (gate order, bias
    placement, sequence direction and where the backward layer's output lands, final-state
    selection) is cross-checked against an implementation the build did not write: torch.nn.LSTM
    in float64, on several heads / windows including padded ones."""
    from poreplex_b200 import params
    p = orc_stock.preset
    sc = params.load_scaler_model(p['signal_processing']['scaler_model'])
    dm = params.load_demux_model(p['demultiplexing']['demux_model'])
    rng = np.random.default_rng(11)
    heads = np.zeros((3, 2000), np.float32)
    for k, n in enumerate((600, 1500, 2000)):
        heads[k, 2000 - n:] = (90 + 12 * rng.standard_normal(n)).astype(np.float32)
    z = orc_stock.scaler_predict(heads)
    for k in range(len(heads)):
        s1, _ = _torch_lstm(sc.l1, heads[k][:, None].astype(np.float64))
        _, h2 = _torch_lstm(sc.l2, s1)
        z64 = h2 @ sc.dense_kernel.astype(np.float64) + sc.dense_bias
        assert np.abs(z[k] - z64).max() < 2e-3, k
    wins = rng.normal(0, 1.2, (4, 300)).astype(np.float32)
    wins[1, :40] = -1000.0
    wins[2, :5] = -1000.0
    wins[3] = np.sort(wins[3])                       # a strongly direction-dependent window
    pr = orc_stock.demux_predict(wins)
    for k in range(len(wins)):
        x = wins[k][:, None].astype(np.float64)
        f, _ = _torch_lstm(dm.fwd, x)
        b, _ = _torch_lstm(dm.bwd, x, reverse=True)
        _, hl = _torch_lstm(dm.l2, np.concatenate([f, b], axis=1))
        logit = hl @ dm.dense_kernel.astype(np.float64) + dm.dense_bias
        p64 = np.exp(logit - logit.max())
        p64 /= p64.sum()
        assert np.abs(pr[k] - p64).max() < 2e-3, k
        assert int(np.argmax(pr[k])) == int(np.argmax(p64)), k


@pytest.mark.parametrize('name', ['stock16k', 'short4k', 'chimera40k'])
def test_oracle_pipeline_reproduces_reference_run(oracle_mod, name):
    """The standalone oracle pipeline (orc_process_batch) must reproduce what was captured
    from INSIDE the reference's own Python when it ran over the same reads: scaling
    params and barcode windows bit for bit, segment tables, statuses and barcode fields."""
    z, doc = load_golden(name)
    orc = oracle_mod.default_oracle(bench_short=(doc['preset'] == 'bench-short'))
    raw, off, ln = pack_golden(z)
    res = orc.process_batch(raw, off, ln, z['range'] / z['digitisation'], z['offset'])
    ids = [str(s) for s in z['read_ids']]
    by_id = {r['read_id']: r for r in doc['results_trim_barcoding'] if 'read_id' in r}
    scal = {str(k): v for k, v in zip(z['scaling_ids'], z['scaling_bits'])}
    wins = {str(k): v for k, v in zip(z['window_ids'], z['window_bits'])}
    n_seg = n_win = 0
    for i, rid in enumerate(ids):
        ref = by_id[rid]
        st = oracle_mod.STATUS_NAMES[res['status'][i]]
        if ref['status'] in ('scaler_signal_too_short', 'scaling_qc_fail', 'adapter_not_detected'):
            assert st == ref['status']
        else:       # statuses decided later (basecall table etc.) are 'okay' at this stage
            assert st == 'okay'
        if rid in scal:
            got = np.array([res['scale'][i], res['shift'][i]], np.float32).view(np.uint32)
            assert np.array_equal(got, scal[rid])
            seg = {orc.seg_names[s]: [int(res['seg'][i][s][0]), int(res['seg'][i][s][1])]
                   for s in range(6) if res['seg'][i][s][0] >= 0}
            assert seg == doc['segments'][rid]
            n_seg += 1
        assert bool(res['pushed'][i]) == (rid in wins)
        if rid in wins:
            a0, a1 = doc['segments'][rid]['adapter']
            pa = orc.dac_to_pa(z['raw'][i][:ln[i]], z['range'][i] / z['digitisation'][i], z['offset'][i])
            sig = orc.scale(orc.pool_mean(pa[:len(pa) // 15 * 15]), res['scale'][i], res['shift'][i])
            w = orc.barcode_window(sig[a0:a1 + 1])
            assert np.array_equal(w.view(np.uint32), wins[rid])
            bc = None if res['barcode'][i] < 0 else int(res['barcode'][i])
            assert ref.get('barcode') == bc
            if bc is not None:
                assert ref['barcode_guess'] == res['guess'][i] and ref['barcode_score'] == res['phred'][i]
            n_win += 1
    assert n_seg >= 15 and n_win >= 15


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
def test_unsplit_restatement_matches_reference(oracle_mod):
    """oracle/unsplit_restated.py vs SignalAnalysis.load_events + detect_unsplit_read running
    verbatim on the chimera fixture (event tables served through the fake FAST5)."""
    import tempfile
    from golden_util import golden_reads, golden_basecalls
    from oracle import refshim, fake_fast5, unsplit_restated as UR
    sa, sl, _, _, _ = refshim.reference_modules()
    from poreplex import worker_persistence as wp
    z, doc = load_golden('chimera40k')
    orc = oracle_mod.default_oracle()
    preset = orc.preset
    raw, off, ln = pack_golden(z)
    res = orc.process_batch(raw, off, ln, z['range'] / z['digitisation'], z['offset'])
    ids = [str(s) for s in z['read_ids']]
    tmp = tempfile.mkdtemp()
    refshim.clear_fast5()
    fake_fast5.build_fast5(tmp, 'reads.fast5', golden_reads(z), ids, golden_basecalls(z))

    class Analyzer:
        pass
    an = Analyzer()
    an.unsplitmodel = wp.load_segmentation_model(preset['unsplit_read_detection_model'])
    an.config = dict(preset, albacore_onthefly=False)
    by_id = {r['read_id']: r for r in doc['results_all_switches'] if 'read_id' in r}
    n_true = n_false = 0
    for i, rid in enumerate(ids):
        if by_id[rid]['status'] not in ('okay', 'unsplit_read'):
            continue
        npread = sl.NanoporeRead('reads.fast5', tmp, rid)
        npread.set_scaling_params(np.array([res['scale'][i], res['shift'][i]], np.float32))
        s = sa.SignalAnalysis(npread, an)
        events = s.load_events()
        seg = {orc.seg_names[k]: (int(res['seg'][i][k][0]), int(res['seg'][i][k][1]))
               for k in range(6) if res['seg'][i][k][0] >= 0}
        want = bool(s.detect_unsplit_read(events, seg, 15))
        scaled, pos, end = UR.derive_event_columns(events['start'].values, events['mean'].values,
                                                   events['move'].values, res['scale'][i],
                                                   res['shift'][i])
        assert np.array_equal(scaled.view(np.uint32),
                              np.asarray(events['scaled_mean'].values, np.float32).view(np.uint32))
        assert np.array_equal(pos, events['pos'].values) and np.array_equal(end, events['end'].values)
        got = UR.detect_unsplit_read(
            preset['unsplit_read_detection'], lambda x: orc.viterbi(x, 'unsplit')[1],
            orc.unsplit_names, events['start'].values.astype(np.int64), end, scaled, pos,
            events['p_model_state'].values.astype(np.float64), seg['adapter'][1],
            npread.sampling_rate)
        assert want == got == (by_id[rid]['status'] == 'unsplit_read')
        n_true += want
        n_false += not want
        npread.close()
    assert n_true >= 4 and n_false >= 6


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
@pytest.mark.parametrize('name', ['stock16k', 'chimera40k'])
def test_event_table_restatement_matches_reference(oracle_mod, name):
    """oracle/events_restated.py vs the reference's own Fast5Reader.get_basecall (Move table ->
    events, fast5_file.py:183-230) + SignalAnalysis.load_events (signal_analyzer.py:311-326)
    running verbatim over the fake FAST5: every column, floats as raw bit patterns."""
    import tempfile
    from golden_util import golden_reads, golden_basecalls
    from oracle import refshim, fake_fast5, events_restated as ER
    sa, sl, _, _, _ = refshim.reference_modules()
    z, doc = load_golden(name)
    ids = [str(s) for s in z['read_ids']]
    bcs = golden_basecalls(z)
    tmp = tempfile.mkdtemp()
    refshim.clear_fast5()
    fake_fast5.build_fast5(tmp, 'reads.fast5', golden_reads(z), ids, bcs)

    class Analyzer:
        pass
    an = Analyzer()
    an.config = {'albacore_onthefly': False}
    checked = 0
    for i, rid in enumerate(ids):
        if bcs[i] is None or i % 3:
            continue
        npread = sl.NanoporeRead('reads.fast5', tmp, rid)
        ss = np.array([0.93 + 0.001 * i, 4.5 - 0.01 * i], np.float32)
        npread.set_scaling_params(ss)
        events = sa.SignalAnalysis(npread, an).load_events()
        raw = z['raw'][i][:int(z['length'][i])]
        got = ER.derive_event_table(raw, z['range'][i], z['digitisation'][i], z['offset'][i],
                                    bcs[i]['moves'], bcs[i]['sequence'], bcs[i]['qstring'],
                                    bcs[i]['first_sample'], bcs[i]['block_stride'], ss)
        for col in ('mean', 'stdv', 'scaled_mean'):
            a = np.asarray(events[col].values, np.float32)
            assert events[col].values.dtype == np.float32, col
            assert np.array_equal(a.view(np.uint32), np.asarray(got[col], np.float32).view(np.uint32)), col
        for col in ('start', 'length', 'move', 'pos', 'end'):
            assert np.array_equal(np.asarray(events[col].values, np.int64), np.asarray(got[col], np.int64)), col
        assert np.array_equal(np.asarray(events['p_model_state'].values, np.float64).view(np.uint64),
                              got['p_model_state'].view(np.uint64))
        assert [s.encode() if isinstance(s, str) else s for s in events['model_state'].values] == \
            list(got['model_state'])
        npread.close()
        checked += 1
    assert checked >= 5
