"""Tensor-core LSTM path (csrc/kernels_lstm_tc.cu) against the exact f32 kernels, which the
other GPU tests tie to the CPU oracle bit for bit.

The tensor-core kernels are approximate by construction (split-fp16 products, fp32
accumulation in TMEM, MUFU gate functions).  What must hold:
  * their class logits stay within the per-window error bound the margin test assumes
    (delta0 + gain * shift under the coarse probe evaluation);
  * every window the margin test calls safe gets exactly the exact path's
    (barcode, guess, score);
  * the default path (tensor-core + exact re-run of the unsafe windows) returns the exact
    path's integer outputs for every window.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DELTA0, GAIN = 1e-3, 0.1     # pb2_context::demux_margin_delta / demux_probe_gain defaults


def _windows(n, seed, T=300, min_len=60):
    """Normalised adapter-like windows: piecewise-constant levels + noise, median/MAD
    normalised, left-padded with -1000 to T (barcoding.py:77-98)."""
    rng = np.random.default_rng(seed)
    win = np.full((n, T), -1000.0, np.float32)
    for i in range(n):
        L = int(rng.integers(min_len, T + 1))
        dwell = int(rng.integers(3, 12))
        lv = np.repeat(rng.normal(0, 1.2, size=L // dwell + 1), dwell)[:L] + rng.normal(0, 0.4, size=L)
        med = np.median(lv)
        mad = np.median(np.abs(lv - med))
        win[i, T - L:] = ((lv - med) / max(0.01, 1.4826 * mad)).astype(np.float32)
    return win


def _exact(eng, win_dev):
    import torch
    eng.set_fast_lstm(False)
    try:
        out = eng.demux_predict(win_dev)
        torch.cuda.synchronize()
    finally:
        eng.set_fast_lstm(True)
    return [o.cpu().numpy() for o in out]


def test_tc_demux_close_to_exact_and_safe_calls_identical(eng_stock):
    import torch
    dev = torch.device('cuda', 0)
    win = _windows(6000, seed=3)
    win[0, :] = 0.0                       # flat window (MAD = 0 upstream)
    win[1, :] = -1000.0                   # all padding
    wd = torch.from_numpy(win).to(dev)
    p_ex, bc_ex, g_ex, s_ex = _exact(eng_stock, wd)
    p_tc, lg_tc, bc_tc, g_tc, s_tc, unsafe, sens = [o.cpu().numpy() for o in eng_stock.demux_predict_tc(wd)]
    torch.cuda.synchronize()
    assert eng_stock.recheck_stats()[1] == 0, 'tensor-core kernel barrier time-out'
    # class logits relative to the called class, recovered from the probabilities
    pe, pt = p_ex[:, :5].astype(np.float64), p_tc[:, :5].astype(np.float64)
    am = pe.argmax(1)
    with np.errstate(divide='ignore', invalid='ignore'):
        le, lt = np.log(pe), np.log(pt)
    d = np.abs((le - le[np.arange(len(am)), am][:, None]) - (lt - lt[np.arange(len(am)), am][:, None]))
    d[~(pe > 1e-30)] = 0
    err = d.max(1)
    bound = DELTA0 + GAIN * sens.astype(np.float64)
    dp = np.abs(pe - pt).max()
    print('tc vs exact: logit error median %.2e max %.2e; max err/bound %.3f; max |dp| %.3e; '
          'unsafe %.2f%%' % (np.median(err), err.max(), (err / bound).max(), dp, 100.0 * unsafe.mean()))
    assert np.median(err) < 1e-4
    assert (err <= bound).all(), 'approximation error exceeds the bound the margin test assumes'
    assert dp < 2e-3
    safe = unsafe == 0
    assert safe.mean() > 0.7
    assert np.array_equal(bc_tc[safe], bc_ex[safe])
    assert np.array_equal(g_tc[safe], g_ex[safe])
    assert np.array_equal(s_tc[safe], s_ex[safe])


def test_tc_demux_with_recheck_equals_exact(eng_stock):
    import torch
    dev = torch.device('cuda', 0)
    win = _windows(5000, seed=5, min_len=100)
    wd = torch.from_numpy(win).to(dev)
    p_ex, bc_ex, g_ex, s_ex = _exact(eng_stock, wd)
    p, bc, g, s = [o.cpu().numpy() for o in eng_stock.demux_predict(wd)]
    torch.cuda.synchronize()
    rechecked, timeouts = eng_stock.recheck_stats()
    assert timeouts == 0
    assert 0 < rechecked < 0.3 * len(win)
    assert np.array_equal(bc, bc_ex)
    assert np.array_equal(g, g_ex)
    assert np.array_equal(s, s_ex)
    assert np.abs(p - p_ex).max() < 2e-3
    # a tile boundary case: 1, 127, 128, 129 rows
    for n in (1, 127, 128, 129):
        sub = wd[:n].contiguous()
        _, b1, g1, s1 = [o.cpu().numpy() for o in eng_stock.demux_predict(sub)]
        assert np.array_equal(b1, bc_ex[:n]) and np.array_equal(g1, g_ex[:n]) and np.array_equal(s1, s_ex[:n])


def test_tc_whole_path_integer_outputs_equal_exact(eng_short, preset_short, monkeypatch):
    """Default path vs exact-only path through pb2_analyze_host on calibrated synthetic reads."""
    from poreplex_b200 import synth
    rd = synth.to_numpy(synth.generate_reads(3000, synth.SynthSpec.for_length(4000), preset_short, seed=31))
    n, L = rd['raw'].shape
    args = (rd['raw'].reshape(-1), np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int64),
            rd['range'], rd['digitisation'], rd['offset'])
    fast = eng_short.analyze_host(*args)
    rerun, timeouts = eng_short.recheck_stats()
    assert timeouts == 0
    eng_short.set_fast_lstm(False)
    try:
        exact = eng_short.analyze_host(*args)
    finally:
        eng_short.set_fast_lstm(True)
    for k in ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'counts', 'label'):
        assert np.array_equal(fast[k], exact[k]), k
    d = np.abs(fast['scale_shift'].astype(np.float64) - exact['scale_shift'])
    print('scale/shift: max |d| %.2e / %.2e; exactly re-run reads %d of %d'
          % (d[:, 0].max(), d[:, 1].max(), rerun, n))
    assert d[:, 0].max() <= 3.4e-5 and d[:, 1].max() <= 1.5e-2      # inside the assumed box
    assert np.median(d[:, 0]) < 1e-6 and np.median(d[:, 1]) < 5e-5
    assert (d.max(1) == 0).sum() >= rerun                            # re-run reads are exact
    assert 0 < rerun < 0.35 * n
    assert (fast['barcode_score'] >= 0).sum() > 2000
    # chunked host paths: same integer outputs whatever the tiling of reads.  `streamed` keeps the
    # whole batch resident and resolves the unsafe reads of all chunks as one sub-batch at the end;
    # `arena` resolves them chunk by chunk.
    other = synth.to_numpy(synth.generate_reads(n, synth.SynthSpec.for_length(4000), preset_short, seed=32))
    for pipeline in ('streamed', 'streamed-early', 'arena'):
        # a different batch first: nothing may be taken from scratch a previous call left behind
        eng_short.analyze_host(other['raw'].reshape(-1), *args[1:3], other['range'], other['digitisation'],
                               other['offset'])
        monkeypatch.setenv('POREPLEX_B200_HOST_CHUNK_ELEMS', str(1_000_000))
        monkeypatch.setenv('POREPLEX_B200_HOST_PIPELINE', pipeline.split('-')[0])
        # streamed-early: the flagged reads of the first half are re-run half way through (what the
        # path does on its own when it finds the GPU waiting for the bus there), the rest at the end
        monkeypatch.setenv('POREPLEX_B200_HOST_EARLY_RESOLVE', '1' if pipeline.endswith('early') else '0')
        piped = eng_short.analyze_host(*args)
        rerun_p = eng_short.recheck_stats()[0]
        monkeypatch.delenv('POREPLEX_B200_HOST_CHUNK_ELEMS')
        monkeypatch.delenv('POREPLEX_B200_HOST_PIPELINE')
        monkeypatch.delenv('POREPLEX_B200_HOST_EARLY_RESOLVE')
        for k in ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'counts', 'label'):
            assert np.array_equal(piped[k], exact[k]), (pipeline, k)
        dp = np.abs(piped['scale_shift'].astype(np.float64) - exact['scale_shift'])
        assert dp[:, 0].max() <= 3.4e-5 and dp[:, 1].max() <= 1.5e-2
        if pipeline.startswith('streamed'):
            assert 0 < rerun_p < 0.35 * n and (dp.max(1) == 0).sum() >= rerun_p


def test_tc_demux_extreme_windows_equal_exact(eng_stock):
    """Saturating and degenerate windows through the default path: huge normalised values
    (|x| up to 100: MAD clamp at 0.01), constant windows, alternating signs, windows that are
    all padding or have a single real sample -- every call must equal the exact kernels'."""
    import torch
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(11)
    n, T = 1024, 300
    win = _windows(n, seed=13)
    win[0:64] = rng.normal(0, 30, (64, T)).astype(np.float32)                 # wild amplitudes
    win[64:96] = np.where(rng.random((32, T)) < 0.5, 100.0, -100.0)           # clamp-level values
    win[96:128] = 0.0
    win[128:160] = -1000.0                                                    # all padding
    win[160:192] = -1000.0
    win[160:192, -1] = rng.normal(0, 1, 32).astype(np.float32)                # one real sample
    win[192:224] = np.tile(np.where(np.arange(T) % 2 == 0, 1.0, -1.0), (32, 1))
    wd = torch.from_numpy(np.ascontiguousarray(win, np.float32)).to(dev)
    p_ex, bc_ex, g_ex, s_ex = _exact(eng_stock, wd)
    p, bc, g, s = [o.cpu().numpy() for o in eng_stock.demux_predict(wd)]
    torch.cuda.synchronize()
    assert eng_stock.recheck_stats()[1] == 0
    assert np.isfinite(p).all()
    assert np.array_equal(bc, bc_ex) and np.array_equal(g, g_ex) and np.array_equal(s, s_ex)


def test_rerun_causes_are_reported(eng_short, preset_short):
    from poreplex_b200 import synth
    rd = synth.to_numpy(synth.generate_reads(2000, synth.SynthSpec.for_length(4000), preset_short, seed=5))
    n, L = rd['raw'].shape
    eng_short.analyze_host(rd['raw'].reshape(-1), np.arange(n, dtype=np.int64) * L,
                           np.full(n, L, np.int64), rd['range'], rd['digitisation'], rd['offset'])
    rerun, timeouts = eng_short.recheck_stats()
    causes = eng_short.rerun_causes()
    assert timeouts == 0 and rerun > 0
    assert set(causes) == {'qc_edge', 'segmentation', 'barcode_call'}
    assert max(causes.values()) <= rerun <= sum(causes.values())


def test_audit_mode_counts_and_finds_no_disagreement(eng_short, preset_short):
    """pb2_set_audit_fraction: a sample of the reads that passed every guard is re-run through
    the exact kernels as well; none of them may disagree, and the outputs stay the exact ones."""
    from poreplex_b200 import synth
    rd = synth.to_numpy(synth.generate_reads(3000, synth.SynthSpec.for_length(4000), preset_short, seed=9))
    n, L = rd['raw'].shape
    args = (rd['raw'].reshape(-1), np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int64),
            rd['range'], rd['digitisation'], rd['offset'])
    plain = eng_short.analyze_host(*args)
    rerun_plain = eng_short.recheck_stats()[0]
    eng_short.audit_stats()
    eng_short.set_audit_fraction(0.25)
    try:
        audited_run = eng_short.analyze_host(*args)
        rerun_audit = eng_short.recheck_stats()[0]
        audited, mismatched = eng_short.audit_stats()
    finally:
        eng_short.set_audit_fraction(0.0)
    assert 0.15 * n < audited < 0.35 * n
    assert mismatched == 0
    # (accepted windows are compacted in read order, so the unsafe set is the same in both runs)
    assert rerun_audit == rerun_plain + audited
    for k in ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'label', 'counts'):
        assert np.array_equal(plain[k], audited_run[k]), k


def test_fast_path_is_reproducible(eng_short, preset_short):
    """Two runs over the same batch give the same bits in EVERY output, floats included: tiles of
    the tensor-core kernels are formed in read order (k_window_accept), so neither the
    approximate values nor the set of re-run reads depend on warp scheduling."""
    from poreplex_b200 import synth
    rd = synth.to_numpy(synth.generate_reads(3000, synth.SynthSpec.for_length(4000), preset_short, seed=21))
    n, L = rd['raw'].shape
    args = (rd['raw'].reshape(-1), np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int64),
            rd['range'], rd['digitisation'], rd['offset'])
    first = {k: np.array(v, copy=True) for k, v in eng_short.analyze_host(*args).items()
             if isinstance(v, np.ndarray)}
    reruns = eng_short.recheck_stats()[0]
    for _ in range(2):
        again = eng_short.analyze_host(*args)
        assert eng_short.recheck_stats()[0] == reruns
        for k, v in first.items():
            assert np.array_equal(v, again[k], equal_nan=True) if v.dtype.kind == 'f' \
                else np.array_equal(v, again[k]), k


def test_fused_probes_equal_separate_probes(eng_stock, monkeypatch):
    """k_lstm_tc_probes runs both sensitivity probes as one ring of work items; its outputs must be
    the bits of the two separate k_lstm_tc launches it replaces (same MMAs per accumulator column,
    same gate code, same pseudo-random rounding), so that the guard statistics carry over.  The
    logit shift each probe measures is compared per probe and combined, on ragged tile counts."""
    import torch
    dev = torch.device('cuda', 0)
    win = _windows(148 * 128 + 77, seed=17, min_len=40)       # a full wave of tiles and a ragged one
    win[5, :] = -1000.0
    wd = torch.from_numpy(win).to(dev)
    for which in ('', '1', '2'):
        if which:
            monkeypatch.setenv('POREPLEX_B200_SENS_PROBE', which)
        fused = [o.cpu().numpy() for o in eng_stock.demux_predict_tc(wd)]
        monkeypatch.setenv('POREPLEX_B200_SPLIT_PROBES', '1')
        split = [o.cpu().numpy() for o in eng_stock.demux_predict_tc(wd)]
        monkeypatch.delenv('POREPLEX_B200_SPLIT_PROBES')
        if which:
            monkeypatch.delenv('POREPLEX_B200_SENS_PROBE')
        assert eng_stock.recheck_stats()[1] == 0, 'tensor-core kernel barrier time-out'
        assert fused[6].max() > 0
        assert np.array_equal(fused[6].view(np.uint32), split[6].view(np.uint32)), 'probe %r' % which
        assert np.array_equal(fused[5], split[5])
        for a, b in zip(fused[:5], split[:5]):
            assert np.array_equal(a, b)


def test_fused_scaler_equals_two_kernel_scaler(eng_short, preset_short, monkeypatch):
    """k_lstm_tc_scaler2 steps both scaler layers in one kernel (layer 2 one step behind layer 1,
    the first layer's sequence never leaves TMEM); every output of the whole path -- the approximate
    (scale, shift) of guard-passing reads included -- must be the bits of the two-launch version
    whose error statistics DESIGN.md 3a validates.  Ragged lengths: tiles start at different steps
    of the zero head, some reads are too short to be scaled at all."""
    from poreplex_b200 import synth
    rd = synth.to_numpy(synth.generate_reads(3000, synth.SynthSpec.for_length(4000), preset_short, seed=77))
    n, L = rd['raw'].shape
    rng = np.random.default_rng(5)
    ln = np.full(n, L, np.int64)
    ln[::3] = rng.integers(600, L, size=len(ln[::3]))
    ln[:130] = np.sort(rng.integers(905, 1500, size=130))      # a whole tile of short heads
    args = (rd['raw'].reshape(-1), np.arange(n, dtype=np.int64) * L, ln,
            rd['range'], rd['digitisation'], rd['offset'])
    fused = {k: np.array(v, copy=True) for k, v in eng_short.analyze_host(*args).items()
             if isinstance(v, np.ndarray)}
    reruns = eng_short.recheck_stats()
    assert reruns[1] == 0, 'tensor-core kernel barrier time-out'
    monkeypatch.setenv('POREPLEX_B200_SPLIT_SCALER', '1')
    split = eng_short.analyze_host(*args)
    monkeypatch.delenv('POREPLEX_B200_SPLIT_SCALER')
    assert eng_short.recheck_stats() == reruns
    assert (fused['status'] == 0).sum() > 1500 and 0 < reruns[0] < 0.5 * n
    for k, v in fused.items():
        assert np.array_equal(v, split[k], equal_nan=True) if v.dtype.kind == 'f' \
            else np.array_equal(v, split[k]), k

