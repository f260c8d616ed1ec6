"""The poly(A) kernel core (poreplex_b200/csrc/polya_core.cuh), compiled for the HOST,
against the oracle: event stream vs the reference's own scrappie C, interval search and
the whole PolyASignalAnalyzer vs the pandas-free restatement and -- when the reference tree
is present -- vs poreplex/polya.py running verbatim."""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

import hostcheck_util as H

REF = '/root/reference'


@pytest.fixture(scope='module')
def hc():
    return H.load()


def test_median7_network_exhaustive(hc):
    for bits in itertools.product([0.0, 1.0], repeat=7):
        v = np.array(bits, np.float32)
        assert hc.hc_median7(v.ctypes.data_as(C.POINTER(C.c_float))) == np.median(v)
    rng = np.random.default_rng(0)
    for _ in range(500):
        v = rng.normal(0, 1, 7).astype(np.float32)
        assert hc.hc_median7(v.ctypes.data_as(C.POINTER(C.c_float))) == np.median(v)


def test_median5_network_exhaustive(hc):
    for bits in itertools.product([0.0, 1.0], repeat=5):
        v = np.array(bits, np.float32)
        assert hc.hc_median5(v.ctypes.data_as(C.POINTER(C.c_float))) == np.median(v)
    rng = np.random.default_rng(0)
    for _ in range(300):
        v = rng.normal(0, 1, 5).astype(np.float32)
        assert hc.hc_median5(v.ctypes.data_as(C.POINTER(C.c_float))) == np.median(v)


def test_pairwise_sum_matches_numpy(hc):
    rng = np.random.default_rng(1)
    for n in list(range(0, 20)) + [127, 128, 129, 255, 256, 257, 1000, 4097, 8192, 8193, 70001]:
        a = rng.normal(100, 5, n).astype(np.float32)
        got = hc.hc_pairwise_sum(a.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(n))
        assert np.float32(got) == (np.sum(a) if n else np.float32(0)), n


def _window_inputs(rng, n):
    levels = np.repeat(rng.normal(100, 12, n // 25 + 1), 25)[:n] + rng.normal(0, 2.0, n)
    gain, off = 1443.0 / 8192.0, 9.0
    raw = np.clip(np.round(levels / gain - off), -32768, 32767).astype(np.int16)
    return raw, gain, off


def test_event_stream_matches_reference_scrappie(hc, oracle_mod, preset):
    """raw -> pA -> scale -> medfilt(7) -> detect_events as ONE stream == the reference's
    event_detection.c on scipy.signal.medfilt output."""
    from scipy.signal import medfilt
    P = H.polya_params(preset['polya_dwell'])
    ed = preset['polya_dwell']['event_detection']
    detect = oracle_mod.detect_events_ref if oracle_mod.have_ref_scrappie() \
        else oracle_mod.detect_events_restated
    rng = np.random.default_rng(2)
    for n in (1, 5, 13, 14, 39, 40, 41, 100, 1500, 9000):
        raw, gain, off = _window_inputs(rng, n)
        scale, shift = np.float32(0.97), np.float32(3.5)
        pa = np.array(gain * (raw + off), dtype=np.float32)
        sig = medfilt(np.poly1d(np.array([scale, shift], np.float32))(pa), 7) if n >= 1 else pa
        want = detect(sig, **ed)
        cap = n + 2
        st = np.zeros(cap, np.uint64); ln = np.zeros(cap, np.float32)
        mn = np.zeros(cap, np.float32); sd = np.zeros(cap, np.float32)
        k = hc.hc_detect_events(raw.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_double(gain),
                                C.c_double(off), C.c_float(scale), C.c_float(shift), C.byref(P),
                                st.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p),
                                mn.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p),
                                C.c_int64(cap))
        assert k == len(want), n
        assert np.array_equal(st[:k], want['start'])
        assert np.array_equal(ln[:k], want['length'])
        assert np.array_equal(mn[:k], want['mean'], equal_nan=True)
        assert np.array_equal(sd[:k], want['stdv'], equal_nan=True)


def test_plain_event_stream_matches_reference_scrappie(hc, oracle_mod):
    """The stream behind pb2_detect_events (float32 signal in, no filtering) == the reference's
    event_detection.c for the csupport default windows (30, 120: the 512-entry ring), the
    preset's (7, 20: both rings) and a short-detector window larger than the long one."""
    detect = oracle_mod.detect_events_ref if oracle_mod.have_ref_scrappie() \
        else oracle_mod.detect_events_restated
    rng = np.random.default_rng(5)
    cases = [(dict(window_length1=30, window_length2=120, threshold1=3.0, threshold2=9.0, peak_height=8.0), (512,)),
             (dict(window_length1=7, window_length2=20, threshold1=3.0, threshold2=8.0, peak_height=4.0), (64, 512)),
             (dict(window_length1=25, window_length2=10, threshold1=2.0, threshold2=5.0, peak_height=1.0), (64, 512)),
             (dict(window_length1=2, window_length2=255, threshold1=1.5, threshold2=4.0, peak_height=0.5), (512,))]
    for kw, rings in cases:
        P = H.PolyaParamsC()
        P.w1, P.w2 = kw['window_length1'], kw['window_length2']
        P.thr1, P.thr2, P.peak_height = kw['threshold1'], kw['threshold2'], kw['peak_height']
        for n in (1, 2, 13, 59, 60, 61, 239, 240, 241, 1000, 20000):
            step = int(rng.integers(5, 60))
            sig = (np.repeat(rng.normal(100, 12, n // step + 1), step)[:n] +
                   rng.normal(0, 2.0, n)).astype(np.float32)
            want = detect(sig, **kw)
            for ring in rings:
                cap = n + 2
                st = np.zeros(cap, np.uint64); ln = np.zeros(cap, np.float32)
                mn = np.zeros(cap, np.float32); sd = np.zeros(cap, np.float32)
                k = hc.hc_detect_events_plain(sig.ctypes.data_as(C.c_void_p), C.c_int64(n), C.byref(P),
                                              C.c_int(ring), st.ctypes.data_as(C.c_void_p),
                                              ln.ctypes.data_as(C.c_void_p), mn.ctypes.data_as(C.c_void_p),
                                              sd.ctypes.data_as(C.c_void_p), C.c_int64(cap))
                assert k == len(want), (kw, n, ring)
                assert np.array_equal(st[:k], want['start']), (kw, n, ring)
                assert np.array_equal(ln[:k], want['length']), (kw, n, ring)
                assert np.array_equal(mn[:k], want['mean'], equal_nan=True), (kw, n, ring)
                assert np.array_equal(sd[:k], want['stdv'], equal_nan=True), (kw, n, ring)


def _make_case(rng, kind):
    na = int(rng.integers(3000, 6000)); npa = int(rng.integers(100, 6000)); nt = int(rng.integers(500, 8000))
    lvl, sd = 108.95, 1.8
    if kind == 'shift':
        lvl = 108.95 + rng.choice([-9, -6, 6, 9, 14])
    if kind == 'noisy':
        sd = 7.0
    a = rng.normal(80, 7, na)
    pa = rng.normal(lvl, sd, npa)
    if kind == 'spikes':
        for _ in range(int(rng.integers(1, 6))):
            p0 = int(rng.integers(10, max(11, npa - 60))); w = int(rng.integers(5, 140))
            pa[p0:p0 + w] = rng.normal(80, 4, len(pa[p0:p0 + w]))
    tl = np.repeat(rng.normal(95, 12, nt // 12 + 1), 12)[:nt] + rng.normal(0, 2, nt)
    sig = np.concatenate([a, pa, tl])
    b, e = na // 15, (na + npa) // 15
    if kind == 'open':
        rr = (b, None)
    elif kind == 'short_rough':
        rr = (b, b + max(1, (e - b) // 3))
    elif kind == 'notail':
        rr = (b, None); sig = np.concatenate([a, tl])
    else:
        rr = (b + int(rng.integers(-3, 4)), e + int(rng.integers(-3, 4)))
    # express the scaled-space signal as int16 DAC + calibration + (scale, shift)
    scale, shift = np.float32(0.93 + 0.1 * rng.random()), np.float32(rng.normal(5, 3))
    gain, off = (1200.0 + 250 * rng.random()) / 8192.0, float(rng.integers(0, 20))
    raw = np.clip(np.round(((sig - shift) / scale) / gain - off), -32768, 32767).astype(np.int16)
    return raw, gain, off, scale, shift, rr


def _same(want, got):
    if want is None or got is None:
        return want is None and got is None
    return (want['begin'] == got['begin'] and want['end'] == got['end'] and
            want['dwell_time'] == got['dwell_time'] and
            [tuple(float(x) for x in s) for s in want['spikes']] == [tuple(s) for s in got['spikes']])


def test_polya_core_matches_restatement_and_reference(hc, oracle_mod, preset):
    from oracle import polya_restated as PR
    Pr = PR.PolyAParams(preset['polya_dwell'])
    Pc = H.polya_params(preset['polya_dwell'])
    detect = oracle_mod.detect_events_ref if oracle_mod.have_ref_scrappie() \
        else oracle_mod.detect_events_restated
    ref_analyzer = None
    if os.path.isdir(REF):
        from oracle import refshim
        _, _, _, polya_mod, _ = refshim.reference_modules()
        ref_analyzer = polya_mod.PolyASignalAnalyzer(preset['polya_dwell'])

    class FakeRead:
        def __init__(self, sig):
            self.sig, self.sampling_rate, self.polya = sig, 3012.0, None

        def load_signal(self, pool=None, pad=False):
            return self.sig

        def set_polya_tail(self, d):
            self.polya = d

    rng = np.random.default_rng(3)
    found = none = extended = spiky = 0
    for kind in ['plain', 'shift', 'noisy', 'spikes', 'open', 'short_rough', 'notail']:
        for rep in range(12):
            raw, gain, off, scale, shift, rr = _make_case(rng, kind)
            pa = np.array(gain * (raw + off), dtype=np.float32)
            scaled = np.poly1d(np.array([scale, shift], np.float32))(pa)
            want = PR.analyze(Pr, scaled, 3012.0, rr, 15, detect_events=detect)
            if ref_analyzer is not None:
                fr = FakeRead(scaled)
                ref_analyzer(fr, rr, 15)
                assert _same(fr.polya, want), (kind, rep, 'restatement vs reference')
            R = H.PolyaResultC()
            hc.hc_polya(C.byref(Pc), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(raw)),
                        C.c_double(gain), C.c_double(off), C.c_float(scale), C.c_float(shift),
                        C.c_int32(rr[0]), C.c_int32(-1 if rr[1] is None else rr[1]), C.byref(R))
            got = H.result_to_dict(R, 3012.0)
            assert _same(want, got), (kind, rep, want, got)
            for cap in (8, 4096):          # replay cache: overflowing and roomy
                R2 = H.PolyaResultC()
                hc.hc_polya_cached(C.byref(Pc), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(raw)),
                                   C.c_double(gain), C.c_double(off), C.c_float(scale),
                                   C.c_float(shift), C.c_int32(rr[0]),
                                   C.c_int32(-1 if rr[1] is None else rr[1]), C.byref(R2),
                                   C.c_int(cap))
                assert _same(want, H.result_to_dict(R2, 3012.0)), (kind, rep, cap)
            for cap in (0, 8, 4096):       # the literal one-loop-per-walk formulation: same bytes
                R3 = H.PolyaResultC()
                hc.hc_polya_nested(C.byref(Pc), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(raw)),
                                   C.c_double(gain), C.c_double(off), C.c_float(scale),
                                   C.c_float(shift), C.c_int32(rr[0]),
                                   C.c_int32(-1 if rr[1] is None else rr[1]), C.byref(R3),
                                   C.c_int(cap))
                assert _same(want, H.result_to_dict(R3, 3012.0)), (kind, rep, cap, 'nested')
                assert (R3.found, R3.n_spikes, R3.begin, R3.end, R3.dwell_samples, R3.extensions,
                        R3.flags) == (R.found, R.n_spikes, R.begin, R.end, R.dwell_samples,
                                      R.extensions, R.flags), (kind, rep, cap)
            found += got is not None
            none += got is None
            extended += R.extensions > 0
            spiky += bool(got and got['spikes'])
    assert found > 40 and none > 10 and extended > 5 and spiky > 15


def test_polya_single_loop_equals_literal_formulation(hc, preset):
    """polya_analyze (one loop, one event-stream step: what k_polya runs) against
    polya_analyze_nested (the reference's control flow, walk by walk) on a few hundred random
    reads of every kind, with and without the replay cache: the whole result record, byte for
    byte (spike table, extension count and overflow flag included)."""
    Pc = H.polya_params(preset['polya_dwell'])
    rng = np.random.default_rng(11)
    seen = {'found': 0, 'none': 0, 'extended': 0, 'spikes': 0}
    for kind in ['plain', 'shift', 'noisy', 'spikes', 'open', 'short_rough', 'notail']:
        for rep in range(60):
            raw, gain, off, scale, shift, rr = _make_case(rng, kind)
            args = (C.byref(Pc), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(raw)), C.c_double(gain),
                    C.c_double(off), C.c_float(scale), C.c_float(shift), C.c_int32(rr[0]),
                    C.c_int32(-1 if rr[1] is None else rr[1]))
            for cap in (0, 16, 4096):
                A, B = H.PolyaResultC(), H.PolyaResultC()
                hc.hc_polya_cached(*args, C.byref(A), C.c_int(cap))
                hc.hc_polya_nested(*args, C.byref(B), C.c_int(cap))
                assert bytes(A) == bytes(B), (kind, rep, cap)
            seen['found'] += A.found
            seen['none'] += 1 - A.found
            seen['extended'] += A.extensions > 0
            seen['spikes'] += A.n_spikes > 0
    assert seen['found'] > 150 and seen['none'] > 50 and seen['extended'] > 20 and seen['spikes'] > 60, seen
