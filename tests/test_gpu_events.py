"""pb2_detect_events / poreplex_b200.csupport.detect_events (SURVEY.md 8a row A9, 8b "native
ABI today") against the reference's own scrappie event detector compiled from
/root/reference (oracle/_ref) or, where that build is absent, its C restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _signal(rng, n):
    if n == 0:
        return np.zeros(0, np.float32)
    step = int(rng.integers(5, 80))
    return (np.repeat(rng.normal(100, 12, n // step + 1), step)[:n] +
            rng.normal(0, 2.0, n)).astype(np.float32)


def _detect(oracle_mod):
    return oracle_mod.detect_events_ref if oracle_mod.have_ref_scrappie() \
        else oracle_mod.detect_events_restated


def _same(got, want):
    assert got.dtype == want.dtype
    assert len(got) == len(want)
    for f in ('start', 'length', 'pos', 'state'):
        assert np.array_equal(got[f], want[f]), f
    for f in ('mean', 'stdv'):
        assert np.array_equal(got[f], want[f], equal_nan=True), f


@pytest.mark.parametrize('kw', [
    dict(),                                                              # csupport defaults (30, 120)
    dict(window_length1=7, window_length2=20, threshold1=3.0, threshold2=8.0, peak_height=4.0),
    dict(window_length1=25, window_length2=10, threshold1=2.0, threshold2=5.0, peak_height=1.0),
])
def test_detect_events_batch_equals_reference(oracle_mod, kw):
    from poreplex_b200 import csupport
    detect = _detect(oracle_mod)
    rng = np.random.default_rng(11)
    lengths = [1, 2, 13, 0, 59, 60, 61, 239, 240, 241, 1000, 0, 7777, 30000] + \
        [int(x) for x in rng.integers(1, 5000, 200)]
    sigs = [_signal(rng, n) for n in lengths]
    got = csupport.detect_events_batch(sigs, **kw)
    assert len(got) == len(sigs)
    for s, g in zip(sigs, got):
        if s.size == 0:
            assert len(g) == 0 and g.dtype == csupport.EVENT_DTYPE
            continue
        _same(g, detect(s, **kw))


def test_detect_events_single_call_contract(oracle_mod):
    """csupport.c:70-124: float64 / list input is cast, 2-D input is a ValueError, a signal
    without events raises csupport.error."""
    from poreplex_b200 import csupport
    detect = _detect(oracle_mod)
    rng = np.random.default_rng(3)
    sig = _signal(rng, 4000)
    want = detect(sig)
    _same(csupport.detect_events(sig.astype(np.float64)), want)
    _same(csupport.detect_events(list(sig[:500])), detect(sig[:500]))
    _same(csupport.detect_events(sig, 7, 20, 3.0, 8.0, 4.0), detect(sig, 7, 20, 3.0, 8.0, 4.0))
    with pytest.raises(ValueError):
        csupport.detect_events(sig.reshape(2, -1))
    with pytest.raises(csupport.error):
        csupport.detect_events(np.zeros(0, np.float32))
    with pytest.raises(Exception):
        csupport.detect_events(sig, window_length2=300)                 # > 255: unsupported
