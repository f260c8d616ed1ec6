"""The compressed upload (pb2_batch.packed): streamvbyte-16 bodies of VBZ chunks cross the bus and
the GPU rebuilds the int16 samples (k_svb16_decode).  Bit-exact by construction: the decoder must
return the samples the host encoder was given, and the whole path must not notice the detour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pack(sigs):
    lens = np.array([len(s) for s in sigs], np.int64)
    pad = (lens + 7) // 8 * 8
    off = np.zeros(len(sigs), np.int64)
    off[1:] = np.cumsum(pad)[:-1]
    raw = np.zeros(int(pad.sum()) + 8, np.int16)
    for s, o in zip(sigs, off):
        raw[o:o + len(s)] = s
    return raw, off, lens


def test_decoder_returns_the_samples(eng_short):
    import torch
    from poreplex_b200 import fast5_loader
    fast5_loader.build()
    rng = np.random.default_rng(5)
    sigs = [rng.integers(-32768, 32768, size=n).astype(np.int16)
            for n in (0, 1, 7, 8, 9, 255, 256, 257, 4000, 4001, 12345)]
    sigs += [(np.cumsum(rng.integers(-25, 26, size=n)) + 480).astype(np.int16) for n in (4000, 100003)]
    sigs += [np.full(999, -32768, np.int16), np.tile(np.array([32767, -32768], np.int16), 600)]
    raw, off, lens = _pack(sigs)
    pk, po = fast5_loader.svb16_encode(raw, off, lens)
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    got, err = eng_short.svb16_decode(t(pk), t(po), t(off), t(lens), raw.size)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert int(err.item()) == 0
    for s, o in zip(sigs, off):
        assert np.array_equal(got[o:o + len(s)], s)
    # a stream cut short must be reported, not read past its end
    po_bad = po.copy()
    po_bad[-1] -= 16 * 40
    _, err = eng_short.svb16_decode(t(pk), t(po_bad), t(off), t(lens), raw.size)
    assert int(err.item()) == 1


def test_packed_host_path_equals_int16_host_path(eng_short, preset_short, monkeypatch):
    from poreplex_b200 import fast5_loader, synth
    rd = synth.to_numpy(synth.generate_reads(5000, synth.SynthSpec.for_length(4000), preset_short, seed=77))
    n, L = rd['raw'].shape
    raw = rd['raw'].reshape(-1)
    off = np.arange(n, dtype=np.int64) * L
    ln = np.full(n, L, np.int64)
    ln[::7] -= np.arange(len(ln[::7])) % 60                      # ragged tails
    cal = (rd['range'], rd['digitisation'], rd['offset'])
    pk, po = fast5_loader.svb16_encode(raw, off, ln)
    assert po[-1] < 0.6 * raw.nbytes                             # the point: fewer bytes to upload
    plain = eng_short.analyze_host(raw, off, ln, *cal, polya=True)
    packed = eng_short.analyze_host(None, off, ln, *cal, polya=True, packed=(pk, po))
    monkeypatch.setenv('POREPLEX_B200_HOST_CHUNK_ELEMS', str(3_000_000))     # pipelined: 7 chunks
    piped = eng_short.analyze_host(None, off, ln, *cal, polya=True, packed=(pk, po))
    monkeypatch.delenv('POREPLEX_B200_HOST_CHUNK_ELEMS')
    for other in (packed, piped):
        for k in plain:
            if k == 'class_probs' and other is piped:
                continue        # approximate floats of the tensor-core classifier depend on which
                                # windows share a 128-read tile, i.e. on the chunking
            a, b = plain[k], other[k]
            if a.dtype.fields:
                for f in ('found', 'n_spikes', 'begin', 'end', 'dwell_samples', 'extensions'):
                    assert np.array_equal(a[f], b[f]), (k, f)
            elif a.dtype.kind == 'f':
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), k
            else:
                assert np.array_equal(a, b), k
    # damaged input is a loud error
    po_bad = po.copy()
    po_bad[-1] -= 1600
    from poreplex_b200._native import NativeError
    with pytest.raises(NativeError):
        eng_short.analyze_host(None, off, ln, *cal, packed=(pk, po_bad))
