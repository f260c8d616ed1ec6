"""Host-side logic: parameter loading (HDF5 walker, presets, HMM baking), read packing,
synthetic generator, the reference-over-shims oracle harness, world_size-2 count reduce."""
import hashlib
import os

import numpy as np
import pytest

from golden_util import load_golden, golden_reads, golden_basecalls, normalise_result

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_packaged_weights_match_reference_files():
    from poreplex_b200 import params
    p = params.load_preset()
    sc = params.load_scaler_model(p['signal_processing']['scaler_model'])
    dm = params.load_demux_model(p['demultiplexing']['demux_model'])
    assert sc.l1.kernel.shape == (1, 192) and sc.l2.recurrent.shape == (48, 192)
    assert sc.input_defs == {'dtype': 'float32', 'stride': 15, 'length': 30000, 'min_length': 9000}
    assert abs(sc.output_transform['scale_mean'] - 0.9553510773987666) < 1e-15
    assert np.allclose(sc.dense_bias, [0.05570572, -0.0642097])
    assert dm.l2.kernel.shape == (96, 256) and dm.n_classes == 5 and len(dm.calibration) == 29
    assert abs(dm.calibration[18] - 0.9797275074316903) < 1e-15
    assert (sc.l1.implementation, dm.fwd.implementation) == (1, 2)
    if os.path.isdir(REF):       # byte-level check of the HDF5 walker against SURVEY App. C
        path = os.path.join(REF, 'poreplex/presets/MIN106-RNA001/scaler-r3.hdf5')
        blob = open(path, 'rb').read()
        assert hashlib.sha256(blob).hexdigest() == p['_provenance']['scaler-r3.hdf5']
        sc2 = params.load_scaler_model(path)
        assert np.array_equal(sc2.l1.recurrent, sc.l1.recurrent)
        direct = np.frombuffer(blob[13600:13600 + 48 * 192 * 4], '<f4').reshape(48, 192)
        assert np.array_equal(direct, sc.l1.recurrent)


def test_hmm_tables_baked_order_and_topology(preset):
    from poreplex_b200 import params
    seg = params.HmmTables(preset['segmentation_model'])
    assert seg.names == ['adapter', 'leader-high', 'leader-low', 'polya-tail', 'pre-leader',
                         'transcript']
    assert seg.left_to_right and seg.n_edges == 12
    assert np.isclose(seg.log_start[4], np.log(0.99)) and np.isneginf(seg.log_start[0])
    un = params.HmmTables(preset['unsplit_read_detection_model'])
    assert not un.left_to_right and un.n_edges == 16
    # the oracle's independent baking agrees bit for bit
    from oracle import oracle as O
    h, names = O.bake_hmm(preset['segmentation_model'])
    assert names == seg.names
    assert [h.in_src[k] for k in range(12)] == list(seg.in_src[:12])
    assert [h.in_logp[k] for k in range(12)] == list(seg.in_logp[:12])
    assert [h.lsp[s][0] for s in range(6)] == list(seg.lsp[:6, 0])


def test_pack_reads_alignment():
    from poreplex_b200.engine import SignalEngine
    sigs = [np.arange(n, dtype=np.int16) for n in (0, 1, 7, 8, 9, 4000, 16001)]
    raw, off, ln = SignalEngine.pack_reads(sigs)
    assert np.all(off % 8 == 0) and list(ln) == [0, 1, 7, 8, 9, 4000, 16001]
    for s, o in zip(sigs, off):
        assert np.array_equal(raw[o:o + len(s)], s)
    assert len(raw) >= off[-1] + 16001


def test_synthetic_generator_is_seeded(preset):
    from poreplex_b200 import synth
    spec = synth.SynthSpec.for_length(4000)
    a = synth.to_numpy(synth.generate_reads(8, spec, preset, seed=3))
    b = synth.to_numpy(synth.generate_reads(8, spec, preset, seed=3))
    c = synth.to_numpy(synth.generate_reads(8, spec, preset, seed=4))
    assert np.array_equal(a['raw'], b['raw']) and not np.array_equal(a['raw'], c['raw'])
    assert a['raw'].dtype == np.int16 and a['raw'].shape == (8, 4000)
    assert np.array_equal(a['gain'], a['range'] / a['digitisation'])


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_reference_over_shims_reproduces_committed_golden():
    """Re-run the reference's process_batch verbatim (config 1: --trim-adapter only, and
    with --barcoding) and compare with the committed golden dicts: the fixtures are
    reproducible and the shim/oracle stack is deterministic."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden
    z, doc = load_golden('short4k')
    preset, rd, read_ids, basecalls = make_golden.build_inputs('bench-short', 4000, 88, 202)
    assert np.array_equal(rd['raw'], z['raw'])
    _, res, cap = make_golden.run_reference(preset, 'bench-short', rd, read_ids, basecalls,
                                            trim_adapter=True, barcoding=True)
    got = make_golden.jsonable(res)
    assert [normalise_result(r) for r in got] == \
        [normalise_result(r) for r in doc['results_trim_barcoding']]
    assert cap['segments'] == doc['segments']


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from poreplex_b200.sharding import shard_range, reduce_counts
    lo, hi = shard_range(1001, rank, world)
    rng = np.random.default_rng(0)
    status = rng.integers(0, 11, 1001); label = rng.integers(0, 4, 1001); bc = rng.integers(-1, 4, 1001)
    local = np.zeros((4, 5, 11), np.int64)
    np.add.at(local, (label[lo:hi], bc[lo:hi] + 1, status[lo:hi]), 1)
    total = reduce_counts(torch.from_numpy(local)).numpy()
    want = np.zeros((4, 5, 11), np.int64)
    np.add.at(want, (label, bc + 1, status), 1)
    q.put((rank, lo, hi, bool(np.array_equal(total, want))))
    dist.destroy_process_group()


def test_world_size_2_count_reduce_gloo():
    """N > 1 path on CPU: block partition of read indices + the one all-reduce of the
    per-(label, barcode, status) counts, gloo backend."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1:3] == (0, 500) and res[1][1:3] == (500, 1001)
    assert all(r[3] for r in res)


def test_config_digest_is_content_based():
    """get_engine keys on the CONTENT of the configuration (pipeline.py:204 re-pickles it for
    every batch), not on object identity."""
    import pickle
    from poreplex_b200 import engine, params
    cfg = dict(params.load_preset(), barcoding=True, barcoding_quality_filter=18, inputdir='/a')
    again = pickle.loads(pickle.dumps(cfg))
    assert again['segmentation_model'] is not cfg['segmentation_model']
    assert engine.config_digest(again) == engine.config_digest(cfg)
    again['inputdir'] = '/b'                                   # not an engine parameter
    assert engine.config_digest(again) == engine.config_digest(cfg)
    for change in (lambda c: c.__setitem__('barcoding_quality_filter', 20),
                   lambda c: c['signal_processing'].__setitem__('scaler_qc_threshold', 0.05),
                   lambda c: c['polya_dwell'].__setitem__('spike_weight', 2.0),
                   lambda c: c['unsplit_read_detection'].__setitem__('window_size', 9.0),
                   lambda c: c['demultiplexing'].__setitem__('minimum_dna_length', 100),
                   lambda c: c.__setitem__('barcoding', False)):
        other = pickle.loads(pickle.dumps(cfg))
        change(other)
        assert engine.config_digest(other) != engine.config_digest(cfg)
