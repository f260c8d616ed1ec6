"""poreplex_b200.summary against the reference's own writers (poreplex/io.py:120-184, 236-332):
golden text produced by running the reference classes (tests/golden/make_summary_golden.py),
and the live reference when /root/reference is mounted."""
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_summary_golden as G                     # noqa: E402
from poreplex_b200 import summary                  # noqa: E402
from poreplex_b200.params import STATUS_CODES      # noqa: E402


def _mine(results, config, labels, barcodes, tmp):
    w = summary.SequencingSummaryWriter(config, str(tmp), labels, barcodes)
    w.write_results(results)
    w.close()
    t = summary.FinalSummaryTracker(labels, barcodes)
    t.feed_results(results)
    buf = io.StringIO()
    t.print_results(buf)
    return open(os.path.join(str(tmp), 'sequencing_summary.txt')).read(), buf.getvalue()


def _case(name):
    config, labels, barcodes = G.CONFIGS[name]
    results = G.make_results()
    if not config['barcoding']:
        results = [{k: v for k, v in e.items() if not k.startswith('barcode')} for e in results]
    return results, config, labels, barcodes


@pytest.mark.parametrize('name', sorted(G.CONFIGS))
def test_writers_match_golden_reference_output(name, tmp_path):
    results, config, labels, barcodes = _case(name)
    seq, final = _mine(results, config, labels, barcodes, tmp_path)
    gdir = os.path.join(HERE, 'golden')
    assert seq == open(os.path.join(gdir, 'summary_%s_sequencing_summary.txt' % name)).read()
    assert final == open(os.path.join(gdir, 'summary_%s_final.txt' % name)).read()


@pytest.mark.parametrize('name', sorted(G.CONFIGS))
def test_writers_match_live_reference(name, tmp_path):
    if not os.path.isdir('/root/reference/poreplex'):
        pytest.skip('reference tree not mounted')
    results, config, labels, barcodes = _case(name)
    assert _mine(results, config, labels, barcodes, tmp_path) == G.reference_outputs(results, config, labels, barcodes)


def test_tracker_fed_by_device_histogram_and_batch_writer(tmp_path):
    """feed_counts(int64[4][5][11]) -- what k_counts / the all-reduce produce -- gives the same
    table as feeding the dicts (counts made distinct so that no tie order is involved), and
    write_batch over result arrays gives the same rows as write_results over dicts."""
    results, config, labels, barcodes = _case('barcoding_polya')
    lab = {'pass': 0, 'fail': 1, 'artifact': 2}
    counts = np.zeros((4, 5, 11), np.int64)
    for e in results:
        counts[lab.get(e.get('label'), 3), e.get('barcode', -1) + 1, STATUS_CODES[e['status']]] += 1
    # make every non-zero count distinct without changing which cells are populated
    nz = np.argwhere(counts > 0)
    for k, (a, b, c) in enumerate(nz):
        counts[a, b, c] = counts[a, b, c] * 1000 + k
    exploded = []
    names = ['pass', 'fail', 'artifact']
    from poreplex_b200.params import STATUS_NAMES
    for a, b, c in nz:
        e = {'status': STATUS_NAMES[c]}
        if a < 3:
            e['label'] = names[a]
        if b > 0:
            e['barcode'] = int(b - 1)
        exploded += [e] * int(counts[a, b, c])
    t1 = summary.FinalSummaryTracker(labels, barcodes)
    t1.feed_results(exploded)
    t2 = summary.FinalSummaryTracker(labels, barcodes)
    t2.feed_counts(counts)
    b1, b2 = io.StringIO(), io.StringIO()
    t1.print_results(b1)
    t2.print_results(b2)
    assert b1.getvalue() == b2.getvalue()

    # write_batch: arrays in, same text out
    labelled = [e for e in results if 'read_id' in e]
    out = {'status': np.array([STATUS_CODES[e['status']] for e in labelled], np.int32),
           'label': np.array([lab.get(e.get('label'), 3) for e in labelled], np.int32),
           'barcode': np.array([e.get('barcode', -1) for e in labelled], np.int32),
           'barcode_score': np.array([e.get('barcode_score', -1) for e in labelled], np.int32)}
    meta = [{k: e[k] for k in summary.SequencingSummaryWriter.SUMMARY_OUTPUT_FIELDS[:10]} for e in labelled]
    dwell = [e['polya']['dwell_time'] if 'polya' in e else None for e in labelled]
    d1, d2 = tmp_path / 'a', tmp_path / 'b'
    d1.mkdir(); d2.mkdir()
    w1 = summary.SequencingSummaryWriter(config, str(d1), labels, barcodes)
    w1.write_results(labelled); w1.close()
    w2 = summary.SequencingSummaryWriter(config, str(d2), labels, barcodes)
    w2.write_batch(meta, out, polya_dwell=dwell); w2.close()
    assert (d1 / 'sequencing_summary.txt').read_text() == (d2 / 'sequencing_summary.txt').read_text()


def _hist(results):
    lab = {'pass': 0, 'fail': 1, 'artifact': 2}
    counts = np.zeros((4, 5, 11), np.int64)
    for e in results:
        counts[lab.get(e.get('label'), 3), e.get('barcode', -1) + 1, STATUS_CODES[e['status']]] += 1
    return counts


def _gloo_table_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from poreplex_b200.sharding import shard_range, reduce_counts
    results, config, labels, barcodes = _case('barcoding_polya')
    lo, hi = shard_range(len(results), rank, world)
    total = reduce_counts(torch.from_numpy(_hist(results[lo:hi]))).numpy()
    t = summary.FinalSummaryTracker(labels, barcodes)
    t.feed_counts(total)
    buf = io.StringIO()
    t.print_results(buf)
    q.put((rank, buf.getvalue()))
    dist.destroy_process_group()


def test_final_table_from_all_reduced_histogram_gloo():
    """Multi-GPU aggregation as the product does it, on CPU: every rank counts
    its shard, ONE all-reduce of int64[4][5][11], and each rank can print the reference's
    final table from the reduced tensor -- equal to the table of the whole run."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_table_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    results, config, labels, barcodes = _case('barcoding_polya')
    t = summary.FinalSummaryTracker(labels, barcodes)
    t.feed_counts(_hist(results))
    buf = io.StringIO()
    t.print_results(buf)
    assert got[0] == got[1] == buf.getvalue()
    assert 'Successfully processed' in got[0] and 'BC4' in got[0]
