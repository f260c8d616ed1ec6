"""The whole-CLI drop-in check (SURVEY.md section 8d, config 1 "one level up"; pipeline.py:40,204).

tests/golden/cli_*.json hold what the UNMODIFIED reference command line writes
(sequencing_summary.txt, FASTQ per label / barcode) for the golden read sets with the
REFERENCE's own process_batch (tests/golden/make_cli_golden.py).

* here, where the reference tree exists but no GPU: the reference CLI is run again and must
  reproduce the committed files (so the fixtures are what the reference writes today);
* on the GPU box, where the reference tree does not exist: poreplex_b200's process_batch is
  driven the way pipeline.py drives it (batches of --batch-size reads) and its result dicts go
  through the mirrors of io.py's writers; the files must equal the committed ones;
* where both exist: the literal test -- the unmodified CLI with
  poreplex_b200.signal_analyzer.process_batch bound at pipeline.py:40.
"""
import json
import os

import pytest

import cli_util
from golden_util import GOLDEN_DIR

HAVE_REF = os.path.isdir('/root/reference/poreplex')


def _golden(key):
    with open(os.path.join(GOLDEN_DIR, 'cli_%s.json' % key)) as f:
        return json.load(f)


def _same(got, want):
    assert got['summary_header'] == want['summary_header']
    assert got['summary_rows'] == want['summary_rows']
    assert {k: v for k, v in got['fastq'].items() if v} == {k: v for k, v in want['fastq'].items() if v}
    assert sorted(got['fastq']) == sorted(want['fastq'])        # same output files, empty ones too


@pytest.mark.skipif(not HAVE_REF, reason='reference tree not present')
@pytest.mark.parametrize('key', sorted(cli_util.SWITCH_SETS))
def test_reference_cli_reproduces_committed_outputs(key, tmp_path):
    ind = str(tmp_path / 'in')
    cli_util.build_input_dir(ind)
    cli_util.run_cli(ind, str(tmp_path / 'out'), cli_util.SWITCH_SETS[key], 'reference')
    _same(cli_util.collect_outputs(str(tmp_path / 'out')), _golden(key))


def test_cli_goldens_cover_the_switches():
    doc = _golden('all')
    assert doc['summary_header'].split('\t')[-3:] == ['barcode', 'barcode_score', 'polya_dwell']
    bcs = {row.split('\t')[-3] for row in doc['summary_rows']}
    assert {'BC1', 'BC2', 'BC3', 'BC4', 'undetermined'} <= bcs
    labels = {row.split('\t')[11] for row in doc['summary_rows']}
    assert labels == {'pass', 'fail', 'artifact'}
    assert any(k.startswith('artifact/') and v for k, v in doc['fastq'].items())


@pytest.mark.gpu
@pytest.mark.parametrize('key', sorted(cli_util.SWITCH_SETS))
def test_b200_process_batch_writes_the_reference_cli_outputs(key, preset, tmp_path, monkeypatch):
    """process_batch driven like pipeline.py:204 drives it, outputs through the writer mirrors."""
    import sys
    from poreplex_b200 import signal_analyzer as sa
    monkeypatch.setitem(sys.modules, 'h5py', None)          # real files through hdf5_min / the native loader
    ind = str(tmp_path / 'in')
    listing = cli_util.build_input_dir(ind)
    got = cli_util.b200_outputs(ind, str(tmp_path / 'out'), listing, cli_util.SWITCH_SETS[key], preset,
                                sa.process_batch)
    _same(got, _golden(key))


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason='reference tree not present (the GPU box has none)')
@pytest.mark.parametrize('key', sorted(cli_util.SWITCH_SETS))
def test_unmodified_cli_with_b200_process_batch(key, tmp_path):
    """The literal check: commandline.__main__() -> pipeline.py with the replacement bound."""
    ind = str(tmp_path / 'in')
    cli_util.build_input_dir(ind)
    cli_util.run_cli(ind, str(tmp_path / 'out'), cli_util.SWITCH_SETS[key], 'b200', parallel=1)
    _same(cli_util.collect_outputs(str(tmp_path / 'out')), _golden(key))
