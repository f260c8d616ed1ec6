"""Shared by tests/golden/make_cli_golden.py and tests/test_cli_dropin.py: the input directory
of the whole-CLI drop-in check (real FAST5 files written from the golden read sets) and the
normal form CLI outputs are compared in."""
import gzip
import os
import subprocess
import sys

import numpy as np

from golden_util import load_golden, golden_reads, golden_basecalls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI_SETS = ('stock16k', 'chimera40k')            # both under the stock preset
SWITCH_SETS = {
    'trim': ['--trim-adapter'],
    'all': ['--trim-adapter', '--barcoding', '--filter-chimera', '--polya'],
}
BATCH_SIZE = 16


def build_input_dir(path):
    """One multi-read FAST5 per golden set, gzip + shuffle chunked Signal / Move datasets.
    Returns {filename: [read_id, ...]}."""
    from oracle import fake_fast5, refshim
    from fast5_files import write_fast5
    os.makedirs(path, exist_ok=True)
    listing = {}
    for name in CLI_SETS:
        z, _ = load_golden(name)
        ids = [str(s) for s in z['read_ids']]
        rd, bcs = golden_reads(z), golden_basecalls(z)
        tree = refshim.FakeFile()
        for i, rid in enumerate(ids):
            raw = rd['raw'][i][:int(rd['length'][i])]
            fake_fast5.add_read(tree, rid, raw, rd['digitisation'][i], rd['range'][i], rd['offset'][i],
                                rd['sampling_rate'][i], channel=str(1 + i % 512), start_time=1000 * i,
                                basecall=bcs[i])
        fn = name + '.fast5'
        write_fast5(os.path.join(path, fn), tree, signal_kw=dict(chunks=4096, gzip=1, shuffle=True),
                    move_kw=dict(chunks=512, gzip=1))
        listing[fn] = ids
    return listing


def run_cli(inputdir, outputdir, switches, process_batch='reference', parallel=2):
    """The reference's command line, unmodified, in a fresh interpreter (oracle/refcli.py)."""
    cmd = [sys.executable, '-m', 'oracle.refcli', '--process-batch', process_batch, '--',
           '-i', inputdir, '-o', outputdir, '-q', '-y', '-p', str(parallel),
           '--batch-size', str(BATCH_SIZE)] + list(switches)
    subprocess.run(cmd, cwd=ROOT, check=True, timeout=1200)


def collect_outputs(outputdir):
    """{'summary_header': str, 'summary_rows': sorted rows, 'fastq': {relative path: sorted
    records}} -- batches finish in any order, so rows and records are compared as sorted lists."""
    with open(os.path.join(outputdir, 'sequencing_summary.txt')) as f:
        lines = f.read().splitlines()
    doc = {'summary_header': lines[0], 'summary_rows': sorted(lines[1:]), 'fastq': {}}
    fq = os.path.join(outputdir, 'fastq')
    for dirpath, _, files in os.walk(fq):
        for fn in files:
            with gzip.open(os.path.join(dirpath, fn), 'rt') as f:
                txt = f.read().splitlines()
            recs = sorted('\n'.join(txt[i:i + 4]) for i in range(0, len(txt), 4))
            doc['fastq'][os.path.relpath(os.path.join(dirpath, fn), fq)] = recs
    return doc


def b200_outputs(inputdir, outputdir, listing, switches, preset, process_batch):
    """What pipeline.py does with the result dicts, without pipeline.py (the GPU box has no
    reference tree): process_batch over --batch-size chunks of each file's reads, results into
    the mirrors of io.py's SequencingSummaryWriter / FASTQWriter (poreplex_b200/summary.py)."""
    from poreplex_b200 import summary
    cfg = dict(preset)
    cfg.update({'inputdir': inputdir, 'outputdir': outputdir,
                'barcoding': '--barcoding' in switches, 'measure_polya': '--polya' in switches,
                'trim_adapter': '--trim-adapter' in switches,
                'filter_unsplit_reads': '--filter-chimera' in switches,
                'minimum_sequence_length': 10, 'dump_adapter_signals': False,
                'dump_basecalls': False, 'albacore_onthefly': False, 'barcoding_quality_filter': 18,
                'fast5_output': False})
    labels, barcodes, layout = summary.output_name_mapping(cfg)
    os.makedirs(outputdir, exist_ok=True)
    sw = summary.SequencingSummaryWriter(cfg, outputdir, labels, barcodes)
    fw = summary.FASTQWriter(outputdir, layout)
    batchid = 0
    for fn, ids in listing.items():
        for i in range(0, len(ids), BATCH_SIZE):
            res = process_batch(batchid, [(fn, rid) for rid in ids[i:i + BATCH_SIZE]], cfg)
            assert not isinstance(res, tuple), res
            sw.write_results(res)
            fw.write_sequences([r for r in res if 'label' in r])
            batchid += 1
    sw.close()
    fw.close()
    return collect_outputs(outputdir)
