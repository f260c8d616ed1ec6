#!/usr/bin/env python3
"""Golden outputs of the WHOLE reference command line (SURVEY.md 8d, config 1 "one level up").

Run in the build container only (needs /root/reference):  python tests/golden/make_cli_golden.py

The golden read sets are written as real FAST5 files; the unmodified reference CLI
(commandline.__main__ -> ProcessingSession.run -> forked ProcessPoolExecutor workers ->
process_batch -> FASTQWriter / SequencingSummaryWriter) runs over them through
oracle/refcli.py with the REFERENCE's process_batch, once with --trim-adapter and once with
all four switches.  sequencing_summary.txt rows and FASTQ records go to
tests/golden/cli_<switches>.json (sorted: batches finish in any order)."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cli_util                                   # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix='cli_golden_')
    ind = os.path.join(tmp, 'in')
    cli_util.build_input_dir(ind)
    for key, switches in cli_util.SWITCH_SETS.items():
        out = os.path.join(tmp, 'out_' + key)
        cli_util.run_cli(ind, out, switches, 'reference')
        doc = cli_util.collect_outputs(out)
        doc['switches'] = switches
        with open(os.path.join(HERE, 'cli_%s.json' % key), 'w') as f:
            json.dump(doc, f, indent=0, sort_keys=True)
        print(key, len(doc['summary_rows']), 'summary rows;',
              {k: len(v) for k, v in doc['fastq'].items()})


if __name__ == '__main__':
    main()
