"""Run the REFERENCE's command line unmodified (TEST INFRASTRUCTURE).

``python -m oracle.refcli [--process-batch reference|b200] -- <poreplex arguments>``

executes ``poreplex.commandline.__main__()`` from /root/reference -- argument parsing,
``ProcessingSession.run`` (pipeline.py), the forked ``ProcessPoolExecutor`` workers, the
FASTQ and sequencing-summary writers (io.py) -- over the shims of oracle/refshim.py
(SURVEY.md section 8c item 6).  With ``--process-batch b200`` the one name the drop-in
replaces, ``process_batch`` as imported at pipeline.py:40 and called at pipeline.py:204,
is bound to ``poreplex_b200.signal_analyzer.process_batch`` first; nothing else changes.
This is the literal "drops into pipeline.py unchanged" check; it needs both the reference
tree and (for ``b200``) a CUDA device.

Extra stubs beyond refshim.install(): ``asyncio.Task.all_tasks`` (removed in Python 3.9,
used at pipeline.py:179,558), ``pysam.BGZFile`` -> gzip, an empty ``mappy``.
"""
import argparse
import asyncio
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _patch_asyncio():
    if hasattr(asyncio.Task, 'all_tasks'):
        return

    class _Task(asyncio.Task):
        @staticmethod
        def all_tasks(loop=None):
            try:
                return asyncio.all_tasks(loop)
            except RuntimeError:
                return set()
    asyncio.Task = _Task


def run(argv, process_batch='reference'):
    from oracle import refshim
    refshim.install()
    _patch_asyncio()
    from poreplex import commandline, pipeline
    if process_batch == 'b200':
        from poreplex_b200 import signal_analyzer as b200
        pipeline.process_batch = b200.process_batch          # pipeline.py:40 binding
    elif process_batch != 'reference':
        raise ValueError(process_batch)
    old = sys.argv
    sys.argv = ['poreplex'] + list(argv)
    try:
        commandline.__main__()
    finally:
        sys.argv = old


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--process-batch', default='reference', choices=['reference', 'b200'])
    ap.add_argument('rest', nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest[:1] == ['--'] else a.rest
    run(rest, a.process_batch)


if __name__ == '__main__':
    main()
