#!/usr/bin/env python3
"""Read-length sweep of BASELINE configs[4] (1k-16k samples) on one GPU: reads/s of the
device-resident path in the three LSTM modes, every integer output of the fast mode compared
with the exact-only kernels over ALL reads, and a sample of each length compared with the CPU
ORACLE (status, segments, barcode / best guess / phred).  One JSON document on stdout.

    python tools/sweep_lengths.py [--bytes 4e9] [--lengths 1000,2000,4000,8000,16000]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oracle import oracle as O
    from poreplex_b200 import params, synth
    from poreplex_b200.engine import SignalEngine
    from poreplex_b200.params import STATUS_NAMES
    ap = argparse.ArgumentParser()
    ap.add_argument('--bytes', type=float, default=4e9, help='raw bytes per length (> L2)')
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--lengths', default='1000,2000,4000,8000,16000')
    ap.add_argument('--oracle-reads', type=int, default=2048)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    base = params.load_preset()
    engines, oracles, doc = {}, {}, []
    keys = ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'label')
    for L in [int(x) for x in a.lengths.split(',')]:
        pname = 'bench-short' if L < 10500 else 'stock'
        preset = params.bench_short_preset(base) if pname == 'bench-short' else base
        if pname not in engines:
            engines[pname] = SignalEngine(dict(preset, barcoding=True), device=0)
            oracles[pname] = O.default_oracle(bench_short=(pname == 'bench-short'))
        eng, orc = engines[pname], oracles[pname]
        n = int(a.bytes / (2 * L)) // 128 * 128
        Lp = (L + 7) // 8 * 8
        rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=L, device=dev)
        raw = torch.zeros((n, Lp), dtype=torch.int16, device=dev)
        raw[:, :L] = rd['raw']
        work = (raw.reshape(-1), torch.arange(n, dtype=torch.int64, device=dev) * Lp,
                torch.full((n,), L, dtype=torch.int64, device=dev), rd['range'], rd['digitisation'],
                rd['offset'])
        res = eng.alloc_results(n)
        ent = {'read_length': L, 'preset': pname, 'reads': n, 'modes': {}}
        outs = {}
        for mode in ('fast', 'strict', 'exact'):
            eng.set_fast_lstm(mode)
            eng.analyze_device(*work, out=res, max_raw_length=L)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                eng.analyze_device(*work, out=res, max_raw_length=L)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            ent['modes'][mode] = {'ms_per_step': ms, 'reads_per_s': n / ms * 1e3,
                                  'raw_GB_per_s': n * L * 2 / ms / 1e6}
            if mode == 'fast':
                ent['exact_rerun_fraction'] = eng.recheck_stats()[0] / float(n)
            outs[mode] = {k: res[k].clone() for k in keys}
        eng.set_fast_lstm('fast')
        for mode in ('fast', 'strict'):
            ent['modes'][mode]['mismatches_vs_exact_kernels_all_reads'] = \
                {k: int((outs[mode][k] != outs['exact'][k]).sum().item()) for k in keys}
        st = outs['fast']['status'].cpu().numpy()
        ent['status_mix'] = {STATUS_NAMES[s]: int(c) for s, c in zip(*np.unique(st, return_counts=True))}
        sc = outs['fast']['barcode_score'].cpu().numpy()
        bc = outs['fast']['barcode'].cpu().numpy()
        ent['classified'] = int((sc >= 0).sum())
        ent['accepted_by_barcode'] = [int(((bc == k) & (sc >= 0)).sum()) for k in range(4)]
        # ---- a sample against the CPU oracle
        m = min(a.oracle_reads, n)
        hraw = raw[:m].cpu().numpy().reshape(-1)
        ref = orc.process_batch(hraw, np.arange(m, dtype=np.int64) * Lp, np.full(m, L, np.int64),
                                (rd['range'][:m] / rd['digitisation'][:m]).cpu().numpy(),
                                rd['offset'][:m].cpu().numpy())
        ok = np.isin(ref['status'], [0, 5])
        p = ref['pushed'] == 1
        f = {k: v[:m].cpu().numpy() for k, v in outs['fast'].items()}
        ent['oracle_sample'] = {
            'reads': m,
            'status_mismatches': int((f['status'] != ref['status']).sum()),
            'segment_mismatches': int((f['segments'][ok][:, :6] != ref['seg'][ok][:, :6]).any(axis=(1, 2)).sum()),
            'barcode_mismatches': int((f['barcode'][p] != ref['barcode'][p]).sum()),
            'guess_mismatches': int((f['barcode_guess'][p] != ref['guess'][p]).sum()),
            'phred_mismatches': int((f['barcode_score'][p] != ref['phred'][p]).sum()),
            'classified_in_sample': int(p.sum())}
        doc.append(ent)
        print(json.dumps(ent), file=sys.stderr)
        del raw, rd, work, res, outs
        torch.cuda.empty_cache()
    print(json.dumps(doc, indent=1))


if __name__ == '__main__':
    main()
