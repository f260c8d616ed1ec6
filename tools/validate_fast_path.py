#!/usr/bin/env python3
"""Large-scale check of the default (tensor-core + guards + exact re-run) path against the
exact-only kernels: K independently seeded batches of synthetic reads, every integer output of
every read compared.  Writes one JSON document to stdout.

    python tools/validate_fast_path.py [--batches 10] [--reads 1000000] [--length 4000]
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
    from poreplex_b200 import params, synth
    from poreplex_b200.engine import SignalEngine
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, default=10)
    ap.add_argument('--reads', type=int, default=1000000)
    ap.add_argument('--length', type=int, default=4000)
    ap.add_argument('--seed0', type=int, default=777000)
    ap.add_argument('--mode', default='fast', choices=['fast', 'strict'])
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    base = params.load_preset()
    preset = params.bench_short_preset(base) if args.length < 10500 else base
    eng = SignalEngine(dict(preset, barcoding=True), device=0)
    eng.set_fast_lstm(args.mode)
    n, L = args.reads, args.length
    Lp = (L + 7) // 8 * 8
    keys = ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'label')
    total = {k: 0 for k in keys}
    doc = {'reads_per_batch': n, 'read_length': L, 'batches': [], 'preset': 'bench-short' if L < 10500 else 'stock'}
    out = eng.alloc_results(n)
    for b in range(args.batches):
        rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=args.seed0 + b, device=dev)
        raw = torch.zeros((n, Lp), dtype=torch.int16, device=dev)
        raw[:, :L] = rd['raw']
        work = (raw.reshape(-1), torch.arange(n, dtype=torch.int64, device=dev) * Lp,
                torch.full((n,), L, dtype=torch.int64, device=dev), rd['range'], rd['digitisation'], rd['offset'])
        eng.analyze_device(*work, out=out, barcoding=True, max_raw_length=L)
        torch.cuda.synchronize()
        rerun, timeouts = eng.recheck_stats()
        causes = eng.rerun_causes()
        fast = {k: out[k].clone() for k in keys}
        fast_ss = out['scale_shift'].clone()
        eng.set_fast_lstm(False)
        eng.analyze_device(*work, out=out, barcoding=True, max_raw_length=L)
        torch.cuda.synchronize()
        eng.set_fast_lstm(args.mode)
        mism = {k: int((fast[k] != out[k]).sum().item()) for k in keys}
        okay = out['status'] == 0
        d = (fast_ss.double() - out['scale_shift'].double()).abs()[okay]
        bc = out['barcode'][out['barcode_score'] >= 0]
        ex = out['scale_shift'].double()[okay]
        rel = (d[:, 0] * 100.0 + d[:, 1]) / (ex[:, 0] * 100.0 + ex[:, 1]).abs()
        ent = {'seed': args.seed0 + b, 'exact_reruns': rerun, 'causes': causes, 'tc_timeouts': timeouts,
               'accepted_by_barcode': [int((bc == k).sum().item()) for k in range(4)],
               'undetermined': int((bc < 0).sum().item()),
               'max_rel_error_normalised_signal_at_100pA': float(rel.max().item()) if rel.numel() else 0.0,
               'reads_over_1e-5_rel': int((rel > 1e-5).sum().item()),
               'mismatches': mism, 'classified': int((out['barcode_score'] >= 0).sum().item()),
               'max_scaler_z0_error_okay_reads': float((d[:, 0] / 0.13295630234669656).max().item()),
               'max_scaler_z1_error_okay_reads': float((d[:, 1] / 9.82564593783874).max().item())}
        doc['batches'].append(ent)
        for k in keys:
            total[k] += mism[k]
        print(json.dumps(ent), file=sys.stderr)
        del raw, rd, work
    doc['total_reads'] = n * args.batches
    doc['mode'] = args.mode
    doc['accepted_by_barcode'] = [sum(e['accepted_by_barcode'][k] for e in doc['batches']) for k in range(4)]
    doc['exact_rerun_fraction'] = sum(e['exact_reruns'] for e in doc['batches']) / float(n * args.batches)
    doc['max_rel_error_normalised_signal_at_100pA'] = max(e['max_rel_error_normalised_signal_at_100pA'] for e in doc['batches'])
    doc['reads_over_1e-5_rel'] = sum(e['reads_over_1e-5_rel'] for e in doc['batches'])
    doc['total_mismatches'] = total
    doc['max_scaler_z0_error'] = max(e['max_scaler_z0_error_okay_reads'] for e in doc['batches'])
    doc['max_scaler_z1_error'] = max(e['max_scaler_z1_error_okay_reads'] for e in doc['batches'])
    doc['scaler_margins_z0_z1'] = [2.5e-4, 1.5e-3]
    print(json.dumps(doc, indent=1))


if __name__ == '__main__':
    main()
