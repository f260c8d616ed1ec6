#!/usr/bin/env python3
"""Read-length sweep of the device-resident path on one GPU (BASELINE config 5's sweep) and
the cost of adding poly(A) (config 4 without the chimera filter, which needs basecall
tables).  Writes one JSON document to stdout.

    python tools/sweep.py [--bytes 2e9]
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
    from poreplex_b200.params import STATUS_NAMES
    ap = argparse.ArgumentParser()
    ap.add_argument('--bytes', type=float, default=2e9, help='raw bytes per configuration')
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--lengths', default='1000,2000,4000,8000,16000,32000')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    base = params.load_preset()
    out = []
    engines = {}
    for L in [int(x) for x in args.lengths.split(',')]:
        for preset_name in (['bench-short'] if L < 10500 else ['stock']):
            preset = params.bench_short_preset(base) if preset_name == 'bench-short' else base
            if preset_name not in engines:
                engines[preset_name] = SignalEngine(dict(preset, barcoding=True), device=0)
            eng = engines[preset_name]
            n = int(args.bytes / (2 * L)) // 64 * 64
            Lp = (L + 7) // 8 * 8
            rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=L, device=dev)
            raw = torch.zeros((n, Lp), dtype=torch.int16, device=dev)
            raw[:, :L] = rd['raw']
            work = (raw.reshape(-1), torch.arange(n, dtype=torch.int64, device=dev) * Lp,
                    torch.full((n,), L, dtype=torch.int64, device=dev), rd['range'],
                    rd['digitisation'], rd['offset'])
            for polya in (False, True):
                res = eng.alloc_results(n, polya=polya)
                for _ in range(2):
                    eng.analyze_device(*work, out=res, max_raw_length=L, polya=polya)
                torch.cuda.synchronize()
                eng.profile_enable(True); eng.profile_read()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    eng.analyze_device(*work, out=res, max_raw_length=L, polya=polya)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                prof = eng.profile_read(); eng.profile_enable(False)
                st = res['status'].cpu().numpy()
                mix = {STATUS_NAMES[s]: int(c) for s, c in zip(*np.unique(st, return_counts=True))}
                ent = {'read_length': L, 'preset': preset_name, 'reads': n, 'polya': polya,
                       'ms_per_step': ms, 'reads_per_s': n / ms * 1e3,
                       'raw_GBps': n * L * 2 / ms / 1e6,
                       'classified': int((res['barcode_score'] >= 0).sum().item()),
                       'status_mix': mix,
                       'kernels_ms': {k: v[0] / args.steps for k, v in prof.items()}}
                if not polya:
                    # default path = tensor-core LSTMs + exact re-runs; compare with the exact-only
                    # kernels: time and every integer output
                    ent['exact_reruns'] = eng.recheck_stats()[0]
                    fast = {k: res[k].clone() for k in ('status', 'segments', 'barcode',
                                                        'barcode_guess', 'barcode_score', 'label')}
                    eng.set_fast_lstm(False)
                    eng.analyze_device(*work, out=res, max_raw_length=L)
                    torch.cuda.synchronize()
                    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    x0.record()
                    eng.analyze_device(*work, out=res, max_raw_length=L)
                    x1.record()
                    torch.cuda.synchronize()
                    eng.set_fast_lstm(True)
                    ent['exact_only_ms_per_step'] = x0.elapsed_time(x1)
                    ent['mismatches_vs_exact_only'] = {k: int((fast[k] != res[k]).sum().item()) for k in fast}
                if polya:
                    pol = res['polya'].cpu().numpy().view(np.dtype([('found', 'i4'), ('rest', 'V804')]))
                    ent['polya_found'] = int(pol['found'].sum())
                if polya and eng.unsplit_ready:
                    # chimera filter on synthetic guppy-style tables: one row per 15 samples,
                    # random moves / qualities, means derived on the device
                    E = L // 15
                    g = torch.Generator(device=dev); g.manual_seed(L)
                    ev_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * E
                    start = (torch.arange(E, dtype=torch.int64, device=dev) * 15).repeat(n)
                    move = (torch.rand(n * E, generator=g, device=dev) < 0.3).to(torch.int32)
                    pst = torch.rand(n * E, generator=g, device=dev, dtype=torch.float64)
                    rate = torch.full((n,), 3012.0, dtype=torch.float64, device=dev)
                    first = torch.zeros(n, dtype=torch.int64, device=dev)
                    maxw = max(1, -(-L // int(3 * 3012)))
                    for rep in range(2):
                        torch.cuda.synchronize()
                        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        u0.record()
                        flag = eng.detect_unsplit_device(work, ev_off, start, move, pst, rate, first, 15,
                                                         res['scale_shift'], res['status'],
                                                         res['segments'], maxw)
                        u1.record()
                        torch.cuda.synchronize()
                    ent['unsplit_ms'] = u0.elapsed_time(u1)
                    ent['unsplit_flagged'] = int((flag == 1).sum().item())
                    ent['unsplit_errors'] = int((flag < 0).sum().item())
                out.append(ent)
                print(json.dumps(ent), file=sys.stderr)
            del raw, rd, work
            torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
