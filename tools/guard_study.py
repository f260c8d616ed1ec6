#!/usr/bin/env python3
"""How good is the classifier guard on BARCODED windows?  

The tensor-core classifier's call is kept only if it survives the error bound
delta = delta0 + gain * s (s = logit shift under the two coarse probes).  Round 1 validated
that bound on windows that were all decoys.  This tool measures it on the population the
guard exists for: synthetic reads carrying the four class prototypes at full and reduced
strength (scores from 0.3 to 0.999, all calibration bins, both sides of the acceptance
threshold).  For every window: the true error of the tensor-core logits against the exact
kernels, the bound the guard assumed, and whether a call that differs slipped through.

    python tools/guard_study.py [--reads 400000] [--length 4000] [--batches 2]  > guard.json
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
    ap.add_argument('--reads', type=int, default=400000)
    ap.add_argument('--length', type=int, default=4000)
    ap.add_argument('--batches', type=int, default=2)
    ap.add_argument('--seed0', type=int, default=4100)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    base = params.load_preset()
    preset = params.bench_short_preset(base) if a.length < 10500 else base
    eng = SignalEngine(dict(preset, barcoding=True), device=0)
    delta0, gain = 1e-3, 0.1                   # pb_internal.h defaults (demux_margin_delta / probe_gain)
    thr = float(eng.demux_model.calibration[18])
    n, L = a.reads, a.length
    Lp = (L + 7) // 8 * 8
    doc = {'read_length': L, 'reads_per_batch': n, 'delta0': delta0, 'probe_gain': gain, 'strata': {},
           'batches': []}
    acc = {}

    def add(name, mask, ratio, err, bad, unsafe):
        s = acc.setdefault(name, {'windows': 0, 'guard_passing': 0, 'max_err_over_bound_passing': 0.0,
                                  'max_err': 0.0, 'calls_differ_before_recheck': 0,
                                  'calls_differ_and_passed_guard': 0, 'err_over_bound_q': []})
        if not mask.any():
            return
        ok = mask & ~unsafe
        s['windows'] += int(mask.sum())
        s['guard_passing'] += int(ok.sum())
        if ok.any():
            s['max_err_over_bound_passing'] = max(s['max_err_over_bound_passing'], float(ratio[ok].max()))
            s['err_over_bound_q'].append(np.quantile(ratio[ok], [0.5, 0.99, 0.9999]).tolist())
        s['max_err'] = max(s['max_err'], float(err[mask].max()))
        s['calls_differ_before_recheck'] += int((bad & mask).sum())
        s['calls_differ_and_passed_guard'] += int((bad & ok).sum())

    for b in range(a.batches):
        rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=a.seed0 + b, device=dev)
        raw = torch.zeros((n, Lp), dtype=torch.int16, device=dev)
        raw[:, :L] = rd['raw']
        w = (raw.reshape(-1), torch.arange(n, dtype=torch.int64, device=dev) * Lp,
             torch.full((n,), L, dtype=torch.int64, device=dev), rd['range'], rd['digitisation'], rd['offset'])
        pooled = eng.pool_signal(*w, max_raw_length=L)
        eng.set_fast_lstm('exact')
        status, ss, _ = eng.fit_scalers(*w, pooled)
        seg, _ = eng.detect_segments(*w, pooled, ss, status, max_raw_length=L)
        win, pushed = eng.barcode_windows(*w, pooled, ss, status, seg)
        keep = pushed.bool()
        wd = win[keep].contiguous()
        planted = rd['planted']['barcode'][keep].cpu().numpy()
        m = wd.shape[0]
        p_ex, bc_ex, g_ex, s_ex = [o.cpu().numpy() for o in eng.demux_predict(wd)]
        eng.set_fast_lstm('fast')
        p_tc, lg_tc, bc_tc, g_tc, s_tc, unsafe, sens = [o.cpu().numpy() for o in eng.demux_predict_tc(wd)]
        pe, pt = p_ex[:, :5].astype(np.float64), p_tc[:, :5].astype(np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            le, lt = np.log(pe), np.log(pt)
        am = pe.argmax(1)
        d = np.abs((le - le[np.arange(m), am][:, None]) - (lt - lt[np.arange(m), am][:, None]))
        d[~(pe > 1e-30)] = 0
        err = d.max(1)
        bound = delta0 + gain * sens
        ratio = err / bound
        unsafe = unsafe != 0
        bad = (bc_tc != bc_ex) | (g_tc != g_ex) | (s_tc != s_ex)
        score = pe.max(1)
        add('all', np.ones(m, bool), ratio, err, bad, unsafe)
        add('accepted_barcode', bc_ex >= 0, ratio, err, bad, unsafe)
        add('barcode_guess_below_threshold', (g_ex >= 0) & (bc_ex < 0), ratio, err, bad, unsafe)
        add('within_0.01_of_threshold', np.abs(score - thr) < 0.01, ratio, err, bad, unsafe)
        add('decoy_call', g_ex < 0, ratio, err, bad, unsafe)
        for k in range(4):
            add('accepted_BC%d' % (k + 1), bc_ex == k, ratio, err, bad, unsafe)
        # what other (delta0, gain) pairs would have done on these windows: largest error / bound,
        # windows whose error exceeds the bound, and windows whose call would be flagged
        lg = lg_tc[:, :5].astype(np.float64)
        top2 = np.sort(lg, axis=1)
        gap = top2[:, -1] - top2[:, -2]
        pbest = pt.max(1)
        edges = np.concatenate([[thr], np.asarray(eng.demux_model.calibration, np.float64)])
        edge_dist = np.abs(pbest[:, None] - edges[None, :]).min(1)
        sweep = []
        for d0 in (1e-3, 5e-4, 2e-4, 1e-4):
            for gn in (0.1, 0.05):
                bnd = d0 + gn * sens
                flagged = (gap <= bnd) | (edge_dist <= bnd * pbest * (1 - pbest) + 1e-6)
                sweep.append({'delta0': d0, 'gain': gn, 'max_err_over_bound': float((err / bnd).max()),
                              'err_exceeds_bound': int((err > bnd).sum()),
                              'flagged_fraction': float(flagged.mean()),
                              'differing_calls_not_flagged': int((bad & ~flagged).sum())})
        ent = {'seed': a.seed0 + b, 'windows': int(m), 'unsafe_fraction': float(unsafe.mean()),
               'bound_sweep': sweep,
               'err_quantiles_50_99_9999_max': np.quantile(err, [0.5, 0.99, 0.9999, 1.0]).tolist(),
               'sens_quantiles_50_99_9999_max': np.quantile(sens, [0.5, 0.99, 0.9999, 1.0]).tolist(),
               'accepted_by_class': [int((bc_ex == k).sum()) for k in range(4)],
               'planted_by_class': [int((planted == k + 1).sum()) for k in range(4)],
               'phred_hist': np.bincount(s_ex[s_ex >= 0], minlength=30).tolist(),
               'timeouts': eng.recheck_stats()[1]}
        doc['batches'].append(ent)
        print(json.dumps(ent), file=sys.stderr)
        del raw, rd, w, pooled, win, wd
    doc['strata'] = acc
    print(json.dumps(doc, indent=1))


if __name__ == '__main__':
    main()
