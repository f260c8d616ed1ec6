"""Diagnostics for the tensor-core demultiplexer: error distribution against the exact
kernels, margin-test statistics, kernel timings.  Run on a GPU box:
    python tools/tc_diag.py [n_windows]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from poreplex_b200 import params                      # noqa: E402
from poreplex_b200.engine import SignalEngine         # noqa: E402
from test_gpu_tc import _windows                      # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    preset = params.load_preset()
    cfg = dict(preset)
    cfg['barcoding'] = True
    eng = SignalEngine(cfg, device=0)
    dev = torch.device('cuda', 0)
    win = _windows(n, seed=3)
    win[0, :] = 0.0
    win[1, :] = -1000.0
    wd = torch.from_numpy(win).to(dev)
    eng.set_fast_lstm(False)
    p_ex, bc_ex, g_ex, s_ex = [o.cpu().numpy() for o in eng.demux_predict(wd)]
    eng.set_fast_lstm(True)
    p_tc, lg_tc, bc_tc, g_tc, s_tc, unsafe, sens = [o.cpu().numpy() for o in eng.demux_predict_tc(wd)]
    print('timeouts', eng.recheck_stats())
    pe, pt = p_ex[:, :5].astype(np.float64), p_tc[:, :5].astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        le, lt = np.log(pe), np.log(pt)
    # logit differences up to a per-row shift: use the exact argmax class as the anchor
    am = pe.argmax(1)
    de = le - le[np.arange(n), am][:, None]
    dt = lt - lt[np.arange(n), am][:, None]
    d = np.abs(de - dt)
    d[~(pe > 1e-30)] = 0
    rowmax = d.max(1)
    print('rel-logit error quantiles (50,90,99,99.9,max):',
          np.quantile(rowmax, [0.5, 0.9, 0.99, 0.999, 1.0]))
    dp = np.abs(pe - pt).max(1)
    print('|dp| quantiles (50,90,99,99.9,max):', np.quantile(dp, [0.5, 0.9, 0.99, 0.999, 1.0]))
    npad = (win == -1000.0).sum(1)
    for i in np.argsort(-rowmax)[:8]:
        print('row %d pad %d rowmax %.3e dp %.3e\n   p_ex %s\n   p_tc %s\n   logit_tc %s' % (
            i, npad[i], rowmax[i], dp[i], pe[i], pt[i], lg_tc[i, :5]))
    # per-window sensitivity probe vs the true error
    print('probe shift quantiles (50,90,99,99.9,max):', np.quantile(sens, [0.5, 0.9, 0.99, 0.999, 1.0]))
    ratio = rowmax / np.maximum(sens, 1e-12)
    print('err / shift quantiles (50,90,99,99.9,max):', np.quantile(ratio, [0.5, 0.9, 0.99, 0.999, 1.0]))
    for d0 in (5e-4, 1e-3, 2e-3):
        for gain in (0.03, 0.05, 0.08, 0.1, 0.25):
            bound = d0 + gain * sens
            viol = rowmax > bound
            print('  delta0 %.0e gain %.2f: max err/bound %.3f, windows with err > bound: %d, > bound/4: %d'
                  % (d0, gain, (rowmax / bound).max(), viol.sum(), (rowmax > bound / 4).sum()))
    i = int(np.argmax(rowmax / (2e-3 + 0.08 * sens)))
    print('  worst window for (1e-3, 0.1): row %d err %.3e shift %.3e pad %d' % (i, rowmax[i], sens[i], npad[i]))
    # error vs smallest prob involved
    for thr in (1e-2, 1e-4, 1e-6, 1e-10):
        m = pe > thr
        dd = np.where(m, d, 0).max()
        print('max rel-logit err over classes with p > %g: %.3e' % (thr, dd))
    safe = unsafe == 0
    print('unsafe %.2f%%; safe calls identical: bc %s guess %s score %s' % (
        100 * unsafe.mean(), np.array_equal(bc_tc[safe], bc_ex[safe]),
        np.array_equal(g_tc[safe], g_ex[safe]), np.array_equal(s_tc[safe], s_ex[safe])))
    bad = (bc_tc != bc_ex) | (g_tc != g_ex) | (s_tc != s_ex)
    print('calls that differ before re-check: %d (of which flagged unsafe: %d)' % (
        bad.sum(), (bad & ~safe).sum()))
    # timings
    for mode in ('exact', 'tc+recheck'):
        eng.set_fast_lstm(mode != 'exact')
        eng.demux_predict(wd)
        torch.cuda.synchronize()
        eng.profile_enable(True)
        eng.profile_read()
        t0 = time.perf_counter()
        eng.demux_predict(wd)
        torch.cuda.synchronize()
        dt_ = time.perf_counter() - t0
        prof = eng.profile_read()
        eng.profile_enable(False)
        print('%s: %.2f ms for %d windows (%.0f windows/s)' % (mode, dt_ * 1e3, n, n / dt_))
        for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            print('   %-24s %8.3f ms  %d launches' % (k, ms, c))
        if mode != 'exact':
            print('   rechecked, timeouts:', eng.recheck_stats())
    eng.set_fast_lstm(True)


if __name__ == '__main__':
    main()
