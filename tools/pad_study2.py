"""How many -1000 pad steps until the backward layer-1 state of the exact kernels stops
changing bit for bit (per read), and does it ever change again?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from poreplex_b200 import params
from poreplex_b200.engine import SignalEngine
from test_gpu_tc import _windows

preset = params.load_preset()
cfg = dict(preset); cfg['barcoding'] = True
eng = SignalEngine(cfg, device=0)
dev = torch.device('cuda', 0)
ks, bad = [], 0
dev_after = []
K = 32
for seed in (17, 18, 19, 20):
    n = 4096
    win = _windows(n, seed=seed, min_len=20)
    npad = (win == -1000.0).sum(1)
    G = eng.debug_demux_l1(torch.from_numpy(win).to(dev)).cpu().numpy()
    hb = G[:, :, 48:].view(np.uint32)
    for r in range(n):
        p = int(npad[r])
        if p < 3:
            continue
        # backward walks t = p-1, p-2, ..., 0 through the pad; frozen from step k on if
        # hb[t] == hb[t-1] for all t <= p-1-k
        same = np.all(hb[r, 1:p] == hb[r, 0:p - 1], axis=1)      # same[t-1]: hb[t] == hb[t-1]
        changed = np.nonzero(~same)[0]
        k = 0 if len(changed) == 0 else (p - 1) - int(changed.min())   # pad steps until frozen for good
        ks.append(k)
        if len(changed) and changed.min() < p - 1 - 40:
            bad += 1
        if p > K + 2:
            f = G[r, :, 48:]
            dev_after.append(float(np.abs(f[0:p - K] - f[p - 1 - K][None, :]).max()))
ks = np.array(ks)
print('reads', len(ks), 'pad steps until the backward state is frozen for good: quantiles (50, 99, 99.9, max):',
      np.quantile(ks, [0.5, 0.99, 0.999, 1.0]), 'still changing after 40 pad steps:', bad)
dev_after = np.array(dev_after)
print('max |hb(t) - hb(pad-1-%d)| over later pad steps: quantiles (50, 99, 99.9, max):' % K, np.quantile(dev_after, [0.5, 0.99, 0.999, 1.0]), 'reads', len(dev_after), 'nonzero', int((dev_after > 0).sum()))
