"""Does the backward layer-1 state converge bit for bit while it walks through the -1000 left
padding, and how many steps does that take?  (Exact kernels; GPU box.)"""
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
n = 4096
win = _windows(n, seed=17, min_len=20)
npad = (win == -1000.0).sum(1)
G = eng.debug_demux_l1(torch.from_numpy(win).to(dev)).cpu().numpy()      # [n][T][96]
hf, hb = G[:, :, :48], G[:, :, 48:]
T = win.shape[1]
# reference: the read with the longest pad; its hb at t = 0
ref = int(np.argmax(npad))
hstar = hb[ref, 0]
print('longest pad', npad[ref], 'hb[ref, 0..3] identical:', [bool(np.array_equal(hb[ref, t].view(np.uint32), hstar.view(np.uint32))) for t in range(4)])
# per read: number of pad steps until hb equals h* bitwise (k_r), and whether it stays there
ks, stays, never = [], 0, 0
for r in range(n):
    p = int(npad[r])
    if p < 2:
        continue
    eq = np.all(hb[r, :p].view(np.uint32) == hstar.view(np.uint32)[None, :], axis=1)   # t = 0..p-1
    # walking backwards in t: first pad step is t = p-1
    if not eq.any():
        never += 1
        continue
    t_first = int(np.max(np.nonzero(eq)[0]))           # largest t that equals h*
    k = p - 1 - t_first                                # pad steps needed
    ks.append(k)
    if eq[:t_first + 1].all():
        stays += 1
ks = np.array(ks)
print('reads with pad >= 2: %d; never reach h*: %d; reach and stay: %d' % (len(ks) + never, never, stays))
if len(ks):
    print('pad steps to reach h* bitwise: quantiles (50,90,99,max):', np.quantile(ks, [0.5, 0.9, 0.99, 1.0]))
# never-converged reads: how far are they
if never:
    for r in range(n):
        p = int(npad[r])
        if p >= 60:
            d = np.abs(hb[r, 0] - hstar).max()
            if d > 0:
                print('row %d pad %d |hb(0) - h*| = %.3e; distinct values at t=0..5: %s' % (
                    r, p, d, [float(np.abs(hb[r, t] - hstar).max()) for t in range(6)]))
                break
# forward pad states: hf(t) for t < pad equals the same table for every read
eqf = all(np.array_equal(hf[r, :min(int(npad[r]), int(npad[ref]))].view(np.uint32),
                         hf[ref, :min(int(npad[r]), int(npad[ref]))].view(np.uint32)) for r in range(0, n, 7))
print('forward pad states universal:', eqf)
# does the forward pad state itself converge?
dfw = [float(np.abs(hf[ref, t] - hf[ref, t - 1]).max()) for t in (5, 20, 50, 100, 150, 200, int(npad[ref]) - 1)]
print('forward |h(t) - h(t-1)| at t=5,20,50,100,150,200,last:', dfw)
