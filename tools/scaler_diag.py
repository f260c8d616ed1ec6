"""Error of the tensor-core scaler against the exact kernels on the bench workload
(GPU box): python tools/scaler_diag.py [reads] [length]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poreplex_b200 import params, synth
from poreplex_b200.engine import SignalEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
dev = torch.device('cuda', 0)
preset = params.bench_short_preset(params.load_preset()) if L < 9000 else params.load_preset()
cfg = dict(preset); cfg['barcoding'] = True
eng = SignalEngine(cfg, device=0)
rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=99, device=dev)
Lp = (L + 7) // 8 * 8
raw = torch.zeros((n, Lp), dtype=torch.int16, device=dev); raw[:, :L] = rd['raw']
off = torch.arange(n, dtype=torch.int64, device=dev) * Lp
ln = torch.full((n,), L, dtype=torch.int64, device=dev)
args = (raw.reshape(-1), off, ln, rd['range'], rd['digitisation'], rd['offset'])
# huge margin => nothing about the scaler is re-run because of its own uncertainty box
fast = {k: v.clone() for k, v in eng.analyze_device(*args, barcoding=False, max_raw_length=L).items()}
torch.cuda.synchronize()
print('reruns (barcoding off):', eng.recheck_stats())
eng.set_fast_lstm(False)
exact = {k: v.clone() for k, v in eng.analyze_device(*args, barcoding=False, max_raw_length=L).items()}
torch.cuda.synchronize()
eng.set_fast_lstm(True)
st = exact['status'].cpu().numpy()
d = (fast['scale_shift'].double() - exact['scale_shift'].double()).abs().cpu().numpy()
ez = np.maximum(d[:, 0] / 0.13295630234669656, d[:, 1] / 9.82564593783874)
for name, m in (('okay', st == 0), ('qc_fail', st == 4), ('no adapter', st == 5)):
    e = ez[m & (ez > 0)]
    if len(e):
        print('%-10s n=%d  z-error quantiles (50, 99, 99.9, 99.99, max): %s' % (
            name, m.sum(), np.quantile(e, [0.5, 0.99, 0.999, 0.9999, 1.0])))
for k in ('status', 'segments'):
    print(k, 'mismatches:', int((fast[k] != exact[k]).sum().item()))
i = int(np.argmax(np.where(st == 0, ez, 0)))
print('worst okay read', i, 'z err', ez[i], 'scale/shift exact', exact['scale_shift'][i].cpu().numpy(), 'fast', fast['scale_shift'][i].cpu().numpy())
