"""Reproduce one batch of tools/validate_fast_path.py and inspect the reads whose integer
outputs differ between the default path and the exact-only kernels."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poreplex_b200 import params, synth
from poreplex_b200.engine import SignalEngine

seed = int(sys.argv[1]); n = 1000000; L = 4000
dev = torch.device('cuda', 0)
preset = params.bench_short_preset(params.load_preset())
eng = SignalEngine(dict(preset, barcoding=True), device=0)
rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=seed, device=dev)
raw = rd['raw'].contiguous()
work = (raw.reshape(-1), torch.arange(n, dtype=torch.int64, device=dev) * L,
        torch.full((n,), L, dtype=torch.int64, device=dev), rd['range'], rd['digitisation'], rd['offset'])
fast = {k: v.clone() for k, v in eng.analyze_device(*work, barcoding=True, max_raw_length=L).items()}
torch.cuda.synchronize()
eng.set_fast_lstm(False)
exact = {k: v.clone() for k, v in eng.analyze_device(*work, barcoding=True, max_raw_length=L).items()}
torch.cuda.synchronize()
eng.set_fast_lstm(True)
bad = torch.nonzero((fast['barcode_score'] != exact['barcode_score']) | (fast['barcode'] != exact['barcode']) |
                    (fast['barcode_guess'] != exact['barcode_guess'])).flatten().cpu().numpy()
print('differing reads:', bad)
calib = np.array(eng.demux_model.calibration)
for r in bad:
    print('read', r, 'fast score/guess/bc', fast['barcode_score'][r].item(), fast['barcode_guess'][r].item(), fast['barcode'][r].item(),
          'exact', exact['barcode_score'][r].item(), exact['barcode_guess'][r].item(), exact['barcode'][r].item())
    print('  probs fast ', fast['class_probs'][r, :5].cpu().numpy())
    print('  probs exact', exact['class_probs'][r, :5].cpu().numpy())
    pe = float(exact['class_probs'][r, :5].max()); pf = float(fast['class_probs'][r, :5].max())
    k = int(np.argmin(np.abs(calib - pe)))
    print('  nearest calibration edge', k, calib[k], 'exact score - edge', pe - calib[k], 'fast - edge', pf - calib[k])
    print('  scale/shift fast', fast['scale_shift'][r].cpu().numpy(), 'exact', exact['scale_shift'][r].cpu().numpy())
    print('  segments equal', bool((fast['segments'][r] == exact['segments'][r]).all()))
    # the window the exact path classifies, through the stage API, then the tensor-core verdict on it
    one = tuple(t[r:r + 1] if t.dim() == 1 and t.numel() == n else t for t in work[1:])
    sub = (raw[r].contiguous(), torch.zeros(1, dtype=torch.int64, device=dev), work[2][r:r + 1],
           work[3][r:r + 1], work[4][r:r + 1], work[5][r:r + 1])
    pooled = eng.pool_signal(*sub, max_raw_length=L)
    win, pushed = eng.barcode_windows(*sub, pooled, exact['scale_shift'][r:r + 1].contiguous(),
                                      exact['status'][r:r + 1].contiguous(), exact['segments'][r:r + 1].contiguous())
    p_tc, lg, bc, g, s, unsafe, sens = eng.demux_predict_tc(win)
    eng.set_fast_lstm(False)
    p_ex = eng.demux_predict(win)[0]
    eng.set_fast_lstm(True)
    pe5, pt5 = p_ex[0, :5].double().cpu().numpy(), p_tc[0, :5].double().cpu().numpy()
    am = pe5.argmax()
    err = np.abs((np.log(pe5) - np.log(pe5[am])) - (np.log(pt5) - np.log(pt5[am]))).max()
    print('  on the exact window: tc unsafe', unsafe.item(), 'probe shift', sens.item(), 'logit error', err,
          'bound', 2e-3 + 0.25 * sens.item(), 'pad', int((win[0] == -1000).sum().item()))
