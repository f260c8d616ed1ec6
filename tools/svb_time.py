"""Timing study of the compressed upload path (run on a GPU box): python tools/svb_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poreplex_b200 import params, synth, fast5_loader
from poreplex_b200.engine import SignalEngine
fast5_loader.build()
preset = params.bench_short_preset(params.load_preset())
eng = SignalEngine(dict(preset, barcoding=True), device=0)
n, L = 1000000, 4000
dev = torch.device('cuda', 0)
rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=1, device=dev)
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
raw = pin(rd['raw'].reshape(-1)).numpy()
cal = [pin(rd[k]).numpy() for k in ('range', 'digitisation', 'offset')]
off = pin(torch.arange(n, dtype=torch.int64) * L).numpy(); ln = pin(torch.full((n,), L, dtype=torch.int64)).numpy()
pk, po = fast5_loader.svb16_encode(raw, off, ln, pinned=True)
out = eng.alloc_host_results(n, pinned=True)
for name, fn in (('int16', lambda: eng.analyze_host(raw, off, ln, *cal, out=out)),
                 ('packed', lambda: eng.analyze_host(None, off, ln, *cal, out=out, packed=(pk, po)))):
    fn()
    eng.profile_enable(True); eng.profile_read()
    t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
    prof = eng.profile_read(); eng.profile_enable(False)
    print(name, 'ms %.1f' % (dt * 1e3), 'kernel ms total %.1f' % sum(v[0] for v in prof.values()),
          {k: round(v[0], 1) for k, v in prof.items() if v[0] > 4})
    t0 = time.perf_counter(); fn(); fn(); print('   unprofiled ms per call %.1f' % ((time.perf_counter() - t0) * 500))
