"""Host-path chunking study (run on a GPU box): python tools/chunk_time.py"""
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
for chunks in (4, 3, 2, 6):
    os.environ['POREPLEX_B200_HOST_CHUNKS'] = str(chunks)
    for name, fn in (('int16', lambda: eng.analyze_host(raw, off, ln, *cal, out=out)),
                     ('packed', lambda: eng.analyze_host(None, off, ln, *cal, out=out, packed=(pk, po)))):
        fn()
        t0 = time.perf_counter(); fn(); fn(); fn()
        print('large chunks', chunks, name, 'ms per call %.1f' % ((time.perf_counter() - t0) * 1e3 / 3), flush=True)
