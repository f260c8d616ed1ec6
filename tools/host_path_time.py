"""Host-path study (run on a GPU box): python tools/host_path_time.py [reads]
pb2_analyze_host from pinned buffers, `streamed` (whole batch resident, one exact re-run at the end)
against `arena` (two chunk-sized arenas, a re-run per chunk), int16 and compressed uploads."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poreplex_b200 import params, synth, fast5_loader
from poreplex_b200.engine import SignalEngine
fast5_loader.build()
preset = params.bench_short_preset(params.load_preset())
eng = SignalEngine(dict(preset, barcoding=True), device=0)
n, L = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000000), 4000
dev = torch.device('cuda', 0)
rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=1, device=dev)
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
raw = pin(rd['raw'].reshape(-1)).numpy()
cal = [pin(rd[k]).numpy() for k in ('range', 'digitisation', 'offset')]
off = pin(torch.arange(n, dtype=torch.int64) * L).numpy(); ln = pin(torch.full((n,), L, dtype=torch.int64)).numpy()
pk, po = fast5_loader.svb16_encode(raw, off, ln, pinned=True)
out = eng.alloc_host_results(n, pinned=True)
res = {}
ref = None
for pipeline in ('streamed', 'streamed-early', 'arena', 'streamed'):
    os.environ['POREPLEX_B200_HOST_PIPELINE'] = pipeline.split('-')[0]
    os.environ['POREPLEX_B200_HOST_EARLY_RESOLVE'] = '1' if pipeline.endswith('early') else '-1'
    for name, fn in (('int16', lambda: eng.analyze_host(raw, off, ln, *cal, out=out)),
                     ('packed', lambda: eng.analyze_host(None, off, ln, *cal, out=out, packed=(pk, po)))):
        r = fn()
        ints = {k: np.array(r[k], copy=True) for k in ('status', 'segments', 'barcode', 'barcode_guess', 'barcode_score', 'label', 'counts')}
        if ref is None:
            ref = ints
        same = all(np.array_equal(ref[k], ints[k]) for k in ref)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); fn(); fn()
        ms = (time.perf_counter() - t0) * 1e3 / 3
        res.setdefault(pipeline + '/' + name, []).append(round(ms, 1))
        print(pipeline, name, 'ms per call %.1f' % ms, 'reads/s %.3e' % (n / ms * 1e3), 'same integers', same,
              'reruns', eng.recheck_stats()[0], flush=True)
print(json.dumps({'reads': n, 'ms_per_call': res}))
