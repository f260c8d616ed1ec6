"""Where the host path spends its kernel time (run on a GPU box): per-kernel sums (CUDA events around
every launch) of one resident call and of one streamed host call over the same 1 M reads, with and
without the pool for partly filled classifier waves."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poreplex_b200 import params, synth
from poreplex_b200.engine import SignalEngine
preset = params.bench_short_preset(params.load_preset())
eng = SignalEngine(dict(preset, barcoding=True), device=0)
n, L = 1000000, 4000
dev = torch.device('cuda', 0)
rd = synth.generate_reads(n, synth.SynthSpec.for_length(L), preset, seed=1, device=dev)
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
raw = pin(rd['raw'].reshape(-1)).numpy()
cal = [pin(rd[k]).numpy() for k in ('range', 'digitisation', 'offset')]
off_t = torch.arange(n, dtype=torch.int64, device=dev) * L
ln_t = torch.full((n,), L, dtype=torch.int64, device=dev)
off = pin(off_t).numpy(); ln = pin(ln_t).numpy()
hout = eng.alloc_host_results(n, pinned=True)
dout = eng.alloc_results(n)
work = (rd['raw'].reshape(-1), off_t, ln_t, rd['range'], rd['digitisation'], rd['offset'])
res = {}
def timed(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); fn(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / 2
    eng.profile_enable(True); eng.profile_read()
    fn(); torch.cuda.synchronize()
    prof = eng.profile_read(); eng.profile_enable(False)
    return wall, {k: round(v[0], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}, sum(v[1] for v in prof.values())
w, p, c = timed(lambda: eng.analyze_device(*work, out=dout, barcoding=True, max_raw_length=L))
res['resident'] = {'wall_ms': w, 'kernel_ms_sum': sum(p.values()), 'launches': c, 'kernels': p}
for name, env in (('streamed', {}),):
    os.environ.pop('POREPLEX_B200_HOST_NO_TAIL_POOL', None)
    os.environ.update(env)
    w, p, c = timed(lambda: eng.analyze_host(raw, off, ln, *cal, out=hout))
    res[name] = {'wall_ms': w, 'kernel_ms_sum': sum(p.values()), 'launches': c, 'kernels': p}
# time line of one streamed call: where the compute stream sat idle
os.environ.pop('POREPLEX_B200_HOST_NO_TAIL_POOL', None)
eng.profile_enable(True); eng.profile_read()
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.analyze_host(raw, off, ln, *cal, out=hout)
wall = (time.perf_counter() - t0) * 1e3
tl = eng.profile_timeline(); eng.profile_enable(False)
gaps = []
for (n0, a0, b0), (n1, a1, b1) in zip(tl[:-1], tl[1:]):
    if a1 - b0 > 0.05:
        gaps.append({'after': n0, 'before': n1, 'at_ms': round(b0, 2), 'idle_ms': round(a1 - b0, 3)})
busy = sum(b - a for _, a, b in tl)
res['timeline'] = {'wall_ms_with_profiling': wall, 'first_launch_to_last_end_ms': tl[-1][2] if tl else 0,
                   'busy_ms': busy, 'idle_between_launches_ms': sum(g['idle_ms'] for g in gaps),
                   'gaps_over_50us': sorted(gaps, key=lambda g: -g['idle_ms'])[:24]}
print('timeline: wall %.1f, first launch -> last end %.1f, busy %.1f, idle %.1f' %
      (wall, res['timeline']['first_launch_to_last_end_ms'], busy, res['timeline']['idle_between_launches_ms']), file=sys.stderr)
for g in res['timeline']['gaps_over_50us'][:16]:
    print('   ', g, file=sys.stderr)
for k, v in res.items():
    if k == 'timeline':
        continue
    print(k, 'wall %.1f' % v['wall_ms'], 'kernel sum %.1f' % v['kernel_ms_sum'], 'launches', v['launches'], file=sys.stderr)
    print('   ', {a: b for a, b in list(v['kernels'].items())[:12]}, file=sys.stderr)
print(json.dumps(res))
