"""Seeded synthetic direct-RNA reads for benchmarks and parity tests.

Reads are sampled from the preset's own segmentation HMM emissions
(presets/rna-r941.cfg:61-101 in the reference) at the pooled rate, up-sampled by the
pooling stride with per-sample noise, pushed through the inverse of a planted
(scale, shift) and quantised to int16 DAC counts with typical MinION calibration.
The recipe follows SURVEY.md section 8(d) "Generator calibration": the scaled-space
signal is compressed about 100 pA by ``k`` before the planted transform is inverted,
otherwise the scaler network sends most reads to ``scaling_qc_fail``.

Written with torch ops only so the same code fills a CUDA buffer in the benchmark
(torch is plumbing here: RNG + device memory) and a small CPU batch in the tests.
"""
import numpy as np
import torch

__all__ = ['SynthSpec', 'generate_reads']

# state order used by the generator (time order of a direct-RNA read)
_ORDER = ['pre-leader', 'leader-low', 'leader-high', 'adapter', 'polya-tail', 'transcript']


# pA centre / spread a prototype unit maps to: 74 - 3.0 * 5.5 .. 74 + 2.2 * 5.5 stays where the adapter emission of
# the segmentation HMM beats the transcript mixture (above ~90 pA the transcript state wins)
ADAPTER_LEVEL = (74.0, 5.5)

class SynthSpec:
    def __init__(self, read_length=4000, adapter_pooled=(110, 190), polya_pooled=(8, 40),
                 lead_pooled=((4, 12), (6, 14), (6, 14)), compress=0.78,
                 scale_dist=(0.955, 0.05), shift_dist=(5.5, 4.0), sample_noise=1.5,
                 frac_no_adapter=0.01, frac_qc_fail=0.01, stride=15, transcript_level=None,
                 frac_barcoded=0.8, frac_weak_barcode=0.3, weak_alpha=(0.35, 0.95),
                 prototypes=None, adapter_level=None):
        self.read_length = read_length
        self.adapter_pooled = adapter_pooled
        self.polya_pooled = polya_pooled
        self.lead_pooled = lead_pooled
        self.compress = compress
        self.scale_dist = scale_dist
        self.shift_dist = shift_dist
        self.sample_noise = sample_noise
        self.frac_no_adapter = frac_no_adapter
        self.frac_qc_fail = frac_qc_fail
        self.stride = stride
        # (mean, sd) of a single-Gaussian transcript level, or None to sample the
        # preset's two-component transcript mixture
        self.transcript_level = transcript_level
        # Barcodes: a fraction of the reads carries one of the four class prototypes of
        # presets/synth_barcode_prototypes.npz (tools/make_barcode_prototypes.py) in its
        # adapter, the rest an adapter sampled from the HMM emission (which the
        # demultiplexer calls decoy).  A planted adapter is alpha * prototype +
        # sqrt(1 - alpha^2) * noise: alpha = 1 for most, drawn from `weak_alpha` for
        # `frac_weak_barcode` of them so that scores spread over the calibration bins and
        # both sides of the acceptance threshold (barcoding.py:108-118).
        self.frac_barcoded = frac_barcoded
        self.frac_weak_barcode = frac_weak_barcode
        self.weak_alpha = weak_alpha
        self.prototypes = prototypes              # 'short' | 'stock' | None (by read length)
        self.adapter_level = adapter_level or ADAPTER_LEVEL   # (pA centre, pA per unit)

    @classmethod
    def for_length(cls, L, **kw):
        """Sensible segment dwell ranges for read length L (raw samples)."""
        T = L // 15
        # Shape parameters were calibrated against the scaler network with the CPU
        # oracle (DESIGN.md "Synthetic reads"): they are the values for which >= 95 %
        # of reads come back `okay` with the planted adapter boundaries recovered.
        if T >= 700:                       # stock preset: adapter must be 260..3000 pooled
            d = dict(adapter_pooled=(270, max(280, min(330, T // 3))), polya_pooled=(20, 60),
                     compress=1.0, transcript_level=(105.0, 12.0))
        elif T >= 250:                     # bench-short: 100..3000 pooled
            # (adapters of 172..178 pooled samples: the barcode prototypes need that much
            # signal, shorter windows are decoys to the network whatever they hold)
            d = dict(adapter_pooled=(172, min(178, T - 80)), polya_pooled=(15, 30),
                     lead_pooled=((3, 6), (4, 8), (4, 8)), adapter_level=(76.0, 3.5),
                     compress=0.85, transcript_level=(112.0, 8.0))
        elif T >= 200:
            d = dict(adapter_pooled=(105, min(135, T - 110)), polya_pooled=(20, 50),
                     compress=0.85, transcript_level=(112.0, 8.0))
        else:                              # too short for demux under any preset
            d = dict(adapter_pooled=(max(8, T // 3), max(10, T // 2)),
                     polya_pooled=(4, max(5, T // 8)))
        d.update(kw)
        return cls(read_length=L, **d)


def _emission_table(preset, transcript_level=None):
    by_name = {s['name']: s for s in preset['segmentation_model']}
    mu = torch.zeros(len(_ORDER), 2)
    sd = torch.zeros(len(_ORDER), 2)
    w0 = torch.ones(len(_ORDER))
    for i, n in enumerate(_ORDER):
        em = by_name[n]['emission']
        mu[i, 0], sd[i, 0] = em[0][0], em[0][1]
        if len(em) > 1:
            mu[i, 1], sd[i, 1] = em[1][0], em[1][1]
            w0[i] = em[0][2] / (em[0][2] + em[1][2])
        else:
            mu[i, 1], sd[i, 1] = em[0][0], em[0][1]
    if transcript_level is not None:
        i = _ORDER.index('transcript')
        mu[i, :] = transcript_level[0]
        sd[i, :] = transcript_level[1]
        w0[i] = 1.0
    return mu, sd, w0


_PROTO_FILE = 'synth_barcode_prototypes.npz'
_proto_cache = {}


def load_prototypes(which):
    """[4][P] float32 class prototypes in normalised units, right-aligned at the adapter end."""
    import os
    if which not in _proto_cache:
        from .params import PRESET_DIR
        z = np.load(os.path.join(PRESET_DIR, _PROTO_FILE))
        _proto_cache[which] = np.ascontiguousarray(z[which], np.float32)
    return _proto_cache[which]


def generate_reads(n, spec, preset, seed=0, device='cpu', chunk=32768):
    """Return a dict of tensors on ``device``:

    ``raw`` int16 [n, L]; ``range``/``digitisation``/``offset`` f64 [n] (channel_id
    attributes); ``gain`` f64 [n] (= range / digitisation, fast5_file.py:130);
    ``sampling_rate`` f64 [n]; ``planted`` dict with the planted scale/shift and the
    pooled segment boundaries (for diagnostics only -- never used by the kernels).
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    L, st = spec.read_length, spec.stride
    T = L // st
    mu, sd, w0 = (t.to(dev) for t in _emission_table(preset, spec.transcript_level))

    raw = torch.empty((n, L), dtype=torch.int16, device=dev)
    gain = torch.empty(n, dtype=torch.float64, device=dev)
    rng_pa = torch.empty(n, dtype=torch.float64, device=dev)
    offset = torch.empty(n, dtype=torch.float64, device=dev)
    p_scale = torch.empty(n, dtype=torch.float32, device=dev)
    p_shift = torch.empty(n, dtype=torch.float32, device=dev)
    bounds = torch.empty((n, 5), dtype=torch.int32, device=dev)
    planted_bc = torch.zeros(n, dtype=torch.int32, device=dev)     # 0 none, 1..4 class
    which = spec.prototypes or ('stock' if T >= 700 else 'short')
    proto = torch.from_numpy(load_prototypes(which)).to(dev) if spec.frac_barcoded > 0 else None

    def randint(lo, hi, m):
        return torch.randint(int(lo), int(hi) + 1, (m,), generator=g, device=dev)

    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        # segment boundaries in pooled samples
        d = [randint(lo, hi, m) for lo, hi in spec.lead_pooled]
        d.append(randint(*spec.adapter_pooled, m))
        d.append(randint(*spec.polya_pooled, m))
        b = torch.cumsum(torch.stack(d, dim=1), dim=1)              # [m, 5]
        # reads with no adapter: leader-high never ends
        no_ad = torch.rand(m, generator=g, device=dev) < spec.frac_no_adapter
        b = torch.where(no_ad[:, None] & (torch.arange(5, device=dev)[None, :] >= 2),
                        torch.full_like(b, T + 1), b)
        t = torch.arange(T, device=dev)[None, :]
        state = (t[:, :, None] >= b[:, None, :]).sum(dim=2)         # [m, T] in 0..5
        comp = (torch.rand((m, T), generator=g, device=dev) >= w0[state]).long()
        level = mu[state, comp] + sd[state, comp] * torch.randn((m, T), generator=g, device=dev)
        if proto is not None:
            # class prototype over the last P pooled samples of the adapter
            P = proto.shape[1]
            cls = torch.randint(1, proto.shape[0] + 1, (m,), generator=g, device=dev)
            cls = torch.where(torch.rand(m, generator=g, device=dev) < spec.frac_barcoded,
                              cls, torch.zeros_like(cls))
            weak = torch.rand(m, generator=g, device=dev) < spec.frac_weak_barcode
            alpha = spec.weak_alpha[0] + (spec.weak_alpha[1] - spec.weak_alpha[0]) * \
                torch.rand(m, generator=g, device=dev)
            alpha = torch.where(weak, alpha, torch.ones_like(alpha))
            j = P - (b[:, 3:4] - t)                                     # [m, T] prototype index
            inside = (state == 3) & (j >= 0) & (j < P) & (cls[:, None] > 0)
            pv = proto[(cls[:, None] - 1).clamp(min=0), j.clamp(0, P - 1)]
            mix = alpha[:, None] * pv + torch.sqrt(1 - alpha[:, None] ** 2) * \
                torch.randn((m, T), generator=g, device=dev)
            level = torch.where(inside, spec.adapter_level[0] + spec.adapter_level[1] * mix, level)
            planted_bc[c0:c0 + m] = torch.where(no_ad, torch.zeros_like(cls), cls).to(torch.int32)
        # up-sample to the raw rate (+ remainder) and add per-sample noise
        sig = level.repeat_interleave(st, dim=1)
        if L > T * st:
            sig = torch.cat([sig, sig[:, -1:].expand(m, L - T * st)], dim=1)
        sig = sig + spec.sample_noise * torch.randn((m, L), generator=g, device=dev)
        sig = 100.0 + spec.compress * (sig - 100.0)
        sc = spec.scale_dist[0] + spec.scale_dist[1] * torch.randn(m, generator=g, device=dev)
        sh = spec.shift_dist[0] + spec.shift_dist[1] * torch.randn(m, generator=g, device=dev)
        qc_fail = torch.rand(m, generator=g, device=dev) < spec.frac_qc_fail
        pa = (sig - sh[:, None]) / sc[:, None]
        # deliberately failing stratum: a 2.2x amplitude error drives the predicted
        # scale below the QC window (signal_loader.py:100-109)
        pa = torch.where(qc_fail[:, None], pa * 2.2, pa)
        rng = 1180.0 + 290.0 * torch.rand(m, generator=g, device=dev, dtype=torch.float64)
        gn = rng / 8192.0
        off = randint(0, 20, m).to(torch.float64)
        dac = torch.round(pa.to(torch.float64) / gn[:, None] - off[:, None])
        raw[c0:c0 + m] = dac.clamp_(-32768, 32767).to(torch.int16)
        gain[c0:c0 + m] = gn
        rng_pa[c0:c0 + m] = rng
        offset[c0:c0 + m] = off
        p_scale[c0:c0 + m] = sc
        p_shift[c0:c0 + m] = sh
        bounds[c0:c0 + m] = b.to(torch.int32)
        del sig, pa, dac, level, state, comp
    return {
        'raw': raw, 'gain': gain, 'offset': offset, 'range': rng_pa,
        'digitisation': torch.full((n,), 8192.0, dtype=torch.float64, device=dev),
        'sampling_rate': torch.full((n,), 3012.0, dtype=torch.float64, device=dev),
        'planted': {'scale': p_scale, 'shift': p_shift, 'bounds': bounds,
                    'barcode': planted_bc},
    }


def to_numpy(reads):
    out = {}
    for k, v in reads.items():
        if isinstance(v, dict):
            out[k] = to_numpy(v)
        else:
            out[k] = v.detach().cpu().numpy()
    return out
