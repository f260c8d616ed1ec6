"""Pandas-free restatement of ``poreplex/polya.py`` (TEST INFRASTRUCTURE).

Follows PolyASignalAnalyzer line by line (citations in each function) with the numeric
semantics this container's numpy 2 / pandas 3 give the original: float32 Series
arithmetic, NEP-50 weak Python scalars (bounds are rounded to float32 before the
comparison), numpy's pairwise float32 sums, int64 score matrices with truncation on
store.  ``tests/test_oracle_cpu.py`` runs the reference's own polya.py verbatim next to
this on the same windows and requires identical results; the CUDA kernel is then held to
this restatement.
"""
import numpy as np
from scipy.signal import medfilt

from . import oracle as O

f32 = np.float32


class PolyAParams:
    """config['polya_dwell'] (rna-r941.cfg:38-59) with the derived constants of
    PolyASignalAnalyzer.__init__ (polya.py:39-48)."""

    def __init__(self, cfg):
        self.refinement_expansion = cfg['refinement_expansion']
        self.openend_expansion = cfg['openend_expansion']
        self.median_pre_filter = cfg['median_pre_filter']
        self.maximum_openend_extension = cfg['maximum_openend_extension']
        self.event_detection = dict(cfg['event_detection'])
        self.polya_mean_dist = list(cfg['polya_mean_dist'])
        self.polya_mean_z_cutoff = cfg['polya_mean_z_cutoff']
        self.polya_stdv_max = cfg['polya_stdv_max']
        self.polya_stdv_range = list(cfg['polya_stdv_range'])
        self.spike_tolerance = cfg['spike_tolerance']
        self.spike_weight = cfg['spike_weight']
        self.recal = dict(cfg['recalibrate_shifted_signal'])
        loc, scale = self.polya_mean_dist
        self.polya_mean_cutoff = (loc - scale * self.polya_mean_z_cutoff,
                                  loc + scale * self.polya_mean_z_cutoff)
        self.trigger = cfg['polya_mean_trigger_recalibration'] * scale


def between_f32(x, lo, hi):
    """Series(float32).between(lo, hi): both bounds act as float32 (NEP 50)."""
    return (x >= f32(lo)) & (x <= f32(hi))


def find_best_polya_interval(is_polya, length, spike_tolerance, spike_weight):
    """polya.py:156-187.  Returns (start, end) inclusive or None."""
    E = len(length)
    v = (is_polya.astype(np.int64) * 2 - 1) * length.astype(np.float64)
    m = np.where(v > 0, v, v * spike_weight).astype(np.int64)          # trunc toward zero
    s = np.where(is_polya, 1.0, -length.astype(np.float64)).astype(np.int64)
    best, best_ij = 0, None
    for i in range(E):
        M, S = 0, 0
        for j in range(i, E):
            M += int(m[j])
            S = -1 if S < 0 else (spike_tolerance if s[j] > 0 else S + int(s[j]))
            val = M if S > 0 else 0
            if val > best:                      # first row-major maximum
                best, best_ij = val, (i, j)
    return best_ij


def internal_stdv(signal, start, length, lo, hi):
    """calc_internal_polya_stdv (polya.py:150-154)."""
    length = int(length)
    begin = int(np.float64(start) + length * lo)
    end = int(np.float64(start) + length * hi)
    return signal[begin:end].std() if end - begin > 2 else np.nan


def analyze(P, scaled_signal, sampling_rate, rough_range, stride, detect_events=None):
    """PolyASignalAnalyzer.__call__ / call_polya / try_recalibrate_shifted_signal with the
    mutual recursion written as a state machine (polya.py:50-148).

    ``scaled_signal`` = npread.load_signal(pool=None) (float32, full resolution);
    ``rough_range`` = (begin, end-or-None) in pooled samples.  Returns the dict handed to
    set_polya_tail, or None."""
    detect = detect_events or O.detect_events_restated
    full_length = len(scaled_signal)
    unit = P.openend_expansion // stride                       # 1000 // 15 = 66
    rough_begin, rough_end_cur = rough_range
    polya_range = None
    ext_depth = 0
    while True:
        # ---- __call__ (polya.py:50-73): window, median filter, events
        rough_end = rough_end_cur
        if rough_end is None or rough_end - rough_begin < unit:
            rough_end = rough_begin + unit
        insp_begin = max(0, rough_begin * stride - P.refinement_expansion)
        insp_end = min(full_length, (rough_end + 1) * stride + P.refinement_expansion)
        adapter_end = rough_begin * stride - insp_begin
        sig = scaled_signal[insp_begin:insp_end]
        if P.median_pre_filter > 1:
            sig = medfilt(sig, P.median_pre_filter)
        ev = detect(sig, **P.event_detection)
        start = ev['start'].astype(np.uint64)
        length = ev['length'].astype(f32)
        mean = ev['mean'].astype(f32)
        stdv = ev['stdv'].astype(f32)
        end = (start.astype(np.float64) + length.astype(np.float64)).astype(np.int64)
        lo, hi = polya_range or P.polya_mean_cutoff
        is_polya = between_f32(mean, lo, hi)
        mode = 'call' if rough_end_cur is not None else 'recal'
        extend = False
        while True:
            if mode == 'recal':
                # ---- try_recalibrate_shifted_signal (polya.py:127-148)
                anchor = ((start <= np.uint64(adapter_end + P.recal['max_dist_from_adapter']))
                          & (end > adapter_end) & (stdv < f32(P.recal['max_stdv'])))
                if not anchor.any():
                    return None
                pm = (mean[anchor] * length[anchor]).sum() / length[anchor].sum()     # float32
                half = P.polya_mean_dist[1] * P.polya_mean_z_cutoff
                polya_range = (pm - f32(half), pm + f32(half))                        # float32
                is_polya = between_f32(mean, polya_range[0], polya_range[1])
                if not (length[is_polya].sum() >= f32(P.recal['min_length'])):
                    return None
                mode = 'call'
            # ---- call_polya (polya.py:75-125)
            best = find_best_polya_interval(is_polya, length, P.spike_tolerance, P.spike_weight)
            if (best is not None and best[1] == len(length) - 1 and insp_end < full_length
                    and ext_depth < P.maximum_openend_extension):
                extend = True
                break
            if best is None:
                mode = 'recal'
                continue
            i, j = best
            if polya_range is None:
                lvl = (mean[i:j + 1] * length[i:j + 1]).sum() / length[i:j + 1].sum()
                if abs(lvl - f32(P.polya_mean_dist[0])) > f32(P.trigger):
                    mode = 'recal'
                    continue
            k = i + int(np.argmax(length[i:j + 1]))                   # idxmax: first maximum
            sd = internal_stdv(sig, start[k], length[k], *P.polya_stdv_range)
            if sd < P.polya_stdv_max:
                pol = is_polya[i:j + 1]
                sub_mean, sub_len = mean[i:j + 1], length[i:j + 1]
                spikes = []
                for spk in np.where(~pol)[0]:
                    spikes.append((float(sub_len[spk]),) +
                                  tuple(float(x) for x in sub_mean[spk - 1:spk + 2]))
                return {'begin': int(start[i]) + insp_begin,
                        'end': int(np.float64(start[j]) + np.float64(length[j])) + insp_begin,
                        'dwell_time': int(sub_len[pol].sum()) / sampling_rate,
                        'spikes': spikes}
            if polya_range is None:
                mode = 'recal'
                continue
            return None
        # extension: self(npread, (base_range[0], base_range[1] + unit), ..., ext_depth + 1)
        rough_end_cur = rough_end + unit
        ext_depth += 1
