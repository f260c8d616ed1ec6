"""Pandas-free restatement of ``SignalAnalysis.load_events`` derived columns and
``SignalAnalysis.detect_unsplit_read`` (poreplex/signal_analyzer.py:311-326, 366-443) plus
``utils.union_intervals`` (utils.py:28-39).  TEST INFRASTRUCTURE: checked against the
reference running verbatim (tests/test_oracle_cpu.py), then used as the yardstick for the
CUDA chimera kernels."""
import numpy as np

f32 = np.float32


def derive_event_columns(start, mean, move, scale, shift):
    """load_events (signal_analyzer.py:320-325): scaled_mean (float32 Horner, unfused),
    pos = cumsum(move), end = start + diff(start) (last + 1)."""
    start = np.asarray(start, np.int64)
    scaled = np.poly1d(np.array([scale, shift], f32))(np.asarray(mean, f32))
    pos = np.cumsum(np.asarray(move).astype(np.int64))
    dur = np.hstack((np.diff(start), [1])).astype(np.int64)
    return scaled, pos, start + dur


def union_intervals(iset):
    merged = []
    for begin, end in sorted(iset):
        if merged:
            if merged[-1][-1] >= begin:
                if merged[-1][-1] < end:
                    merged[-1][-1] = end
                continue
        merged.append([begin, end])
    return merged


def detect_unsplit_read(cfg, viterbi, state_names, start, end, scaled_mean, pos, p_model_state,
                        adapter_last_pooled, sampling_rate, elspan=15):
    """Returns True when the read holds two or more molecules.
    ``viterbi(x) -> path`` decodes with the unsplit-read model (baked state indices)."""
    payload_start = (adapter_last_pooled + 1) * elspan
    _ = lambda name: int(cfg[name] * sampling_rate)
    window_size, window_step = _('window_size'), _('window_step')
    strict_duration = _('strict_duration')
    duration_cutoffs = [(_('loosen_full_length'), _('loosen_dna_length')),
                        (_('strict_full_length'), _('strict_dna_length'))]
    leaderish = {state_names.index(n) for n in ('adapter', 'leader-high', 'leader-low')}
    adapter = state_names.index('adapter')
    excessive = []
    for left in range(payload_start, int(end[-1]), window_step):
        sel = np.nonzero((start >= left) & (start <= left + window_size))[0]
        if len(sel) < 1:
            break
        path = viterbi(scaled_mean[sel])
        leader_start = None
        t, T = 0, len(path)
        while t < T:
            s, first = path[t], t
            while t + 1 < T and path[t + 1] == s:
                t += 1
            last = t
            t += 1
            if s not in leaderish:
                leader_start = None
                continue
            if leader_start is None:
                leader_start = first
            if s != adapter:
                continue
            adapter_end = int(end[sel[last]])
            leader_start_in_read = int(start[sel[leader_start]])
            total_duration = adapter_end - leader_start_in_read
            adapter_duration = adapter_end - int(start[sel[first]])
            total_cutoff, adapter_cutoff = duration_cutoffs[
                (leader_start_in_read - payload_start) <= strict_duration]
            if total_duration >= total_cutoff and adapter_duration >= adapter_cutoff:
                excessive.append([leader_start_in_read, 1 + adapter_end])
            leader_start = None
    if not excessive:
        return False
    intervals = [[0, payload_start]] + union_intervals(excessive) + [[np.inf, np.inf]]

    def count_hq(lo, hi):
        sel = (start >= lo) & (start <= hi)
        if not sel.any():
            return 0
        p, q = pos[sel], p_model_state[sel]
        n = 0
        i = 0
        while i < len(p):
            j = i
            best = q[i]
            while j + 1 < len(p) and p[j + 1] == p[i]:
                j += 1
                best = max(best, q[j])
            n += bool(best > cfg['basecount_quality_limit'])
            i = j + 1
        return n

    sub = [count_hq(a[1], b[0]) for a, b in zip(intervals[0:], intervals[1:])]
    total = sum(sub[1:])
    return bool(total > cfg['subread_basecount_limit'] or
                (total + 1) / (sub[0] + 1) > cfg['subread_baseratio_limit'])
