"""ctypes access to tests/hostcheck/libpb_hostcheck.so: a TEST-ONLY host (g++) build of the
__host__ __device__ cores in poreplex_b200/csrc, so the CPU suite can run the exact
kernel code against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostcheck')
LIB = os.path.join(HERE, 'libpb_hostcheck.so')
MAX_SPIKES = 48


class PolyaParamsC(C.Structure):
    _fields_ = [('stride', C.c_int32), ('refinement_expansion', C.c_int32),
                ('openend_unit', C.c_int32), ('max_extension', C.c_int32),
                ('w1', C.c_int32), ('w2', C.c_int32),
                ('thr1', C.c_float), ('thr2', C.c_float), ('peak_height', C.c_float),
                ('cutoff_lo', C.c_float), ('cutoff_hi', C.c_float), ('mean_loc', C.c_float),
                ('trigger', C.c_float), ('half_range', C.c_float), ('stdv_max', C.c_float),
                ('stdv_lo', C.c_double), ('stdv_hi', C.c_double),
                ('spike_tolerance', C.c_int32), ('spike_weight', C.c_double),
                ('recal_max_dist', C.c_int32), ('recal_min_length', C.c_float),
                ('recal_max_stdv', C.c_float)]


class PolyaResultC(C.Structure):
    _fields_ = [('found', C.c_int32), ('n_spikes', C.c_int32), ('begin', C.c_int64),
                ('end', C.c_int64), ('dwell_samples', C.c_int64), ('extensions', C.c_int32),
                ('flags', C.c_int32), ('spikes', (C.c_float * 4) * MAX_SPIKES)]


def polya_params(cfg, stride=15):
    """config['polya_dwell'] -> the kernel's parameter block (same rounding rules as
    PolyASignalAnalyzer.__init__, polya.py:39-48)."""
    f = np.float32
    loc, sd = cfg['polya_mean_dist']
    z = cfg['polya_mean_z_cutoff']
    ed = cfg['event_detection']
    rc = cfg['recalibrate_shifted_signal']
    return PolyaParamsC(stride, cfg['refinement_expansion'], cfg['openend_expansion'] // stride,
                        cfg['maximum_openend_extension'], ed['window_length1'], ed['window_length2'],
                        ed['threshold1'], ed['threshold2'], ed['peak_height'],
                        f(loc - sd * z), f(loc + sd * z), f(loc),
                        f(cfg['polya_mean_trigger_recalibration'] * sd), f(sd * z),
                        f(cfg['polya_stdv_max']), cfg['polya_stdv_range'][0],
                        cfg['polya_stdv_range'][1], cfg['spike_tolerance'], cfg['spike_weight'],
                        rc['max_dist_from_adapter'], f(rc['min_length']), f(rc['max_stdv']))


def load():
    src = os.path.join(HERE, 'hostcheck.cpp')
    deps = [src] + [os.path.join(HERE, '..', '..', 'poreplex_b200', 'csrc', f)
                    for f in ('polya_core.cuh', 'pb_math.cuh')]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off',
                               '-mavx2', '-mfma', '-o', LIB, src])
    L = C.CDLL(LIB)
    L.hc_median7.restype = C.c_float
    L.hc_median5.restype = C.c_float
    L.hc_pairwise_sum.restype = C.c_float
    L.hc_detect_events.restype = C.c_int64
    L.hc_detect_events_plain.restype = C.c_int64
    assert L.hc_sizeof_params() == C.sizeof(PolyaParamsC)
    assert L.hc_sizeof_result() == C.sizeof(PolyaResultC)
    return L


def result_to_dict(R, sampling_rate):
    """PolyaResult -> the dict polya.py hands to set_polya_tail (polya.py:116-121)."""
    if not R.found:
        return None
    spikes = []
    for k in range(min(R.n_spikes, MAX_SPIKES)):
        vals = [R.spikes[k][0]] + [R.spikes[k][j] for j in (1, 2, 3) if not np.isnan(R.spikes[k][j])]
        spikes.append(tuple(float(v) for v in vals))
    return {'begin': int(R.begin), 'end': int(R.end),
            'dwell_time': int(R.dwell_samples) / sampling_rate, 'spikes': spikes}
