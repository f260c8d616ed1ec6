"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- never imported by
poreplex_b200/; only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it).

Wraps oracle/libpb_oracle.so (pb_oracle.c) and, when present, the reference's own
event detector compiled into oracle/_ref/libscrappie_ref.so.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libpb_oracle.so')
REF_SCRAPPIE_PATH = os.path.join(HERE, '_ref', 'libscrappie_ref.so')

MAX_STATES, MAX_COMP, MAX_EDGES = 8, 4, 64
MAX_UNITS, MAX_CLASSES, MAX_CALIB = 64, 8, 64
SQRT_2_PI = 2.50662827463      # pomegranate NormalDistribution.pyx DEF

STATUS_NAMES = ['okay', 'disappeared', 'irregular_fast5', 'scaler_signal_too_short',
                'scaling_qc_fail', 'adapter_not_detected', 'not_basecalled',
                'basecall_table_incomplete', 'unsplit_read', 'sequence_too_short',
                'unknown_error']
FLAG_BARCODING = 1

_fp = C.POINTER(C.c_float)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, 'pb_oracle.c')):
        subprocess.check_call(['make', '-C', HERE, 'CC=gcc', 'libpb_oracle.so'])
    if os.path.isdir('/root/reference') and (force or not os.path.exists(REF_SCRAPPIE_PATH)):
        subprocess.check_call(['make', '-C', HERE, 'CC=gcc', 'ref'])


class LstmC(C.Structure):
    _fields_ = [('in_dim', C.c_int32), ('units', C.c_int32), ('impl', C.c_int32),
                ('W', _fp), ('U', _fp), ('b', _fp)]


class ScalerC(C.Structure):
    _fields_ = [('l1', LstmC), ('l2', LstmC), ('Wd', _fp), ('bd', _fp)]


class DemuxC(C.Structure):
    _fields_ = [('fwd', LstmC), ('bwd', LstmC), ('l2', LstmC), ('Wd', _fp), ('bd', _fp),
                ('n_classes', C.c_int32)]


class HmmC(C.Structure):
    _fields_ = [('n_states', C.c_int32),
                ('n_comp', C.c_int32 * MAX_STATES),
                ('mu', (C.c_double * MAX_COMP) * MAX_STATES),
                ('lsp', (C.c_double * MAX_COMP) * MAX_STATES),
                ('inv2s2', (C.c_double * MAX_COMP) * MAX_STATES),
                ('logw', (C.c_double * MAX_COMP) * MAX_STATES),
                ('log_start', C.c_double * MAX_STATES),
                ('in_begin', C.c_int32 * (MAX_STATES + 1)),
                ('in_src', C.c_int32 * MAX_EDGES),
                ('in_logp', C.c_double * MAX_EDGES)]


class EventC(C.Structure):
    _fields_ = [('start', C.c_uint64), ('length', C.c_float), ('mean', C.c_float),
                ('stdv', C.c_float), ('pos', C.c_int32), ('state', C.c_int32)]


class ModelC(C.Structure):
    _fields_ = [('scaler', ScalerC), ('demux', DemuxC), ('seg', HmmC),
                ('stride', C.c_int32), ('scaler_length', C.c_int32),
                ('scaler_min_length', C.c_int32), ('scan_limit', C.c_int32),
                ('scale_std', C.c_double), ('scale_mean', C.c_double),
                ('shift_std', C.c_double), ('shift_mean', C.c_double),
                ('qc_scale', C.c_double * 2), ('qc_shift', C.c_double * 2),
                ('adapter_state', C.c_int32),
                ('demux_minlen', C.c_int32), ('demux_maxlen', C.c_int32),
                ('demux_trimlen', C.c_int32), ('n_decoy', C.c_int32),
                ('demux_pad', C.c_float), ('n_calib', C.c_int32),
                ('calib', C.c_double * MAX_CALIB), ('score_threshold', C.c_double)]


class ResultC(C.Structure):
    _fields_ = [('status', C.c_int32), ('z', C.c_float * 2),
                ('scale', C.c_float), ('shift', C.c_float),
                ('viterbi_logp', C.c_double),
                ('seg', (C.c_int32 * 2) * MAX_STATES),
                ('pushed', C.c_int32),
                ('barcode', C.c_int32), ('guess', C.c_int32), ('phred', C.c_int32),
                ('probs', C.c_float * MAX_CLASSES)]


RESULT_DTYPE = np.dtype([('status', 'i4'), ('z', 'f4', 2), ('scale', 'f4'), ('shift', 'f4'),
                         ('viterbi_logp', 'f8'), ('seg', 'i4', (MAX_STATES, 2)),
                         ('pushed', 'i4'), ('barcode', 'i4'), ('guess', 'i4'),
                         ('phred', 'i4'), ('probs', 'f4', MAX_CLASSES)], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(ResultC)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_tanhf.restype = C.c_float
        L.orc_tanhf.argtypes = [C.c_float]
        L.orc_sigmoidf.restype = C.c_float
        L.orc_sigmoidf.argtypes = [C.c_float]
        L.orc_expf.restype = C.c_float
        L.orc_expf.argtypes = [C.c_float]
        L.orc_exp_neg.restype = C.c_double
        L.orc_exp_neg.argtypes = [C.c_double]
        L.orc_log_1to2.restype = C.c_double
        L.orc_log_1to2.argtypes = [C.c_double]
        L.orc_pair_lse.restype = C.c_double
        L.orc_pair_lse.argtypes = [C.c_double, C.c_double]
        L.orc_viterbi.restype = C.c_double
        L.orc_detect_events.restype = C.c_int64
        L.orc_barcode_window.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def bake_hmm(modeldata):
    """YAML state list -> HmmC in pomegranate's baked order (the oracle's own
    restatement of bake()): states sorted by name, log
    probabilities, out-edge re-weighting only if round(sum, 8) != 1."""
    names = sorted(s['name'] for s in modeldata)
    idx = {n: i for i, n in enumerate(names)}
    S = len(names)
    h = HmmC()
    h.n_states = S
    logtrans = [[-math.inf] * S for _ in range(S)]
    for s in modeldata:
        i = idx[s['name']]
        em = s['emission']
        h.n_comp[i] = len(em)
        ws = [(e[2] if len(e) > 2 else 1.0) for e in em]
        wsum = float(np.sum(np.array(ws, dtype=np.float64)))
        for j, e in enumerate(em):
            mu, sigma = float(e[0]), float(e[1])
            h.mu[i][j] = mu
            h.lsp[i][j] = -math.log(sigma * SQRT_2_PI)
            h.inv2s2[i][j] = 1. / (2 * sigma ** 2)
            h.logw[i][j] = math.log(ws[j] / wsum)
        h.log_start[i] = math.log(s['start_prob']) if s.get('start_prob', 0) > 0 else -math.inf
        for nxt, prob in s['transition']:
            if prob > 0:
                logtrans[i][idx[nxt]] = math.log(prob)
    for i in range(S):
        tot = round(sum(math.e ** lp for lp in logtrans[i] if lp > -math.inf), 8)
        if tot != 1. and tot > 0:
            logtrans[i] = [lp - math.log(tot) for lp in logtrans[i]]
    tot = round(sum(math.e ** h.log_start[i] for i in range(S) if h.log_start[i] > -math.inf), 8)
    if tot != 1. and tot > 0:
        for i in range(S):
            h.log_start[i] = h.log_start[i] - math.log(tot)
    for i in range(S, MAX_STATES):
        h.log_start[i] = -math.inf
    k = 0
    for l in range(S):
        h.in_begin[l] = k
        for src in range(S):
            if logtrans[src][l] > -math.inf:
                h.in_src[k] = src
                h.in_logp[k] = logtrans[src][l]
                k += 1
    for l in range(S, MAX_STATES + 1):
        h.in_begin[l] = k
    return h, names


def _lstm_c(layer, keep):
    W = np.ascontiguousarray(layer.kernel, np.float32)
    U = np.ascontiguousarray(layer.recurrent, np.float32)
    b = np.ascontiguousarray(layer.bias, np.float32)
    keep += [W, U, b]
    return LstmC(layer.in_dim, layer.units, layer.implementation, _p(W), _p(U), _p(b))


class Oracle:
    """The CPU oracle for one parameter set (preset dict + the two weight files)."""

    def __init__(self, preset, scaler_model, demux_model, barcoding_quality_filter=18):
        from scipy.stats import norm
        self._keep = []
        self.preset = preset
        m = ModelC()
        # scaler
        m.scaler.l1 = _lstm_c(scaler_model.l1, self._keep)
        m.scaler.l2 = _lstm_c(scaler_model.l2, self._keep)
        Wd = np.ascontiguousarray(scaler_model.dense_kernel, np.float32)
        bd = np.ascontiguousarray(scaler_model.dense_bias, np.float32)
        self._keep += [Wd, bd]
        m.scaler.Wd, m.scaler.bd = _p(Wd), _p(bd)
        sp = preset['signal_processing']
        idef = scaler_model.input_defs
        m.stride = sp['rough_signal_stride']
        m.scaler_length = idef['length']
        m.scaler_min_length = sp.get('scaler_min_length_override', idef['min_length'])
        m.scan_limit = preset['segmentation']['segmentation_scan_limit']
        xf = scaler_model.output_transform
        m.scale_std, m.scale_mean = xf['scale_std'], xf['scale_mean']
        m.shift_std, m.shift_mean = xf['shift_std'], xf['shift_mean']
        # signal_loader.py:62-68
        q = sp['scaler_qc_threshold']
        qs = norm.ppf([q, 1 - q], xf['scale_mean'], xf['scale_std'])
        qh = norm.ppf([q, 1 - q], xf['shift_mean'], xf['shift_std'])
        m.qc_scale[0], m.qc_scale[1] = qs
        m.qc_shift[0], m.qc_shift[1] = qh
        # hmm
        m.seg, self.seg_names = bake_hmm(preset['segmentation_model'])
        self.unsplit_hmm, self.unsplit_names = bake_hmm(preset['unsplit_read_detection_model'])
        m.adapter_state = self.seg_names.index('adapter')
        # demux
        dm = preset['demultiplexing']
        m.demux.fwd = _lstm_c(demux_model.fwd, self._keep)
        m.demux.bwd = _lstm_c(demux_model.bwd, self._keep)
        m.demux.l2 = _lstm_c(demux_model.l2, self._keep)
        Wd2 = np.ascontiguousarray(demux_model.dense_kernel, np.float32)
        bd2 = np.ascontiguousarray(demux_model.dense_bias, np.float32)
        self._keep += [Wd2, bd2]
        m.demux.Wd, m.demux.bd = _p(Wd2), _p(bd2)
        m.demux.n_classes = demux_model.n_classes
        m.demux_minlen = dm['minimum_dna_length']
        m.demux_maxlen = dm['maximum_dna_length']
        m.demux_trimlen = dm['signal_trim_length']
        m.n_decoy = dm['number_of_decoy_labels']
        m.demux_pad = -1000.0
        calib = list(demux_model.calibration)
        if len(calib) - 1 < barcoding_quality_filter:     # barcoding.py:41-45
            raise ValueError('The current demultiplexer does not support calibrated score '
                             'of {}.'.format(barcoding_quality_filter))
        m.n_calib = len(calib)
        for i, v in enumerate(calib):
            m.calib[i] = v
        m.score_threshold = calib[barcoding_quality_filter]
        self.model = m
        self.L = lib()

    # ---- element kernels -------------------------------------------------
    def dac_to_pa(self, raw, gain, offset):
        raw = np.ascontiguousarray(raw, np.int16)
        out = np.empty(len(raw), np.float32)
        self.L.orc_dac_to_pa(_p(raw, C.c_int16), C.c_int64(len(raw)), C.c_double(gain),
                             C.c_double(offset), _p(out))
        return out

    def pool_mean(self, x, stride=15):
        x = np.ascontiguousarray(x, np.float32)
        n = len(x) // stride
        out = np.empty(n, np.float32)
        self.L.orc_pool_mean(_p(x), C.c_int64(n), C.c_int(stride), _p(out))
        return out

    def scale(self, x, scale, shift):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self.L.orc_scale(_p(x), C.c_int64(len(x)), C.c_float(scale), C.c_float(shift), _p(out))
        return out

    def scaler_predict(self, heads):
        heads = np.ascontiguousarray(heads, np.float32)
        heads = heads.reshape(heads.shape[0], -1)
        out = np.empty((heads.shape[0], 2), np.float32)
        for i in range(heads.shape[0]):
            self.L.orc_scaler_predict(C.byref(self.model.scaler), _p(heads[i]),
                                      C.c_int(heads.shape[1]), _p(out[i]))
        return out

    def demux_predict(self, windows):
        windows = np.ascontiguousarray(windows, np.float32)
        windows = windows.reshape(windows.shape[0], -1)
        n = self.model.demux.n_classes
        out = np.empty((windows.shape[0], n), np.float32)
        for i in range(windows.shape[0]):
            self.L.orc_demux_predict(C.byref(self.model.demux), _p(windows[i]),
                                     C.c_int(windows.shape[1]), _p(out[i]))
        return out

    def viterbi(self, signal, which='seg'):
        hmm = self.model.seg if which == 'seg' else self.unsplit_hmm
        x = np.ascontiguousarray(signal, np.float32)
        path = np.empty(max(len(x), 1), np.int32)
        logp = self.L.orc_viterbi(C.byref(hmm), _p(x), C.c_int(len(x)), _p(path, C.c_int32))
        return logp, path[:len(x)]

    def emissions(self, x, which='seg'):
        hmm = self.model.seg if which == 'seg' else self.unsplit_hmm
        e = np.empty(MAX_STATES, np.float64)
        self.L.orc_hmm_emissions(C.byref(hmm), C.c_double(x), _p(e, C.c_double))
        return e[:hmm.n_states]

    def segments(self, path, n_states=None):
        path = np.ascontiguousarray(path, np.int32)
        n_states = n_states or self.model.seg.n_states
        seg = np.empty((n_states, 2), np.int32)
        self.L.orc_segments_from_path(_p(path, C.c_int32), C.c_int(len(path)),
                                      C.c_int(n_states), _p(seg, C.c_int32))
        return seg

    def barcode_window(self, adapter_signal):
        m = self.model
        x = np.ascontiguousarray(adapter_signal, np.float32)
        out = np.empty(m.demux_trimlen, np.float32)
        ok = self.L.orc_barcode_window(_p(x), C.c_int(len(x)), m.demux_minlen, m.demux_maxlen,
                                       m.demux_trimlen, C.c_float(m.demux_pad), _p(out))
        return out if ok else None

    def barcode_decide(self, probs):
        m = self.model
        probs = np.ascontiguousarray(probs, np.float32)
        b, g, p = C.c_int32(), C.c_int32(), C.c_int32()
        self.L.orc_barcode_decide(_p(probs), m.demux.n_classes, m.n_decoy, m.calib, m.n_calib,
                                  C.c_double(m.score_threshold), C.byref(b), C.byref(g),
                                  C.byref(p))
        return (None if b.value < 0 else b.value), g.value, p.value

    def detect_events(self, signal, window_length1=30, window_length2=120, threshold1=3.0,
                      threshold2=9.0, peak_height=8.0):
        return detect_events_restated(signal, window_length1, window_length2, threshold1,
                                      threshold2, peak_height)

    # ---- whole pipeline ---------------------------------------------------
    def process_batch(self, raw, offsets, lengths, gain, offset, barcoding=True, nthreads=0):
        raw = np.ascontiguousarray(raw, np.int16)
        offsets = np.ascontiguousarray(offsets, np.int64)
        lengths = np.ascontiguousarray(lengths, np.int64)
        gain = np.ascontiguousarray(gain, np.float64)
        offset = np.ascontiguousarray(offset, np.float64)
        N = len(lengths)
        res = np.zeros(N, RESULT_DTYPE)
        self.L.orc_process_batch(C.byref(self.model), _p(raw, C.c_int16), _p(offsets, C.c_int64),
                                 _p(lengths, C.c_int64), _p(gain, C.c_double),
                                 _p(offset, C.c_double), C.c_int64(N),
                                 C.c_int(FLAG_BARCODING if barcoding else 0),
                                 C.c_int(nthreads), res.ctypes.data_as(C.POINTER(ResultC)))
        return res

    def max_threads(self):
        return self.L.orc_max_threads()


EVENT_DTYPE = np.dtype([('start', 'u8'), ('length', 'f4'), ('mean', 'f4'), ('stdv', 'f4'),
                        ('pos', 'i4'), ('state', 'i4')])      # csupport.c:156-159 (28 B)
_EVENT_C_DTYPE = np.dtype([('start', 'u8'), ('length', 'f4'), ('mean', 'f4'), ('stdv', 'f4'),
                           ('pos', 'i4'), ('state', 'i4')], align=True)   # C event_t, 32 B


def _as_signal(signal):
    sig = np.ascontiguousarray(signal, dtype=np.float32)   # csupport.c:92-93
    if sig.ndim != 1:
        raise ValueError('Expects an 1-dimensional array.')  # csupport.c:97-101
    return sig


def detect_events_restated(signal, window_length1=30, window_length2=120, threshold1=3.0,
                           threshold2=9.0, peak_height=8.0):
    """csupport.detect_events through the restated detector (pb_oracle.c)."""
    sig = _as_signal(signal)
    n = len(sig)
    buf = np.zeros(max(n, 1), _EVENT_C_DTYPE)
    ne = lib().orc_detect_events(_p(sig), C.c_int64(n), C.c_int64(window_length1),
                                 C.c_int64(window_length2), C.c_float(threshold1),
                                 C.c_float(threshold2), C.c_float(peak_height),
                                 buf.ctypes.data_as(C.POINTER(EventC)), C.c_int64(len(buf)))
    out = np.empty(ne, EVENT_DTYPE)
    for f in EVENT_DTYPE.names:
        out[f] = buf[f][:ne]
    out['pos'] = -1
    out['state'] = -1
    return out


# --- the reference's own scrappie build (oracle/_ref) ------------------------
class _RawTable(C.Structure):
    _fields_ = [('n', C.c_size_t), ('start', C.c_size_t), ('end', C.c_size_t), ('raw', _fp)]


class _EventTable(C.Structure):
    _fields_ = [('n', C.c_size_t), ('start', C.c_size_t), ('end', C.c_size_t),
                ('event', C.POINTER(EventC))]


class _DetectorParam(C.Structure):
    _fields_ = [('window_length1', C.c_size_t), ('window_length2', C.c_size_t),
                ('threshold1', C.c_float), ('threshold2', C.c_float),
                ('peak_height', C.c_float)]


_ref_lib = None


def have_ref_scrappie():
    return os.path.exists(REF_SCRAPPIE_PATH)


def detect_events_ref(signal, window_length1=30, window_length2=120, threshold1=3.0,
                      threshold2=9.0, peak_height=8.0):
    """csupport.detect_events semantics (src/csupport.c:70-124) over the REFERENCE's
    own event_detection.c compiled into oracle/_ref (structs by value via ctypes)."""
    global _ref_lib
    if _ref_lib is None:
        _ref_lib = C.CDLL(REF_SCRAPPIE_PATH)
        _ref_lib.detect_events.restype = _EventTable
        _ref_lib.detect_events.argtypes = [_RawTable, _DetectorParam]
        _ref_lib_c = C.CDLL(None)
        _ref_lib._free = _ref_lib_c.free
        _ref_lib._free.argtypes = [C.c_void_p]
    sig = _as_signal(signal)
    rt = _RawTable(len(sig), 0, len(sig), _p(sig))
    et = _ref_lib.detect_events(rt, _DetectorParam(window_length1, window_length2, threshold1,
                                                   threshold2, peak_height))
    if et.n <= 0:
        raise RuntimeError('Event detection failed.')
    buf = np.ctypeslib.as_array(C.cast(et.event, C.POINTER(C.c_uint8)),
                                shape=(et.n * C.sizeof(EventC),)).view(_EVENT_C_DTYPE)
    out = np.empty(et.n, EVENT_DTYPE)
    for f in EVENT_DTYPE.names:
        out[f] = buf[f]
    _ref_lib._free(C.cast(et.event, C.c_void_p))
    return out


def default_oracle(bench_short=False, barcoding_quality_filter=18):
    from poreplex_b200 import params
    preset = params.load_preset()
    if bench_short:
        preset = params.bench_short_preset(preset)
    sc = params.load_scaler_model(preset['signal_processing']['scaler_model'])
    dm = params.load_demux_model(preset['demultiplexing']['demux_model'])
    return Oracle(preset, sc, dm, barcoding_quality_filter)
