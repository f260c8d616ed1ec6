"""SignalEngine: the per-process GPU state behind ``SignalAnalyzer``.

Plays the role of the objects the reference keeps alive per worker process in
``WorkerPersistenceStorage`` (worker_persistence.py:35-90): the segmentation HMM, the
scaler network (``SignalLoader``) and the barcode demultiplexer
(``BarcodeDemultiplexer``), here as one native context holding their parameters in
HBM.  Every numeric stage is a call through the C ABI (include/poreplex_b200.h);
torch is used only to own device buffers / streams in the device-resident API.
"""
import ctypes as C

import numpy as np

from . import _native as N
from . import params as P

__all__ = ['SignalEngine', 'get_engine', 'close_engines', 'config_digest', 'POLYA_DTYPE',
           'polya_to_dict']

POLYA_DTYPE = np.dtype([('found', 'i4'), ('n_spikes', 'i4'), ('begin', 'i8'), ('end', 'i8'),
                        ('dwell_samples', 'i8'), ('extensions', 'i4'), ('flags', 'i4'),
                        ('spikes', 'f4', (N.POLYA_MAX_SPIKES, 4))])
assert POLYA_DTYPE.itemsize == C.sizeof(N.PolyaResult)


def polya_to_dict(rec, sampling_rate):
    """pb2_polya_result -> the dict of NanoporeRead.set_polya_tail (polya.py:116-121)."""
    # Capacity limits of the kernel's fixed-size buffers are never silent: the reference would
    # report the full record, so a read that exceeds them becomes an error for that read
    # (the drop-in turns it into `unknown_error` with this message).
    if int(rec['flags']) & 1:
        raise OverflowError('poly(A) recalibration met more anchor events than the kernel '
                            'buffers (polya_core.cuh POLYA_MAX_ANCHORS)')
    if not rec['found']:
        return None
    if int(rec['n_spikes']) > N.POLYA_MAX_SPIKES:
        raise OverflowError('poly(A) tail holds {} spikes, more than the {} the result record '
                            'carries'.format(int(rec['n_spikes']), N.POLYA_MAX_SPIKES))
    spikes = []
    for k in range(min(int(rec['n_spikes']), N.POLYA_MAX_SPIKES)):
        row = rec['spikes'][k]
        vals = [row[0]] + [row[j] for j in (1, 2, 3) if not np.isnan(row[j])]
        spikes.append(tuple(float(v) for v in vals))
    return {'begin': int(rec['begin']), 'end': int(rec['end']),
            'dwell_time': int(rec['dwell_samples']) / sampling_rate, 'spikes': spikes}


def polya_params_struct(cfg, stride):
    """config['polya_dwell'] -> pb2_polya_params, with the reference's rounding
    (PolyASignalAnalyzer.__init__, polya.py:39-48; float32 where the reference compares
    against float32 Series)."""
    f = np.float32
    loc, sd = cfg['polya_mean_dist']
    z = cfg['polya_mean_z_cutoff']
    ed = cfg['event_detection']
    rc = cfg['recalibrate_shifted_signal']
    if cfg['median_pre_filter'] != 7:
        raise ValueError('poly(A) kernel is built for median_pre_filter = 7')
    return N.PolyaParams(stride, cfg['refinement_expansion'], cfg['openend_expansion'] // stride,
                         cfg['maximum_openend_extension'], ed['window_length1'],
                         ed['window_length2'], ed['threshold1'], ed['threshold2'],
                         ed['peak_height'], f(loc - sd * z), f(loc + sd * z), f(loc),
                         f(cfg['polya_mean_trigger_recalibration'] * sd), f(sd * z),
                         f(cfg['polya_stdv_max']), cfg['polya_stdv_range'][0],
                         cfg['polya_stdv_range'][1], cfg['spike_tolerance'], cfg['spike_weight'],
                         rc['max_dist_from_adapter'], f(rc['min_length']), f(rc['max_stdv']))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _lstm_struct(layer, keep):
    W = np.ascontiguousarray(layer.kernel, np.float32)
    U = np.ascontiguousarray(layer.recurrent, np.float32)
    b = np.ascontiguousarray(layer.bias, np.float32)
    keep += [W, U, b]
    fp = C.POINTER(C.c_float)
    return N.LstmWeights(layer.in_dim, layer.units, layer.implementation,
                         W.ctypes.data_as(fp), U.ctypes.data_as(fp), b.ctypes.data_as(fp))


def _hmm_struct(tables):
    h = N.HmmParams()
    h.n_states = tables.n_states
    for s in range(N.MAX_STATES):
        h.n_comp[s] = int(tables.n_comp[s])
        h.log_start[s] = float(tables.log_start[s])
        for j in range(N.MAX_COMP):
            h.mu[s][j] = float(tables.mu[s, j])
            h.log_norm[s][j] = float(tables.lsp[s, j])
            h.inv_two_var[s][j] = float(tables.inv2s2[s, j])
            h.log_weight[s][j] = float(tables.logw[s, j])
    for i in range(N.MAX_STATES + 1):
        h.in_begin[i] = int(tables.in_begin[i])
    for k in range(N.MAX_EDGES):
        h.in_src[k] = int(tables.in_src[k])
        h.in_logp[k] = float(tables.in_logp[k])
    return h


class SignalEngine:
    """One native context configured from the reference's flat ``config`` dict
    (commandline.py:268-296) or from a preset dict (params.load_preset())."""

    def __init__(self, config, device=0, barcoding=None, barcoding_quality_filter=None):
        from scipy.stats import norm
        self.lib = N.load()
        self._keep = []
        handle = C.c_void_p()
        rc = self.lib.pb2_create(int(device), C.byref(handle))
        if rc != 0 or not handle.value:
            raise N.NativeError('poreplex_b200: cannot open CUDA device %d (rc=%d); this '
                                'path has no CPU fallback' % (device, rc))
        self.handle = handle
        self.device = int(device)

        sp = config['signal_processing']
        self.stride = int(sp['rough_signal_stride'])
        # --- scaler: SignalLoader.load_scaler_model (signal_loader.py:49-75)
        self.scaler_model = sm = P.load_scaler_model(sp['scaler_model'])
        idef, xf = sm.input_defs, sm.output_transform
        if idef['stride'] != self.stride:
            raise ValueError('scaler stride differs from rough_signal_stride')
        q = sp['scaler_qc_threshold']
        qs = norm.ppf([q, 1 - q], xf['scale_mean'], xf['scale_std'])
        qh = norm.ppf([q, 1 - q], xf['shift_mean'], xf['shift_std'])
        self.scaler_length = int(idef['length'])
        self.scaler_min_length = int(sp.get('scaler_min_length_override', idef['min_length']))
        fp = C.POINTER(C.c_float)
        dk = np.ascontiguousarray(sm.dense_kernel, np.float32)
        db = np.ascontiguousarray(sm.dense_bias, np.float32)
        self._keep += [dk, db]
        spar = N.ScalerParams(_lstm_struct(sm.l1, self._keep), _lstm_struct(sm.l2, self._keep),
                              dk.ctypes.data_as(fp), db.ctypes.data_as(fp), self.stride,
                              self.scaler_length, self.scaler_min_length,
                              xf['scale_std'], xf['scale_mean'], xf['shift_std'],
                              xf['shift_mean'], qs[0], qs[1], qh[0], qh[1])
        self._check(self.lib.pb2_set_scaler(self.handle, C.byref(spar)))

        # --- segmentation HMM: load_segmentation_model (worker_persistence.py:95-121)
        self.seg_tables = st = P.HmmTables(config['segmentation_model'])
        self.state_names = st.names
        self.scan_limit_pooled = int(config['segmentation']['segmentation_scan_limit']) // self.stride
        self.adapter_state = st.index_of('adapter')
        if self.adapter_state < 0:
            raise ValueError("segmentation model has no 'adapter' state")
        hs = _hmm_struct(st)
        self._check(self.lib.pb2_set_segmentation_hmm(self.handle, C.byref(hs),
                                                      self.scan_limit_pooled, self.adapter_state))

        # --- poly(A) analyzer: PolyASignalAnalyzer.__init__ (polya.py:39-48)
        self.polya_ready = False
        if 'polya_dwell' in config:
            pp = polya_params_struct(config['polya_dwell'], self.stride)
            self._check(self.lib.pb2_set_polya(self.handle, C.byref(pp), st.index_of('polya-tail')))
            self.polya_ready = True

        # --- unsplit-read (chimera) model + switches (worker_persistence.py:63,
        #     rna-r941.cfg:17-27,103-151)
        self.unsplit_ready = False
        if 'unsplit_read_detection_model' in config and 'unsplit_read_detection' in config:
            ut = P.HmmTables(config['unsplit_read_detection_model'])
            uc = config['unsplit_read_detection']
            self.unsplit_cfg = uc
            up = N.UnsplitParams(*(float(uc[k]) for k in (
                'window_size', 'window_step', 'strict_duration', 'strict_full_length',
                'strict_dna_length', 'loosen_full_length', 'loosen_dna_length',
                'basecount_quality_limit', 'subread_basecount_limit', 'subread_baseratio_limit')))
            uh = _hmm_struct(ut)
            self._check(self.lib.pb2_set_unsplit(self.handle, C.byref(uh), C.byref(up),
                                                 ut.index_of('adapter'), ut.index_of('leader-high'),
                                                 ut.index_of('leader-low')))
            self.unsplit_ready = True

        # --- demultiplexer: BarcodeDemultiplexer (barcoding.py:34-70)
        if barcoding is None:
            barcoding = bool(config.get('barcoding', True))
        self.barcoding = barcoding
        self.demux_model = None
        if barcoding:
            if barcoding_quality_filter is None:
                barcoding_quality_filter = config.get('barcoding_quality_filter', 18)
            dc = config['demultiplexing']
            self.demux_model = dm = P.load_demux_model(dc['demux_model'])
            calib = np.ascontiguousarray(dm.calibration, np.float64)
            if len(calib) - 1 < barcoding_quality_filter:       # barcoding.py:41-45
                raise ValueError('The current demultiplexer does not support calibrated score '
                                 'of {}. Consider lowering --barcoding-quality-filter value.'
                                 .format(barcoding_quality_filter))
            if dm.n_classes != dc['number_of_decoy_labels'] + dc['number_of_barcodes']:
                raise ValueError('demux model classes do not match the preset')
            dk2 = np.ascontiguousarray(dm.dense_kernel, np.float32)
            db2 = np.ascontiguousarray(dm.dense_bias, np.float32)
            self._keep += [dk2, db2, calib]
            self.trim_length = int(dc['signal_trim_length'])
            dpar = N.DemuxParams(_lstm_struct(dm.fwd, self._keep), _lstm_struct(dm.bwd, self._keep),
                                 _lstm_struct(dm.l2, self._keep), dk2.ctypes.data_as(fp),
                                 db2.ctypes.data_as(fp), dm.n_classes,
                                 int(dc['number_of_decoy_labels']),
                                 int(dc['minimum_dna_length']), int(dc['maximum_dna_length']),
                                 self.trim_length, -1000.0, len(calib),
                                 calib.ctypes.data_as(C.POINTER(C.c_double)),
                                 float(calib[barcoding_quality_filter]))
            self._check(self.lib.pb2_set_demux(self.handle, C.byref(dpar)))

    # ------------------------------------------------------------------ util
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.pb2_last_error(self.handle)
            raise N.NativeError('poreplex_b200 native error %d: %s'
                                % (rc, msg.decode() if msg else '?'))

    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle.value:
            self.lib.pb2_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self):
        return int(self.lib.pb2_kernel_launches(self.handle))

    def set_exact_division(self, on=True):
        """Verification mode: IEEE division in the LSTM kernels (same outputs, slower)."""
        self._check(self.lib.pb2_set_exact_division(self.handle, 1 if on else 0))

    def set_fast_lstm(self, on=True, demux_margin_delta=0.0, demux_probe_gain=0.0):
        """True / 'fast': tensor-core scaler and classifier with margin tests + exact re-run (the
        library default); 'strict': exact scaler, segmentation and windows for every read (all
        float outputs but the class probabilities bit-exact), tensor-core classifier with
        margin test + exact re-run; False / 'exact': exact f32 kernels only."""
        mode = {'fast': 1, 'strict': 2, 'exact': 0}.get(on, 1 if on else 0) if isinstance(on, str) \
            else (2 if on == 2 else (1 if on else 0))
        self.lstm_mode = ('exact', 'fast', 'strict')[mode]
        self._check(self.lib.pb2_set_fast_lstm(self.handle, mode,
                                               float(demux_margin_delta), float(demux_probe_gain)))

    def recheck_stats(self):
        """(windows re-run exactly by the last demultiplexer launch, tensor-core time-outs)"""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pb2_recheck_stats(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_audit_fraction(self, fraction):
        """Also re-run this fraction of the guard-passing reads exactly and count disagreements."""
        self._check(self.lib.pb2_set_audit_fraction(self.handle, float(fraction)))

    def audit_stats(self):
        """(reads audited, audited reads whose exact integer outputs differed) since the last call."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pb2_audit_stats(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def probe2_rows(self):
        """Windows of the last classifier launch that needed the second sensitivity probe."""
        a = C.c_int64(0)
        self._check(self.lib.pb2_probe2_rows(self.handle, C.byref(a)))
        return int(a.value)

    def rerun_causes(self):
        """{cause: reads} behind the exact re-runs of the last whole-path call."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pb2_rerun_causes(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {'qc_edge': int(a.value), 'segmentation': int(b.value), 'barcode_call': int(c.value)}

    def profile_enable(self, on=True):
        self._check(self.lib.pb2_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        """{kernel name: (total_ms, launches)} since the last read (synchronises)."""
        k = self.lib.pb2_profile_kernel_count()
        ms = (C.c_double * k)()
        cnt = (C.c_int64 * k)()
        self._check(self.lib.pb2_profile_read(self.handle, ms, cnt, k))
        return {self.lib.pb2_profile_kernel_name(i).decode(): (ms[i], int(cnt[i]))
                for i in range(k) if cnt[i]}

    def profile_timeline(self, capacity=4096):
        """[(kernel name, start ms, end ms)] of every launch since the last read, in launch order,
        times relative to the first launch (synchronises; consumes the records)."""
        ids = (C.c_int32 * capacity)()
        a = (C.c_double * capacity)()
        b = (C.c_double * capacity)()
        n = C.c_int64(0)
        self._check(self.lib.pb2_profile_timeline(self.handle, ids, a, b, capacity, C.byref(n)))
        return [(self.lib.pb2_profile_kernel_name(ids[i]).decode(), a[i], b[i]) for i in range(n.value)]

    def pooled_offsets(self, raw_offsets):
        """Element offset of each read inside a pooled buffer (same rule as the kernels)."""
        return (np.asarray(raw_offsets, np.int64) + self.stride - 1) // self.stride

    # -------------------------------------------------------- host-buffer API
    @staticmethod
    def pack_reads(signals):
        """Concatenate int16 signals with every read 16-byte aligned.
        Returns (raw, offsets, lengths)."""
        lengths = np.array([len(s) for s in signals], np.int64)
        padded = (lengths + 7) // 8 * 8
        offsets = np.zeros(len(signals), np.int64)
        if len(signals):
            offsets[1:] = np.cumsum(padded)[:-1]
        raw = np.zeros(int(padded.sum()) + 8, np.int16)
        for s, o, n in zip(signals, offsets, lengths):
            raw[o:o + n] = s
        return raw, offsets, lengths

    def alloc_host_results(self, n, pinned=True):
        """Result buffers for ``analyze_host(..., out=...)``: reusable across calls and, when
        pinned, written by direct DMA instead of through the driver's staging buffer."""
        import torch
        def mk(shape, dtype):
            return torch.zeros(shape, dtype=dtype, pin_memory=pinned).numpy()
        return {
            'status': mk(n, torch.int32), 'label': mk(n, torch.int32),
            'scale_shift': mk((n, 2), torch.float32),
            'segments': mk((n, N.MAX_STATES, 2), torch.int32),
            'barcode': mk(n, torch.int32), 'barcode_guess': mk(n, torch.int32),
            'barcode_score': mk(n, torch.int32),
            'class_probs': mk((n, N.MAX_CLASSES), torch.float32),
            'counts': mk((N.N_LABEL, N.N_BARCODE_SLOTS, N.N_STATUS), torch.int64),
        }

    def analyze_host(self, raw, offsets, lengths, rng, digitisation, offset, barcoding=None,
                     keep_pooled=False, polya=False, exact_scaler=False, out=None, packed=None):
        """SignalAnalyzer.process stages A-D over HOST numpy buffers (H2D, kernels, D2H).

        Returns a dict of numpy arrays: status, label, scale_shift [n,2], segments
        [n,8,2] (baked state order = ``state_names``), barcode, barcode_guess,
        barcode_score, class_probs [n,8], counts [4,5,11], and ``pooled`` when asked.

        ``packed`` = (uint8 buffer, packed_offsets[n + 1]) from ``fast5_loader.svb16_encode``: the
        compressed upload form of the same batch (streamvbyte-16 bodies of VBZ chunks; ``raw`` may
        then be None, ``offsets`` / ``lengths`` still describe the int16 layout the streams decode
        into on the device).
        """
        if barcoding is None:
            barcoding = self.barcoding
        if barcoding and self.demux_model is None:
            raise ValueError('engine was created without the demultiplexer')
        offsets = np.ascontiguousarray(offsets, np.int64)
        lengths = np.ascontiguousarray(lengths, np.int64)
        if packed is not None:
            pk = np.ascontiguousarray(packed[0], np.uint8)
            pko = np.ascontiguousarray(packed[1], np.int64)
            if len(pko) != len(lengths) + 1 or np.any(pko % 16):
                raise ValueError('packed_offsets must hold n + 1 multiples of 16')
            n_raw_total = int((offsets + (lengths + 7) // 8 * 8).max()) + 8 if len(lengths) else 0
            raw = None if raw is None else np.ascontiguousarray(raw, np.int16).reshape(-1)
        else:
            raw = np.ascontiguousarray(raw, np.int16).reshape(-1)
            n_raw_total = raw.size
        rng = np.ascontiguousarray(rng, np.float64)
        digitisation = np.ascontiguousarray(digitisation, np.float64)
        offset = np.ascontiguousarray(offset, np.float64)
        n = len(lengths)
        if np.any(offsets % 8):
            raise ValueError('raw_offsets must be multiples of 8 samples (see pack_reads)')
        if n and int((offsets + lengths).max()) > n_raw_total:
            raise ValueError('read extends past the raw buffer')
        if out is not None:
            # caller-owned buffers (alloc_host_results): same keys, shapes and dtypes
            if out['status'].shape != (n,) or out['segments'].shape != (n, N.MAX_STATES, 2):
                raise ValueError('out buffers do not match the batch size')
            out = dict(out)
        else:
            out = {
                'status': np.empty(n, np.int32), 'label': np.empty(n, np.int32),
                'scale_shift': np.zeros((n, 2), np.float32),
                'segments': np.empty((n, N.MAX_STATES, 2), np.int32),
                'barcode': np.empty(n, np.int32), 'barcode_guess': np.empty(n, np.int32),
                'barcode_score': np.empty(n, np.int32),
                'class_probs': np.zeros((n, N.MAX_CLASSES), np.float32),
                'counts': np.zeros((N.N_LABEL, N.N_BARCODE_SLOTS, N.N_STATUS), np.int64),
            }
        if keep_pooled:
            out['pooled'] = np.zeros(n_raw_total // self.stride + 2, np.float32)
        if polya:
            if not self.polya_ready:
                raise ValueError("config has no 'polya_dwell' section")
            if 'polya' not in out or out['polya'].shape != (n,):
                out['polya'] = np.zeros(n, POLYA_DTYPE)
        b = N.Batch(n, n_raw_total, int(lengths.max()) if n else 0,
                    _np_ptr(raw) if raw is not None else None, _np_ptr(offsets),
                    _np_ptr(lengths), _np_ptr(rng), _np_ptr(digitisation), _np_ptr(offset),
                    _np_ptr(pk) if packed is not None else None,
                    _np_ptr(pko) if packed is not None else None)
        r = N.Results(_np_ptr(out['status']), _np_ptr(out['label']), _np_ptr(out['scale_shift']),
                      _np_ptr(out['segments']), _np_ptr(out['barcode']),
                      _np_ptr(out['barcode_guess']), _np_ptr(out['barcode_score']),
                      _np_ptr(out['class_probs']),
                      _np_ptr(out['pooled']) if keep_pooled else None, _np_ptr(out['counts']),
                      _np_ptr(out['polya']) if polya else None)
        flags = (N.FLAG_BARCODING if barcoding else 0) | (N.FLAG_KEEP_POOLED if keep_pooled else 0) \
            | (N.FLAG_POLYA if polya else 0) | (N.FLAG_EXACT_SCALER if exact_scaler else 0)
        self._check(self.lib.pb2_analyze_host(self.handle, C.byref(b), C.byref(r), flags))
        return out

    def detect_unsplit_host(self, tables, sampling_rate, scale_shift, status, segments,
                            batch=None):
        """SignalAnalysis.detect_unsplit_read for a batch (host buffers).

        ``tables``: per read either None or a dict with 'start', 'move', 'p_model_state' and
        either 'mean' (albacore-style tables) or 'first_sample' + 'block_stride' (guppy Move
        tables: the means are then derived on the device from ``batch`` = (raw, offsets,
        lengths, range, digitisation, offset) of the same reads).  Returns int32 flags:
        1 unsplit, 0 not, < 0 internal error for that read."""
        if not self.unsplit_ready:
            raise ValueError('config has no unsplit-read detection model')
        # guppy Move tables as Fast5Source hands them over (move + FASTQ strings): the event
        # columns are derived on the device first
        need = [i for i, t in enumerate(tables)
                if t is not None and t.get('guppy_move') and 'start' not in t]
        if need:
            if batch is None:
                raise ValueError('guppy Move tables need the batch of the same reads')
            raw, roff, rlen, rng, dig, off = (np.ascontiguousarray(a) for a in batch)
            sub = (raw, roff[need], rlen[need], rng[need], dig[need], off[need])
            strides = {int(tables[i]['block_stride']) for i in need}
            if len(strides) != 1:
                raise ValueError('reads of one batch must share block_stride')
            derived, err = self.derive_event_tables_host(
                sub, [tables[i]['move'] for i in need], [int(tables[i]['first_sample']) for i in need],
                strides.pop(), sequences=[tables[i]['sequence'] for i in need],
                qstrings=[tables[i]['qstring'] for i in need],
                columns=('mean', 'start', 'p_model_state'))
            if err.any():
                raise ValueError('event-table derivation failed for reads %r' % (np.nonzero(err)[0].tolist(),))
            tables = list(tables)
            for k, i in enumerate(need):
                tables[i] = dict(tables[i], **derived[k])
        has = [t is not None and len(t['start']) > 0 for t in tables]
        with_mean = [h and 'mean' in t for h, t in zip(has, tables)]
        derived = [h and 'mean' not in t for h, t in zip(has, tables)]
        flag = np.zeros(len(tables), np.int32)
        for group in (with_mean, derived):
            if any(group):
                sub = [t if g else None for t, g in zip(tables, group)]
                f = self._detect_unsplit_group(sub, sampling_rate, scale_shift, status, segments,
                                               batch if group is derived else None)
                flag = np.where(group, f, flag).astype(np.int32)
        return flag

    def _detect_unsplit_group(self, tables, sampling_rate, scale_shift, status, segments, batch):
        n = len(tables)
        counts = np.array([0 if t is None else len(t['start']) for t in tables], np.int64)
        offs = np.zeros(n + 1, np.int64)
        offs[1:] = np.cumsum(counts)
        total = int(offs[-1])
        cat = lambda key, dt: (np.concatenate([np.asarray(t[key]).astype(dt) for t in tables
                                               if t is not None and len(t['start'])])
                               if total else np.zeros(0, dt))
        start = cat('start', np.int64)
        move, pstate = cat('move', np.int32), cat('p_model_state', np.float64)
        derive = batch is not None
        mean = None if derive else cat('mean', np.float32)
        first = np.array([0 if t is None else int(t.get('first_sample', 0)) for t in tables], np.int64)
        strides = {int(t['block_stride']) for t in tables if t is not None and 'block_stride' in t}
        if derive and len(strides) != 1:
            raise ValueError('reads of one batch must share block_stride')
        rate = np.ascontiguousarray(sampling_rate, np.float64)
        scale_shift = np.ascontiguousarray(scale_shift, np.float32)
        status = np.ascontiguousarray(status, np.int32)
        segments = np.ascontiguousarray(segments, np.int32)
        # windows needed: range(payload_start, last_end, int(window_step * rate))
        maxw = 1
        ia = self.adapter_state
        for i, t in enumerate(tables):
            if t is None or not len(t['start']):
                continue
            step = int(self.unsplit_cfg['window_step'] * rate[i])
            payload = (int(segments[i, ia, 1]) + 1) * self.stride
            span = int(t['start'][-1]) + 1 - payload
            if step > 0 and span > 0:
                maxw = max(maxw, -(-span // step))
        ev = N.EventTables(total, _np_ptr(offs), _np_ptr(start),
                           None if derive else _np_ptr(mean), _np_ptr(move), _np_ptr(pstate),
                           _np_ptr(rate), _np_ptr(first), strides.pop() if derive else 0)
        bptr = None
        if derive:
            raw, roff, rlen, rng, dig, off = (np.ascontiguousarray(a) for a in batch)
            b = N.Batch(n, raw.size, int(rlen.max()) if n else 0, _np_ptr(raw), _np_ptr(roff),
                        _np_ptr(rlen), _np_ptr(rng), _np_ptr(dig), _np_ptr(off))
            bptr = C.byref(b)
        flag = np.zeros(n, np.int32)
        self._check(self.lib.pb2_detect_unsplit_host(
            self.handle, bptr, C.byref(ev), n, _np_ptr(scale_shift), _np_ptr(status),
            _np_ptr(segments), int(maxw), _np_ptr(flag)))
        return flag

    @staticmethod
    def quality_table():
        """qual[byte] = 1 - 10 ** -((byte - 33) / 10) with the reference's own numpy expression
        (fast5_file.py:188) for every byte value; the device only gathers from it."""
        with np.errstate(over='ignore'):
            return np.ascontiguousarray(
                1 - 10 ** -((np.arange(256, dtype=np.uint8) - 33) / 10), np.float64)

    def derive_event_tables_host(self, batch, moves, first_sample, block_stride, sequences=None,
                                 qstrings=None, scale_shift=None,
                                 columns=('mean', 'stdv', 'start', 'end', 'length', 'pos',
                                          'p_model_state', 'model_state', 'scaled_mean')):
        """Guppy ``Move`` tables -> event-table columns for a batch of reads, on the device
        (Fast5Reader.construct_events_from_moves / convert_events_guppy, fast5_file.py:183-230,
        plus the derived columns of SignalAnalysis.load_events, signal_analyzer.py:311-326).

        ``batch`` = (raw, offsets, lengths, range, digitisation, offset) host arrays; ``moves``:
        per read a uint8 / int array; ``sequences`` / ``qstrings``: per read ``str`` (FASTQ
        lines 2 and 4).  Returns (list of per-read dicts of numpy columns, int32 error codes):
        1 = unknown k-mer size, 2 = events / raw strides mismatch (the reference raises)."""
        n = len(moves)
        counts = np.array([len(m) for m in moves], np.int64)
        ev_off = np.zeros(n + 1, np.int64)
        ev_off[1:] = np.cumsum(counts)
        E = int(ev_off[-1])
        move = (np.concatenate([np.asarray(m).astype(np.int32) for m in moves])
                if E else np.zeros(0, np.int32))
        first = np.ascontiguousarray(first_sample, np.int64)
        raw, roff, rlen, rng, dig, off = (np.ascontiguousarray(a) for a in batch)
        b = N.Batch(n, raw.size, int(rlen.max()) if n else 0, _np_ptr(raw), _np_ptr(roff),
                    _np_ptr(rlen), _np_ptr(rng), _np_ptr(dig), _np_ptr(off))
        ev = N.EventTables(E, _np_ptr(ev_off), None, None, _np_ptr(move), None, None,
                           _np_ptr(first), int(block_stride))
        bc = None
        keep = []
        if sequences is not None:
            seq_off = np.zeros(n + 1, np.int64)
            seq_off[1:] = np.cumsum([len(s) for s in sequences])
            seq = np.frombuffer(''.join(sequences).encode('ascii'), np.uint8).copy()
            qs = np.frombuffer(''.join(qstrings).encode('ascii'), np.uint8).copy()
            if len(qs) != len(seq):
                raise ValueError('sequence and quality strings differ in length')
            qt = self.quality_table()
            keep += [seq_off, seq, qs, qt]
            bc = N.Basecalls(_np_ptr(seq), _np_ptr(qs), _np_ptr(seq_off), _np_ptr(qt))
        dt = {'mean': np.float32, 'stdv': np.float32, 'scaled_mean': np.float32, 'start': np.int64,
              'end': np.int64, 'length': np.int64, 'pos': np.int64, 'p_model_state': np.float64}
        cols = {}
        for c in columns:
            if c in ('pos', 'p_model_state', 'model_state') and bc is None and c != 'pos':
                continue
            if c == 'scaled_mean' and scale_shift is None:
                continue
            cols[c] = np.zeros((E, 5), np.uint8) if c == 'model_state' else np.zeros(E, dt[c])
        err = np.zeros(n, np.int32)
        out = N.EventColumns(*[(_np_ptr(cols[k]) if k in cols else None) for k in
                               ('mean', 'stdv', 'scaled_mean', 'start', 'end', 'length', 'pos',
                                'p_model_state', 'model_state')], _np_ptr(err))
        ss = None
        if scale_shift is not None:
            ss = np.ascontiguousarray(scale_shift, np.float32)
        self._check(self.lib.pb2_derive_event_tables_host(
            self.handle, C.byref(b), C.byref(ev), C.byref(bc) if bc is not None else None,
            _np_ptr(ss) if ss is not None else None, C.byref(out)))
        if 'model_state' in cols:
            cols['model_state'] = cols['model_state'].reshape(E, 5).view('S5').reshape(E)
        tables = [{k: v[ev_off[i]:ev_off[i + 1]] for k, v in cols.items()} for i in range(n)]
        return tables, err

    # ----------------------------------------------------- device-resident API
    def _batch_from_tensors(self, raw, offsets, lengths, rng, digitisation, offset,
                            max_raw_length=0):
        return N.Batch(int(lengths.numel()), int(raw.numel()), int(max_raw_length),
                       raw.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), rng.data_ptr(),
                       digitisation.data_ptr(), offset.data_ptr())

    def alloc_results(self, n, n_raw_total=0, keep_pooled=False, polya=False):
        import torch
        dev = torch.device('cuda', self.device)
        out = {
            'status': torch.empty(n, dtype=torch.int32, device=dev),
            'label': torch.empty(n, dtype=torch.int32, device=dev),
            'scale_shift': torch.zeros((n, 2), dtype=torch.float32, device=dev),
            'segments': torch.empty((n, N.MAX_STATES, 2), dtype=torch.int32, device=dev),
            'barcode': torch.full((n,), -1, dtype=torch.int32, device=dev),
            'barcode_guess': torch.full((n,), -1, dtype=torch.int32, device=dev),
            'barcode_score': torch.full((n,), -1, dtype=torch.int32, device=dev),
            'class_probs': torch.zeros((n, N.MAX_CLASSES), dtype=torch.float32, device=dev),
            'counts': torch.zeros((N.N_LABEL, N.N_BARCODE_SLOTS, N.N_STATUS),
                                  dtype=torch.int64, device=dev),
        }
        if keep_pooled:
            out['pooled'] = torch.zeros(n_raw_total // self.stride + 2, dtype=torch.float32,
                                        device=dev)
        if polya:
            out['polya'] = torch.zeros((n, POLYA_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        return out

    def analyze_device(self, raw, offsets, lengths, rng, digitisation, offset, out=None,
                       barcoding=None, keep_pooled=False, max_raw_length=0, stream=None,
                       polya=False, exact_scaler=False):
        """Same path over tensors already resident in HBM; enqueued on ``stream`` (default:
        torch's current stream).  In the ``exact`` and ``strict`` modes the call only enqueues; in
        the ``fast`` mode it waits once on that stream, for the number of reads the guards sent
        to the exact re-run (the size of the sub-batch it then enqueues), so the kernels before
        that point have finished when it returns -- the results are complete only after the
        caller synchronises the stream, as usual."""
        import torch
        if barcoding is None:
            barcoding = self.barcoding
        n = int(lengths.numel())
        if out is None:
            out = self.alloc_results(n, int(raw.numel()), keep_pooled, polya)
        b = self._batch_from_tensors(raw, offsets, lengths, rng, digitisation, offset,
                                     max_raw_length)
        r = N.Results(out['status'].data_ptr(), out['label'].data_ptr(),
                      out['scale_shift'].data_ptr(), out['segments'].data_ptr(),
                      out['barcode'].data_ptr(), out['barcode_guess'].data_ptr(),
                      out['barcode_score'].data_ptr(), out['class_probs'].data_ptr(),
                      out['pooled'].data_ptr() if keep_pooled else None,
                      out['counts'].data_ptr(), out['polya'].data_ptr() if polya else None)
        flags = (N.FLAG_BARCODING if barcoding else 0) | (N.FLAG_KEEP_POOLED if keep_pooled else 0) \
            | (N.FLAG_POLYA if polya else 0) | (N.FLAG_EXACT_SCALER if exact_scaler else 0)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self.lib.pb2_analyze_device(self.handle, C.byref(b), C.byref(r), flags,
                                                C.c_void_p(st.cuda_stream)))
        return out

    def detect_unsplit_device(self, batch, ev_offsets, start, move, p_model_state, sampling_rate,
                              first_sample, block_stride, scale_shift, status, segments,
                              max_windows, mean=None, stream=None):
        """pb2_detect_unsplit over tensors resident in HBM.  ``batch`` = (raw, offsets, lengths,
        range, digitisation, offset) tensors; ``mean=None`` derives the event means on the
        device.  Returns an int32 flag tensor."""
        import torch
        n = int(sampling_rate.numel())
        ev = N.EventTables(int(start.numel()), ev_offsets.data_ptr(), start.data_ptr(),
                           mean.data_ptr() if mean is not None else None, move.data_ptr(),
                           p_model_state.data_ptr(), sampling_rate.data_ptr(),
                           first_sample.data_ptr(), int(block_stride))
        b = self._batch_from_tensors(*batch)
        flag = torch.zeros(n, dtype=torch.int32, device=status.device)
        self._check(self.lib.pb2_detect_unsplit(
            self.handle, C.byref(b), C.byref(ev), n, scale_shift.data_ptr(), status.data_ptr(),
            segments.data_ptr(), int(max_windows), flag.data_ptr(), self._stream(stream)))
        return flag

    # ------------------------------------------------ single stages (tensors)
    def _stream(self, stream):
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(st.cuda_stream)

    def pool_signal(self, raw, offsets, lengths, rng, digitisation, offset, max_raw_length=0,
                    stream=None):
        import torch
        pooled = torch.zeros(int(raw.numel()) // self.stride + 2, dtype=torch.float32,
                             device=raw.device)
        b = self._batch_from_tensors(raw, offsets, lengths, rng, digitisation, offset,
                                     max_raw_length)
        self._check(self.lib.pb2_pool_signal(self.handle, C.byref(b), pooled.data_ptr(),
                                             self._stream(stream)))
        return pooled

    def fit_scalers(self, raw, offsets, lengths, rng, digitisation, offset, pooled, stream=None):
        import torch
        n = int(lengths.numel())
        status = torch.empty(n, dtype=torch.int32, device=raw.device)
        ss = torch.zeros((n, 2), dtype=torch.float32, device=raw.device)
        z = torch.zeros((n, 2), dtype=torch.float32, device=raw.device)
        b = self._batch_from_tensors(raw, offsets, lengths, rng, digitisation, offset)
        self._check(self.lib.pb2_fit_scalers(self.handle, C.byref(b), pooled.data_ptr(),
                                             status.data_ptr(), ss.data_ptr(), z.data_ptr(),
                                             self._stream(stream)))
        return status, ss, z

    def detect_segments(self, raw, offsets, lengths, rng, digitisation, offset, pooled,
                        scale_shift, status, keep_pooled=False, max_raw_length=0, stream=None):
        import torch
        n = int(lengths.numel())
        seg = torch.empty((n, N.MAX_STATES, 2), dtype=torch.int32, device=raw.device)
        scaled = torch.zeros_like(pooled) if keep_pooled else None
        b = self._batch_from_tensors(raw, offsets, lengths, rng, digitisation, offset,
                                     max_raw_length)
        self._check(self.lib.pb2_detect_segments(
            self.handle, C.byref(b), pooled.data_ptr(), scale_shift.data_ptr(),
            status.data_ptr(), seg.data_ptr(), scaled.data_ptr() if keep_pooled else None,
            self._stream(stream)))
        return seg, scaled

    def viterbi_paths(self, x, lengths, stream=None):
        import torch
        n, ld = x.shape
        path = torch.full((n, ld), -1, dtype=torch.int32, device=x.device)
        logp = torch.empty(n, dtype=torch.float64, device=x.device)
        self._check(self.lib.pb2_viterbi_paths(self.handle, 0, x.data_ptr(), lengths.data_ptr(),
                                               n, ld, path.data_ptr(), logp.data_ptr(),
                                               self._stream(stream)))
        return path, logp

    def barcode_windows(self, raw, offsets, lengths, rng, digitisation, offset, pooled,
                        scale_shift, status, segments, stream=None):
        import torch
        n = int(lengths.numel())
        win = torch.zeros((n, self.trim_length), dtype=torch.float32, device=raw.device)
        pushed = torch.zeros(n, dtype=torch.int32, device=raw.device)
        b = self._batch_from_tensors(raw, offsets, lengths, rng, digitisation, offset)
        self._check(self.lib.pb2_barcode_windows(
            self.handle, C.byref(b), pooled.data_ptr(), scale_shift.data_ptr(),
            status.data_ptr(), segments.data_ptr(), win.data_ptr(), pushed.data_ptr(),
            self._stream(stream)))
        return win, pushed

    def demux_predict(self, windows, pushed=None, stream=None):
        import torch
        n = windows.shape[0]
        dev = windows.device
        probs = torch.zeros((n, N.MAX_CLASSES), dtype=torch.float32, device=dev)
        bc = torch.full((n,), -1, dtype=torch.int32, device=dev)
        guess = torch.full((n,), -1, dtype=torch.int32, device=dev)
        score = torch.full((n,), -1, dtype=torch.int32, device=dev)
        self._check(self.lib.pb2_demux_predict(
            self.handle, windows.data_ptr(), pushed.data_ptr() if pushed is not None else None,
            n, probs.data_ptr(), bc.data_ptr(), guess.data_ptr(), score.data_ptr(),
            self._stream(stream)))
        return probs, bc, guess, score

    def demux_predict_tc(self, windows, stream=None):
        """Verification: tensor-core demultiplexer without the exact re-run ->
        (probs, logits, barcode, guess, score, unsafe, sensitivity)."""
        import torch
        n = windows.shape[0]
        dev = windows.device
        probs = torch.zeros((n, N.MAX_CLASSES), dtype=torch.float32, device=dev)
        logits = torch.zeros((n, N.MAX_CLASSES), dtype=torch.float32, device=dev)
        bc = torch.full((n,), -1, dtype=torch.int32, device=dev)
        guess = torch.full((n,), -1, dtype=torch.int32, device=dev)
        score = torch.full((n,), -1, dtype=torch.int32, device=dev)
        unsafe = torch.zeros((n,), dtype=torch.int32, device=dev)
        sens = torch.zeros((n,), dtype=torch.float32, device=dev)
        self._check(self.lib.pb2_demux_predict_tc(
            self.handle, windows.data_ptr(), n, probs.data_ptr(), logits.data_ptr(),
            bc.data_ptr(), guess.data_ptr(), score.data_ptr(), unsafe.data_ptr(),
            sens.data_ptr(), self._stream(stream)))
        return probs, logits, bc, guess, score, unsafe, sens

    def debug_demux_l1(self, windows, stream=None):
        """Verification: exact layer-1 outputs [n][T][2*units] (forward | backward)."""
        import torch
        n = windows.shape[0]
        out = torch.zeros((n, self.trim_length, 96), dtype=torch.float32, device=windows.device)
        self._check(self.lib.pb2_debug_demux_l1(self.handle, windows.data_ptr(), n, out.data_ptr(),
                                                self._stream(stream)))
        return out

    def scaler_predict(self, heads, stream=None):
        import torch
        n = heads.shape[0]
        z = torch.zeros((n, 2), dtype=torch.float32, device=heads.device)
        self._check(self.lib.pb2_scaler_predict(self.handle, heads.data_ptr(), n, z.data_ptr(),
                                                self._stream(stream)))
        return z

    def svb16_decode(self, packed, packed_offsets, raw_offsets, raw_lengths, n_raw_total, stream=None):
        """Device half of the VBZ decoder over tensors in HBM (pb2_svb16_decode): streamvbyte-16
        bodies -> int16 samples in the layout raw_offsets / raw_lengths describe.  Returns
        (raw int16 tensor, error flag tensor)."""
        import torch
        raw = torch.zeros(int(n_raw_total), dtype=torch.int16, device=packed.device)
        err = torch.zeros(1, dtype=torch.int32, device=packed.device)
        self._check(self.lib.pb2_svb16_decode(
            self.handle, packed.data_ptr(), packed_offsets.data_ptr(), raw_offsets.data_ptr(),
            raw_lengths.data_ptr(), int(raw_lengths.numel()), raw.data_ptr(), err.data_ptr(),
            self._stream(stream)))
        return raw, err

    def count_results(self, status, label, barcode, stream=None):
        import torch
        counts = torch.zeros((N.N_LABEL, N.N_BARCODE_SLOTS, N.N_STATUS), dtype=torch.int64,
                             device=status.device)
        self._check(self.lib.pb2_count_results(
            self.handle, status.data_ptr(), label.data_ptr(),
            barcode.data_ptr() if barcode is not None else None, int(status.numel()),
            counts.data_ptr(), self._stream(stream)))
        return counts


_engines = {}          # device -> (config digest, SignalEngine)

# every config entry SignalEngine.__init__ (and the analyzer's engine set-up) reads
_ENGINE_KEYS = ('signal_processing', 'segmentation', 'segmentation_model', 'polya_dwell',
                'unsplit_read_detection_model', 'unsplit_read_detection', 'demultiplexing',
                'barcoding', 'barcoding_quality_filter', 'fast_lstm')


def config_digest(config):
    """Content hash of the parameters an engine is built from.  ``pipeline.py:204`` pickles
    ``config`` anew for every batch, so object identity says nothing: two configs with the
    same content must map to the same engine, and a changed parameter must not."""
    import hashlib
    import json
    doc = {k: config.get(k) for k in _ENGINE_KEYS}
    doc['barcoding'] = bool(doc['barcoding'])
    blob = json.dumps(doc, sort_keys=True, default=repr)
    return hashlib.sha1(blob.encode()).hexdigest()


def get_engine(config, device=0):
    """Process-lifetime engine, the analogue of the reference's
    ``sys.modules['__poreplex_persistence']`` singleton (worker_persistence.py:46-58):
    one engine per (process, device), rebuilt -- and the superseded one closed, its device
    arenas freed -- only when the content of the configuration changes."""
    digest = config_digest(config)
    held = _engines.get(device)
    if held is not None and held[0] == digest:
        return held[1]
    if held is not None:
        held[1].close()
        del _engines[device]
    eng = SignalEngine(config, device=device)
    # The drop-in defaults to the exact kernels: every float and integer output is then the
    # oracle's bit for bit.  config['fast_lstm'] opts into the tensor-core path, whose
    # integer outputs are guarded and whose floats are approximate.
    eng.set_fast_lstm(config.get('fast_lstm') or False)
    _engines[device] = (digest, eng)
    return eng


def close_engines():
    """Release every cached engine (tests; worker shutdown)."""
    for dev in list(_engines):
        _engines.pop(dev)[1].close()
