"""ctypes binding of libporeplex_b200.so (include/poreplex_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device can be opened,
``load()`` / ``Context()`` raise.  ``build()`` compiles the library in-tree with nvcc
for sm_100a (works without a GPU).
"""
import ctypes as C
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# POREPLEX_B200_LIB: load another build of the same ABI (kernel tuning experiments)
LIB_PATH = os.environ.get('POREPLEX_B200_LIB') or os.path.join(HERE, 'libporeplex_b200.so')
HEADER = os.path.join(HERE, '..', 'include', 'poreplex_b200.h')

MAX_STATES, MAX_COMP, MAX_EDGES, MAX_CLASSES, MAX_CALIB = 8, 4, 64, 8, 64
N_LABEL, N_BARCODE_SLOTS, N_STATUS = 4, 5, 11
FLAG_BARCODING, FLAG_KEEP_POOLED, FLAG_POLYA, FLAG_EXACT_SCALER = 1, 2, 4, 8
POLYA_MAX_SPIKES = 48
LABEL_NAMES = ['pass', 'fail', 'artifact', None]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-fmad=false',
              '-std=c++17', '-Xcompiler', '-fPIC', '-shared']

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class LstmWeights(C.Structure):
    _fields_ = [('in_dim', C.c_int32), ('units', C.c_int32), ('implementation', C.c_int32),
                ('kernel', _fp), ('recurrent', _fp), ('bias', _fp)]


class DetectorParams(C.Structure):
    _fields_ = [('window_length1', C.c_int64), ('window_length2', C.c_int64),
                ('threshold1', C.c_float), ('threshold2', C.c_float), ('peak_height', C.c_float)]


class ScalerParams(C.Structure):
    _fields_ = [('l1', LstmWeights), ('l2', LstmWeights), ('dense_kernel', _fp),
                ('dense_bias', _fp), ('stride', C.c_int32), ('length', C.c_int32),
                ('min_length', C.c_int32),
                ('scale_std', C.c_double), ('scale_mean', C.c_double),
                ('shift_std', C.c_double), ('shift_mean', C.c_double),
                ('qc_scale_lo', C.c_double), ('qc_scale_hi', C.c_double),
                ('qc_shift_lo', C.c_double), ('qc_shift_hi', C.c_double)]


class HmmParams(C.Structure):
    _fields_ = [('n_states', C.c_int32),
                ('n_comp', C.c_int32 * MAX_STATES),
                ('mu', (C.c_double * MAX_COMP) * MAX_STATES),
                ('log_norm', (C.c_double * MAX_COMP) * MAX_STATES),
                ('inv_two_var', (C.c_double * MAX_COMP) * MAX_STATES),
                ('log_weight', (C.c_double * MAX_COMP) * MAX_STATES),
                ('log_start', C.c_double * MAX_STATES),
                ('in_begin', C.c_int32 * (MAX_STATES + 1)),
                ('in_src', C.c_int32 * MAX_EDGES),
                ('in_logp', C.c_double * MAX_EDGES)]


class DemuxParams(C.Structure):
    _fields_ = [('fwd', LstmWeights), ('bwd', LstmWeights), ('l2', LstmWeights),
                ('dense_kernel', _fp), ('dense_bias', _fp), ('n_classes', C.c_int32),
                ('n_decoy', C.c_int32), ('min_length', C.c_int32), ('max_length', C.c_int32),
                ('trim_length', C.c_int32), ('pad_value', C.c_float),
                ('n_calibration', C.c_int32), ('calibration', _dp),
                ('score_threshold', C.c_double)]


class PolyaParams(C.Structure):
    _fields_ = [('stride', C.c_int32), ('refinement_expansion', C.c_int32),
                ('openend_unit', C.c_int32), ('max_extension', C.c_int32),
                ('window_length1', C.c_int32), ('window_length2', C.c_int32),
                ('threshold1', C.c_float), ('threshold2', C.c_float), ('peak_height', C.c_float),
                ('cutoff_lo', C.c_float), ('cutoff_hi', C.c_float), ('mean_loc', C.c_float),
                ('trigger', C.c_float), ('half_range', C.c_float), ('stdv_max', C.c_float),
                ('stdv_lo', C.c_double), ('stdv_hi', C.c_double),
                ('spike_tolerance', C.c_int32), ('spike_weight', C.c_double),
                ('recal_max_dist', C.c_int32), ('recal_min_length', C.c_float),
                ('recal_max_stdv', C.c_float)]


class PolyaResult(C.Structure):
    _fields_ = [('found', C.c_int32), ('n_spikes', C.c_int32), ('begin', C.c_int64),
                ('end', C.c_int64), ('dwell_samples', C.c_int64), ('extensions', C.c_int32),
                ('flags', C.c_int32), ('spikes', (C.c_float * 4) * POLYA_MAX_SPIKES)]


class UnsplitParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        'window_size', 'window_step', 'strict_duration', 'strict_full_length',
        'strict_dna_length', 'loosen_full_length', 'loosen_dna_length',
        'basecount_quality_limit', 'subread_basecount_limit', 'subread_baseratio_limit')]


class EventTables(C.Structure):
    _fields_ = [('n_events_total', C.c_int64), ('event_offsets', C.c_void_p),
                ('start', C.c_void_p), ('mean', C.c_void_p), ('move', C.c_void_p),
                ('p_model_state', C.c_void_p), ('sampling_rate', C.c_void_p),
                ('first_sample', C.c_void_p), ('block_stride', C.c_int32)]


class Basecalls(C.Structure):
    _fields_ = [('sequence', C.c_void_p), ('qstring', C.c_void_p), ('seq_offsets', C.c_void_p),
                ('qual_table', C.c_void_p)]


class EventColumns(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('mean', 'stdv', 'scaled_mean', 'start', 'end', 'length',
                                           'pos', 'p_model_state', 'model_state', 'error')]


class Batch(C.Structure):
    _fields_ = [('n_reads', C.c_int64), ('n_raw_total', C.c_int64),
                ('max_raw_length', C.c_int64),
                ('raw', C.c_void_p), ('raw_offsets', C.c_void_p), ('raw_lengths', C.c_void_p),
                ('range', C.c_void_p), ('digitisation', C.c_void_p), ('offset', C.c_void_p),
                ('packed', C.c_void_p), ('packed_offsets', C.c_void_p)]


class Results(C.Structure):
    _fields_ = [('status', C.c_void_p), ('label', C.c_void_p), ('scale_shift', C.c_void_p),
                ('segments', C.c_void_p), ('barcode', C.c_void_p),
                ('barcode_guess', C.c_void_p), ('barcode_score', C.c_void_p),
                ('class_probs', C.c_void_p), ('pooled', C.c_void_p), ('counts', C.c_void_p),
                ('polya', C.c_void_p)]


# every symbol include/poreplex_b200.h declares
EXPORTS = ['pb2_abi_version', 'pb2_create', 'pb2_destroy', 'pb2_last_error', 'pb2_set_scaler',
           'pb2_set_segmentation_hmm', 'pb2_set_demux', 'pb2_analyze_device',
           'pb2_analyze_host', 'pb2_pool_signal', 'pb2_fit_scalers', 'pb2_detect_segments',
           'pb2_viterbi_paths', 'pb2_barcode_windows', 'pb2_demux_predict',
           'pb2_scaler_predict', 'pb2_count_results', 'pb2_kernel_launches',
           'pb2_profile_enable', 'pb2_profile_kernel_count', 'pb2_profile_kernel_name',
           'pb2_profile_read', 'pb2_profile_timeline', 'pb2_set_exact_division', 'pb2_set_polya', 'pb2_measure_polya',
           'pb2_set_unsplit', 'pb2_detect_unsplit', 'pb2_detect_unsplit_host',
           'pb2_set_fast_lstm', 'pb2_demux_predict_tc', 'pb2_recheck_stats', 'pb2_debug_demux_l1', 'pb2_rerun_causes', 'pb2_set_audit_fraction',
           'pb2_audit_stats', 'pb2_detect_events', 'pb2_derive_event_tables',
           'pb2_derive_event_tables_host', 'pb2_probe2_rows', 'pb2_svb16_decode']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + \
        glob.glob(os.path.join(CSRC, '*.cuh')) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> poreplex_b200/libporeplex_b200.so"""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
        ['-o', LIB_PATH] + sources()
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'poreplex_b200: CUDA library %s is not built (run __graft_entry__.build()); '
            'there is no CPU fallback' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.pb2_abi_version.restype = C.c_int
    L.pb2_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.pb2_destroy.argtypes = [vp]
    L.pb2_destroy.restype = None
    L.pb2_last_error.argtypes = [vp]
    L.pb2_last_error.restype = C.c_char_p
    L.pb2_set_scaler.argtypes = [vp, C.POINTER(ScalerParams)]
    L.pb2_set_segmentation_hmm.argtypes = [vp, C.POINTER(HmmParams), C.c_int32, C.c_int32]
    L.pb2_set_demux.argtypes = [vp, C.POINTER(DemuxParams)]
    L.pb2_analyze_device.argtypes = [vp, C.POINTER(Batch), C.POINTER(Results), C.c_uint32, vp]
    L.pb2_analyze_host.argtypes = [vp, C.POINTER(Batch), C.POINTER(Results), C.c_uint32]
    L.pb2_pool_signal.argtypes = [vp, C.POINTER(Batch), vp, vp]
    L.pb2_fit_scalers.argtypes = [vp, C.POINTER(Batch), vp, vp, vp, vp, vp]
    L.pb2_detect_segments.argtypes = [vp, C.POINTER(Batch), vp, vp, vp, vp, vp, vp]
    L.pb2_viterbi_paths.argtypes = [vp, C.c_int, vp, vp, C.c_int64, C.c_int32, vp, vp, vp]
    L.pb2_barcode_windows.argtypes = [vp, C.POINTER(Batch), vp, vp, vp, vp, vp, vp, vp]
    L.pb2_demux_predict.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp]
    L.pb2_scaler_predict.argtypes = [vp, vp, C.c_int64, vp, vp]
    L.pb2_count_results.argtypes = [vp, vp, vp, vp, C.c_int64, vp, vp]
    L.pb2_kernel_launches.argtypes = [vp]
    L.pb2_kernel_launches.restype = C.c_int64
    L.pb2_set_exact_division.argtypes = [vp, C.c_int]
    L.pb2_set_fast_lstm.argtypes = [vp, C.c_int, C.c_double, C.c_double]
    L.pb2_demux_predict_tc.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.pb2_recheck_stats.argtypes = [vp, _i64p, _i64p]
    L.pb2_debug_demux_l1.argtypes = [vp, vp, C.c_int64, vp, vp]
    L.pb2_rerun_causes.argtypes = [vp, _i64p, _i64p, _i64p]
    L.pb2_set_audit_fraction.argtypes = [vp, C.c_double]
    L.pb2_audit_stats.argtypes = [vp, _i64p, _i64p]
    L.pb2_set_polya.argtypes = [vp, C.POINTER(PolyaParams), C.c_int32]
    L.pb2_detect_events.argtypes = [vp, vp, vp, vp, C.c_int64, C.POINTER(DetectorParams), vp, vp,
                                    vp, vp]
    L.pb2_measure_polya.argtypes = [vp, C.POINTER(Batch), vp, vp, vp, vp, vp]
    L.pb2_set_unsplit.argtypes = [vp, C.POINTER(HmmParams), C.POINTER(UnsplitParams), C.c_int32,
                                  C.c_int32, C.c_int32]
    L.pb2_detect_unsplit.argtypes = [vp, C.POINTER(Batch), C.POINTER(EventTables), C.c_int64, vp,
                                     vp, vp, C.c_int32, vp, vp]
    L.pb2_detect_unsplit_host.argtypes = [vp, C.POINTER(Batch), C.POINTER(EventTables), C.c_int64,
                                          vp, vp, vp, C.c_int32, vp]
    L.pb2_derive_event_tables.argtypes = [vp, C.POINTER(Batch), C.POINTER(EventTables),
                                          C.POINTER(Basecalls), vp, C.POINTER(EventColumns), vp]
    L.pb2_derive_event_tables_host.argtypes = [vp, C.POINTER(Batch), C.POINTER(EventTables),
                                               C.POINTER(Basecalls), vp, C.POINTER(EventColumns)]
    L.pb2_probe2_rows.argtypes = [vp, _i64p]
    L.pb2_svb16_decode.argtypes = [vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp]
    L.pb2_profile_enable.argtypes = [vp, C.c_int]
    L.pb2_profile_kernel_count.restype = C.c_int
    L.pb2_profile_kernel_name.argtypes = [C.c_int]
    L.pb2_profile_kernel_name.restype = C.c_char_p
    L.pb2_profile_read.argtypes = [vp, _dp, _i64p, C.c_int]
    L.pb2_profile_timeline.argtypes = [vp, C.POINTER(C.c_int32), _dp, _dp, C.c_int64, _i64p]
    if L.pb2_abi_version() != 1:
        raise RuntimeError('poreplex_b200: ABI version mismatch')
    _lib = L
    return L


class NativeError(RuntimeError):
    pass
