"""Model-parameter loading for the B200 signal path.

Mirrors what the reference builds once per worker process in
``WorkerPersistenceStorage.init_persistence_objects`` (worker_persistence.py:60-90):

* the two Keras weight files (``SignalLoader.load_scaler_model`` signal_loader.py:49-75,
  ``BarcodeDemultiplexer.load_model`` barcoding.py:51-70), read here either straight
  from the reference's ``.hdf5`` (via :mod:`hdf5_min`) or from the ``.npz`` extracts
  shipped in ``poreplex_b200/presets`` (made by ``tools/import_reference_preset.py``);
* the two HMMs described in the preset YAML (``load_segmentation_model``
  worker_persistence.py:95-121), turned into the flat log-space tables the CUDA
  Viterbi kernel consumes ("baking").

Nothing here does per-read arithmetic; it only prepares constants.
"""
import json
import math
import os

import numpy as np

from .hdf5_min import Hdf5File

__all__ = ['LstmLayer', 'ScalerModel', 'DemuxModel', 'HmmTables', 'load_scaler_model',
           'load_demux_model', 'load_preset', 'resolve_model_file', 'PRESET_DIR',
           'STATUS_NAMES', 'STATUS_CODES', 'bench_short_preset']

PRESET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'presets')

# io.py:245-260
STATUS_NAMES = ['okay', 'disappeared', 'irregular_fast5', 'scaler_signal_too_short',
                'scaling_qc_fail', 'adapter_not_detected', 'not_basecalled',
                'basecall_table_incomplete', 'unsplit_read', 'sequence_too_short',
                'unknown_error']
STATUS_CODES = {n: i for i, n in enumerate(STATUS_NAMES)}

# pomegranate/distributions/NormalDistribution.pyx: DEF SQRT_2_PI = 2.50662827463
SQRT_2_PI = 2.50662827463


class LstmLayer:
    """One Keras LSTM / LSTMCell: kernel [in,4H], recurrent_kernel [H,4H], bias [4H];
    gate column blocks ordered i|f|c|o; ``implementation`` 1 or 2 (bias add order)."""

    def __init__(self, kernel, recurrent, bias, implementation):
        self.kernel = np.ascontiguousarray(kernel, np.float32)
        self.recurrent = np.ascontiguousarray(recurrent, np.float32)
        self.bias = np.ascontiguousarray(bias, np.float32)
        self.units = self.recurrent.shape[0]
        self.in_dim = self.kernel.shape[0]
        self.implementation = int(implementation)
        if self.kernel.shape[1] != 4 * self.units or self.bias.shape != (4 * self.units,):
            raise ValueError('inconsistent LSTM weight shapes')


class ScalerModel:
    def __init__(self, l1, l2, dense_kernel, dense_bias, input_defs, output_transform,
                 model_version):
        self.l1, self.l2 = l1, l2
        self.dense_kernel = np.ascontiguousarray(dense_kernel, np.float32)
        self.dense_bias = np.ascontiguousarray(dense_bias, np.float32)
        self.input_defs = dict(input_defs)
        self.output_transform = dict(output_transform)
        self.model_version = model_version


class DemuxModel:
    def __init__(self, fwd, bwd, l2, dense_kernel, dense_bias, calibration, loss_weights):
        self.fwd, self.bwd, self.l2 = fwd, bwd, l2
        self.dense_kernel = np.ascontiguousarray(dense_kernel, np.float32)
        self.dense_bias = np.ascontiguousarray(dense_bias, np.float32)
        self.calibration = np.ascontiguousarray(calibration, np.float64)
        self.loss_weights = np.asarray(loss_weights, np.float32)
        self.n_classes = self.dense_bias.shape[0]


def _literal(b):
    import ast
    return ast.literal_eval(b.decode().strip())


def _model_layers(h5):
    cfg = json.loads(h5.attrs['model_config'].decode())
    return cfg['config']['layers']


def _load_scaler_hdf5(path):
    with Hdf5File(path) as f:
        layers = _model_layers(f)
        kinds = [l['class_name'] for l in layers if l['class_name'] != 'Dropout']
        if kinds != ['LSTM', 'LSTM', 'Dense']:
            raise ValueError('unsupported scaler architecture: %r' % kinds)
        lstm_cfgs = [l['config'] for l in layers if l['class_name'] == 'LSTM']
        for c in lstm_cfgs:
            if (c['activation'], c['recurrent_activation']) != ('tanh', 'sigmoid'):
                raise ValueError('unsupported LSTM activations')
        mw = f['model_weights']
        names = [n.decode() for n in mw.attrs['layer_names']]
        lstm_names = [n for n in names if n.startswith('lstm')]
        dense_name = [n for n in names if n.startswith('dense')][0]

        def lstm(name, cfg):
            g = mw['%s/%s' % (name, name)]
            return LstmLayer(g['kernel:0'][:], g['recurrent_kernel:0'][:], g['bias:0'][:],
                             cfg['implementation'])
        l1 = lstm(lstm_names[0], lstm_cfgs[0])
        l2 = lstm(lstm_names[1], lstm_cfgs[1])
        dg = mw['%s/%s' % (dense_name, dense_name)]
        return ScalerModel(l1, l2, dg['kernel:0'][:], dg['bias:0'][:],
                           _literal(mw.attrs['input_defs']),
                           _literal(mw.attrs['output_transform']),
                           mw.attrs['model_version'].decode())


def _load_demux_hdf5(path):
    with Hdf5File(path) as f:
        layers = _model_layers(f)
        kinds = [l['class_name'] for l in layers
                 if l['class_name'] not in ('Dropout', 'GaussianNoise')]
        if kinds != ['Bidirectional', 'RNN', 'Dense']:
            raise ValueError('unsupported demux architecture: %r' % kinds)
        bidi = [l for l in layers if l['class_name'] == 'Bidirectional'][0]['config']
        if bidi['merge_mode'] != 'concat':
            raise ValueError('unsupported Bidirectional merge mode')
        cell1 = bidi['layer']['config']['cell']['config']
        cell2 = [l for l in layers if l['class_name'] == 'RNN'][0]['config']['cell']['config']
        mw = f['model_weights']
        names = [n.decode() for n in mw.attrs['layer_names']]
        bname = [n for n in names if n.startswith('bidirectional')][0]
        rname = [n for n in names if n.startswith('rnn')][0]
        dname = [n for n in names if n.startswith('dense')][0]
        bg = mw['%s/%s' % (bname, bname)]

        def cell(g, cfg):
            return LstmLayer(g['kernel:0'][:], g['recurrent_kernel:0'][:], g['bias:0'][:],
                             cfg['implementation'])
        fwd = cell(bg['forward_rnn'], cell1)
        bwd = cell(bg['backward_rnn'], cell1)
        l2 = cell(mw['%s/%s' % (rname, rname)], cell2)
        dg = mw['%s/%s' % (dname, dname)]
        calib = f['poreplex_params/calibration'].read()
        # barcoding.py:57-59
        if np.any(calib['phred'] != np.arange(len(calib))):
            raise RuntimeError('Calibration table in {} is not continuous.'.format(path))
        return DemuxModel(fwd, bwd, l2, dg['kernel:0'][:], dg['bias:0'][:],
                          calib['pred_score'].astype(np.float64),
                          f['poreplex_params/loss_weights'][:])


def _pack_lstm(prefix, l, out):
    out[prefix + '_kernel'] = l.kernel
    out[prefix + '_recurrent'] = l.recurrent
    out[prefix + '_bias'] = l.bias
    out[prefix + '_impl'] = np.int32(l.implementation)


def _unpack_lstm(prefix, z):
    return LstmLayer(z[prefix + '_kernel'], z[prefix + '_recurrent'], z[prefix + '_bias'],
                     int(z[prefix + '_impl']))


def save_scaler_npz(model, path):
    out = {}
    _pack_lstm('l1', model.l1, out)
    _pack_lstm('l2', model.l2, out)
    out['dense_kernel'] = model.dense_kernel
    out['dense_bias'] = model.dense_bias
    out['meta'] = np.array(json.dumps({'input_defs': model.input_defs,
                                       'output_transform': model.output_transform,
                                       'model_version': model.model_version}))
    np.savez(path, **out)


def save_demux_npz(model, path):
    out = {}
    _pack_lstm('fwd', model.fwd, out)
    _pack_lstm('bwd', model.bwd, out)
    _pack_lstm('l2', model.l2, out)
    out['dense_kernel'] = model.dense_kernel
    out['dense_bias'] = model.dense_bias
    out['calibration'] = model.calibration
    out['loss_weights'] = model.loss_weights
    np.savez(path, **out)


def resolve_model_file(name):
    """Find ``name`` (as written in the preset, e.g. ``MIN106-RNA001/scaler-r3.hdf5``):
    $POREPLEX_B200_PRESETS first, then the packaged presets; ``.hdf5`` preferred,
    ``.npz`` extract accepted."""
    dirs = []
    if os.environ.get('POREPLEX_B200_PRESETS'):
        dirs.append(os.environ['POREPLEX_B200_PRESETS'])
    dirs.append(PRESET_DIR)
    stem = os.path.splitext(name)[0]
    for d in dirs:
        for cand in (name, stem + '.npz'):
            p = os.path.join(d, cand)
            if os.path.exists(p):
                return p
    raise FileNotFoundError('model file %r not found in %r' % (name, dirs))


def load_scaler_model(name_or_path):
    path = name_or_path if os.path.exists(name_or_path) else resolve_model_file(name_or_path)
    if path.endswith('.npz'):
        z = np.load(path)
        meta = json.loads(str(z['meta']))
        return ScalerModel(_unpack_lstm('l1', z), _unpack_lstm('l2', z), z['dense_kernel'],
                           z['dense_bias'], meta['input_defs'], meta['output_transform'],
                           meta['model_version'])
    return _load_scaler_hdf5(path)


def load_demux_model(name_or_path):
    path = name_or_path if os.path.exists(name_or_path) else resolve_model_file(name_or_path)
    if path.endswith('.npz'):
        z = np.load(path)
        return DemuxModel(_unpack_lstm('fwd', z), _unpack_lstm('bwd', z), _unpack_lstm('l2', z),
                          z['dense_kernel'], z['dense_bias'], z['calibration'],
                          z['loss_weights'])
    return _load_demux_hdf5(path)


class HmmTables:
    """Flat log-space tables of one HMM in pomegranate's *baked* state order.

    ``modeldata`` is the list of state dicts from the preset
    (``segmentation_model`` / ``unsplit_read_detection_model``).  Following
    ``bake()``: non-silent states are sorted by name; edges are
    stored as natural logs and a state's out-edges are re-weighted only when their
    probabilities, rounded to 8 decimals, do not sum to 1.  In-edges of each state are enumerated in baked source
    order (ties between finite candidates are measure-zero; the order is fixed so
    that oracle and kernel agree).
    """
    MAX_STATES = 8
    MAX_COMP = 4
    MAX_EDGES = 64

    def __init__(self, modeldata):
        names_yaml = [s['name'] for s in modeldata]
        order = sorted(range(len(modeldata)), key=lambda i: names_yaml[i])
        self.names = [names_yaml[i] for i in order]
        self.yaml_index = order                       # baked index -> preset index
        idx = {n: i for i, n in enumerate(self.names)}
        S = self.n_states = len(self.names)
        if S > self.MAX_STATES:
            raise ValueError('too many HMM states')
        self.n_comp = np.zeros(self.MAX_STATES, np.int32)
        self.mu = np.zeros((self.MAX_STATES, self.MAX_COMP))
        self.lsp = np.zeros((self.MAX_STATES, self.MAX_COMP))
        self.inv2s2 = np.zeros((self.MAX_STATES, self.MAX_COMP))
        self.logw = np.zeros((self.MAX_STATES, self.MAX_COMP))
        self.log_start = np.full(self.MAX_STATES, -np.inf)
        trans = np.zeros((S, S))
        start = np.zeros(S)
        for s in modeldata:
            i = idx[s['name']]
            em = s['emission']
            if len(em) > self.MAX_COMP:
                raise ValueError('too many mixture components')
            self.n_comp[i] = len(em)
            w = np.array([e[2] if len(e) > 2 else 1.0 for e in em], dtype=np.float64)
            w = w / w.sum()
            for j, e in enumerate(em):
                mu, sigma = float(e[0]), float(e[1])
                self.mu[i, j] = mu
                self.lsp[i, j] = -math.log(sigma * SQRT_2_PI)
                self.inv2s2[i, j] = 1. / (2 * sigma ** 2)
                self.logw[i, j] = math.log(w[j])
            if 'start_prob' in s:
                start[i] = float(s['start_prob'])
            for nxt, prob in s['transition']:
                trans[i, idx[nxt]] += float(prob)
        # bake(): out-edges of a state are re-weighted only when
        # round(sum(e**logp), 8) != 1, by subtracting log(sum) in log space.
        with np.errstate(divide='ignore'):
            logtrans = np.log(trans)
            logstart = np.log(start)
        for i in range(S):
            tot = round(float(np.sum(np.e ** logtrans[i][trans[i] > 0])), 8)
            if tot != 1. and tot > 0:
                logtrans[i] = logtrans[i] - math.log(tot)
        tot = round(float(np.sum(np.e ** logstart[start > 0])), 8)
        if tot != 1. and tot > 0:
            logstart = logstart - math.log(tot)
        self.log_start[:S] = logstart
        self.in_begin = np.zeros(self.MAX_STATES + 1, np.int32)
        self.in_src = np.zeros(self.MAX_EDGES, np.int32)
        self.in_logp = np.zeros(self.MAX_EDGES)
        k = 0
        for l in range(S):
            self.in_begin[l] = k
            for src in range(S):
                if trans[src, l] > 0:
                    if k >= self.MAX_EDGES:
                        raise ValueError('too many HMM edges')
                    self.in_src[k] = src
                    self.in_logp[k] = logtrans[src, l]
                    k += 1
        self.in_begin[S:] = k
        self.n_edges = k
        self.trans = trans
        # left-to-right: no cycle other than self loops -> segments can be carried
        # forward without a traceback matrix
        self.left_to_right = self._is_dag(trans)

    @staticmethod
    def _is_dag(trans):
        S = trans.shape[0]
        adj = (trans > 0) & ~np.eye(S, dtype=bool)
        indeg = adj.sum(axis=0)
        alive = np.ones(S, bool)
        for _ in range(S):
            free = [i for i in range(S) if alive[i] and indeg[i] == 0]
            if not free:
                break
            for i in free:
                alive[i] = False
                indeg -= adj[i]
        return not alive.any()

    def index_of(self, name):
        return self.names.index(name) if name in self.names else -1


def load_preset(path=None):
    """Load a preset into the flat config fragments the reference keeps in its YAML
    (``presets/rna-r941.cfg``; commandline.py:60-76).  Our packaged copy is JSON."""
    if path is None:
        path = os.path.join(PRESET_DIR, 'rna_r941.json')
    if path.endswith('.json'):
        with open(path) as f:
            return json.load(f)
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def bench_short_preset(preset):
    """The ``bench-short`` variant: two data-only changes so
    that 4000-sample reads traverse every stage -- scaler min_length 9000 -> 900 raw
    samples and demultiplexing.minimum_dna_length 260 -> 100 pooled samples."""
    import copy
    p = copy.deepcopy(preset)
    p['demultiplexing']['minimum_dna_length'] = 100
    p.setdefault('signal_processing', {})['scaler_min_length_override'] = 900
    p['preset_name'] = p.get('preset_name', 'rna-r941') + '+bench-short'
    return p
