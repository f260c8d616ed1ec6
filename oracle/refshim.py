"""Shims that let the REFERENCE's own Python run verbatim in this container
(TEST INFRASTRUCTURE; used only by tests/golden/make_golden.py and the CPU tests).

The reference imports three packages that do not exist here and cannot be installed
(no network): h5py, pomegranate, tensorflow -- plus its own compiled extension
``poreplex.csupport``.  ``install()`` puts stand-ins into ``sys.modules``:

  h5py            in-memory FAST5 trees (registered with ``register_fast5``) and
                  read-only access to real ``.hdf5`` weight files through
                  poreplex_b200.hdf5_min
  pomegranate     HiddenMarkovModel/State/NormalDistribution/GeneralMixtureModel whose
                  ``bake()``/``viterbi()`` call the oracle's C restatement
  tensorflow      ``keras.models.load_model(path).predict(x, batch_size)`` running the
                  oracle's C LSTM restatement on the real weights
  poreplex.csupport   ``detect_events`` over the reference's own scrappie C compiled
                  into oracle/_ref (falls back to the restatement if that is absent)
  pysam, mappy    import-only stubs (io.py:23, alignment_writer.py:23-25)

With these, ``import poreplex.signal_analyzer`` from /root/reference works unmodified
and ``process_batch`` produces the golden result dicts.
"""
import os
import sys
import types

import numpy as np

from . import oracle as O

REFERENCE_ROOT = os.environ.get('POREPLEX_REFERENCE', '/root/reference')

_fast5_registry = {}


# ------------------------------------------------------------------ fake h5py
class FakeAttrs(dict):
    pass


class FakeDataset:
    def __init__(self, data, name='', attrs=None):
        self._data = data
        self.name = name
        self.attrs = FakeAttrs(attrs or {})

    def __len__(self):
        return len(self._data)

    def __getitem__(self, key):
        if isinstance(key, tuple) and key == ():
            return self._data
        return self._data[key]

    @property
    def shape(self):
        return np.shape(self._data)

    @property
    def dtype(self):
        return np.asarray(self._data).dtype


class FakeGroup:
    def __init__(self, name='/', attrs=None):
        self.name = name
        self.attrs = FakeAttrs(attrs or {})
        self._children = {}

    # construction helpers (not part of the h5py API)
    def add_group(self, name, attrs=None):
        g = FakeGroup(self.name.rstrip('/') + '/' + name, attrs)
        self._children[name] = g
        return g

    def add_dataset(self, name, data, attrs=None):
        d = FakeDataset(data, self.name.rstrip('/') + '/' + name, attrs)
        self._children[name] = d
        return d

    # h5py API subset (fast5_file.py:43-58,65-131,133-230)
    def _lookup(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, FakeGroup) or part not in node._children:
                raise KeyError("Unable to open object (object '%s' doesn't exist)" % part)
            node = node._children[part]
        return node

    def __getitem__(self, path):
        return self._lookup(path)

    def __contains__(self, path):
        try:
            self._lookup(path)
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self._children)

    def keys(self):
        return self._children.keys()

    def values(self):
        return self._children.values()

    def items(self):
        return self._children.items()


class FakeFile(FakeGroup):
    def __init__(self):
        super().__init__('/')

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass


def register_fast5(path, tree):
    _fast5_registry[os.path.abspath(path)] = tree


def clear_fast5():
    _fast5_registry.clear()


def _h5py_File(path, mode='r', *a, **kw):
    ap = os.path.abspath(path)
    if ap in _fast5_registry:
        return _fast5_registry[ap]
    if os.path.exists(path) and os.path.getsize(path) > 0:
        from poreplex_b200.hdf5_min import Hdf5File
        return Hdf5File(path)
    raise OSError('Unable to open file (unregistered fake FAST5): ' + path)


def make_h5py():
    m = types.ModuleType('h5py')
    m.File = _h5py_File
    m.__version__ = '0.0-oracle-shim'
    return m


# ------------------------------------------------------------ fake pomegranate
class NormalDistribution:
    def __init__(self, mu, sigma):
        self.parameters = (float(mu), float(sigma))


class GeneralMixtureModel:
    def __init__(self, distributions, weights=None):
        self.distributions = list(distributions)
        self.weights = np.asarray(weights, dtype=np.float64)


class State:
    def __init__(self, distribution, name=None):
        self.distribution = distribution
        self.name = name


class HiddenMarkovModel:
    def __init__(self, name=None):
        self.name = name
        self.start = State(None, name=str(name) + '-start')
        self.end = State(None, name=str(name) + '-end')
        self._states = []
        self._edges = []
        self._baked = None

    def add_state(self, state):
        self._states.append(state)

    def add_transition(self, a, b, probability):
        self._edges.append((a, b, float(probability)))

    def bake(self):
        modeldata = []
        for st in self._states:
            d = st.distribution
            if isinstance(d, NormalDistribution):
                emission = [list(d.parameters)]
            else:
                emission = [list(c.parameters) + [float(w)]
                            for c, w in zip(d.distributions, d.weights)]
            entry = {'name': st.name, 'emission': emission, 'transition': []}
            modeldata.append(entry)
        by_obj = {id(st): e for st, e in zip(self._states, modeldata)}
        for a, b, p in self._edges:
            if a is self.start:
                by_obj[id(b)]['start_prob'] = p
            else:
                by_obj[id(a)]['transition'].append([b.name, p])
        self._hmm, names = O.bake_hmm(modeldata)
        by_name = {st.name: st for st in self._states}
        self.states = [by_name[n] for n in names] + [self.start, self.end]
        self._baked = True

    def viterbi(self, sequence):
        import ctypes as C
        x = np.ascontiguousarray(np.asarray(sequence, dtype=np.float64).astype(np.float32))
        if not np.array_equal(x.astype(np.float64), np.asarray(sequence, dtype=np.float64)):
            raise ValueError('oracle viterbi expects float32-representable input')
        path = np.empty(max(len(x), 1), np.int32)
        logp = O.lib().orc_viterbi(C.byref(self._hmm), x.ctypes.data_as(C.POINTER(C.c_float)),
                                   C.c_int(len(x)), path.ctypes.data_as(C.POINTER(C.c_int32)))
        if logp == -np.inf:
            return logp, None
        start_idx = len(self.states) - 2
        calls = [(start_idx, self.start)]
        calls += [(int(s), self.states[int(s)]) for s in path[:len(x)]]
        return logp, calls


def make_pomegranate():
    m = types.ModuleType('pomegranate')
    m.__version__ = '0.10.0-oracle-shim'
    m.HiddenMarkovModel = HiddenMarkovModel
    m.GeneralMixtureModel = GeneralMixtureModel
    m.State = State
    m.NormalDistribution = NormalDistribution
    return m


# ------------------------------------------------------------- fake tensorflow
class _KerasModel:
    def __init__(self, path):
        from poreplex_b200 import params
        import ctypes as C
        base = os.path.basename(path)
        self._keep = []
        if 'scaler' in base:
            self.kind = 'scaler'
            mdl = params.load_scaler_model(path)
            s = O.ScalerC()
            s.l1 = O._lstm_c(mdl.l1, self._keep)
            s.l2 = O._lstm_c(mdl.l2, self._keep)
            Wd = np.ascontiguousarray(mdl.dense_kernel, np.float32)
            bd = np.ascontiguousarray(mdl.dense_bias, np.float32)
            self._keep += [Wd, bd]
            s.Wd, s.bd = O._p(Wd), O._p(bd)
            self._c, self.n_out = s, 2
        else:
            self.kind = 'demux'
            mdl = params.load_demux_model(path)
            d = O.DemuxC()
            d.fwd = O._lstm_c(mdl.fwd, self._keep)
            d.bwd = O._lstm_c(mdl.bwd, self._keep)
            d.l2 = O._lstm_c(mdl.l2, self._keep)
            Wd = np.ascontiguousarray(mdl.dense_kernel, np.float32)
            bd = np.ascontiguousarray(mdl.dense_bias, np.float32)
            self._keep += [Wd, bd]
            d.Wd, d.bd = O._p(Wd), O._p(bd)
            d.n_classes = mdl.n_classes
            self._c, self.n_out = d, mdl.n_classes
        self._C = C

    def predict(self, x, batch_size=None, verbose=0):
        C = self._C
        x = np.ascontiguousarray(x, np.float32)
        n, T = x.shape[0], x.shape[1]
        x = x.reshape(n, T)
        out = np.empty((n, self.n_out), np.float32)
        fn = O.lib().orc_scaler_predict if self.kind == 'scaler' else O.lib().orc_demux_predict
        for i in range(n):
            fn(C.byref(self._c), O._p(x[i]), C.c_int(T), O._p(out[i]))
        return out


def make_tensorflow():
    tf = types.ModuleType('tensorflow')
    keras = types.ModuleType('tensorflow.keras')
    models = types.ModuleType('tensorflow.keras.models')
    utils = types.ModuleType('tensorflow.keras.utils')
    backend = types.ModuleType('tensorflow.keras.backend')
    losses = types.ModuleType('tensorflow.keras.losses')
    metrics = types.ModuleType('tensorflow.keras.metrics')
    _custom = {}
    models.load_model = lambda path, *a, **kw: _KerasModel(path)
    utils.get_custom_objects = lambda: _custom
    backend.cast_to_floatx = lambda x: np.asarray(x, np.float32)

    class CategoricalCrossentropy:
        def __init__(self, *a, **kw):
            pass

    class CategoricalAccuracy:
        def __init__(self, *a, **kw):
            pass
    losses.CategoricalCrossentropy = CategoricalCrossentropy
    metrics.CategoricalAccuracy = CategoricalAccuracy

    class _Logger:
        def setLevel(self, *_):
            pass
    tf.get_logger = lambda: _Logger()
    tf.__version__ = '0.0-oracle-shim'
    tf.keras = keras
    keras.models, keras.utils, keras.backend = models, utils, backend
    keras.losses, keras.metrics = losses, metrics
    return {'tensorflow': tf, 'tensorflow.keras': keras, 'tensorflow.keras.models': models,
            'tensorflow.keras.utils': utils, 'tensorflow.keras.backend': backend,
            'tensorflow.keras.losses': losses, 'tensorflow.keras.metrics': metrics}


# ------------------------------------------------------------ csupport + stubs
def make_csupport():
    m = types.ModuleType('poreplex.csupport')
    if O.have_ref_scrappie():
        m.detect_events = O.detect_events_ref
        m.__oracle_backend__ = 'reference scrappie C (oracle/_ref)'
    else:
        m.detect_events = O.detect_events_restated
        m.__oracle_backend__ = 'restated (pb_oracle.c)'

    class error(Exception):
        pass
    m.error = error
    return m


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_fake_h5py():
    """Only the in-memory h5py (what the GPU-box tests need to feed FAST5 trees to the
    product's process_batch; no reference tree required)."""
    if 'h5py' not in sys.modules or not hasattr(sys.modules['h5py'], '__oracle_shim__'):
        m = make_h5py()
        m.__oracle_shim__ = True
        sys.modules['h5py'] = m
    return sys.modules['h5py']


_installed = False


def install():
    """Install every shim and make ``import poreplex`` resolve to the reference tree."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'poreplex')):
        raise RuntimeError('reference tree not found at ' + REFERENCE_ROOT)
    install_fake_h5py()
    sys.modules['pomegranate'] = make_pomegranate()
    sys.modules.update(make_tensorflow())
    import gzip
    sys.modules.setdefault('pysam', _stub(
        'pysam', BGZFile=lambda path, mode='r': gzip.open(path, mode), faidx=None,
        FUNMAP=4, FREVERSE=16, FSECONDARY=256, FSUPPLEMENTARY=2048,
        AlignmentFile=None, AlignedSegment=None))
    sys.modules.setdefault('mappy', _stub('mappy'))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import poreplex                                     # the reference package
    cs = make_csupport()
    sys.modules['poreplex.csupport'] = cs
    poreplex.csupport = cs
    _installed = True


def reference_modules():
    """(signal_analyzer, signal_loader, barcoding, polya, fast5_file) of the reference."""
    install()
    from poreplex import signal_analyzer, signal_loader, barcoding, polya, fast5_file
    return signal_analyzer, signal_loader, barcoding, polya, fast5_file


def reset_reference_persistence():
    """Forget the per-process model cache (worker_persistence.py:37-58)."""
    sys.modules.pop('__poreplex_persistence', None)
