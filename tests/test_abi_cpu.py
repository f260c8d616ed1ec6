"""The C-ABI shared library: builds for sm_100a without a GPU, loads, and exports every
symbol include/poreplex_b200.h declares.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'poreplex_b200.h')


@pytest.fixture(scope='module')
def native():
    from poreplex_b200 import _native
    _native.build()
    return _native


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pb2_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_all_exported(native):
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(native.EXPORTS) == syms, 'python binding list out of date with the header'
    lib = C.CDLL(native.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.run(['nm', '-D', '--defined-only', native.LIB_PATH], capture_output=True,
                         text=True).stdout
    exported = set(re.findall(r' T (pb2_\w+)', out))
    assert exported == set(syms)


def test_library_is_sm100a_and_has_no_torch_types(native):
    out = subprocess.run(['cuobjdump', '-lelf', native.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    text = open(HEADER).read()
    assert 'torch' not in text.lower().replace('no python,\n * torch', '').replace('torch or cuda types', '')


def test_library_sass_holds_tcgen05_tmem_and_tma_instructions(native):
    """The built cubin is what the design says it is: tcgen05 products (UTCHMMA) with TMEM
    traffic (LDTM / STTM) in every tensor-core LSTM kernel, bulk copies by the TMA unit
    (UBLKCP + mbarrier waits) in k_pool, and no legacy HMMA tensor-core path anywhere
    (tools/sass_histogram.py; B200_PROFILING.md's mnemonics)."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'sass_histogram.py')], capture_output=True,
                         text=True, check=True).stdout
    doc = json.loads(out)
    assert doc['totals'].get('UTCHMMA', 0) >= 100 and doc['totals'].get('STTM', 0) >= 50
    assert doc['totals'].get('UBLKCP', 0) >= 1
    assert 'HMMA_legacy' not in doc['totals']
    by = {r['kernel']: r for r in doc['kernels']}
    tc = [k for k in by if k.startswith('pb::k_lstm_tc')]
    assert len(tc) >= 7
    for k in tc:
        assert by[k].get('UTCHMMA', 0) > 0 and by[k].get('LDTM', 0) > 0 and by[k].get('STTM', 0) > 0, k
        assert by[k].get('UTCBAR', 0) > 0 and by[k].get('SYNCS', 0) > 0, k
    pool = [k for k in by if k.startswith('pb::k_pool')]
    assert pool and all(by[k].get('UBLKCP', 0) > 0 and by[k].get('SYNCS', 0) > 0 for k in pool)


def test_struct_layouts_match_header_sizes(native):
    """ctypes mirrors must have the C layout (checked against a tiny C program)."""
    src = r'''
#include <stdio.h>
#include "poreplex_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(pb2_lstm_weights), sizeof(pb2_scaler_params),
         sizeof(pb2_hmm_params), sizeof(pb2_demux_params), sizeof(pb2_batch), sizeof(pb2_results));
  printf("%zu %zu %zu %zu %zu\n", sizeof(pb2_polya_params), sizeof(pb2_polya_result),
         sizeof(pb2_unsplit_params), sizeof(pb2_event_tables), sizeof(pb2_detector_params));
  return 0; }
'''
    import tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, 't.c'), 'w').write(src)
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', os.path.join(d, 't'),
                           os.path.join(d, 't.c')])
    sizes = [int(x) for x in subprocess.check_output([os.path.join(d, 't')]).split()]
    N = native
    assert sizes == [C.sizeof(N.LstmWeights), C.sizeof(N.ScalerParams), C.sizeof(N.HmmParams),
                     C.sizeof(N.DemuxParams), C.sizeof(N.Batch), C.sizeof(N.Results),
                     C.sizeof(N.PolyaParams), C.sizeof(N.PolyaResult), C.sizeof(N.UnsplitParams),
                     C.sizeof(N.EventTables), C.sizeof(N.DetectorParams)]


def test_fast5_struct_layout_matches_header():
    from poreplex_b200 import fast5_loader as FL
    src = r'''
#include <stdio.h>
#include "poreplex_b200_fast5.h"
int main(void) { printf("%zu\n", sizeof(pb2f_read_meta)); return 0; }
'''
    import tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, 't.c'), 'w').write(src)
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', os.path.join(d, 't'),
                           os.path.join(d, 't.c')])
    assert int(subprocess.check_output([os.path.join(d, 't')])) == C.sizeof(FL.ReadMeta)


def test_no_gpu_means_loud_failure(native):
    """Without a CUDA device the context cannot be created and the Python engine raises;
    there is no CPU fallback in the product path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    lib = native.load()
    h = C.c_void_p()
    assert lib.pb2_create(0, C.byref(h)) != 0 and not h.value
    from poreplex_b200.engine import SignalEngine
    from poreplex_b200 import params
    with pytest.raises(native.NativeError):
        SignalEngine(dict(params.load_preset(), barcoding=True))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'poreplex_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, fn
                assert 'pb_oracle' not in text.replace('oracle/pb_oracle.c', ''), fn
