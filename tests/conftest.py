import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope='session')
def preset():
    from poreplex_b200 import params
    return params.load_preset()


@pytest.fixture(scope='session')
def preset_short(preset):
    from poreplex_b200 import params
    return params.bench_short_preset(preset)


@pytest.fixture(scope='session')
def orc_stock(oracle_mod):
    return oracle_mod.default_oracle(bench_short=False)


@pytest.fixture(scope='session')
def orc_short(oracle_mod):
    return oracle_mod.default_oracle(bench_short=True)


def _engine(preset):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from poreplex_b200 import _native
    from poreplex_b200.engine import SignalEngine
    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    cfg = dict(preset)
    cfg['barcoding'] = True
    return SignalEngine(cfg, device=0)


@pytest.fixture(scope='session')
def eng_stock(preset):
    return _engine(preset)


@pytest.fixture(scope='session')
def eng_short(preset_short):
    return _engine(preset_short)
