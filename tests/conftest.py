import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def engine():
    """One engine for the whole GPU session (the CUDA library must exist: no fallback)."""
    from context_adaptive_neural_network_based_prediction_b200 import Engine
    eng = Engine()
    yield eng
    eng.close()


@pytest.fixture(scope='session')
def weights_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp('pnnw'))
