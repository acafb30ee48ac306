import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def pkg():
    import __graft_entry__ as ge
    p = ge.load_package()
    if not os.path.exists(p.LIB_PATH):
        ge.build()
    return p


@pytest.fixture(scope='session')
def synth(pkg):
    import __graft_entry__ as ge
    return ge.load_synth()


@pytest.fixture(scope='session')
def ob():
    import oracle_binding
    oracle_binding.lib()
    return oracle_binding
