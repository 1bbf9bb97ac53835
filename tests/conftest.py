import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session')
def golden():
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'ref_kat.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def built_lib():
    """The C-ABI library; built on demand so the CPU suite can check its exports."""
    so = os.path.join(ROOT, 'dis-yolo_b200', 'libdisyolo_b200.so')
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    return so
