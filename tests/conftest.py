import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the Lightspinner reference at /root/reference (CPU container only)')


@pytest.fixture(scope='session')
def oracle():
    from oracle import mali_oracle
    mali_oracle.build()
    return mali_oracle
