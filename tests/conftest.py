import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "refcheck: cross-check against the live reference tree "
                                       "(only where /root/reference exists)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import fluids2d_oracle as orc
    orc.build()
    return orc
