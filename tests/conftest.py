import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_fns():
    from tests import oracle_loader
    return oracle_loader.load()


@pytest.fixture(scope="session")
def product_fns():
    from mpcholonavigation_b200 import load_product
    return load_product()
