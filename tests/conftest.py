import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def p21():
    from bundleadjustment_benchmarks_b200 import bal
    return bal.load_named("problem-21-11315")


@pytest.fixture(scope="session")
def p39():
    from bundleadjustment_benchmarks_b200 import bal
    return bal.load_named("problem-39-18060")


@pytest.fixture(scope="session")
def tiny():
    from bundleadjustment_benchmarks_b200 import bal
    return bal.synthetic(6, 60, seed=1)


@pytest.fixture(scope="session")
def small():
    from bundleadjustment_benchmarks_b200 import bal
    return bal.synthetic(40, 3000, window=8, seed=2)
