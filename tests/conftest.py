import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rtb():
    import importlib

    mod = importlib.import_module("raytracing-in-one-weekend_b200")
    mod.build.build_host()
    return mod


@pytest.fixture(scope="session")
def oracle(rtb):
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def ctx(rtb):
    """One plugin context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    c = rtb.plugin.Context(0)
    yield c
    c.close()
