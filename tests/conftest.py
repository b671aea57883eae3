import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port():
    from oracle.pyoracle import Port

    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref, ref_available

    if not ref_available() and not os.path.isdir("/root/reference/src"):
        pytest.skip("oracle/_ref/libatk_ref.so not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def ctx():
    from algotoolkit_b200 import capi

    c = capi.Context(0)
    yield c
    c.destroy()


def mm_path(name):
    return os.path.join(GOLDEN, "data", name + ".mtx")
