import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session", params=["int8-from-n0", "default"])
def engine(request):
    """The scoring path picks its factorisation kernels by size: FP64 DMMA below n = 1536, the int8 tensor-core path from
    there on (AGPC_OZ_MIN_N, read when the context is created).  The parity tests run at n = 100 ... 600, so every test
    that takes this fixture runs twice: once with the int8 path forced on at every size, once as shipped."""
    from autogpc_b200.engine import Engine
    old = os.environ.get("AGPC_OZ_MIN_N")
    if request.param == "int8-from-n0":
        os.environ["AGPC_OZ_MIN_N"] = "0"
    else:
        os.environ.pop("AGPC_OZ_MIN_N", None)
    try:
        eng = Engine(0)
    finally:
        if old is None:
            os.environ.pop("AGPC_OZ_MIN_N", None)
        else:
            os.environ["AGPC_OZ_MIN_N"] = old
    eng.test_env = {"AGPC_OZ_MIN_N": "0"} if request.param == "int8-from-n0" else {}   # for tests that build a second engine
    yield eng
    eng.close()
