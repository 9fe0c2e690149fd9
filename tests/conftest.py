import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g
    g.build()
    import rmpe_b200
    return rmpe_b200


@pytest.fixture(scope="session")
def rmpe(built_lib):
    """The package with the library initialised on cuda:0 -- fails loudly without a GPU."""
    built_lib.lib.ensure_init(0)
    return built_lib


@pytest.fixture(scope="session")
def gt_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "gt_golden.npz"))


@pytest.fixture(scope="session")
def decode_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "decode_golden.npz"))
