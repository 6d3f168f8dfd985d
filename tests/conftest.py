import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail loudly (not skip) when the CUDA library is missing on a GPU box;
    # on a box without a GPU they are deselected by `-m "not gpu"`.
    pass
