import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (the .so files are git-ignored): build the in-tree CUDA library once, the
    way __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU).  Only the TESTS do this -- the product
    itself still fails loudly (NativeLibraryMissing) when libgmsm.so is absent."""
    lib = os.path.join(ROOT, "gnark-crypto_b200", "libgmsm.so")
    if not os.path.exists(lib):
        import importlib

        importlib.import_module("gnark-crypto_b200.build").build(force=False, verbose=False)


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail loudly (not skip) when the CUDA library is missing on a GPU box;
    # on a box without a GPU they are deselected by `-m "not gpu"`.
    pass
