import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_ref: needs oracle/_ref (the compiled reference)")


def pytest_collection_modifyitems(config, items):
    import checkers

    if checkers.reference() is None:
        skip = pytest.mark.skip(reason="oracle/_ref not built (reference sources absent)")
        for it in items:
            if "needs_ref" in it.keywords:
                it.add_marker(skip)
