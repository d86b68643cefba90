import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def ref_ext():
    """The unmodified reference extensions compiled into oracle/_ref (None when not built)."""
    from oracle import build_ref

    def get(name):
        try:
            return build_ref.load(name)
        except (FileNotFoundError, ImportError):
            return None
    return get
