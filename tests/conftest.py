import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): compile it once if a fresh checkout has none, so the ABI /
    host-logic tests do not depend on someone having run __graft_entry__.build() first.  The product never builds itself."""
    from maua_b200 import _build

    if not os.path.exists(_build.LIB_PATH):
        try:
            _build.build()
        except Exception as e:  # no nvcc: the tests that need the library will say so themselves
            print(f"conftest: could not build libmaua_b200.so ({e})")


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
