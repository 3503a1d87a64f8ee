import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def emul_engine():
    """Engine bound to the CPU kernel emulator build of csrc/ (tests/emul; test tooling only)."""
    import parops
    from tests.emul.build_emul import build
    lib = parops._lib.bind(build())
    eng = parops.Engine(device="cpu", lib=lib)
    yield eng
    eng.destroy()


@pytest.fixture(scope="session")
def cuda_engine():
    import torch
    import parops
    assert torch.cuda.is_available(), "-m gpu tests need a GPU"
    eng = parops.default_engine()
    # the CUDA path must be the real library, in-tree
    assert eng.lib.path.endswith(os.path.join("csrc", "libparops.so")) and not eng.emulated
    return eng
