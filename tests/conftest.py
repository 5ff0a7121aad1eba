import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture
def cpu_kernels(monkeypatch):
    """Route saragan_b200.kernels to the torch-CPU emulation in tests/cpu_emul.py so the host
    logic above the C ABI can be exercised without a GPU."""
    import torch
    from saragan_b200 import kernels
    from tests import cpu_emul
    for name in cpu_emul.ALL:
        monkeypatch.setattr(kernels, name, getattr(cpu_emul, name))
    torch.manual_seed(0)
    return cpu_emul


@pytest.fixture(autouse=True)
def _default_leaky_slope():
    """The LeakyReLU slope is a process-wide constant of the kernel library (and of its emulation): tests that
    build network_dict.py models change it, every test starts and ends at network.py's 0.2."""
    yield
    import torch
    from tests import cpu_emul
    cpu_emul.set_leaky_slope(0.2)
    if torch.cuda.is_available():
        from saragan_b200 import kernels
        kernels.ensure_leaky_slope(0.2)
