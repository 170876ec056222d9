import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/b2_oracle.c through ctypes): the checker, never the product."""
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the kernel's traversal logic (tests/emu): test infrastructure."""
    from tests.emu import emu_binding
    emu_binding.build()
    return emu_binding


@pytest.fixture(scope="session")
def has_gpu():
    from mujoco_ros2_simulation_b200 import engine
    return engine.device_count() > 0
