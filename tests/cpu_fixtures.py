"""Shared fixture: run the package's host logic on the CPU emulator (tests/emu_backend.py)."""
import pytest

from flox_b200 import kernels
from tests.emu_backend import EmuKernels


@pytest.fixture
def emu():
    backend = EmuKernels()
    kernels.set_backend(backend)
    yield backend
    kernels.set_backend(None)
