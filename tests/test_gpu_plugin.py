"""The plug-in seam inside the UNMODIFIED reference on the GPU: baseline/_ref (installed by
oracle/install_reference.py, shipped with the snapshot) + flox_b200.plugin.install() + the CUDA kernels."""
import numpy as np
import pytest
import torch

from flox_b200 import kernels
from oracle import ref_loader
from tests import plugin_cases as P

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.reference_available(), reason="needs baseline/_ref (oracle/install_reference.py)")]


@pytest.fixture(autouse=True)
def _cuda_backend():
    assert torch.cuda.is_available(), "GPU tests need CUDA"
    kernels.set_backend(None)
    assert kernels.get_kernels().name == "cuda"
    yield


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_reference_groupby_reduce_with_engine_b200(dtype):
    l0 = kernels.get_kernels().launches
    P.check_groupby_reduce(P.load_patched_reference(), dtype)
    assert kernels.get_kernels().launches > l0, "engine='b200' did not reach the CUDA kernels"


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_reference_graph_blocks_with_engine_b200(dtype):
    l0 = kernels.get_kernels().launches
    P.check_blocks_through_reference_graph(P.load_patched_reference(), dtype)
    assert kernels.get_kernels().launches > l0
