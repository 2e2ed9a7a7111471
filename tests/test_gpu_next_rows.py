"""Sorted-segment family on the GPU: device radix sort of the (group, value) keys + csrc/segments.cu against the
reference-made fixtures (bit-exact), and against numpy at a size where the sort matters."""
import numpy as np
import pytest
import torch

import flox_b200
from flox_b200 import kernels
from tests import next_cases as N

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cuda_backend():
    assert torch.cuda.is_available(), "GPU tests need CUDA"
    kernels.set_backend(None)
    assert kernels.get_kernels().name == "cuda"
    yield


@pytest.mark.parametrize("shuffle", [False, True], ids=["sorted", "shuffled"])
@pytest.mark.parametrize("case", N.QUANT, ids=N.ids(N.QUANT))
def test_engine_quantile_family(case, shuffle):
    N.check_engine_quantile(case, shuffle)


@pytest.mark.parametrize("container", ["numpy", "cuda"])
@pytest.mark.parametrize("case", N.API_QUANT[::2], ids=N.ids(N.API_QUANT[::2]))
def test_api_quantile_family(case, container):
    N.check_api_quantile(case, as_torch_device="cuda" if container == "cuda" else None)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_quantiles_vs_numpy_at_size(dtype):
    """2e6 values, 500 random groups, several quantiles: per group np.nanquantile(method="linear") — equal up to the last
    bit of the interpolation weight (the reference adds the running count of earlier groups into the virtual index)."""
    rng = np.random.default_rng(1)
    n, G = 2_000_000, 500
    by = rng.integers(0, G, n)
    if np.dtype(dtype).kind == "f":
        x = (rng.standard_normal(n) * 100).astype(dtype)
        x[rng.random(n) < 0.01] = np.nan
    else:
        x = rng.integers(-10_000, 10_000, n).astype(dtype)
    q = [0.0, 0.01, 0.5, 0.999, 1.0]
    got, _ = flox_b200.groupby_reduce(x, by, func="nanquantile", finalize_kwargs={"q": q}, expected_groups=np.arange(G))
    assert got.shape == (len(q), G) and got.dtype == np.float64
    order = np.argsort(by, kind="stable")
    bounds = np.searchsorted(by[order], np.arange(G + 1))
    xs = x[order].astype(np.float64)
    for g in (0, 1, 17, 250, 499):
        want = np.nanquantile(xs[bounds[g]:bounds[g + 1]], q)
        np.testing.assert_allclose(got[:, g], want, rtol=1e-12, atol=1e-9)
    med, _ = flox_b200.groupby_reduce(x, by, func="nanmedian", expected_groups=np.arange(G))
    np.testing.assert_allclose(med[17], np.nanmedian(xs[bounds[17]:bounds[18]]), rtol=1e-6)
