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


@pytest.mark.parametrize("case", N.SCAN, ids=N.ids(N.SCAN))
def test_engine_scans(case):
    N.check_engine_scan(case)


@pytest.mark.parametrize("container", ["numpy", "cuda"])
@pytest.mark.parametrize("case", N.API_SCAN, ids=N.ids(N.API_SCAN))
def test_api_groupby_scan(case, container):
    N.check_api_scan(case, as_torch_device="cuda" if container == "cuda" else None)


@pytest.mark.parametrize("func", ["cumsum", "nancumsum", "ffill", "bfill"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_scans_across_blocks_vs_numpy(func, dtype):
    """3e6 elements (1465 scan blocks of 2048), 7 groups and 100 000 groups, 3 rows: groups continue across many blocks
    (the carry chain of csrc/segments.cu) and many groups start and end inside one block."""
    rng = np.random.default_rng(3)
    n = 3_000_000
    for G in (7, 100_000):
        by = rng.integers(0, G, n)
        if np.dtype(dtype).kind == "f":
            x = (rng.random((3, n)) - 0.3).astype(dtype)
            x[:, rng.random(n) < 0.05] = np.nan
        else:
            x = rng.integers(-9, 10, (3, n)).astype(dtype)
        got = flox_b200.groupby_scan(x, by, func=func, axis=-1)
        order = np.argsort(by, kind="stable")
        sx, sb = x[:, order].astype(np.float64 if np.dtype(dtype).kind == "f" else np.int64), by[order]
        heads = np.flatnonzero(np.concatenate(([True], sb[1:] != sb[:-1])))
        want_sorted = np.empty_like(sx)
        if func in ("cumsum", "nancumsum"):
            z = np.where(np.isnan(sx), 0.0, sx) if sx.dtype.kind == "f" else sx   # cumsum: NaNs are masked in below
            cs = np.cumsum(z, axis=1)
            start_tot = np.concatenate((np.zeros((3, 1), cs.dtype), cs[:, heads[1:] - 1]), axis=1)
            want_sorted = cs - np.repeat(start_tot, np.diff(np.concatenate((heads, [n]))), axis=1)
            if func == "cumsum" and sx.dtype.kind == "f":
                bad = np.cumsum(np.isnan(sx), axis=1)
                bad_start = np.concatenate((np.zeros((3, 1), bad.dtype), bad[:, heads[1:] - 1]), axis=1)
                want_sorted[(bad - np.repeat(bad_start, np.diff(np.concatenate((heads, [n]))), axis=1)) > 0] = np.nan
        else:
            src = sx if func == "ffill" else None
            if func == "bfill":
                continue_sorted = None
            idx = np.where(np.isnan(sx), 0, np.arange(n)) if sx.dtype.kind == "f" else np.broadcast_to(np.arange(n), sx.shape).copy()
            if sx.dtype.kind == "f":
                if func == "ffill":
                    idx[:, heads] = np.broadcast_to(heads, (3, len(heads)))
                    np.maximum.accumulate(idx, axis=1, out=idx)
                    want_sorted = np.take_along_axis(sx, idx, axis=1)
                else:  # bfill: ffill of the reversed row, groups reversed with it
                    rs, rb = sx[:, ::-1], sb[::-1]
                    rheads = np.flatnonzero(np.concatenate(([True], rb[1:] != rb[:-1])))
                    ridx = np.where(np.isnan(rs), 0, np.arange(n))
                    ridx[:, rheads] = np.broadcast_to(rheads, (3, len(rheads)))
                    np.maximum.accumulate(ridx, axis=1, out=ridx)
                    want_sorted = np.take_along_axis(rs, ridx, axis=1)[:, ::-1]
            else:
                want_sorted = sx
        want = np.empty_like(want_sorted)
        want[:, order] = want_sorted
        if np.dtype(dtype).kind == "f" and "cumsum" in func:
            # the numpy CHECKER here subtracts group-start totals from one global f64 cumsum (3e6 sequential adds up to
            # ~6e5): its own error is ~1e-8 absolute, the kernel's per-group f64 scan is tighter
            np.testing.assert_allclose(got, want.astype(dtype), rtol=1e-6 if dtype == np.float32 else 1e-10,
                                       atol=1e-2 if dtype == np.float32 else 1e-6, equal_nan=True)
        else:
            np.testing.assert_array_equal(got, want.astype(got.dtype))


@pytest.mark.parametrize("container", ["numpy", "cuda"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_first_last_family(dtype, container):
    N.check_first_last(dtype, as_torch_device="cuda" if container == "cuda" else None)
