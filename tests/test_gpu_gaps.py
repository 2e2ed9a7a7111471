"""GPU-tier coverage of paths that only the CPU emulator exercised in round 1:
chunked method="blockwise" / "cohorts", chunked arg-reductions with ties across chunk boundaries, the explicit
combine kernels (floxb_combine / floxb_var_combine) behind `_simple_combine`, sorted labels at large G."""
import numpy as np
import pandas as pd
import pytest
import torch

import flox_b200
from flox_b200 import ChunkedArray, kernels
from flox_b200.aggregations import _initialize_aggregation
from flox_b200.core import chunk_reduce
from flox_b200.mapreduce import _aggregate, _expand_dims, _simple_combine
from flox_b200.reindex import ReindexStrategy
from oracle import flox_oracle as O
from tests.golden_io import assert_same, cond_atol, load_golden

pytestmark = pytest.mark.gpu
G = load_golden()
MAPRED = G.family("mapred")


@pytest.fixture(autouse=True)
def _cuda_backend():
    assert torch.cuda.is_available(), "GPU tests need CUDA"
    kernels.set_backend(None)
    k = kernels.get_kernels()
    assert k.name == "cuda"
    k.force_path = 0
    yield k
    k.force_path = 0


def _tols(func, arr):
    r64 = 1e-6 if arr.dtype == np.float32 else 1e-12
    return dict(rtol_f32=1e-6, rtol_f64=r64, atol_f32=cond_atol(func, arr, 1e-6), atol_f64=cond_atol(func, arr, r64))


@pytest.mark.parametrize("method", ["blockwise", "cohorts"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_chunked_blockwise_and_cohorts_vs_oracle(method, dtype):
    """Chunked inputs with method="blockwise" / "cohorts" finish as the numpy-path value (dask.py:390-392,469-518): final
    fills and dtypes, absent groups included — against the oracle's single-shot result."""
    rng = np.random.default_rng(8)
    n, ngroups = 400_003, 5000
    by = np.sort(rng.integers(0, ngroups, n))
    by[by == 44] = 45
    by[::997] = -1
    if np.dtype(dtype).kind == "f":
        x = (rng.random(n) * 6 - 2).astype(dtype)
        x[rng.random(n) < 0.02] = np.nan
    else:
        x = rng.integers(-500, 500, n).astype(dtype)
    for container in ("numpy", "cuda"):
        xa, ba = (x, by) if container == "numpy" else (torch.as_tensor(x, device="cuda"), torch.as_tensor(by, device="cuda"))
        arr, lab = ChunkedArray.from_array(xa, 50_000), ChunkedArray.from_array(ba, 50_000)
        for func in ("sum", "nansum", "count", "mean", "nanmean", "max", "nanmin", "var", "nanvar", "nanstd", "prod"):
            if func == "prod" and np.dtype(dtype).kind != "f":
                continue
            fk = {"ddof": 1} if "var" in func else None
            xin = np.where(np.abs(x) > 1.001, 1.0, x).astype(dtype) if func == "prod" else x
            arr_f = arr if func != "prod" else ChunkedArray.from_array(xin if container == "numpy" else torch.as_tensor(xin, device="cuda"), 50_000)
            want, _ = O.groupby_reduce(xin, by, func=func, expected_groups=np.arange(ngroups), ddof=1 if fk else 0)
            got, _ = flox_b200.groupby_reduce(arr_f, lab, func=func, expected_groups=np.arange(ngroups), method=method,
                                              finalize_kwargs=fk)
            got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
            assert_same(got, want, what=f"{method}/{func}/{np.dtype(dtype)}/{container}", **_tols(func, xin))


@pytest.mark.parametrize("func", ["argmax", "argmin", "nanargmax", "nanargmin"])
@pytest.mark.parametrize("method", ["map-reduce", "cohorts"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_chunked_arg_reductions_ties_across_chunks(func, method, dtype):
    """Positions are global along the reduced axis (chunk_argreduce's index fix-up, core.py:157-211; _grouped_combine,
    dask.py:198-322) and the FIRST occurrence wins even when the same extreme value sits in several chunks."""
    rng = np.random.default_rng(5)
    n, ngroups = 120_000, 61
    by = rng.integers(0, ngroups, n)
    by[by == 9] = 10
    x = rng.integers(0, 6, n).astype(dtype)          # 6 distinct values: every chunk holds every group's extreme
    if np.dtype(dtype).kind == "f":
        x[rng.random(n) < 0.03] = np.nan
        x[by == 20] = np.nan                         # an all-NaN group
    want, _ = O.groupby_reduce(x, by, func=func, expected_groups=np.arange(ngroups))
    for chunk in (7_001, 40_000):
        got, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(x, chunk), by, func=func, expected_groups=np.arange(ngroups),
                                          method=method)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)
        xt, bt = torch.as_tensor(x, device="cuda"), torch.as_tensor(by, device="cuda")
        got_t, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(xt, chunk), ChunkedArray.from_array(bt, chunk), func=func,
                                            expected_groups=np.arange(ngroups), method=method)
        np.testing.assert_array_equal(got_t.cpu().numpy(), want)


@pytest.mark.parametrize("case", MAPRED[::3], ids=[c["id"] for c in MAPRED[::3]])
def test_explicit_combine_kernels_golden(case):
    """chunk_reduce per slab -> _expand_dims -> _simple_combine (floxb_combine / floxb_var_combine: dask.py:131-143,
    aggregations.py:392-451) -> _aggregate, exactly the reference's graph, against the reference-made goldens —
    combined intermediates included."""
    a, by = G.arr(case["array"]), G.arr(case["by"])
    bounds = np.linspace(0, a.shape[-1], case["nchunks"] + 1).astype(int)
    fk = {"ddof": case["ddof"]} if case["ddof"] else None
    dt = None if case["dtype"] is None else np.dtype(case["dtype"])
    agg = _initialize_aggregation(case["func"], dt, a.dtype, case["fill_value"], case["min_count"], fk)
    if agg.fill_value["user"] is None:
        agg.fill_value["user"] = agg.fill_value[agg.name]
    expected = pd.RangeIndex(case["ngroups"])
    axis = (a.ndim - 1,)
    parts = [_expand_dims(chunk_reduce(a[..., lo:hi], by[lo:hi], func=agg.chunk, expected_groups=expected, axis=axis,
                                       fill_value=agg.fill_value["intermediate"], dtype=agg.dtype["intermediate"],
                                       reindex=True)) for lo, hi in zip(bounds[:-1], bounds[1:])]
    launches0 = kernels.get_kernels().launches
    combined = _simple_combine(parts, agg, axis, keepdims=True, is_aggregate=True)
    assert kernels.get_kernels().launches > launches0, "the combine kernels did not run"
    flat = []
    for c in combined["intermediates"]:
        flat.extend(c.arrays if hasattr(c, "arrays") else [c])
    for got, key in zip(flat, case["intermediates"]):
        want = G.arr(key)
        assert_same(np.asarray(got).reshape(want.shape), want, what=case["id"] + "/inter", **_tols(case["func"], a))
    out = _aggregate(parts, _simple_combine, agg, expected, axis, keepdims=True, fill_value=None,
                     reindex=ReindexStrategy(blockwise=True))
    assert_same(out[agg.name], G.arr(case["result"]), what=case["id"], **_tols(case["func"], a))


@pytest.mark.parametrize("func", ["sum", "count", "nanmean", "nanmax", "nanvar"])
def test_sorted_labels_large_G(func):
    """BASELINE config 5's shape at oracle size: 5e6 rows, G = 1e6, sorted labels, chunked, method="cohorts" — the
    run-merging L2-atomic kernels (k_moments_global_runs) and their min/max / sqdev siblings."""
    rng = np.random.default_rng(55)
    n, ngroups = 5_000_000, 1_000_000
    by = np.sort(rng.integers(0, ngroups, n))
    x = rng.random(n, dtype=np.float32)
    x[::53] = np.nan
    fk = {"ddof": 1} if "var" in func else None
    want, _ = O.groupby_reduce(x, by, func=func, expected_groups=pd.RangeIndex(ngroups), ddof=1 if fk else 0)
    xt, bt = torch.as_tensor(x, device="cuda"), torch.as_tensor(by, device="cuda")
    got, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(xt, 1_000_000), ChunkedArray.from_array(bt, 1_000_000), func=func,
                                      expected_groups=pd.RangeIndex(ngroups), method="cohorts", finalize_kwargs=fk)
    assert_same(got.cpu().numpy(), want, what=f"c5/{func}", **_tols(func, x))
