"""Shared bodies of the sorted-segment-family tests, CPU-emulator tier and GPU tier: the engine
functions and the public API against the reference-made fixtures tests/golden/golden_next.* — bit-exact."""
import json
import os

import numpy as np

import flox_b200
from tests.golden_io import assert_same

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(_DIR, "golden_next.json")) as _f:
    CASES = json.load(_f)
ARR = np.load(os.path.join(_DIR, "golden_next.npz"), allow_pickle=False)

QUANT = [c for c in CASES if c["family"] == "quantile"]
SCAN = [c for c in CASES if c["family"] == "scan"]
API_QUANT = [c for c in CASES if c["family"] == "api_quantile"]
API_SCAN = [c for c in CASES if c["family"] == "api_scan"]


def ids(cases):
    return [c["id"] for c in cases]


def check_engine_quantile(case, shuffle: bool):
    """aggregate_b200.<quantile family>(codes, array, q=, size=, fill_value=): the engine contract.  The reference needs
    rows sorted by group (chunk_reduce sorts for engine="flox"); the b200 engine does not — `shuffle` permutes them."""
    codes, arr = ARR[case["codes"]], ARR[case["array"]]
    if shuffle:
        perm = np.random.default_rng(7).permutation(codes.size)
        codes, arr = codes[perm], arr[..., perm]
    fn = getattr(flox_b200.aggregate_b200, case["func"])
    res = fn(codes, arr, size=case["size"], fill_value=np.nan, **case["kw"])
    assert_same(res, ARR[case["result"]], what=case["id"])


def check_api_quantile(case, as_torch_device=None):
    kw = {}
    if "expected_groups" in case:
        kw["expected_groups"] = ARR[case["expected_groups"]]
    if case["fill_value"] is not None:
        kw["fill_value"] = case["fill_value"]
    arr, by = ARR[case["array"]], ARR[case["by"]]
    if as_torch_device is not None:
        import torch

        arr, by = torch.as_tensor(arr, device=as_torch_device), torch.as_tensor(by, device=as_torch_device)
    res, groups = flox_b200.groupby_reduce(arr, by, func=case["func"], finalize_kwargs=case["finalize_kwargs"], **kw)
    if as_torch_device is not None:
        res = res.cpu().numpy()
    assert_same(res, ARR[case["result"]], what=case["id"])
    np.testing.assert_array_equal(np.asarray(groups), ARR[case["groups"]])


def _scan_tolerance(arr, func):
    """cumsum / nancumsum of floats: the reference forms ONE running sum over the code-sorted row and subtracts the total
    preceding each group (aggregate_flox.py:308-316), so its values carry the rounding of that GLOBAL total; the engine
    scans every group on its own in f64 and rounds once.  The two agree within a few ulps of the largest running total."""
    if arr.dtype.kind != "f" or "cumsum" not in func:
        return None
    eps = np.finfo(arr.dtype).eps
    scale = float(np.nanmax(np.nancumsum(np.abs(np.nan_to_num(arr)), axis=-1))) if arr.size else 0.0
    return dict(rtol=4 * eps, atol=4 * eps * scale)


def _assert_scan(res, want, arr, func, what):
    assert res.dtype == want.dtype and res.shape == want.shape, (what, res.dtype, want.dtype, res.shape, want.shape)
    tol = _scan_tolerance(arr, func)
    if tol is None:
        np.testing.assert_array_equal(res, want, err_msg=what)        # integers, ffill / bfill: bit-exact
    else:
        np.testing.assert_allclose(res, want, equal_nan=True, err_msg=what, **tol)
        assert np.array_equal(np.isnan(res), np.isnan(want)), what


def check_engine_scan(case):
    codes, arr = ARR[case["codes"]], ARR[case["array"]]
    res = getattr(flox_b200.aggregate_b200, case["func"])(codes, arr, axis=arr.ndim - 1)
    _assert_scan(res, ARR[case["result"]], arr, case["func"], case["id"])


def check_api_scan(case, as_torch_device=None):
    arr, by = ARR[case["array"]], ARR[case["by"]]
    a, b = arr, by
    if as_torch_device is not None:
        import torch

        a, b = torch.as_tensor(arr, device=as_torch_device), torch.as_tensor(by, device=as_torch_device)
    res = flox_b200.groupby_scan(a, b, func=case["func"], axis=-1)
    if as_torch_device is not None:
        res = res.cpu().numpy()
    _assert_scan(np.asarray(res), ARR[case["result"]], arr, case["func"], case["id"])


def _first_last_numpy(x, by, G, func):
    """numpy semantics: the first / last element of each group in order of appearance (numpy_groupies' "first" / "last",
    the kernels behind flox's first / last); nanfirst / nanlast skip NaNs (flox/xrutils.py:386-412)."""
    out = np.full(x.shape[:-1] + (G,), np.nan)
    for g in range(G):
        sel = x[..., by == g]
        for idx in np.ndindex(x.shape[:-1]):
            row = sel[idx]
            if func.startswith("nan") and row.dtype.kind == "f":
                row = row[~np.isnan(row)]
            if row.size:
                out[idx + (g,)] = row[0] if "first" in func else row[-1]
    return out


def check_first_last(dtype, as_torch_device=None):
    """No in-repo reference kernel exists for these (engine="flox" falls back to numpy_groupies, absent here): pinned to
    numpy semantics like the arg-reductions.  In memory, chunked (map-reduce: positions are global, the
    first occurrence across chunk boundaries wins) and cohorts."""
    rng = np.random.default_rng(0)
    n, G = 20_000, 17
    by = rng.integers(0, G, n)
    by[by == 3] = 4
    by[::401] = -1
    x = (rng.random((3, n)) * 10).astype(dtype)
    if np.dtype(dtype).kind == "f":
        x[:, rng.random(n) < 0.3] = np.nan
        x[1, by == 5] = np.nan
    valid = np.where(by >= 0, by, G + 5)
    for func in ("first", "last", "nanfirst", "nanlast"):
        want = _first_last_numpy(x, valid, G, func)
        present = ~np.isnan(want) if np.dtype(dtype).kind != "f" else None
        for mode in ("memory", "map-reduce", "cohorts"):
            a, b = x, by
            if as_torch_device is not None:
                import torch

                a, b = torch.as_tensor(x, device=as_torch_device), torch.as_tensor(by, device=as_torch_device)
            if mode == "memory":
                got, _ = flox_b200.groupby_reduce(a, b, func=func, expected_groups=np.arange(G))
            else:
                got, _ = flox_b200.groupby_reduce(flox_b200.ChunkedArray.from_array(a, 3001), b, func=func,
                                                  expected_groups=np.arange(G), method=mode)
            got = got.cpu().numpy() if as_torch_device is not None else got
            if present is None:
                assert got.dtype == np.dtype(dtype)
                np.testing.assert_array_equal(got, want.astype(dtype), err_msg=f"{func}/{mode}")
            else:  # integer arrays: absent groups hold the integer fill placeholder
                np.testing.assert_array_equal(got[present], want[present].astype(got.dtype), err_msg=f"{func}/{mode}")
