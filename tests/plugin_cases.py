"""Shared body of the plug-in seam tests (CPU emulator tier and GPU tier): the UNMODIFIED reference, patched at run
time by flox_b200.plugin.install(), must give engine="b200" results equal to its own engine="flox" results."""
import numpy as np
import pandas as pd

FUNCS = ["sum", "nansum", "prod", "nanprod", "max", "nanmax", "min", "nanmin", "count", "mean", "nanmean",
         "var", "nanvar", "std", "nanstd"]


def load_patched_reference():
    from oracle import ref_loader

    import flox_b200.plugin as plugin

    flox = ref_loader.load_reference(prefer="installed")
    plugin.install(flox)
    return flox


def _close(got, want, func, arr):
    assert got.dtype == want.dtype, (func, got.dtype, want.dtype)
    assert got.shape == want.shape, (func, got.shape, want.shape)
    if want.dtype.kind in "iub":
        np.testing.assert_array_equal(got, want, err_msg=func)
        return
    rtol = 1e-6 if arr.dtype == np.float32 else 1e-12
    scale = float(np.nanmax(np.abs(arr))) if arr.size else 1.0
    power = 2 if ("var" in func) else 1
    atol = rtol * (scale ** power) * (arr.shape[-1] if "prod" not in func else 1)
    if "prod" in func:
        atol = 0.0
        rtol = 1e-5 if arr.dtype == np.float32 else 1e-10
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, equal_nan=True, err_msg=func)


def inputs(dtype, seed=12345):
    rng = np.random.default_rng(seed)
    n, G = 6000, 37
    by = rng.integers(0, G, n)
    by[by == 5] = 6
    if np.dtype(dtype).kind == "f":
        x = (rng.random((3, n)) * 2 + 0.5).astype(dtype)
        x[1, rng.random(n) < 0.05] = np.nan
    else:
        x = rng.integers(1, 3, (3, n)).astype(dtype)
    return x, by, G


def check_groupby_reduce(flox, dtype):
    """flox.groupby_reduce(..., engine="b200") == flox.groupby_reduce(..., engine="flox") for the hot-path funcs,
    with and without expected_groups / fill_value / min_count."""
    x, by, G = inputs(dtype)
    for func in FUNCS:
        fk = {"ddof": 1} if ("var" in func or "std" in func) else None
        xin = x if "prod" not in func else np.where(np.isnan(x.astype(np.float64)), x, 1 + (x - 1) * 1e-3).astype(dtype)
        for kw in (dict(), dict(expected_groups=np.arange(G)), dict(expected_groups=np.arange(G), fill_value=-7, min_count=3)):
            want, *wg = flox.groupby_reduce(xin, by, func=func, engine="flox", finalize_kwargs=fk, **kw)
            got, *gg = flox.groupby_reduce(xin, by, func=func, engine="b200", finalize_kwargs=fk, **kw)
            _close(np.asarray(got), np.asarray(want), func, xin)
            for a, b in zip(gg, wg):
                np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


def check_blocks_through_reference_graph(flox, dtype, nchunks=5):
    """What a dask graph executes, minus the scheduler: the reference's chunk_reduce(engine="b200") per block, its
    agg.simple_combine over the stacked intermediates (dask.py:131-143), its _finalize_results — against the same graph
    with engine="flox"."""
    from flox.aggregations import _initialize_aggregation
    from flox.core import DUMMY_AXIS, _finalize_results, chunk_reduce
    from flox.reindex import ReindexStrategy

    x, by, G = inputs(dtype, seed=99)
    expected = pd.RangeIndex(G)
    bounds = np.linspace(0, x.shape[-1], nchunks + 1).astype(int)
    for func in FUNCS:
        if "prod" in func:
            continue
        fk = {"ddof": 1} if ("var" in func or "std" in func) else None
        results = {}
        for engine in ("flox", "b200"):
            agg = _initialize_aggregation(func, None, x.dtype, None, 0, fk)
            if agg.fill_value["user"] is None:
                agg.fill_value["user"] = agg.fill_value[agg.name]
            parts = []
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                r = chunk_reduce(x[..., lo:hi], by[lo:hi], func=agg.chunk, expected_groups=expected, axis=(x.ndim - 1,),
                                 fill_value=agg.fill_value["intermediate"], dtype=agg.dtype["intermediate"], reindex=True,
                                 engine=engine, user_dtype=agg.dtype["user"])
                parts.append([np.expand_dims(v, DUMMY_AXIS) for v in r["intermediates"]])
            combined = []
            for i, comb in enumerate(agg.simple_combine):
                stacked = np.concatenate([p[i] for p in parts], axis=DUMMY_AXIS)
                combined.append(comb(stacked, axis=(stacked.ndim - 2,), keepdims=True).squeeze(stacked.ndim - 2))
            fin = _finalize_results({"groups": expected, "intermediates": combined}, agg, (x.ndim - 1,), expected,
                                    reindex=ReindexStrategy(blockwise=True))
            results[engine] = np.asarray(fin[agg.name])
        _close(results["b200"], results["flox"], func, x)
