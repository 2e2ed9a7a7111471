"""Host logic of flox_b200 (dtype/fill rules, chunk driver, reindex, API surface) against the
golden vectors the reference produced — run on the CPU kernel emulator (tests/emu_backend.py), so
this file checks everything EXCEPT the CUDA kernels themselves (tests/test_gpu_*.py do that).

Tolerances: results must have the reference's exact dtype and shape; integer / count outputs are
bit-exact; float outputs within 1e-6 relative (f32) / 1e-12 (f64) — the engine accumulates in f64
and rounds once, the reference sums pairwise in the array dtype.
"""
import numpy as np
import pytest

import flox_b200
from tests.cpu_fixtures import emu  # noqa: F401
from tests.golden_io import api_kwargs, assert_same, cond_atol, load_golden

G = load_golden()
RTOL32, RTOL64 = 1e-6, 1e-12


def _ids(cases):
    return [c["id"] for c in cases]


def _tols(case):
    # the precision cases subtract 1e12-sized offsets: tolerance from the reference's own test
    # (tests/test_core.py:2287-2319 uses rtol=3e-8 there)
    arr = G.arr(case["array"])
    at = dict(atol_f32=cond_atol(case["func"], arr, RTOL32), atol_f64=cond_atol(case["func"], arr, RTOL64))
    if case["id"].startswith("api/prec"):
        return dict(rtol_f32=RTOL32, rtol_f64=3e-8, **at)
    return dict(rtol_f32=RTOL32, rtol_f64=RTOL64, **at)


@pytest.mark.parametrize("case", G.family("api"), ids=_ids(G.family("api")))
def test_groupby_reduce_matches_reference(case, emu):  # noqa: F811
    kw = api_kwargs(case)
    fk = {"ddof": case["ddof"]} if case["ddof"] else None
    bys = [G.arr(b) for b in case["bys"]]
    if "raises" in case:
        if case["func"].startswith(("arg", "nanarg")):
            pytest.skip("engine='flox' rejects arg-reductions; engine='b200' implements them (tested separately)")
        with pytest.raises(Exception):
            flox_b200.groupby_reduce(G.arr(case["array"]), *bys, func=case["func"], finalize_kwargs=fk, **kw)
        return
    res, *groups = flox_b200.groupby_reduce(G.arr(case["array"]), *bys, func=case["func"], finalize_kwargs=fk, **kw)
    assert_same(res, G.arr(case["result"]), what=case["id"], **_tols(case))
    for g, key in zip(groups, case["groups"]):
        want = G.arr(key)
        g = np.asarray(g)
        if g.dtype == object and len(g) and hasattr(g[0], "left"):
            g = np.asarray([iv.left for iv in g] + [g[-1].right])
        np.testing.assert_array_equal(g, want)


@pytest.mark.parametrize("case", G.family("engine"), ids=_ids(G.family("engine")))
def test_engine_functions_match_reference(case, emu):  # noqa: F811
    """The plug-in boundary: aggregate_b200.<func>(codes, array, size=, fill_value=, dtype=)."""
    fn = getattr(flox_b200.aggregate_b200, case["func"])
    dt = None if case["dtype"] is None else np.dtype(case["dtype"])
    res = fn(G.arr(case["codes"]), G.arr(case["array"]), size=case["size"], fill_value=case["fill_value"], dtype=dt)
    want = G.arr(case["result"])
    if case["func"] == "nanlen" and want.dtype != res.dtype:
        res = res.astype(want.dtype)
    arr = G.arr(case["array"])
    # an f32 input squared / multiplied in f32 by the reference carries f32 rounding even when the
    # requested output dtype is f64
    r64 = RTOL32 if arr.dtype == np.float32 else RTOL64
    assert_same(res, want, what=case["id"], rtol_f32=RTOL32, rtol_f64=r64,
                atol_f32=cond_atol(case["func"], arr, RTOL32), atol_f64=cond_atol(case["func"], arr, r64))


@pytest.mark.parametrize("label_dtype", ["int64", "int32", "uint8", "int16", "datetime64[D]", "datetime64[ns]"])
def test_unknown_integer_labels_are_discovered_without_a_host_hash(label_dtype, emu):  # noqa: F811
    """expected_groups=None with integer / datetime labels: groups and codes come from the min/max -> presence
    table -> lookup-table kernels (flox_b200/factorize.py:_discover_int_labels), not from pd.factorize; the
    groups (values, order, dtype) and results must equal the reference's (oracle, flox/factorize.py:95-99)."""
    from flox_b200 import factorize
    from oracle import flox_oracle as O

    rng = np.random.default_rng(8)
    x = rng.random(5000).astype(np.float32)
    if label_dtype.startswith("datetime"):
        base = np.datetime64("2001-03-01").astype(label_dtype)
        by = base + rng.integers(-40, 40, 5000).astype("timedelta64[D]").astype(f"timedelta64[{label_dtype[11:-1]}]")
    else:
        lo = 0 if label_dtype == "uint8" else -7
        by = rng.integers(lo, 90, 5000).astype(label_dtype)
    calls = []
    orig = factorize._discover_int_labels
    factorize._discover_int_labels = lambda b: calls.append(1) or orig(b)
    try:
        for func in ("sum", "nanmax", "count"):
            got, groups = flox_b200.groupby_reduce(x, by, func=func)
            want, want_groups = O.groupby_reduce(x, by, func=func)
            want_groups = np.asarray(want_groups)
            assert groups.dtype == want_groups.dtype and np.array_equal(groups, want_groups)
            assert_same(got, want, what=f"discover/{label_dtype}/{func}", rtol_f32=RTOL32, rtol_f64=RTOL64)
    finally:
        factorize._discover_int_labels = orig
    assert len(calls) == 3
    # a value span above 2^26 falls back to the host path and still matches
    wide = by.astype("int64").copy() if not label_dtype.startswith("datetime") else None
    if wide is not None:
        wide[0] = 2**40
        got, groups = flox_b200.groupby_reduce(x, wide, func="sum")
        want, want_groups = O.groupby_reduce(x, wide, func="sum")
        assert np.array_equal(groups, np.asarray(want_groups))
        assert_same(got, want, what="discover/wide", rtol_f32=RTOL32, rtol_f64=RTOL64)


def test_concurrent_calls_from_many_threads(emu):  # noqa: F811
    """dask's threaded scheduler calls chunk_reduce from many threads at once (reference tests/test_core.py:1827-1851):
    the host logic keeps no per-call state in module globals (the encoder cache and the transfer counters are the only
    shared objects, both benign), so concurrent calls with different arguments must not disturb each other."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import flox_oracle as O

    rng = np.random.default_rng(3)
    jobs = []
    for i in range(24):
        n, g = int(rng.integers(50, 400)), int(rng.integers(2, 9))
        x = (rng.random((2, n)) * 6 - 2).astype(np.float32 if i % 2 else np.float64)
        x[:, rng.random(n) < 0.1] = np.nan
        by = rng.integers(0, g, n) * (i % 3 + 1)          # different label sets: different encoders per call
        func = ["nansum", "nanmean", "max", "count", "nanvar", "sum"][i % 6]
        eg = None if i % 4 == 0 else np.unique(by)
        jobs.append((x, by, func, eg))

    def run(job):
        x, by, func, eg = job
        return flox_b200.groupby_reduce(x, by, func=func, expected_groups=eg)

    with ThreadPoolExecutor(max_workers=8) as ex:
        for _ in range(3):
            results = list(ex.map(run, jobs))
            for (x, by, func, eg), (got, groups) in zip(jobs, results):
                want, want_groups = O.groupby_reduce(x, by, func=func, expected_groups=eg)
                np.testing.assert_array_equal(np.asarray(groups), np.asarray(want_groups))
                r64 = RTOL32 if x.dtype == np.float32 else RTOL64
                assert_same(got, want, what=f"threads/{func}", rtol_f32=RTOL32, rtol_f64=r64,
                            atol_f32=cond_atol(func, x, RTOL32), atol_f64=cond_atol(func, x, r64))


# ---- advisor findings (round 1): values below are the reference's own answers ----------------------
def _reference_or_none():
    from oracle import ref_loader

    return ref_loader.load_reference() if ref_loader.reference_available() else None


@pytest.mark.parametrize("func,want", [("sum", [16.0, 2.0, 4.0]), ("count", [1, 1, 1]), ("max", [16.0, 2.0, 4.0])])
def test_fractional_float_labels_are_dropped_not_truncated(func, want, emu):  # noqa: F811
    """float `by` + integer 0..G-1 expected_groups: 0.5 / 1.5 belong to NO group (factorize.py:84-94)."""
    x = np.array([1.0, 2, 4, 8, 16])
    by = np.array([0.5, 1, 2, 1.5, 0])
    got, groups = flox_b200.groupby_reduce(x, by, func=func, expected_groups=np.array([0, 1, 2]))
    np.testing.assert_array_equal(got, np.asarray(want))
    np.testing.assert_array_equal(groups, [0, 1, 2])
    ref = _reference_or_none()
    if ref is not None:
        rwant, _ = ref.groupby_reduce(x, by, func=func, expected_groups=np.array([0, 1, 2]), engine="flox")
        np.testing.assert_array_equal(got, rwant)


@pytest.mark.parametrize("kind", ["M8[ns]", "m8[ns]"])
def test_count_of_datetimes_excludes_nat(kind, emu):  # noqa: F811
    d = np.array([1, np.iinfo(np.int64).min, 3, 4], dtype=np.int64).view(kind)  # INT64_MIN is NaT
    by = np.array([0, 0, 1, 1])
    got, _ = flox_b200.groupby_reduce(d, by, func="count")
    assert got.dtype == np.dtype(np.intp)
    # timedelta64 subclasses np.integer: the reference's notnull() counts its NaT as valid (xrutils.py:149-160)
    np.testing.assert_array_equal(got, [1, 2] if kind[0] == "M" else [2, 2])
    ref = _reference_or_none()
    if ref is not None:
        rwant, _ = ref.groupby_reduce(d, by, func="count", engine="flox")
        np.testing.assert_array_equal(got, rwant)
        assert got.dtype == rwant.dtype


def test_in_memory_labels_next_to_chunked_ones(emu):  # noqa: F811
    """a plain `by` mixed with a chunked one is sliced by chunk boundaries (no per-element objects)."""
    rng = np.random.default_rng(3)
    x = rng.random(1000)
    b0, b1 = rng.integers(0, 3, 1000), rng.integers(0, 4, 1000)
    xa = flox_b200.ChunkedArray.from_array(x, 300)
    got, *_ = flox_b200.groupby_reduce(xa, flox_b200.ChunkedArray.from_array(b0, 300), b1, func="sum",
                                       expected_groups=(np.arange(3), np.arange(4)), method="map-reduce")
    want, *_ = flox_b200.groupby_reduce(x, b0, b1, func="sum", expected_groups=(np.arange(3), np.arange(4)))
    np.testing.assert_allclose(np.asarray(got), np.nan_to_num(want), rtol=1e-12)


def test_parallel_context_is_per_thread():
    """`parallel.distributed()` entered by one thread must not change what other threads see (dask's threaded
    scheduler runs chunk tasks concurrently; the reference's race regression is tests/test_core.py:1827-1851)."""
    import threading

    from flox_b200 import parallel

    seen = {}
    inside, release = threading.Event(), threading.Event()

    def holder():
        with parallel.distributed(sharding="local") as ctx:
            seen["holder"] = parallel.current() is ctx and parallel.current().sharding == "local"
            inside.set()
            release.wait(10)

    t = threading.Thread(target=holder)
    t.start()
    assert inside.wait(10)
    seen["other"] = parallel.current().sharding
    release.set()
    t.join()
    assert seen["holder"] and seen["other"] == "replicated"
    assert parallel.current().sharding == "replicated"
