"""Seeded randomized sweep of the public API against the CPU oracle, on the kernel emulator.

The golden vectors (tests/golden) replay what the reference's own tests exercise; this file draws argument
combinations those tables do not enumerate — shapes with kept leading dims, empty inputs, labels that miss
expected groups or fall outside them, bins, fill values, min_count, dtype overrides, ddof, chunked map-reduce —
and requires the same result, dtype and groups as the oracle (itself pinned bit-exactly to the reference,
tests/test_oracle_pinning.py).  Same bars as everywhere: exact for integer results, 1e-6 / 1e-12 relative for floats.
"""
import numpy as np
import pytest

import flox_b200
from flox_b200 import ChunkedArray
from oracle import flox_oracle as O
from tests.cpu_fixtures import emu  # noqa: F401
from tests.golden_io import assert_same, cond_atol

FUNCS = ["count", "sum", "nansum", "prod", "nanprod", "mean", "nanmean", "max", "nanmax", "min", "nanmin",
         "var", "nanvar", "std", "nanstd", "argmax", "nanargmin"]
RTOL32, RTOL64 = 1e-6, 1e-12


def _draw_case(rng):
    dtype = rng.choice(["float32", "float64", "int32", "int64"])
    n = int(rng.choice([0, 1, 2, 7, 33, 120, 257]))
    lead = [(), (3,), (2, 4)][int(rng.integers(0, 3))]
    ngroups = int(rng.integers(1, 9))
    if dtype.startswith("float"):
        x = (rng.random(lead + (n,)) * 8 - 3).astype(dtype)
        if n and rng.random() < 0.6:
            x[..., rng.random(n) < 0.2] = np.nan
    else:
        x = rng.integers(-9, 10, lead + (n,)).astype(dtype)
    by = rng.integers(-1 if rng.random() < 0.5 else 0, ngroups + (2 if rng.random() < 0.3 else 0), n)
    func = str(rng.choice(FUNCS))
    kw = {}
    eg_kind = rng.choice(["none", "arange", "subset", "bins"])
    if eg_kind == "arange":
        kw["expected_groups"] = np.arange(ngroups + int(rng.integers(0, 3)))
    elif eg_kind == "subset":
        kw["expected_groups"] = np.sort(rng.choice(np.arange(-1, ngroups + 2), size=min(ngroups, 3), replace=False))
    elif eg_kind == "bins":
        kw["expected_groups"] = np.array([-0.5, 1.5, 3.5, ngroups + 0.5])
        kw["isbin"] = True
    if rng.random() < 0.35:
        kw["fill_value"] = [0, -1, np.nan, 123][int(rng.integers(0, 4))]
    if rng.random() < 0.3:
        kw["min_count"] = int(rng.integers(1, 4))
    if rng.random() < 0.2 and not func.startswith(("arg", "nanarg")) and func != "count":
        kw["dtype"] = np.dtype(rng.choice(["float64", "float32"]))
    ddof = int(rng.integers(0, 2)) if ("var" in func or "std" in func) else 0
    return x, by, func, kw, ddof


def _compare(got, want, x, func, what):
    for g, w in zip(got[1:], want[1:]):
        g, w = np.asarray(g), np.asarray(w)
        if w.dtype == object or g.dtype == object:  # interval groups
            assert len(g) == len(w), what
        else:
            np.testing.assert_array_equal(g, w, err_msg=what)
    r64 = RTOL32 if x.dtype == np.float32 else RTOL64
    g0, w0 = np.asarray(got[0]), np.asarray(want[0])
    if ("var" in func or "std" in func) and w0.dtype.kind == "f":
        # count == ddof: the reference divides rounding noise (or exactly 0) by 0 and returns inf or nan depending on
        # the last bit of the group mean; neither value means anything, treat them as the same "undefined"
        g0, w0 = np.where(np.isinf(g0), np.nan, g0), np.where(np.isinf(w0), np.nan, w0)
    got, want = (g0,) + tuple(got[1:]), (w0,) + tuple(want[1:])
    assert_same(got[0], want[0], what=what, rtol_f32=RTOL32, rtol_f64=r64,
                atol_f32=cond_atol(func, x, RTOL32), atol_f64=cond_atol(func, x, r64))


@pytest.mark.parametrize("seed", range(12))
def test_random_api_cases_match_oracle(seed, emu):  # noqa: F811
    rng = np.random.default_rng(1000 + seed)
    checked = 0
    for i in range(40):
        x, by, func, kw, ddof = _draw_case(rng)
        what = f"seed{seed}/{i}: {func} {x.dtype}{x.shape} kw={kw} ddof={ddof}"
        try:
            want = O.groupby_reduce(x, by, func=func, ddof=ddof, **kw)
        except Exception as err:  # the reference rejects the combination: so must the package
            with pytest.raises(Exception):
                flox_b200.groupby_reduce(x, by, func=func, finalize_kwargs={"ddof": ddof} if ddof else None, **kw)
            assert not isinstance(err, AssertionError), what
            continue
        got = flox_b200.groupby_reduce(x, by, func=func, finalize_kwargs={"ddof": ddof} if ddof else None, **kw)
        _compare(got, want, x, func, what)
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("seed", range(6))
def test_random_mapreduce_cases_match_oracle(seed, emu):  # noqa: F811
    """method="map-reduce" over ragged chunks against the oracle's chunk_reduce + simple_combine + finalize."""
    rng = np.random.default_rng(5000 + seed)
    for i in range(25):
        x, by, func, kw, ddof = _draw_case(rng)
        if func.startswith(("arg", "nanarg")) or x.shape[-1] < 4 or kw.get("expected_groups") is None:
            continue
        if "min_count" in kw and "fill_value" not in kw:
            # the reference raises "Filling is required" here (core.py:437-447); the oracle's map-reduce helper
            # substitutes the final fill like the golden generator does (tests/golden/make_golden.py:123-124)
            continue
        n = x.shape[-1]
        nchunks = int(rng.integers(2, 5))
        bounds = np.linspace(0, n, nchunks + 1).astype(int)   # the slabs the oracle uses
        arr = ChunkedArray([x[..., lo:hi] for lo, hi in zip(bounds[:-1], bounds[1:])])
        what = f"mr seed{seed}/{i}: {func} {x.dtype}{x.shape} chunks={nchunks} kw={kw} ddof={ddof}"
        try:
            want = O.groupby_reduce_mapreduce(x, by, func=func, nchunks=nchunks, ddof=ddof, **kw)
        except Exception:
            continue
        got = flox_b200.groupby_reduce(arr, by, func=func, method="map-reduce",
                                       finalize_kwargs={"ddof": ddof} if ddof else None, **kw)
        _compare(got, want, x, func, what)


@pytest.mark.parametrize("seed", range(8))
def test_random_multidim_and_multi_by_cases_match_oracle(seed, emu):  # noqa: F811
    """N-d labels, reductions over a subset of the label axes, two groupers, float labels with NaN, bool values."""
    rng = np.random.default_rng(9000 + seed)
    checked = 0
    for i in range(30):
        shape = [(4, 6), (3, 4, 5), (2, 9)][int(rng.integers(0, 3))]
        kind = rng.choice(["f32", "i64", "bool"])
        if kind == "f32":
            x = (rng.random(shape) * 5 - 1).astype(np.float32)
            x[rng.random(shape) < 0.15] = np.nan
        elif kind == "i64":
            x = rng.integers(-5, 6, shape)
        else:
            x = rng.random(shape) < 0.5
        func = str(rng.choice(["sum", "nansum", "count", "mean", "nanmax", "min", "nanvar", "max"]))
        mode = rng.choice(["by_nd_all", "by_nd_axis", "two_by", "float_labels"])
        kw = {}
        if mode == "by_nd_all":
            bys = (rng.integers(0, 4, shape[-2:]),)
        elif mode == "by_nd_axis":
            bys = (rng.integers(0, 3, shape[-2:]),)
            kw["axis"] = -1
        elif mode == "two_by":
            bys = (rng.integers(0, 3, shape[-1]), rng.integers(0, 2, shape[-1]))
            if rng.random() < 0.5:
                kw["expected_groups"] = (np.arange(3), np.arange(2))
        else:
            lab = rng.integers(0, 4, shape[-1]).astype(np.float64)
            lab[rng.random(shape[-1]) < 0.2] = np.nan
            bys = (lab,)
        if rng.random() < 0.3:
            kw["fill_value"] = [0, np.nan, -7][int(rng.integers(0, 3))]
        what = f"nd seed{seed}/{i}: {func} {x.dtype}{x.shape} {mode} kw={kw}"
        try:
            want = O.groupby_reduce(x, *bys, func=func, **kw)
        except Exception:
            with pytest.raises(Exception):
                flox_b200.groupby_reduce(x, *bys, func=func, **kw)
            continue
        got = flox_b200.groupby_reduce(x, *bys, func=func, **kw)
        _compare(got, want, x if x.dtype != bool else x.astype(np.int64), func, what)
        checked += 1
    assert checked >= 15
