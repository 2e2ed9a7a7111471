"""Generate golden vectors from the UNMODIFIED reference (flox @ /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/golden_cases.json (parameters) + golden_arrays.npz (inputs and
the reference's outputs).  The GPU box has no /root/reference; tests there read
only these committed files.

Three families:
  api/*     flox.core.groupby_reduce(..., engine="flox") on numpy inputs
            (reference entry: flox/core.py:739).
  engine/*  flox.aggregate_flox.<func>(sorted_codes, array, size=, fill_value=, dtype=)
            — the engine plug-in boundary (flox/aggregate_flox.py:133-266).
  mapred/*  per-chunk flox.core.chunk_reduce(func=agg.chunk, ...) on slabs, stacked on
            DUMMY_AXIS, reduced with the reference's own agg.simple_combine and
            finished with flox.core._finalize_results — i.e. flox/dask.py:90-175 with the
            dask plumbing (absent here) replaced by a python loop.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import load_reference  # noqa: E402

flox = load_reference()
import pandas as pd  # noqa: E402
from flox import aggregate_flox  # noqa: E402
from flox.aggregations import _initialize_aggregation  # noqa: E402
from flox.core import DUMMY_AXIS, _finalize_results, chunk_reduce, groupby_reduce  # noqa: E402
from flox.reindex import ReindexStrategy  # noqa: E402

RNG = np.random.default_rng(12345)  # the reference tests' seed (tests/test_core.py:211)
ARRAYS: dict[str, np.ndarray] = {}
CASES: list[dict] = []

LABELS = np.array([0, 0, 2, 2, 2, 1, 1, 2, 2, 1, 1, 0])
NAN_LABELS = LABELS.astype(float)
NAN_LABELS[:5] = np.nan

FUNCS = ["count", "sum", "nansum", "prod", "nanprod", "mean", "nanmean", "var", "nanvar", "std", "nanstd",
         "max", "nanmax", "min", "nanmin"]


_BY_HASH: dict[str, str] = {}


def _put(name, arr):
    """Store once per distinct content; return the key of the stored copy."""
    import hashlib

    arr = np.array(arr, order="C", copy=True)  # private copy: the reference mutates some results in place
    h = hashlib.sha1(arr.tobytes() + str(arr.dtype).encode() + str(arr.shape).encode()).hexdigest()
    if h not in _BY_HASH:
        key = f"a{len(_BY_HASH)}"
        _BY_HASH[h] = key
        ARRAYS[key] = arr
    return _BY_HASH[h]


def _jsonable(v):
    if isinstance(v, dict):
        return {k: _jsonable(x) for k, x in v.items()}
    if isinstance(v, type) or isinstance(v, np.dtype):
        return str(np.dtype(v))
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, float) and np.isnan(v):
        return "nan"
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (tuple, list)):
        return [_jsonable(x) for x in v]
    return v


def add_api(cid, array, bys, func, **kw):
    fk = kw.pop("finalize_kwargs", None)
    try:
        res, *groups = groupby_reduce(array, *bys, func=func, engine="flox", finalize_kwargs=fk, **kw)
    except Exception as e:  # record the error contract too
        CASES.append(dict(id=cid, family="api", func=func, kw=_jsonable(kw), ddof=(fk or {}).get("ddof", 0),
                          array=_put(f"{cid}/array", array), bys=[_put(f"{cid}/by{i}", b) for i, b in enumerate(bys)],
                          raises=type(e).__name__))
        return
    groups = [g.to_numpy() if isinstance(g, pd.Index) else np.asarray(g) for g in groups]
    groups = [np.asarray([iv.left for iv in g] + [g[-1].right]) if g.dtype == object and len(g) and hasattr(g[0], "left") else g
              for g in groups]
    CASES.append(dict(
        id=cid, family="api", func=func, kw=_jsonable(kw), ddof=(fk or {}).get("ddof", 0),
        array=_put(f"{cid}/array", array), bys=[_put(f"{cid}/by{i}", b) for i, b in enumerate(bys)],
        result=_put(f"{cid}/result", res), result_dtype=str(res.dtype),
        groups=[_put(f"{cid}/groups{i}", g) for i, g in enumerate(groups)],
    ))


def add_engine(cid, func, codes, array, size, fill_value, dtype):
    order = np.argsort(codes, kind="stable")
    sc, sa = codes[order], array[..., order]
    try:
        res = getattr(aggregate_flox, func)(sc, sa, axis=-1, size=size, fill_value=fill_value,
                                           dtype=None if dtype is None else np.dtype(dtype))
    except Exception:
        return  # e.g. in-place int mean: not a valid engine call, the API never issues it
    CASES.append(dict(
        id=cid, family="engine", func=func, size=int(size), fill_value=_jsonable(fill_value),
        dtype=None if dtype is None else str(np.dtype(dtype)),
        codes=_put(f"{cid}/codes", codes), array=_put(f"{cid}/array", array),
        result=_put(f"{cid}/result", res), result_dtype=str(res.dtype),
    ))


def add_mapred(cid, array, by, func, nchunks, ngroups, fill_value=None, min_count=0, ddof=0, dtype=None):
    fk = {"ddof": ddof} if "var" in func or "std" in func else None
    agg = _initialize_aggregation(func, dtype, array.dtype, fill_value, min_count, fk)
    if agg.fill_value["user"] is None:
        agg.fill_value["user"] = agg.fill_value[agg.name]
    expected = pd.RangeIndex(ngroups)
    bounds = np.linspace(0, array.shape[-1], nchunks + 1).astype(int)
    parts = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        r = chunk_reduce(array[..., lo:hi], by[lo:hi], func=agg.chunk, expected_groups=expected, axis=(array.ndim - 1,),
                         fill_value=agg.fill_value["intermediate"], dtype=agg.dtype["intermediate"], reindex=True,
                         engine="flox", user_dtype=agg.dtype["user"])
        parts.append([np.expand_dims(x, DUMMY_AXIS) for x in r["intermediates"]])
    combined = []
    for i, comb in enumerate(agg.simple_combine):
        stacked = np.concatenate([p[i] for p in parts], axis=DUMMY_AXIS)
        out = comb(stacked, axis=(stacked.ndim - 2,), keepdims=True)
        combined.append(out.squeeze(stacked.ndim - 2))
    # keep the raw combined intermediates too (pins _simple_combine/_var_combine numerics)
    inter_names = []
    for i, c in enumerate(combined):
        if hasattr(c, "arrays"):
            for j, a in enumerate(c.arrays):
                inter_names.append(_put(f"{cid}/inter{i}_{j}", a))
        else:
            inter_names.append(_put(f"{cid}/inter{i}", c))
    fin = _finalize_results({"groups": expected, "intermediates": combined}, agg, (array.ndim - 1,), expected,
                            reindex=ReindexStrategy(blockwise=True))
    res = fin[agg.name]
    CASES.append(dict(
        id=cid, family="mapred", func=func, nchunks=nchunks, ngroups=ngroups, fill_value=_jsonable(fill_value),
        min_count=min_count, ddof=ddof, dtype=None if dtype is None else str(np.dtype(dtype)),
        array=_put(f"{cid}/array", array), by=_put(f"{cid}/by", by), intermediates=inter_names,
        result=_put(f"{cid}/result", res), result_dtype=str(res.dtype),
    ))


def with_nans(a, frac=0.2):
    a = a.copy()
    a[..., RNG.random(a.shape[-1]) < frac] = np.nan
    return a


def main():
    # ---- api: the reference's own known-answer tables (tests/test_core.py:126-206) ----------------
    n = 0
    for dt in (np.float64, np.int64, np.float32, np.int32):
        for eg in (None, [0, 1, 2]):
            for func in ("sum", "count", "nanmean", "mean", "max", "nanmin", "var"):
                for arr, by in ((np.ones(12), LABELS), (np.ones(12), NAN_LABELS), (np.ones((2, 12)), LABELS),
                                (np.ones((2, 12)), NAN_LABELS), (np.ones((2, 12)), np.array([LABELS, LABELS])),
                                (np.ones((2, 12)), np.array([NAN_LABELS, NAN_LABELS]))):
                    add_api(f"api/kat{n}", arr.astype(dt), (by,), func, expected_groups=eg, fill_value=123)
                    n += 1
    # ---- api: seeded sweeps -------------------------------------------------------------------------
    x12 = np.arange(12, dtype=np.float32)
    x12n = x12.copy()
    x12n[[2, 5, 6, 9, 10]] = np.nan
    n = 0
    for func in FUNCS:
        fk = {"ddof": 1} if ("var" in func or "std" in func) else None
        for arr in (x12, x12n, x12.astype(np.float64), x12.astype(np.int32), x12.astype(np.int64), (x12 > 5)):
            for kw in (dict(), dict(expected_groups=np.arange(4)), dict(expected_groups=np.arange(4), fill_value=123),
                       dict(min_count=3), dict(expected_groups=np.arange(4), fill_value=-1, min_count=4)):
                add_api(f"api/small{n}", arr, (LABELS,), func, finalize_kwargs=fk, **kw)
                n += 1
            add_api(f"api/small{n}", arr, (NAN_LABELS,), func, finalize_kwargs=fk)
            n += 1
    # random medium problems, floats with NaNs, several G, 1-D and 2-D, dtype= given
    n = 0
    for N, G in ((1000, 7), (5000, 100), (4096, 1000)):
        by = RNG.integers(0, G, N)
        for dt in (np.float32, np.float64):
            base = RNG.random(N).astype(dt) * 10 - 3
            for arr in (base, with_nans(base), np.stack([base, with_nans(base), base[::-1]])):
                for func in FUNCS:
                    if func in ("prod", "nanprod"):
                        a = np.where(np.isnan(arr), np.nan, 1 + (arr - 2) / 50).astype(dt)
                    else:
                        a = arr
                    fk = {"ddof": 1} if ("var" in func or "std" in func) else None
                    add_api(f"api/rand{n}", a, (by,), func, expected_groups=np.arange(G), finalize_kwargs=fk)
                    n += 1
        ia = RNG.integers(-50, 50, N).astype(np.int32)
        for func in FUNCS:
            if "prod" in func:
                continue
            add_api(f"api/rand{n}", ia, (by,), func, expected_groups=np.arange(G))
            n += 1
            add_api(f"api/rand{n}", ia.astype(np.int64), (by,), func)
            n += 1
    # dtype= requests (tests/test_core.py:1135-1155, 2121-2131)
    n = 0
    base = RNG.random(300).astype(np.float32)
    by = RNG.integers(0, 5, 300)
    for func in ("sum", "nansum", "mean", "nanmean", "var", "max", "count"):
        for dt in (np.float64, np.float32):
            add_api(f"api/dtype{n}", base, (by,), func, dtype=dt)
            n += 1
    add_api(f"api/dtype{n}", base.astype(np.int32), (by,), "sum", dtype=np.float32)
    # axis subset with 2-D labels (tests/test_core.py:582-611) and full 2-D reduce
    n = 0
    lab2 = np.array([LABELS, LABELS[::-1]])
    arr3 = RNG.random((3, 2, 12)).astype(np.float32)
    for func in ("sum", "count", "nanmean", "max", "nanmin", "var"):
        for axis in (None, -1, 1, (1, 2)):
            add_api(f"api/axis{n}", arr3, (lab2,), func, axis=axis, expected_groups=[0, 1, 2])
            n += 1
        add_api(f"api/axis{n}", with_nans(arr3), (lab2,), func, axis=-1, expected_groups=[0, 2], fill_value=123)
        n += 1
    # broadcast labels (by has a size-1 dim): the dim=... case (flox/core.py:298-312)
    labb = LABELS.reshape(1, 12)
    for func in ("sum", "mean", "max"):
        add_api(f"api/axis{n}", arr3, (labb,), func)
        n += 1
    # multiple `by` (flox/factorize.py:102-108,183-184)
    n = 0
    by0 = RNG.integers(0, 3, 200)
    by1 = RNG.integers(10, 14, 200).astype(float)
    by1[::17] = np.nan
    arr = RNG.random(200).astype(np.float64)
    for func in ("sum", "count", "nanmean", "max"):
        add_api(f"api/multi{n}", arr, (by0, by1), func)
        n += 1
        add_api(f"api/multi{n}", arr, (by0, by1), func, expected_groups=(np.arange(4), np.array([10., 11, 12, 13, 14])))
        n += 1
    # binning (flox/factorize.py:55-82; tests/test_core.py:1238-1253,1367-1381)
    n = 0
    vals = RNG.random(100) * 12 - 1
    for func in ("sum", "count", "mean"):
        add_api(f"api/bins{n}", np.ones(100), (vals,), func, expected_groups=np.array([0., 2, 4, 8, 10]), isbin=True)
        n += 1
        add_api(f"api/bins{n}", np.ones(100), (vals,), func, expected_groups=np.array([0., 2, 4, 8, 10]), isbin=True,
                fill_value=0)
        n += 1
    # var precision (tests/test_core.py:2287-2319)
    n = 0
    for off in (1e2, 1e6, 1e9, 1e12):
        a = RNG.standard_normal(600) + off
        by = np.repeat(np.arange(6), 100)
        for func in ("var", "nanstd"):
            add_api(f"api/prec{n}", a, (by,), func, finalize_kwargs={"ddof": 1})
            n += 1
    # all-NaN groups, empty inputs, all labels NaN (tests/test_core.py:814-878)
    n = 0
    allnan = np.full(12, np.nan, dtype=np.float32)
    for func in FUNCS:
        add_api(f"api/edge{n}", allnan, (LABELS,), func)
        n += 1
        add_api(f"api/edge{n}", x12, (np.full(12, np.nan),), func, expected_groups=[0, 1, 2])
        n += 1
        add_api(f"api/edge{n}", x12[:0], (LABELS[:0],), func, expected_groups=[0, 1, 2])
        n += 1
    # arg-reductions are rejected for engine="flox" (flox/core.py:856-859)
    add_api("api/argreject", x12, (LABELS,), "argmax")

    # ---- engine boundary ------------------------------------------------------------------------
    n = 0
    for N, G in ((12, 3), (777, 10), (4000, 64)):
        codes = RNG.integers(0, G, N)
        if N == 12:
            codes = LABELS.copy()
        for dt in (np.float32, np.float64, np.int32, np.int64):
            a = (RNG.random((2, N)) * 20 - 5).astype(dt)
            if np.dtype(dt).kind == "f":
                a[0, RNG.random(N) < 0.15] = np.nan
            for func in ("sum", "nansum", "max", "nanmax", "min", "nanmin", "nanlen", "mean", "nanmean",
                         "sum_of_squares", "nansum_of_squares", "prod", "nanprod"):
                if "prod" in func:
                    aa = (1 + (a.astype(np.float64) % 3) / 7).astype(dt) if np.dtype(dt).kind == "f" else (a % 3 + 1).astype(dt)
                else:
                    aa = a
                for size, fv, odt in ((G, 0, None), (G + 3, -7, None), (G + 1, np.nan, np.float64)):
                    if fv is np.nan and func == "nanlen":
                        continue
                    add_engine(f"engine/e{n}", func, codes, aa, size, fv, odt)
                    n += 1

    # ---- map-reduce -----------------------------------------------------------------------------
    n = 0
    for N, G, nchunks in ((1200, 7, 4), (5000, 100, 8), (3001, 40, 5)):
        by = RNG.integers(0, G, N)
        by[by == 3] = 4  # one group never occurs
        for dt in (np.float32, np.float64):
            a = (RNG.random(N) * 10 - 3).astype(dt)
            an = with_nans(a)
            a2 = np.stack([a, an])
            for func in FUNCS:
                if "prod" in func:
                    continue
                for arr in (a, an, a2):
                    add_mapred(f"mapred/m{n}", arr, by, func, nchunks, G, ddof=1 if "var" in func or "std" in func else 0)
                    n += 1
            add_mapred(f"mapred/m{n}", an, by, "nansum", nchunks, G, fill_value=np.nan, min_count=2)
            n += 1
            add_mapred(f"mapred/m{n}", a, by, "sum", nchunks, G, dtype=np.float64)
            n += 1
        ia = RNG.integers(-9, 9, N).astype(np.int32)
        for func in ("sum", "count", "mean", "max", "nanmin", "var"):
            add_mapred(f"mapred/m{n}", ia, by, func, nchunks, G)
            n += 1

    with open(os.path.join(HERE, "golden_cases.json"), "w") as f:
        json.dump(CASES, f, indent=0)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **ARRAYS)
    print(f"{len(CASES)} cases, {len(ARRAYS)} arrays")


if __name__ == "__main__":
    main()
