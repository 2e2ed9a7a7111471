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
