"""Pin the CPU oracle (oracle/flox_oracle.py) to the reference's own outputs.

The fixtures under tests/golden were produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  The oracle repeats the reference's numpy
operations in the same order, so float results are required to be BIT-EXACT
here, not merely close.
"""
import numpy as np
import pytest

from oracle import flox_oracle as O
from tests.golden_io import api_kwargs, assert_same, load_golden

G = load_golden()


def _ids(cases):
    return [c["id"] for c in cases]


@pytest.mark.parametrize("case", G.family("engine"), ids=_ids(G.family("engine")))
def test_engine_boundary(case):
    dt = None if case["dtype"] is None else np.dtype(case["dtype"])
    res = O.engine_call(case["func"], G.arr(case["codes"]), G.arr(case["array"]), size=case["size"],
                        fill_value=case["fill_value"], dtype=dt)
    assert_same(res, G.arr(case["result"]), what=case["id"])


@pytest.mark.parametrize("case", G.family("api"), ids=_ids(G.family("api")))
def test_groupby_reduce_api(case):
    kw = api_kwargs(case)
    bys = [G.arr(b) for b in case["bys"]]
    if "raises" in case:
        if case["func"].startswith(("arg", "nanarg")):
            pytest.skip("engine='flox' rejects arg-reductions; the oracle pins numpy semantics instead")
        with pytest.raises(Exception):
            O.groupby_reduce(G.arr(case["array"]), *bys, func=case["func"], ddof=case["ddof"], **kw)
        return
    res, *groups = O.groupby_reduce(G.arr(case["array"]), *bys, func=case["func"], ddof=case["ddof"], **kw)
    assert_same(res, G.arr(case["result"]), what=case["id"])
    for g, key in zip(groups, case["groups"]):
        want = G.arr(key)
        if hasattr(g, "left"):
            g = np.asarray(list(g.left) + [g.right[-1]])
        np.testing.assert_array_equal(np.asarray(g), want)


@pytest.mark.parametrize("case", G.family("mapred"), ids=_ids(G.family("mapred")))
def test_mapreduce(case):
    dt = None if case["dtype"] is None else np.dtype(case["dtype"])
    res, _ = O.groupby_reduce_mapreduce(
        G.arr(case["array"]), G.arr(case["by"]), func=case["func"], nchunks=case["nchunks"],
        expected_groups=np.arange(case["ngroups"]), fill_value=case["fill_value"], dtype=dt,
        min_count=case["min_count"], ddof=case["ddof"])
    assert_same(res, G.arr(case["result"]), what=case["id"])
