"""Pin the oracle of the NEXT rows (oracle/flox_oracle_next.py: grouped quantile / median and grouped scans,
SURVEY §8f rank 4) to outputs of the UNMODIFIED reference (tests/golden/golden_next.*, written by
tests/golden/make_golden_next.py).  Bit-exact, like tests/test_oracle_pinning.py.  The CUDA counterparts of the quantile family are
tested in tests/test_next_rows_cpu.py / tests/test_gpu_next_rows.py; the scans have no kernels yet."""
import json
import os

import numpy as np
import pytest

import flox_b200
from oracle import flox_oracle_next as ON
from tests.golden_io import assert_same

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(_DIR, "golden_next.json")) as _f:
    CASES = json.load(_f)
ARR = np.load(os.path.join(_DIR, "golden_next.npz"), allow_pickle=False)


def _ids(cases):
    return [c["id"] for c in cases]


QUANT = [c for c in CASES if c["family"] == "quantile"]
SCAN = [c for c in CASES if c["family"] == "scan"]


@pytest.mark.parametrize("case", QUANT, ids=_ids(QUANT))
def test_quantile_family(case):
    func = case["func"]
    q = case["kw"].get("q", 0.5)
    res = ON.engine_quantile(ARR[case["codes"]], ARR[case["array"]], q=q, skipna=func.startswith("nan"), size=case["size"],
                             fill_value=np.nan, dtype=None)
    assert_same(res, ARR[case["result"]], what=case["id"])


@pytest.mark.parametrize("case", SCAN, ids=_ids(SCAN))
def test_scans(case):
    codes, x = ARR[case["codes"]], ARR[case["array"]]
    if case["func"] == "ffill":
        res = ON.engine_ffill(codes, x)
    else:
        res = ON.engine_cumsum(codes, x, skipna=case["func"] == "nancumsum")
    assert_same(res, ARR[case["result"]], what=case["id"])


@pytest.mark.parametrize("func", ["mode", "nanmode", "all"])
def test_package_says_not_implemented(func):
    """Reductions still outside the engine raise instead of falling back to a CPU engine."""
    with pytest.raises(NotImplementedError):
        flox_b200.groupby_reduce(np.ones(4), np.array([0, 0, 1, 1]), func=func)
