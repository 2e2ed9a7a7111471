"""Sorted-segment family (grouped quantile / median) on the CPU emulator tier: the package's host logic — the
(group, value) row sort, segment bookkeeping, dtype / fill / min_count rules, the quantile axis — against fixtures the
UNMODIFIED reference produced (tests/golden/make_golden_next.py).  Bit-exact."""
import numpy as np
import pytest

import flox_b200
from tests import next_cases as N
from tests.cpu_fixtures import emu  # noqa: F401


@pytest.mark.parametrize("shuffle", [False, True], ids=["sorted", "shuffled"])
@pytest.mark.parametrize("case", N.QUANT, ids=N.ids(N.QUANT))
def test_engine_quantile_family(case, shuffle, emu):  # noqa: F811
    N.check_engine_quantile(case, shuffle)


@pytest.mark.parametrize("case", N.API_QUANT, ids=N.ids(N.API_QUANT))
def test_api_quantile_family(case, emu):  # noqa: F811
    N.check_api_quantile(case)


def test_quantile_argument_errors(emu):  # noqa: F811
    x, by = np.arange(6.0), np.array([0, 0, 1, 1, 2, 2])
    with pytest.raises(ValueError):
        flox_b200.groupby_reduce(x, by, func="quantile")                       # q missing (core.py:869-871)
    with pytest.raises(NotImplementedError):
        flox_b200.groupby_reduce(flox_b200.ChunkedArray.from_array(x, 3), by, func="median", expected_groups=np.arange(3),
                                 method="map-reduce")


@pytest.mark.parametrize("case", N.SCAN, ids=N.ids(N.SCAN))
def test_engine_scans(case, emu):  # noqa: F811
    N.check_engine_scan(case)


@pytest.mark.parametrize("case", N.API_SCAN, ids=N.ids(N.API_SCAN))
def test_api_groupby_scan(case, emu):  # noqa: F811
    N.check_api_scan(case)


def test_groupby_scan_argument_errors(emu):  # noqa: F811
    x, by = np.arange(6.0), np.array([0, 0, 1, 1, 2, 2])
    with pytest.raises(NotImplementedError):
        flox_b200.groupby_scan(x, by, func="cumsum", expected_groups=[0, 1, 2])
    with pytest.raises(NotImplementedError):
        flox_b200.groupby_scan(x, by, func="cumsum", engine="numpy")
    with pytest.raises(NotImplementedError):
        flox_b200.groupby_scan(np.ones((2, 6)), by, func="cumsum", axis=(0, 1))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_first_last_family(dtype, emu):  # noqa: F811
    N.check_first_last(dtype)
