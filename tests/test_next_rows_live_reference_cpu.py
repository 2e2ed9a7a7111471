"""Seeded random sweep of the sorted-segment family against the LIVE unmodified reference (when /root/reference or
baseline/_ref is importable; the committed fixtures of tests/golden/golden_next.* cover the GPU box): argument
combinations the fixtures do not enumerate — empty inputs, single elements, all-NaN groups, labels outside the expected
groups, several `by`, fill values, min_count, dtype overrides, vector q.  Quantile family and ffill / bfill / integer sums:
bit-exact; float cumsum: within a few ulps of the reference's largest global running total (tests/next_cases.py)."""
import numpy as np
import pytest

import flox_b200
from oracle import ref_loader
from tests.cpu_fixtures import emu  # noqa: F401
from tests.golden_io import assert_same
from tests.next_cases import _assert_scan

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference or baseline/_ref")


def _ref():
    return ref_loader.load_reference()


def _draw(rng, scans: bool):
    dtype = str(rng.choice(["float32", "float64", "int32", "int64"]))
    n = int(rng.choice([1, 2, 5, 31, 200]))
    lead = [(), (3,), (2, 2)][int(rng.integers(0, 3))]
    G = int(rng.integers(1, 7))
    if dtype.startswith("float"):
        x = (rng.random(lead + (n,)) * 8 - 3).astype(dtype)
        if rng.random() < 0.7:
            x[..., rng.random(n) < 0.25] = np.nan
        if rng.random() < 0.2:
            x[...] = np.nan
    else:
        x = rng.integers(-9, 10, lead + (n,)).astype(dtype)
    by = rng.integers(0, G + (2 if (rng.random() < 0.3 and not scans) else 0), n)
    return x, by, G


@pytest.mark.parametrize("seed", range(8))
def test_quantile_family_matches_live_reference(seed, emu):  # noqa: F811
    flox = _ref()
    rng = np.random.default_rng(500 + seed)
    for i in range(30):
        x, by, G = _draw(rng, scans=False)
        func = str(rng.choice(["median", "nanmedian", "quantile", "nanquantile"]))
        fk = None
        if "quantile" in func:
            q = [0.5, 0.1, [0.25, 0.75], [0.0, 0.5, 1.0]][int(rng.integers(0, 4))]
            fk = {"q": q}
        kw = {}
        if rng.random() < 0.6:
            kw["expected_groups"] = np.arange(G + int(rng.integers(0, 2)))
            if rng.random() < 0.4:
                kw["fill_value"] = [-1, 0, np.nan][int(rng.integers(0, 3))]
            if rng.random() < 0.3:
                kw["min_count"] = int(rng.integers(1, 4))
        if rng.random() < 0.2:
            kw["dtype"] = np.dtype(rng.choice(["float64", "float32"]))
        what = f"seed{seed}/{i}: {func} {x.dtype}{x.shape} fk={fk} kw={kw}"
        try:
            want = flox.groupby_reduce(x.copy(), by.copy(), func=func, engine="flox", finalize_kwargs=fk, **kw)
        except Exception as exc:  # noqa: BLE001 - the reference rejects the combination: so must we
            with pytest.raises(Exception):
                flox_b200.groupby_reduce(x, by, func=func, finalize_kwargs=fk, **kw)
            continue
        got = flox_b200.groupby_reduce(x, by, func=func, finalize_kwargs=fk, **kw)
        assert_same(np.asarray(got[0]), np.asarray(want[0]), what=what)
        np.testing.assert_array_equal(np.asarray(got[1]), np.asarray(want[1]), err_msg=what)


@pytest.mark.parametrize("seed", range(6))
def test_groupby_scan_matches_live_reference(seed, emu):  # noqa: F811
    flox = _ref()
    rng = np.random.default_rng(900 + seed)
    for i in range(30):
        x, by, G = _draw(rng, scans=True)
        func = str(rng.choice(["cumsum", "nancumsum", "ffill", "bfill"]))
        what = f"seed{seed}/{i}: {func} {x.dtype}{x.shape}"
        want = np.asarray(flox.groupby_scan(x.copy(), by.copy(), func=func, axis=-1))
        got = np.asarray(flox_b200.groupby_scan(x, by, func=func, axis=-1))
        _assert_scan(got, want, x, func, what)
