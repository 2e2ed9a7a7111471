"""Reader for the committed golden vectors (tests/golden/make_golden.py wrote them)."""
from __future__ import annotations

import functools
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, cases, arrays):
        self.cases = cases
        self._arrays = arrays

    def arr(self, key):
        return self._arrays[key]

    def family(self, name):
        return [c for c in self.cases if c["family"] == name]


def _unjson(v):
    if isinstance(v, str) and v == "nan":
        return float("nan")
    if isinstance(v, list):
        return [_unjson(x) for x in v]
    if isinstance(v, dict):
        return {k: _unjson(x) for k, x in v.items()}
    return v


@functools.lru_cache(maxsize=1)
def load_golden() -> Golden:
    with open(os.path.join(_DIR, "golden_cases.json")) as f:
        cases = json.load(f)
    npz = np.load(os.path.join(_DIR, "golden_arrays.npz"), allow_pickle=False)
    arrays = {k: npz[k] for k in npz.files}
    for c in cases:
        for k in ("kw", "fill_value"):
            if k in c:
                c[k] = _unjson(c[k])
    return Golden(cases, arrays)


def api_kwargs(case):
    """Rebuild groupby_reduce kwargs from a recorded api/* case."""
    kw = dict(case["kw"])
    if "dtype" in kw and kw["dtype"] is not None:
        kw["dtype"] = np.dtype(kw["dtype"])
    eg = kw.get("expected_groups")
    if eg is not None:
        if len(case["bys"]) > 1:
            kw["expected_groups"] = tuple(np.asarray(e) for e in eg)
        else:
            kw["expected_groups"] = np.asarray(eg)
    if "axis" in kw and isinstance(kw["axis"], list):
        kw["axis"] = tuple(kw["axis"])
    return kw


def cond_atol(func: str, array, rtol: float) -> float:
    """Conditioning-aware absolute slack: `rtol` times the magnitude of the SUMMANDS.  A group sum
    that cancels to ~0 cannot match to `rtol` of the *result* between two correct float
    implementations (the reference's own f32 pairwise sum is only accurate to eps * sum|x|)."""
    a = np.asarray(array)
    if a.dtype.kind != "f" or a.size == 0:
        return 0.0
    finite = a[np.isfinite(a)]
    if finite.size == 0:
        return 0.0
    m = float(np.max(np.abs(finite)))
    if "var" in func:
        return rtol * m * m
    if "prod" in func:
        return 0.0
    return rtol * m


def assert_same(actual, desired, *, rtol_f32=0.0, rtol_f64=0.0, atol_f32=0.0, atol_f64=0.0, what=""):
    """dtype must match exactly (tests/__init__.py:105-106 of the reference); values
    bit-exact for ints/bools, within rtol (default 0 = exact, NaNs equal) for floats."""
    actual = np.asarray(actual)
    desired = np.asarray(desired)
    assert actual.dtype == desired.dtype, f"{what}: dtype {actual.dtype} != {desired.dtype}"
    assert actual.shape == desired.shape, f"{what}: shape {actual.shape} != {desired.shape}"
    if desired.dtype.kind == "f":
        small = desired.dtype.itemsize <= 4
        np.testing.assert_allclose(actual, desired, rtol=rtol_f32 if small else rtol_f64,
                                   atol=atol_f32 if small else atol_f64, equal_nan=True, err_msg=what)
    else:
        np.testing.assert_array_equal(actual, desired, err_msg=what)
