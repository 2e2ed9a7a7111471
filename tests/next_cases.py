"""Shared bodies of the sorted-segment-family tests, CPU-emulator tier and GPU tier: the engine
functions and the public API against the reference-made fixtures tests/golden/golden_next.* — bit-exact."""
import json
import os

import numpy as np

import flox_b200
from tests.golden_io import assert_same

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(_DIR, "golden_next.json")) as _f:
    CASES = json.load(_f)
ARR = np.load(os.path.join(_DIR, "golden_next.npz"), allow_pickle=False)

QUANT = [c for c in CASES if c["family"] == "quantile"]
SCAN = [c for c in CASES if c["family"] == "scan"]
API_QUANT = [c for c in CASES if c["family"] == "api_quantile"]


def ids(cases):
    return [c["id"] for c in cases]


def check_engine_quantile(case, shuffle: bool):
    """aggregate_b200.<quantile family>(codes, array, q=, size=, fill_value=): the engine contract.  The reference needs
    rows sorted by group (chunk_reduce sorts for engine="flox"); the b200 engine does not — `shuffle` permutes them."""
    codes, arr = ARR[case["codes"]], ARR[case["array"]]
    if shuffle:
        perm = np.random.default_rng(7).permutation(codes.size)
        codes, arr = codes[perm], arr[..., perm]
    fn = getattr(flox_b200.aggregate_b200, case["func"])
    res = fn(codes, arr, size=case["size"], fill_value=np.nan, **case["kw"])
    assert_same(res, ARR[case["result"]], what=case["id"])


def check_api_quantile(case, as_torch_device=None):
    kw = {}
    if "expected_groups" in case:
        kw["expected_groups"] = ARR[case["expected_groups"]]
    if case["fill_value"] is not None:
        kw["fill_value"] = case["fill_value"]
    arr, by = ARR[case["array"]], ARR[case["by"]]
    if as_torch_device is not None:
        import torch

        arr, by = torch.as_tensor(arr, device=as_torch_device), torch.as_tensor(by, device=as_torch_device)
    res, groups = flox_b200.groupby_reduce(arr, by, func=case["func"], finalize_kwargs=case["finalize_kwargs"], **kw)
    if as_torch_device is not None:
        res = res.cpu().numpy()
    assert_same(res, ARR[case["result"]], what=case["id"])
    np.testing.assert_array_equal(np.asarray(groups), ARR[case["groups"]])
