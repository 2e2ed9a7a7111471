"""Map-reduce / blockwise / cohorts host logic on the CPU emulator, against the reference-made
mapred/* goldens (per-chunk flox.core.chunk_reduce + agg.simple_combine + _finalize_results), and a
world-size-2 gloo run of the cross-rank merge."""
import os
import subprocess
import sys

import numpy as np
import pytest

import flox_b200
from flox_b200 import ChunkedArray
from flox_b200.aggregations import _initialize_aggregation
from flox_b200.core import chunk_reduce
from flox_b200.mapreduce import _aggregate, _expand_dims, _simple_combine
from flox_b200.reindex import ReindexStrategy
from tests.cpu_fixtures import emu  # noqa: F401
from tests.golden_io import assert_same, cond_atol, load_golden

G = load_golden()
CASES = G.family("mapred")


def _chunks(case):
    a, by = G.arr(case["array"]), G.arr(case["by"])
    bounds = np.linspace(0, a.shape[-1], case["nchunks"] + 1).astype(int)
    return a, by, bounds


def _kw(case):
    kw = dict(expected_groups=np.arange(case["ngroups"]), fill_value=case["fill_value"], method="map-reduce")
    if case["dtype"] is not None:
        kw["dtype"] = np.dtype(case["dtype"])
    if case["min_count"]:
        kw["min_count"] = case["min_count"]
    if case["ddof"]:
        kw["finalize_kwargs"] = {"ddof": case["ddof"]}
    return kw


def _check(case, res):
    arr = G.arr(case["array"])
    r64 = 1e-6 if arr.dtype == np.float32 else 1e-12
    assert_same(res, G.arr(case["result"]), what=case["id"], rtol_f32=1e-6, rtol_f64=r64,
                atol_f32=cond_atol(case["func"], arr, 1e-6), atol_f64=cond_atol(case["func"], arr, r64))


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_fused_mapreduce_matches_reference(case, emu):  # noqa: F811
    a, by, bounds = _chunks(case)
    arr = ChunkedArray([a[..., lo:hi] for lo, hi in zip(bounds[:-1], bounds[1:])])
    res, groups = flox_b200.groupby_reduce(arr, by, func=case["func"], **_kw(case))
    _check(case, res)
    np.testing.assert_array_equal(groups, np.arange(case["ngroups"]))


@pytest.mark.parametrize("case", CASES[::7], ids=[c["id"] for c in CASES[::7]])
def test_explicit_intermediates_combine(case, emu):  # noqa: F811
    """chunk_reduce per slab -> _expand_dims -> _simple_combine -> _aggregate, like flox/dask.py."""
    import pandas as pd

    a, by, bounds = _chunks(case)
    fk = {"ddof": case["ddof"]} if case["ddof"] else None
    dt = None if case["dtype"] is None else np.dtype(case["dtype"])
    agg = _initialize_aggregation(case["func"], dt, a.dtype, case["fill_value"], case["min_count"], fk)
    if agg.fill_value["user"] is None:
        agg.fill_value["user"] = agg.fill_value[agg.name]
    expected = pd.RangeIndex(case["ngroups"])
    axis = (a.ndim - 1,)
    parts = [_expand_dims(chunk_reduce(a[..., lo:hi], by[lo:hi], func=agg.chunk, expected_groups=expected, axis=axis,
                                       fill_value=agg.fill_value["intermediate"], dtype=agg.dtype["intermediate"],
                                       reindex=True)) for lo, hi in zip(bounds[:-1], bounds[1:])]
    out = _aggregate(parts, _simple_combine, agg, expected, axis, keepdims=True, fill_value=None,
                     reindex=ReindexStrategy(blockwise=True))
    _check(case, out[agg.name])


def test_blockwise_and_cohorts_fill_absent_groups(emu):  # noqa: F811
    rng = np.random.default_rng(3)
    by = np.sort(rng.integers(0, 9, 600))
    by[by == 4] = 5
    x = rng.random(600)
    want, _ = flox_b200.groupby_reduce(x, by, func="sum", expected_groups=np.arange(9))
    for method in ("blockwise", "cohorts"):
        got, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(x, 100), by, func="sum", expected_groups=np.arange(9),
                                          method=method)
        np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True)
        assert np.isnan(got[4])


_WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["FLOXB_ROOT"])
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["FLOXB_PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
from flox_b200 import kernels, groupby_reduce, ChunkedArray, parallel
from tests.emu_backend import EmuKernels
kernels.set_backend(EmuKernels())
rng = np.random.default_rng(7)
n, G = 4000, 13
x = rng.standard_normal(n).astype(np.float32); x[::11] = np.nan
by = rng.integers(0, G, n); by[by == 5] = 6
ok = True
for func, fk in (("sum", None), ("nanmean", None), ("count", None), ("nanmax", None), ("min", None),
                 ("nanvar", {"ddof": 1}), ("nanargmax", None), ("argmin", None)):
    want, _ = groupby_reduce(x, by, func=func, expected_groups=np.arange(G), finalize_kwargs=fk)
    for sharding in ("replicated", "local"):
        arr = ChunkedArray.from_array(x, 500)
        lab = by
        if sharding == "local":   # each rank only holds its half
            half = slice(0, n // 2) if dist.get_rank() == 0 else slice(n // 2, n)
            arr, lab = ChunkedArray.from_array(x[half], 500), by[half]
        with parallel.distributed(sharding=sharding):
            got, _ = groupby_reduce(arr, lab, func=func, expected_groups=np.arange(G), finalize_kwargs=fk,
                                    method="cohorts")
        good = got.dtype == want.dtype and np.allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True)
        if func in ("count", "nanargmax", "argmin"):
            good = good and np.array_equal(got, want)
        if not good:
            ok = False
            print("MISMATCH", func, sharding, dist.get_rank(), got, want, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_two_rank_gloo_merge(tmp_path):
    """World-size-2 gloo run: both sharding modes, sum/max/min/var/arg merges across ranks."""
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + (os.getpid() % 400))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), FLOXB_ROOT=root, FLOXB_PORT=port, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


_SPARSE_WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["FLOXB_ROOT"])
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["FLOXB_PORT"],
                        rank=int(os.environ["RANK"]), world_size=3)
from flox_b200 import kernels, groupby_reduce, ChunkedArray, parallel
from tests.emu_backend import EmuKernels
kernels.set_backend(EmuKernels())
rank = dist.get_rank()
rng = np.random.default_rng(11)
n, G = 9000, 400
x = rng.standard_normal(n).astype(np.float64); x[::13] = np.nan
ok = True
for layout in ("sorted", "random"):
    by = np.sort(rng.integers(0, G, n)) if layout == "sorted" else rng.integers(0, G, n)
    by[by == 7] = 8
    cut = [0, 2950, 6100, n]          # the cuts fall inside groups: boundary groups straddle ranks
    sl = slice(cut[rank], cut[rank + 1])
    mine = np.unique(by[sl])
    for func, fk in (("sum", None), ("count", None), ("nanmean", None), ("nanmax", None), ("nanvar", {"ddof": 1}),
                     ("nanargmax", None), ("nanfirst", None)):
        want, _ = groupby_reduce(x, by, func=func, expected_groups=np.arange(G), finalize_kwargs=fk)
        with parallel.distributed(sharding="local", output="sharded") as ctx:
            got, _ = groupby_reduce(ChunkedArray.from_array(x[sl], 700), by[sl], func=func, expected_groups=np.arange(G),
                                    finalize_kwargs=fk, method="cohorts")
            n_coll = ctx.comm.collectives
        good = got.dtype == want.dtype and np.allclose(got[mine], want[mine], rtol=1e-12, atol=1e-12, equal_nan=True)
        if func in ("count", "nanargmax"):
            good = good and np.array_equal(got[mine], want[mine])
        if not good:
            ok = False
            print("MISMATCH", layout, func, rank, flush=True)
# kept trailing columns (BASELINE config 4's layout): (n, 4) blocks handed over as block.T, method="blockwise"
by = np.sort(rng.integers(0, G, n))
x4 = rng.standard_normal((n, 4)); x4[::9, 2] = np.nan
sl = slice(cut[rank], cut[rank + 1])
mine = np.unique(by[sl])
for func, fk in (("nansum", None), ("nanvar", {"ddof": 1}), ("count", None)):
    want, _ = groupby_reduce(np.ascontiguousarray(x4.T), by, func=func, expected_groups=np.arange(G), finalize_kwargs=fk)
    with parallel.distributed(sharding="local", output="sharded"):
        got, _ = groupby_reduce(ChunkedArray([torch.as_tensor(x4[sl]).t()]), ChunkedArray([torch.as_tensor(by[sl])]), func=func,
                                expected_groups=np.arange(G), finalize_kwargs=fk, method="blockwise")
    got = got.numpy()
    if not (got.shape == want.shape and np.allclose(got[:, mine], want[:, mine], rtol=1e-12, atol=1e-12, equal_nan=True)):
        ok = False
        print("MISMATCH B>1", func, rank, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_three_rank_gloo_sparse_cohort_merge(tmp_path):
    """parallel.distributed(output="sharded"): only the groups that straddle ranks are merged (sorted labels), random
    labels fall back to the dense merge; every rank's result is complete for the groups its rows touch."""
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = tmp_path / "sparse_worker.py"
    script.write_text(_SPARSE_WORKER)
    port = str(29900 + (os.getpid() % 90))
    procs = []
    for rank in range(3):
        env = dict(os.environ, RANK=str(rank), FLOXB_ROOT=root, FLOXB_PORT=port, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
