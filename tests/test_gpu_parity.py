"""Parity tests proper: the CUDA path (through the C ABI) against the reference-made golden
vectors and against the CPU oracle on seeded inputs.  Run on the B200 box: pytest -m gpu.

Bars (BASELINE.json north_star): bit-exact for count / nanlen / argmin / argmax and every integer
result; within 1e-6 relative for float sum / mean / var on f32 inputs (1e-12 on f64 inputs), with
an absolute slack of the same factor times the magnitude of the summands for sums that cancel
(tests/golden_io.py:cond_atol).
"""
import numpy as np
import pytest
import torch

import flox_b200
from flox_b200 import ChunkedArray, kernels
from oracle import flox_oracle as O
from tests.golden_io import api_kwargs, assert_same, cond_atol, load_golden

pytestmark = pytest.mark.gpu

G = load_golden()
RTOL32, RTOL64 = 1e-6, 1e-12


@pytest.fixture(autouse=True)
def _cuda_backend():
    assert torch.cuda.is_available(), "GPU tests need CUDA"
    kernels.set_backend(None)
    k = kernels.get_kernels()
    assert k.name == "cuda"
    k.force_path = 0
    yield k
    k.force_path = 0


def _ids(cases):
    return [c["id"] for c in cases]


def _tols(func, arr, prec=False):
    r64 = RTOL32 if arr.dtype == np.float32 else RTOL64
    if prec:
        r64 = 3e-8  # the reference's own tolerance for the 1e12-offset cases (tests/test_core.py:2287-2319)
    return dict(rtol_f32=RTOL32, rtol_f64=r64, atol_f32=cond_atol(func, arr, RTOL32), atol_f64=cond_atol(func, arr, r64))


# ---- golden vectors -----------------------------------------------------------------------------

@pytest.mark.parametrize("case", G.family("api"), ids=_ids(G.family("api")))
def test_api_golden(case):
    kw = api_kwargs(case)
    fk = {"ddof": case["ddof"]} if case["ddof"] else None
    bys = [G.arr(b) for b in case["bys"]]
    arr = G.arr(case["array"])
    if "raises" in case:
        if case["func"].startswith(("arg", "nanarg")):
            pytest.skip("engine='flox' rejects arg-reductions; b200 implements them (test_arg_reductions)")
        with pytest.raises(Exception):
            flox_b200.groupby_reduce(arr, *bys, func=case["func"], finalize_kwargs=fk, **kw)
        return
    res, *groups = flox_b200.groupby_reduce(arr, *bys, func=case["func"], finalize_kwargs=fk, **kw)
    assert_same(res, G.arr(case["result"]), what=case["id"], **_tols(case["func"], arr, case["id"].startswith("api/prec")))
    for g, key in zip(groups, case["groups"]):
        g = np.asarray(g)
        if g.dtype == object and len(g) and hasattr(g[0], "left"):
            g = np.asarray([iv.left for iv in g] + [g[-1].right])
        np.testing.assert_array_equal(g, G.arr(key))


@pytest.mark.parametrize("path", [1, 2], ids=["smem", "l2"])
@pytest.mark.parametrize("case", G.family("engine"), ids=_ids(G.family("engine")))
def test_engine_golden(case, path, _cuda_backend):
    """aggregate_b200.<func>(codes, array, size=, fill_value=, dtype=) on BOTH kernel paths."""
    arr = G.arr(case["array"])
    if path == 1 and arr.ndim > 1 and arr.shape[0] > 1:
        path = 0  # the shared-memory path takes 1-D problems; let the dispatcher choose
    _cuda_backend.force_path = path
    fn = getattr(flox_b200.aggregate_b200, case["func"])
    dt = None if case["dtype"] is None else np.dtype(case["dtype"])
    res = fn(G.arr(case["codes"]), arr, size=case["size"], fill_value=case["fill_value"], dtype=dt)
    want = G.arr(case["result"])
    if case["func"] == "nanlen":
        res = res.astype(want.dtype)
    assert_same(res, want, what=case["id"], **_tols(case["func"], arr))


@pytest.mark.parametrize("case", G.family("mapred"), ids=_ids(G.family("mapred")))
def test_mapreduce_golden(case):
    a, by = G.arr(case["array"]), G.arr(case["by"])
    bounds = np.linspace(0, a.shape[-1], case["nchunks"] + 1).astype(int)
    arr = ChunkedArray([a[..., lo:hi] for lo, hi in zip(bounds[:-1], bounds[1:])])
    kw = dict(expected_groups=np.arange(case["ngroups"]), fill_value=case["fill_value"], method="map-reduce")
    if case["dtype"] is not None:
        kw["dtype"] = np.dtype(case["dtype"])
    if case["min_count"]:
        kw["min_count"] = case["min_count"]
    if case["ddof"]:
        kw["finalize_kwargs"] = {"ddof": case["ddof"]}
    res, _ = flox_b200.groupby_reduce(arr, by, func=case["func"], **kw)
    assert_same(res, G.arr(case["result"]), what=case["id"], **_tols(case["func"], a))


# ---- seeded problems against the oracle ------------------------------------------------------------

FUNCS = ["count", "sum", "nansum", "mean", "nanmean", "var", "nanvar", "nanstd", "max", "nanmax", "min", "nanmin"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("N,ngroups", [(1_000_003, 100), (2_000_000, 10_000), (500_000, 40_000), (65_537, 1)])
def test_seeded_vs_oracle(N, ngroups, dtype):
    """Sizes the oracle finishes in seconds; 40k groups exercises the L2-atomic path, 1e4 the
    shared-memory path with the exact headline shape, odd N the unaligned tails."""
    rng = np.random.default_rng(12345)
    by = rng.integers(0, ngroups, N)
    if ngroups > 2:
        by[by == 1] = 2       # an absent group
        by[::1001] = -1       # "no group" codes (NaN labels after factorization)
    if np.dtype(dtype).kind == "f":
        x = (rng.random(N) * 4 - 1).astype(dtype)
        x[rng.random(N) < 0.01] = np.nan
    else:
        x = rng.integers(-1000, 1000, N).astype(dtype)
    for func in FUNCS:
        fk = {"ddof": 1} if func in ("var", "nanvar") else None
        kw = dict(expected_groups=np.arange(ngroups))
        got, _ = flox_b200.groupby_reduce(x, by, func=func, finalize_kwargs=fk, **kw)
        want, _ = O.groupby_reduce(x, by, func=func, ddof=1 if fk else 0, **kw)
        assert_same(got, want, what=f"{func}/{np.dtype(dtype)}/{N}/{ngroups}", **_tols(func, x))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_ragged_unaligned_chunk_tables(dtype):
    """The persistent chunk-table walk (csrc/common.cuh:walk_tiles): 300 ragged chunks (more than one
    table of floxb_max_chunks_per_launch), empty chunks, chunks shorter than a tile, chunks that are
    views starting 1..3 elements into an allocation (not 16-byte aligned: scalar path), chunk tails."""
    rng = np.random.default_rng(77)
    ngroups = 1000
    sizes = [0, 1, 3, 4095, 4096, 4097, 8192 + 5, 50_000, 0, 123_457] + list(rng.integers(1, 30_000, 290))
    N = int(sum(sizes))
    if np.dtype(dtype).kind == "f":
        x = (rng.random(N) * 4 - 1).astype(dtype)
        x[rng.random(N) < 0.01] = np.nan
    else:
        x = rng.integers(-1000, 1000, N).astype(dtype)
    by = rng.integers(-1, ngroups, N)
    # device copies with a 1..3-element lead so that chunk k starts (off_k + lead) elements into the buffer
    for lead in (0, 1, 3):
        xd = torch.empty(N + lead, dtype=torch.as_tensor(x[:1]).dtype, device="cuda")
        bd = torch.empty(N + lead, dtype=torch.int64, device="cuda")
        xd[lead:] = torch.as_tensor(x)
        bd[lead:] = torch.as_tensor(by)
        offs = np.concatenate([[0], np.cumsum(sizes)]) + lead
        xs = [xd[offs[i]:offs[i + 1]] for i in range(len(sizes))]
        bs = [bd[offs[i]:offs[i + 1]] for i in range(len(sizes))]
        arr, lab = ChunkedArray(xs), ChunkedArray(bs)
        for func in ("sum", "nanmean", "nanmax", "min", "count", "nanvar"):
            got, _ = flox_b200.groupby_reduce(arr, lab, func=func, expected_groups=np.arange(ngroups), method="map-reduce")
            want, _ = O.groupby_reduce(x, by, func=func, expected_groups=np.arange(ngroups))
            got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
            assert_same(got, want, what=f"ragged/{func}/{np.dtype(dtype)}/lead{lead}", **_tols(func, x))


@pytest.mark.parametrize("ngroups,label_dtype", [(1000, np.int64), (32767, np.int32), (40_000, np.int64), (40_000, np.int32)])
def test_host_label_staging(ngroups, label_dtype):
    """Host-resident chunks: labels are range-checked + narrowed on the host (floxb_host_pack_codes), copied and
    widened on the device (floxb_widen_codes).  Codes outside [0, G) — including ones that would alias a valid
    group after a blind int16/int32 cast — must be ignored exactly as on the device-resident path."""
    from flox_b200.arrays import TransferStats

    rng = np.random.default_rng(31)
    N = 1_200_007
    x = (rng.random(N) * 4 - 1).astype(np.float32)
    by = rng.integers(0, ngroups, N).astype(label_dtype)
    by[::97] = -1
    by[5::1013] = ngroups                      # just out of range
    if label_dtype == np.int64:
        by[7::2029] = 65536 + 3                # aliases group 3 under a blind int16 cast
        by[11::4051] = 2**32 + 5               # aliases group 5 under a blind int32 cast
    want_sum, _ = O.groupby_reduce(x, np.where((by >= 0) & (by < ngroups), by, -1), func="nansum", expected_groups=np.arange(ngroups))
    for container in ("numpy", "pinned"):
        if container == "numpy":
            arr, lab = ChunkedArray.from_array(x, 100_000), ChunkedArray.from_array(by, 100_000)
        else:
            arr = ChunkedArray.from_array(torch.from_numpy(x).pin_memory(), 100_000)
            lab = ChunkedArray.from_array(torch.from_numpy(by).pin_memory(), 100_000)
        h0 = TransferStats.h2d_bytes
        got, _ = flox_b200.groupby_reduce(arr, lab, func="nansum", expected_groups=np.arange(ngroups), method="map-reduce")
        moved = TransferStats.h2d_bytes - h0
        got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
        assert_same(got, want_sum, what=f"staging/{container}/{ngroups}", **_tols("nansum", x))
        cnt, _ = flox_b200.groupby_reduce(arr, lab, func="count", expected_groups=np.arange(ngroups), method="map-reduce")
        cnt = cnt.cpu().numpy() if isinstance(cnt, torch.Tensor) else cnt
        np.testing.assert_array_equal(cnt, np.bincount(by[(by >= 0) & (by < ngroups)], minlength=ngroups))
        packed = 2 if ngroups <= 32767 else (4 if label_dtype == np.int64 else np.dtype(label_dtype).itemsize)
        assert moved == N * (4 + packed), (moved, N, packed)


@pytest.mark.parametrize("container", ["numpy", "cuda"])
@pytest.mark.parametrize("label_dtype", [np.int64, np.int32])
def test_unknown_integer_labels_discovered_on_device(label_dtype, container):
    """expected_groups=None (BASELINE config 1's call): groups and codes from floxb_label_minmax / _presence /
    floxb_lut_codes instead of a host pd.factorize; groups, order and results equal the reference's."""
    rng = np.random.default_rng(17)
    N = 1_000_003
    x = rng.random(N).astype(np.float32)
    by = (rng.integers(0, 100, N) * 3 - 50).astype(label_dtype)   # 100 groups, non-contiguous values, negative ones too
    xa, ba = (x, by) if container == "numpy" else (torch.as_tensor(x, device="cuda"), torch.as_tensor(by, device="cuda"))
    for func in ("sum", "nanmean", "count", "max"):
        got, groups = flox_b200.groupby_reduce(xa, ba, func=func)
        want, want_groups = O.groupby_reduce(x, by, func=func)
        got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
        groups = np.asarray(groups)
        assert groups.dtype == np.asarray(want_groups).dtype and np.array_equal(groups, np.asarray(want_groups))
        assert_same(got, want, what=f"discover/{func}/{container}", **_tols(func, x))


@pytest.mark.parametrize("func", ["argmax", "argmin", "nanargmax", "nanargmin"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_arg_reductions(func, dtype):
    """First-occurrence semantics, bit-exact (reference pins numpy's: tests/test_core.py:388-408)."""
    rng = np.random.default_rng(7)
    N, ngroups = 50_000, 37
    by = rng.integers(0, ngroups, N)
    by[by == 3] = 4
    x = rng.integers(0, 50, N).astype(dtype)  # many ties
    if np.dtype(dtype).kind == "f":
        x[rng.random(N) < 0.02] = np.nan
        x[by == 7] = np.nan  # an all-NaN group
    got, _ = flox_b200.groupby_reduce(x, by, func=func, expected_groups=np.arange(ngroups))
    want, _ = O.groupby_reduce(x, by, func=func, expected_groups=np.arange(ngroups))
    assert got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)


def test_multidim_layouts_vs_oracle():
    """(A, n) row-batched arrays, axis-subset reductions with 2-D labels, broadcast labels."""
    rng = np.random.default_rng(99)
    arr = rng.random((6, 5, 4000)).astype(np.float32)
    arr[0, 0, ::7] = np.nan
    by1 = rng.integers(0, 12, 4000)
    by2 = rng.integers(0, 12, (5, 4000))
    for func in ("sum", "nanmean", "max", "count", "nanvar"):
        for by, kw in ((by1, {}), (by2, {}), (by2, dict(axis=-1)), (by1.reshape(1, -1), {})):
            kw = dict(kw, expected_groups=np.arange(12))
            got, _ = flox_b200.groupby_reduce(arr, by, func=func, **kw)
            want, _ = O.groupby_reduce(arr, by, func=func, **kw)
            assert_same(got, want, what=f"{func}/{by.shape}/{kw}", **_tols(func, arr))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_time_major_layout_column_kernels(dtype):
    """config-3 layout: a (time, lat, lon) array grouped along time, handed over as the
    transposed VIEW (lat, lon, time) — consumed in place by the column-owner kernels."""
    rng = np.random.default_rng(11)
    T, ny, nx = 730, 36, 1024  # ny*nx = 36864 columns >= COLS_MIN_COLUMNS
    base = (rng.random((T, ny, nx)) * 10 - 2).astype(dtype)
    if np.dtype(dtype).kind == "f":
        base[rng.random(T) < 0.05, 3, 5] = np.nan
    months = (np.arange(T) // 31) % 12
    months[months == 7] = 8
    view = base.transpose(1, 2, 0)
    xt = torch.as_tensor(base, device="cuda").permute(1, 2, 0)
    for func in ("sum", "nanmean", "max", "nanmin", "count", "nanvar"):
        want, _ = O.groupby_reduce(np.ascontiguousarray(view), months, func=func, expected_groups=np.arange(12))
        got, _ = flox_b200.groupby_reduce(view, months, func=func, expected_groups=np.arange(12))
        assert_same(got, want, what=f"cols/{func}/{np.dtype(dtype)}", **_tols(func, base))
        got_t, _ = flox_b200.groupby_reduce(xt, torch.as_tensor(months, device="cuda"), func=func,
                                            expected_groups=np.arange(12))
        assert got_t.is_cuda
        assert_same(got_t.cpu().numpy(), want, what=f"cols-torch/{func}/{np.dtype(dtype)}", **_tols(func, base))


@pytest.mark.parametrize("ngroups,T", [(13, 3650), (13, 3651), (40, 3650)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_row_batched_layout_warp_per_row_kernels(dtype, ngroups, T):
    """(lat*lon, time) rows sharing one label vector — the layout flox's own reshape produces.  13 groups take the
    lane-private-bin kernel (k_rows_bins, G <= 32: pair loads for even T, scalar loads for odd T), 40 groups the
    permutation kernel (k_rows)."""
    rng = np.random.default_rng(21)
    A = 700
    arr = (rng.random((A, T)) * 10 - 2).astype(dtype)
    if np.dtype(dtype).kind == "f":
        arr[5, rng.random(T) < 0.1] = np.nan
        arr[6, :] = np.nan
    if ngroups == 13:
        labels = ((np.arange(T) % 365) // 30.42).astype(np.int64).clip(0, 11)
    else:
        labels = ((np.arange(T) % 365) // 9.4).astype(np.int64).clip(0, 38)
    labels[::400] = -1
    for func in ("sum", "nanmean", "max", "nanmin", "count", "nanvar"):
        want, _ = O.groupby_reduce(arr, labels, func=func, expected_groups=np.arange(ngroups))
        got, _ = flox_b200.groupby_reduce(arr, labels, func=func, expected_groups=np.arange(ngroups))
        assert_same(got, want, what=f"rows/{func}/{np.dtype(dtype)}/{ngroups}", **_tols(func, arr))


@pytest.mark.parametrize("labels", ["random", "sorted", "blocked"])
@pytest.mark.parametrize("B", [2, 8, 32, 3])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_small_trailing_dims_huge_G(dtype, B, labels):
    """BASELINE config 4's layout at oracle-sized n: (n, B) values with NaNs handed over as the (B, n) view, G = 1e5,
    labels random / sorted / blocked (method="blockwise": every group in one contiguous block, blocks in random order).
    Power-of-two B <= 32 takes the run-merging kernels (k_moments_global_bruns, k_sqdev_global_bruns), B = 3 the
    per-element L2-atomic ones; both must agree with the oracle."""
    rng = np.random.default_rng(404)
    n, G = 300_007, 100_000
    base = rng.standard_normal((n, B)).astype(dtype) * 3 + 1
    base.reshape(-1)[::7] = np.nan
    by = rng.integers(0, G, n)
    if labels == "sorted":
        by = np.sort(by)
    elif labels == "blocked":
        order = rng.permutation(G)
        by = order[np.sort(by)]
    by[::501] = -1
    view = base.T
    xt = torch.as_tensor(base, device="cuda").t()
    bt = torch.as_tensor(by, device="cuda")
    for func in ("nansum", "sum", "count", "nanvar", "var", "nanmax", "nanmean"):
        fk = {"ddof": 1} if "var" in func else None
        want, _ = O.groupby_reduce(np.ascontiguousarray(view), by, func=func, ddof=1 if fk else 0, expected_groups=np.arange(G))
        got, _ = flox_b200.groupby_reduce(xt, bt, func=func, finalize_kwargs=fk, expected_groups=np.arange(G))
        assert_same(got.cpu().numpy(), want, what=f"c4/{func}/{np.dtype(dtype)}/B{B}/{labels}", **_tols(func, base))
        if func in ("nansum", "nanvar", "count"):  # the same slabs as chunks of a ChunkedArray (consumed in place, B kept columns)
            cuts = [0, 100_003, 200_001, n]
            base_t = torch.as_tensor(base, device="cuda")
            arr = ChunkedArray([base_t[lo:hi].t() for lo, hi in zip(cuts[:-1], cuts[1:])])
            lab = ChunkedArray([bt[lo:hi] for lo, hi in zip(cuts[:-1], cuts[1:])])
            got_c, _ = flox_b200.groupby_reduce(arr, lab, func=func, finalize_kwargs=fk, expected_groups=np.arange(G), method="cohorts")
            assert_same(got_c.cpu().numpy(), want, what=f"c4-chunked/{func}/{np.dtype(dtype)}/B{B}/{labels}", **_tols(func, base))


def test_torch_cuda_inputs_stay_on_device():
    rng = np.random.default_rng(5)
    x = rng.random(300_000).astype(np.float32)
    by = rng.integers(0, 500, 300_000)
    xt, bt = torch.as_tensor(x, device="cuda"), torch.as_tensor(by, device="cuda")
    got, groups = flox_b200.groupby_reduce(xt, bt, func="nanmean", expected_groups=np.arange(500))
    assert isinstance(got, torch.Tensor) and got.is_cuda and got.dtype == torch.float32
    want, _ = O.groupby_reduce(x, by, func="nanmean", expected_groups=np.arange(500))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6)


@pytest.mark.parametrize("ngroups", [12, 100, 3000, 10_000])
def test_f32_two_level_accuracy_under_skew(ngroups):
    """The f32-partial / f64-accumulator kernel must hold the 1e-6 bar for HOT groups (half of all
    rows in one group: thousands of adds per sweep interval) and for addends spanning 12 decades."""
    gen = torch.Generator(device="cuda").manual_seed(99)
    N = 40_000_000
    by = torch.randint(0, ngroups, (N,), device="cuda", generator=gen, dtype=torch.int64)
    by[torch.rand(N, device="cuda", generator=gen) < 0.5] = ngroups // 2      # the hot group
    for name, x in (("uniform", torch.rand(N, device="cuda", generator=gen)),
                    ("lognormal", torch.exp(6.0 * torch.randn(N, device="cuda", generator=gen)).to(torch.float32)),
                    ("signed", torch.randn(N, device="cuda", generator=gen))):
        exact = torch.zeros(ngroups, dtype=torch.float64, device="cuda").index_add_(0, by, x.to(torch.float64))
        scale = torch.zeros(ngroups, dtype=torch.float64, device="cuda").index_add_(0, by, x.abs().to(torch.float64))
        got, _ = flox_b200.groupby_reduce(x, by, func="sum", expected_groups=np.arange(ngroups), dtype=np.float64)
        err = ((got - exact).abs() / scale).max().item()
        # every f32 partial holds <= 32 addends before it moves to f64: the error is ~sqrt(32) * 2^-25 of the
        # partial (a strict bound of 32 * 2^-24 needs every rounding to go the same way).  Well-scaled data
        # measure ~1e-9, the 12-decade lognormal (one huge addend dominates its partial) 2-3e-7; the bar is 1e-6.
        assert err < 5e-7, (name, ngroups, err)
        got32, _ = flox_b200.groupby_reduce(x, by, func="mean", expected_groups=np.arange(ngroups))
        cnt = torch.bincount(by, minlength=ngroups).to(torch.float64)
        rel = ((got32.to(torch.float64) - exact / cnt).abs() / (scale / cnt)).max().item()
        assert rel < 1e-6, (name, ngroups, rel)


@pytest.mark.parametrize("ngroups", [7, 1000, 10_000, 25_000])
@pytest.mark.parametrize("labels", ["random", "one_group", "sorted"])
def test_f32_fixed_point_level1(ngroups, labels, _cuda_backend):
    """k_moments_smem_fx (csrc/moments.cu): the integer level 1 is EXACT or not taken.  Mixed signs, zeros, exact small
    integers far below the window, denormals, +-inf, 30 decades of magnitudes (most of those rows take the direct f64
    path), one group receiving every row (the drain path, thousands of same-address adds in flight), sorted labels —
    against an f64 index_add_, for both level-1 kernels (force_path 3 = the CAS-loop one)."""
    gen = torch.Generator(device="cuda").manual_seed(2024)
    N = 6_000_011
    if labels == "random":
        by = torch.randint(0, ngroups, (N,), device="cuda", generator=gen, dtype=torch.int64)
    elif labels == "one_group":
        by = torch.full((N,), ngroups - 1, device="cuda", dtype=torch.int64)
    else:
        by = torch.sort(torch.randint(0, ngroups, (N,), device="cuda", generator=gen, dtype=torch.int64)).values
    by[::1009] = -1
    cases = {
        "uniform": torch.rand(N, device="cuda", generator=gen),
        "signed": torch.randn(N, device="cuda", generator=gen) * 300.0,
        "kelvin": 250.0 + 70.0 * torch.rand(N, device="cuda", generator=gen),
        "wide": (torch.exp(14.0 * torch.randn(N, device="cuda", generator=gen)) *
                 torch.sign(torch.randn(N, device="cuda", generator=gen))).to(torch.float32),
    }
    mixed = torch.rand(N, device="cuda", generator=gen) * 8.0 - 4.0
    mixed[::3] = 0.0
    mixed[1::7] = torch.randint(-5, 6, (mixed[1::7].numel(),), device="cuda", generator=gen).to(torch.float32) * 2.0 ** -20
    mixed[2::11] = 1e-41            # denormal
    mixed[5::13] = -3e-39           # denormal
    mixed[4::1000] = 2.0 ** 60      # far above the window
    cases["mixed"] = mixed
    valid = by >= 0
    for name, x in cases.items():
        exact = torch.zeros(ngroups, dtype=torch.float64, device="cuda").index_add_(0, by[valid], x[valid].to(torch.float64))
        scale = torch.zeros(ngroups, dtype=torch.float64, device="cuda").index_add_(0, by[valid], x[valid].abs().to(torch.float64))
        cnt = torch.bincount(by[valid], minlength=ngroups)
        for path, tol in ((0, 2e-11), (3, 5e-7)):  # 2e-11: the f64 atomics of the checker itself
            _cuda_backend.force_path = path
            got, _ = flox_b200.groupby_reduce(x, by, func="sum", expected_groups=np.arange(ngroups), dtype=np.float64)
            c, _ = flox_b200.groupby_reduce(x, by, func="count", expected_groups=np.arange(ngroups))
            torch.testing.assert_close(c, cnt, rtol=0, atol=0)
            present = cnt > 0  # absent groups hold the NaN fill
            assert bool(torch.isnan(got[~present]).all())
            err = ((got - exact).abs() / scale.clamp_min(1e-300))[present].max().item()
            assert err < tol, (name, labels, ngroups, path, err)
    _cuda_backend.force_path = 0
    # infinities and NaNs propagate like the reference's (NaN-propagating sum, nansum skips)
    x = torch.rand(N, device="cuda", generator=gen)
    by = torch.randint(0, ngroups, (N,), device="cuda", generator=gen, dtype=torch.int64)
    x[10], x[11], x[12] = float("inf"), float("-inf"), float("nan")
    by[10], by[11], by[12] = 0, 1 % ngroups, 2 % ngroups
    got, _ = flox_b200.groupby_reduce(x, by, func="sum", expected_groups=np.arange(ngroups))
    want = torch.zeros(ngroups, dtype=torch.float64, device="cuda").index_add_(0, by, x.to(torch.float64)).to(torch.float32)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=0, equal_nan=True)


# ---- size-independent properties at (near) full size -----------------------------------------------

def test_full_size_properties():
    """2.5e8 f32 rows x 1e4 groups generated on device (3 GB): totals conserve, chunking and the
    kernel path do not change the answer, counts are exact."""
    gen = torch.Generator(device="cuda").manual_seed(12345)
    N, ngroups = 250_000_000, 10_000
    x = torch.rand(N, device="cuda", generator=gen, dtype=torch.float32)
    by = torch.randint(0, ngroups, (N,), device="cuda", generator=gen, dtype=torch.int64)
    eg = np.arange(ngroups)
    s, _ = flox_b200.groupby_reduce(x, by, func="sum", expected_groups=eg, dtype=np.float64)
    c, _ = flox_b200.groupby_reduce(x, by, func="count", expected_groups=eg)
    assert int(c.sum()) == N
    torch.testing.assert_close(c, torch.bincount(by, minlength=ngroups), rtol=0, atol=0)
    total = x.to(torch.float64).sum()
    # f32 values go through short f32 partials (<= ~8 addends) before the f64 accumulator: per-group
    # error ~1e-8 relative, far inside the 1e-6 bar, and it averages out further over the groups
    assert abs(float(s.sum() - total)) <= 1e-8 * float(total)
    exact = torch.zeros(ngroups, dtype=torch.float64, device="cuda").index_add_(0, by, x.to(torch.float64))
    torch.testing.assert_close(s, exact, rtol=2e-7, atol=0)
    # chunked map-reduce == single shot up to the order of the atomics
    sm, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(x, 10_000_000), by, func="sum", expected_groups=eg,
                                     dtype=np.float64, method="map-reduce")
    torch.testing.assert_close(sm, s, rtol=2e-7, atol=0)
    # L2-atomic path == shared-memory path
    k = kernels.get_kernels()
    k.force_path = 2
    s2, _ = flox_b200.groupby_reduce(x[: N // 10], by[: N // 10], func="sum", expected_groups=eg, dtype=np.float64)
    k.force_path = 0
    s1, _ = flox_b200.groupby_reduce(x[: N // 10], by[: N // 10], func="sum", expected_groups=eg, dtype=np.float64)
    torch.testing.assert_close(s2, s1, rtol=2e-7, atol=0)
    # max over groups == global max; mean within f32 of sum/count
    mx, _ = flox_b200.groupby_reduce(x, by, func="max", expected_groups=eg)
    assert float(mx.max()) == float(x.max())
    mean, _ = flox_b200.groupby_reduce(x, by, func="mean", expected_groups=eg)
    torch.testing.assert_close(mean.to(torch.float64), s / c, rtol=2e-7, atol=0)
