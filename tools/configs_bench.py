"""Time the BASELINE.json configs (1, 3, 4, 5-per-GPU) through the public API on one GPU.

Device-resident inputs, CUDA events, best of N after warm-up.  Writes gpurun_out/configs.json.
Algorithmic bytes per element follow SURVEY §8d.
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import flox_b200  # noqa: E402
from flox_b200 import ChunkedArray  # noqa: E402

PEAK = 6570.3
OUT = []


def timeit(fn, reps=5, warm=2, spin_s=0.4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    # steady state: keep the GPU busy for ~0.4 s first.  After an idle gap (data generation, host work) the
    # first calls run up to 2.4x slower (C5 alone: 3.4-4.1 ms against 1.65 ms when it follows other configs).
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < spin_s:
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.mean(ts))


def report(name, elems, bytes_per_elem, best_ms, mean_ms, note=""):
    gbs = elems * bytes_per_elem / best_ms / 1e6
    row = dict(config=name, elements=elems, best_ms=best_ms, mean_ms=mean_ms, elems_per_s=elems / best_ms * 1e3, algorithmic_GBs=gbs,
               frac_of_measured_hbm=gbs / PEAK, note=note)
    OUT.append(row)
    print(json.dumps(row), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c3", "c4", "c5", "rows"]
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(12345)
    # placement experiment: FLOXB_BENCH_PAD_MB shifts every later allocation by that many MiB
    pad_mb = int(os.environ.get("FLOXB_BENCH_PAD_MB", "0"))
    pad = torch.empty(pad_mb << 20, dtype=torch.uint8, device=dev) if pad_mb else None  # noqa: F841

    if "c1" in which:
        n, G = 1_000_000, 100
        x = torch.rand(n, device=dev, generator=gen)
        by = torch.randint(0, G, (n,), device=dev, generator=gen)
        for func in ("sum", "mean"):
            best, mean = timeit(lambda: flox_b200.groupby_reduce(x, by, func=func, expected_groups=np.arange(G)), reps=20, warm=5)
            report(f"C1 f32[1e6] G=100 {func} (device resident)", n, 12, best, mean, "launch/host-latency bound at this size")
        xn, bn = x.cpu().numpy(), by.cpu().numpy()
        t0 = time.perf_counter()
        for _ in range(10):
            flox_b200.groupby_reduce(xn, bn, func="sum")
        dt = (time.perf_counter() - t0) / 10 * 1e3
        report("C1 f32[1e6] G=100 sum (numpy in/out, incl. pd.factorize + H2D)", n, 12, dt, dt, "wall clock, engine='b200' end to end")
        # G=100 at scale (contention-heavy small-G case)
        n2 = 200_000_000
        x2 = torch.rand(n2, device=dev, generator=gen)
        b2 = torch.randint(0, G, (n2,), device=dev, generator=gen)
        best, mean = timeit(lambda: flox_b200.groupby_reduce(x2, b2, func="mean", expected_groups=np.arange(G)))
        report("f32[2e8] G=100 mean", n2, 12, best, mean)
        del x2, b2

    if "c3" in which:
        T, ny, nx = 3650, 720, 1440
        x = torch.rand((T, ny, nx), device=dev, generator=gen)  # 15.1 GB, native (time, lat, lon)
        days = np.arange(T)
        months = ((days % 365) // 30.42).astype(np.int64).clip(0, 11)
        by = torch.as_tensor(months, device=dev)
        view = x.permute(1, 2, 0)  # what moving the grouped dim to the end gives: a VIEW
        for func in ("sum", "mean", "max"):
            best, mean = timeit(lambda: flox_b200.groupby_reduce(view, by, func=func, expected_groups=np.arange(12)), reps=3, warm=1)
            report(f"C3 f32 (3650,720,1440) by month {func}", T * ny * nx, 4, best, mean, "column-owner kernels, time-major view consumed in place")
        del x, view
        torch.cuda.empty_cache()

    if "c4" in which:
        n, B, G = 100_000_000, 8, 100_000
        x = torch.randn((n, B), device=dev, generator=gen, dtype=torch.float64)
        x.view(-1)[::7] = float("nan")
        by = torch.randint(0, G, (n,), device=dev, generator=gen)
        view = x.t()  # (8, n): reduced axis last, memory is (n, 8)
        # method="blockwise" means every group lives in ONE block (rechunk.py:158-223): sorted labels are that layout;
        # random labels are the worst case for the same call
        for lname, lab in (("sorted (blockwise layout)", torch.sort(by).values), ("random", by)):
            for func in ("nansum", "nanvar"):
                best, mean = timeit(lambda: flox_b200.groupby_reduce(view, lab, func=func, expected_groups=np.arange(G)), reps=3, warm=1)
                passes = 2 if "var" in func else 1
                report(f"C4 f64 (1e8,8) NaNs G=1e5 {func}, labels {lname}", n * B * passes, 9, best, mean,
                       f"{passes} pass(es) over the data, L2-atomic path, run-merging kernels (force_path={os.environ.get('FLOXB_FORCE_PATH', '0')})")
        del x, view
        torch.cuda.empty_cache()

    if "c5" in which:
        n, G = 500_000_000, 1_000_000  # one GPU's share of config 5 (4e9 rows / 8 GPUs), sorted labels
        by = torch.sort(torch.randint(0, G, (n,), device=dev, generator=gen)).values
        x = torch.rand(n, device=dev, generator=gen)
        arr, lab = ChunkedArray.from_array(x, 50_000_000), ChunkedArray.from_array(by, 50_000_000)
        for func in ("sum", "count"):
            best, mean = timeit(lambda: flox_b200.groupby_reduce(arr, lab, func=func, expected_groups=pd.RangeIndex(G), method="cohorts"), reps=5, warm=2)
            report(f"C5/8 f32[5e8] sorted labels G=1e6 {func} cohorts", n, 12, best, mean, "L2-atomic path with run merging")
        # the same rows with RANDOM labels: the unsorted large-G case
        byr = torch.randint(0, G, (n,), device=dev, generator=gen)
        best, mean = timeit(lambda: flox_b200.groupby_reduce(x, byr, func="sum", expected_groups=pd.RangeIndex(G)), reps=5, warm=2)
        report("f32[5e8] RANDOM labels G=1e6 sum", n, 12, best, mean, "L2-atomic path, one RED set per row")
        del x, by, byr

    if "rows" in which:
        # row-batched (A >> 1, B == 1): (lat*lon, time) contiguous, the layout flox produces after its reshape copy
        A, n, G = 720 * 144, 3650, 12
        x = torch.rand((A, n), device=dev, generator=gen)
        by = torch.as_tensor(((np.arange(n) % 365) // 30.42).astype(np.int64).clip(0, 11), device=dev)
        best, mean = timeit(lambda: flox_b200.groupby_reduce(x, by, func="sum", expected_groups=np.arange(G)), reps=3, warm=1)
        report("rows f32 (103680, 3650) by month sum", A * n, 4, best, mean, "warp-per-row kernel")

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
