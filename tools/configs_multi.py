"""BASELINE.json configs 3, 4, 5 on N GPUs of one box (torchrun, one process per GPU), through the public API:

  C3  f32 (time=3650, lat=720, lon=1440) by month, sum/mean/max — LAT-sharded: outputs are disjoint, no reduction
  C4  f64 (1e8, 8) with NaNs, G = 1e5, nanvar, method="blockwise" — rows sharded, sorted (blockwise-layout) labels
  C5  f32 4e9 rows (or what fits), G = 1e6, sum + count, method="cohorts", sorted labels — rows sharded
CUDA events, barrier + synchronize on both sides, max over ranks.  Writes gpurun_out/configs_multi_N.json on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/configs_multi.py [c3 c4 c5]
"""
import json
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import flox_b200  # noqa: E402
from flox_b200 import ChunkedArray, parallel  # noqa: E402

PEAK = 6570.3


def main():
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev).manual_seed(12345 + rank)
    out = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timeit(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        for _ in range(20):
            fn()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record()
            fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return min(ts), float(np.mean(ts))

    def report(name, total_elems, bytes_per_elem, best, mean, note):
        gbs = total_elems * bytes_per_elem / best / 1e6
        row = dict(config=name, n_gpus=world, elements=total_elems, best_ms=best, mean_ms=mean, elems_per_s=total_elems / best * 1e3,
                   algorithmic_GBs_all_gpus=gbs, frac_of_measured_hbm_per_gpu=gbs / world / PEAK, note=note)
        out.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    if "c3" in which:
        T, ny, nx = 3650, 720, 1440
        ny_l = ny // world
        x = torch.rand((T, ny_l, nx), device=dev, generator=gen)
        months = torch.as_tensor(((np.arange(T) % 365) // 30.42).astype(np.int64).clip(0, 11), device=dev)
        view = x.permute(1, 2, 0)
        for func in ("sum", "mean", "max"):
            best, mean = timeit(lambda: flox_b200.groupby_reduce(view, months, func=func, expected_groups=np.arange(12)), reps=3, warm=1)
            report(f"C3 f32 (3650,720,1440) by month {func}", T * ny_l * world * nx, 4, best, mean,
                   "lat-sharded: every GPU reduces its own latitudes, outputs disjoint, no collective")
        del x, view
        torch.cuda.empty_cache()

    if "c4" in which:
        n, B, G = 100_000_000 // world, 8, 100_000
        x = torch.randn((n, B), device=dev, generator=gen, dtype=torch.float64)
        x.view(-1)[::7] = float("nan")
        # blockwise layout: rank r holds the rows of groups [r*G/W, (r+1)*G/W), sorted
        lo, hi = rank * G // world, (rank + 1) * G // world
        by = torch.sort(torch.randint(lo, hi, (n,), device=dev, generator=gen)).values
        arr, lab = ChunkedArray([x.t()]), ChunkedArray([by])
        lab.static_labels = True
        for func in ("nansum", "nanvar"):
          for output in ("sharded", "replicated"):
            def fn():
                with parallel.distributed(sharding="local", output=output):
                    flox_b200.groupby_reduce(arr, lab, func=func, expected_groups=np.arange(G), method="blockwise")
            best, mean = timeit(fn, reps=3, warm=1)
            passes = 2 if "var" in func else 1
            report(f"C4 f64 (1e8,8) NaNs G=1e5 {func} blockwise, output={output}", n * world * B * passes, 9, best, mean,
                   f"{passes} pass(es); rows sharded, every group in one rank's block; " +
                   ("nothing to merge: the result stays sharded by group (blockwise, dask.py:390-392)" if output == "sharded"
                    else "[G,8] accumulators all-reduced"))
        del x, by, arr, lab
        torch.cuda.empty_cache()

    if "c5" in which:
        n, G = 500_000_000, 1_000_000   # 4e9 rows at 8 GPUs
        lo, hi = rank * G // world, (rank + 1) * G // world
        by = torch.sort(torch.randint(lo, min(hi + 1, G), (n,), device=dev, generator=gen)).values  # neighbours share a boundary group
        x = torch.rand(n, device=dev, generator=gen)
        arr, lab = ChunkedArray.from_array(x, 50_000_000), ChunkedArray.from_array(by, 50_000_000)
        lab.static_labels = True   # the labels of this loop never change: the straddling groups are found once
        for output in ("sharded", "replicated"):
            for func in ("sum", "count"):
                def fn():
                    with parallel.distributed(sharding="local", output=output):
                        flox_b200.groupby_reduce(arr, lab, func=func, expected_groups=pd.RangeIndex(G), method="cohorts")
                best, mean = timeit(fn, reps=5, warm=2)
                report(f"C5 f32[{n * world:.1e}] sorted labels G=1e6 {func} cohorts, output={output}", n * world, 12, best, mean,
                       "rows sharded in label order; boundary groups straddle neighbouring ranks; " +
                       ("only the straddling groups are merged, the result stays sharded by group (dask.py:469-518)"
                        if output == "sharded" else "the full [G] accumulators are all-reduced"))
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"configs_multi_{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
