"""Multi-GPU parity check, run under torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank reduces its shard on its GPU; the NCCL-merged result must equal the single-process
oracle result on the concatenated data (bit-exact for counts / arg positions, 1e-6 for floats).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import flox_b200  # noqa: E402
from flox_b200 import ChunkedArray, parallel  # noqa: E402
from oracle import flox_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(2024)
    n, G = 2_000_000, 10_000
    x = (rng.random(n) * 3 - 1).astype(np.float32)
    x[::101] = np.nan
    by = rng.integers(0, G, n)
    by[by == 17] = 18
    per = n // world
    lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
    xl = torch.as_tensor(x[lo:hi], device="cuda")
    bl = torch.as_tensor(by[lo:hi], device="cuda")
    ok = True
    for func, ddof in (("sum", 0), ("nansum", 0), ("count", 0), ("nanmean", 0), ("nanmax", 0), ("min", 0), ("nanvar", 1),
                       ("nanargmax", 0)):
        fk = {"ddof": ddof} if ddof else None
        want, _ = O.groupby_reduce(x, by, func=func, expected_groups=np.arange(G), ddof=ddof)
        with parallel.distributed(sharding="local"):
            got, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(xl, 250_000), ChunkedArray.from_array(bl, 250_000),
                                              func=func, expected_groups=np.arange(G), method="cohorts", finalize_kwargs=fk)
        got = got.cpu().numpy()
        if want.dtype.kind in "iu":
            good = np.array_equal(got, want)
        else:
            good = np.allclose(got, want, rtol=1e-6, atol=3e-6, equal_nan=True)
        good = good and got.dtype == want.dtype
        if not good:
            ok = False
            bad = np.nonzero(~np.isclose(got, want, rtol=1e-6, atol=3e-6, equal_nan=True))[0][:5]
            print(f"[rank {rank}] MISMATCH {func}: idx {bad} got {got[bad]} want {want[bad]}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_check", "OK" if int(flag.item()) else "FAILED", f"world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
