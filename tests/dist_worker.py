"""Multi-GPU parity worker, launched by tests/test_gpu_multi.py (or by hand) under torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py

Every rank reduces its shard on its GPU; the merged result must equal the single-process oracle result on the
concatenated data (bit-exact for counts / arg positions, 1e-6 for floats), must be bit-identical on every rank, for both
cross-rank merges (the NVLink peer-memory exchange of csrc/exchange.cu and NCCL) and for every `method`.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import flox_b200  # noqa: E402
from flox_b200 import ChunkedArray, exchange, parallel  # noqa: E402
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
    # ties for the arg-reductions: the same extreme value in several ranks' shards (first occurrence must win)
    xi = rng.integers(0, 40, n).astype(np.float64)
    per = n // world
    lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
    ok = True
    notes = []

    def check(tag, got, want):
        nonlocal ok
        got = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
        good = got.dtype == want.dtype and got.shape == want.shape
        if good:
            good = np.array_equal(got, want) if want.dtype.kind in "iu" else np.allclose(got, want, rtol=1e-6, atol=3e-6, equal_nan=True)
        if not good:
            ok = False
            notes.append(f"[rank {rank}] MISMATCH {tag}: dtype {got.dtype}/{want.dtype}")
        # bit-identical on every rank
        t = torch.as_tensor(np.nan_to_num(got.astype(np.float64), nan=-1.25), device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        if not torch.equal(tmax, tmin):
            ok = False
            notes.append(f"[rank {rank}] ranks disagree on {tag}")

    for xchg in ("1", "0"):
        os.environ["FLOXB_PEER_XCHG"] = xchg
        for method in ("map-reduce", "cohorts", "blockwise"):
            for func, ddof, data in (("sum", 0, x), ("nansum", 0, x), ("count", 0, x), ("mean", 0, x), ("nanmean", 0, x),
                                     ("nanmax", 0, x), ("min", 0, x), ("nanvar", 1, x), ("nanstd", 0, x),
                                     ("nanargmax", 0, x), ("argmin", 0, xi), ("argmax", 0, xi)):
                if func.startswith(("arg", "nanarg")) and method == "blockwise":
                    continue  # rejected like the reference (core.py:1058-1062)
                fk = {"ddof": ddof} if ddof else None
                xl = torch.as_tensor(data[lo:hi], device="cuda")
                bl = torch.as_tensor(by[lo:hi], device="cuda")
                if method == "map-reduce" and not func.startswith(("arg", "nanarg")):
                    want, _ = O.groupby_reduce_mapreduce(data, by, func=func, nchunks=8, expected_groups=np.arange(G), ddof=ddof)
                else:
                    want, _ = O.groupby_reduce(data, by, func=func, expected_groups=np.arange(G), ddof=ddof)
                with parallel.distributed(sharding="local"):
                    got, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(xl, 250_000), ChunkedArray.from_array(bl, 250_000),
                                                      func=func, expected_groups=np.arange(G), method=method, finalize_kwargs=fk)
                check(f"{func}/{method}/xchg{xchg}", got, want)
    # group-sharded output: sorted labels, only the groups that straddle ranks are merged; every rank's result is complete
    # for the groups its rows touch
    os.environ["FLOXB_PEER_XCHG"] = "1"
    bys = np.sort(by)
    mine = np.unique(bys[lo:hi])
    for func, ddof in (("sum", 0), ("count", 0), ("nanmean", 0), ("nanmax", 0), ("nanvar", 1), ("nanargmax", 0), ("nanlast", 0)):
        fk = {"ddof": ddof} if ddof else None
        want, _ = O.groupby_reduce(x, bys, func=func, expected_groups=np.arange(G), ddof=ddof) if func != "nanlast" else (None, None)
        xl, bl = torch.as_tensor(x[lo:hi], device="cuda"), torch.as_tensor(bys[lo:hi], device="cuda")
        with parallel.distributed(sharding="local", output="sharded"):
            got, _ = flox_b200.groupby_reduce(ChunkedArray.from_array(xl, 250_000), ChunkedArray.from_array(bl, 250_000), func=func,
                                              expected_groups=np.arange(G), method="cohorts", finalize_kwargs=fk)
        got = got.cpu().numpy()
        if want is None:
            continue
        good = got.dtype == want.dtype and (np.array_equal(got[mine], want[mine]) if want.dtype.kind in "iu" else
                                            np.allclose(got[mine], want[mine], rtol=1e-6, atol=3e-6, equal_nan=True))
        if not good:
            ok = False
            notes.append(f"[rank {rank}] MISMATCH sharded/{func}")
    # the exchange was really used, and no barrier ever timed out
    used = [ex for ex in exchange._exchanges.values() if ex]
    if not used or used[0].epoch == 0:
        ok = False
        notes.append(f"[rank {rank}] the peer-memory exchange was never used")
    elif used[0].status() != 0:
        ok = False
        notes.append(f"[rank {rank}] exchange barrier timed out in collective {used[0].status()}")
    for line in notes[:20]:
        print(line, flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ep = used[0].epoch if used else 0
        print("dist_worker", "OK" if int(flag.item()) else "FAILED", f"world={world} peer-exchange collectives={ep}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
