"""Host profile of one C5-shaped call (sorted labels, 10 chunks, cohorts) — where does the wall time go?"""
import cProfile
import os
import pstats
import sys
import time

import pandas as pd
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flox_b200  # noqa: E402
from flox_b200 import ChunkedArray  # noqa: E402

dev = torch.device("cuda")
n, G = 500_000_000, 1_000_000
gen = torch.Generator(device=dev).manual_seed(1)
by = torch.sort(torch.randint(0, G, (n,), device=dev, generator=gen)).values
x = torch.rand(n, device=dev, generator=gen)
arr, lab = ChunkedArray.from_array(x, 50_000_000), ChunkedArray.from_array(by, 50_000_000)
f = lambda: flox_b200.groupby_reduce(arr, lab, func="sum", expected_groups=pd.RangeIndex(G), method="cohorts")  # noqa: E731
for _ in range(20):
    f()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    f()
torch.cuda.synchronize()
print("wall per call ms", (time.perf_counter() - t) / 20 * 1e3)
# host-only time per call (no sync inside?): enqueue 20 calls, measure before the final sync
t = time.perf_counter()
for _ in range(20):
    f()
t_enq = time.perf_counter() - t
torch.cuda.synchronize()
print("enqueue per call ms", t_enq / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    f()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
