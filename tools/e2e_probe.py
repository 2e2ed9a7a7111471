"""Where does the end-to-end step spend its time?  PCIe copy rates for 40 MB f32 / 20 MB int16 pinned chunks,
host pack throughput vs. thread count, with and without a concurrent H2D stream.  Run under gpurun."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from flox_b200.kernels import CudaKernels  # noqa: E402

n, nch, G = 10_000_000, 100, 10_000
dev = torch.device("cuda:0")
hx = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(nch)]
hb = [torch.empty(n, dtype=torch.int64, pin_memory=True) for _ in range(nch)]
rng = np.random.default_rng(0)
for h in hb:
    h.numpy()[...] = rng.integers(0, G, n)
st = [torch.empty(n, dtype=torch.int16, pin_memory=True) for _ in range(nch)]
dx = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(nch)]
d16 = [torch.empty(n, dtype=torch.int16, device=dev) for _ in range(nch)]


def copy_all(srcs, dsts):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for s, d in zip(srcs, dsts):
        d.copy_(s, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    b = sum(s.numel() * s.element_size() for s in srcs)
    return dt, b / dt / 1e9


for rep in range(2):
    print("values  H2D: %.1f ms %.1f GB/s" % (copy_all(hx, dx)[0] * 1e3, copy_all(hx, dx)[1]))
    print("int16   H2D: %.1f ms %.1f GB/s" % (copy_all(st, d16)[0] * 1e3, copy_all(st, d16)[1]))
for threads in (8, 16, 32, 64, 128):
    with ThreadPoolExecutor(threads) as ex:
        for rep in range(2):
            t = time.perf_counter()
            list(ex.map(lambda i: CudaKernels.host_pack_codes(hb[i].numpy(), G, st[i]), range(nch)))
            dt = time.perf_counter() - t
        print("pack %3d threads: %.1f ms (%.1f GB/s read)" % (threads, dt * 1e3, nch * n * 8 / dt / 1e9))
        # with a concurrent value copy
        torch.cuda.synchronize()
        t = time.perf_counter()
        for s, d in zip(hx, dx):
            d.copy_(s, non_blocking=True)
        list(ex.map(lambda i: CudaKernels.host_pack_codes(hb[i].numpy(), G, st[i]), range(nch)))
        tp = time.perf_counter() - t
        torch.cuda.synchronize()
        tc = time.perf_counter() - t
        print("   with concurrent 4 GB H2D: pack done %.1f ms, copy done %.1f ms" % (tp * 1e3, tc * 1e3))
