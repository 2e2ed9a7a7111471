"""Where does the step's non-kernel time go?  Back-to-back timings (CUDA events, 50 repetitions each, no sleeps):
  A  k.moments only (init=False)                 B  k.moments with init=True (fill kernel + moments)
  C  B + finalize(mean)                          D  the public groupby_reduce step
  E  A with the clocks already loaded (after D)  F  C replayed from a CUDA graph
Run under gpurun: python tools/step_breakdown.py [rows]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import flox_b200  # noqa: E402
from flox_b200 import ChunkedArray, kernels  # noqa: E402
from flox_b200 import _lib as L  # noqa: E402
from flox_b200.kernels import PreparedChunks  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    G, chunk = 10_000, 10_000_000
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand(rows, device=dev, generator=gen)
    by = torch.randint(0, G, (rows,), device=dev, generator=gen, dtype=torch.int64)
    arr, lab = ChunkedArray.from_array(x, chunk), ChunkedArray.from_array(by, chunk)
    k = kernels.get_kernels()
    chunks = PreparedChunks([(v.reshape(1, -1, 1), c) for v, c in zip(arr.chunks, lab.chunks)])
    acc = torch.empty(G, dtype=torch.float64, device=dev)
    cnt = torch.empty(G, dtype=torch.int64, device=dev)
    nnan = torch.empty(G, dtype=torch.int64, device=dev)
    out = torch.empty(G, dtype=torch.float32, device=dev)
    eg = np.arange(G)

    def A():
        k.moments(chunks, 1, 1, G, L.OP_SUM, acc, cnt, nnan, init=False)

    def B():
        k.moments(chunks, 1, 1, G, L.OP_SUM, acc, cnt, nnan, init=True)

    def C():
        B()
        k.finalize(L.FIN_MEAN, torch.float32, acc, None, cnt, nnan, False, 0, 0, None, 0, out)

    def D():
        flox_b200.groupby_reduce(arr, lab, func="mean", expected_groups=eg, method="map-reduce")

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / n

    res = {"rows": rows}
    for name, fn in (("A_moments_only", A), ("B_fill+moments", B), ("C_fill+moments+finalize", C), ("D_groupby_reduce", D),
                     ("E_moments_only_again", A)):
        res[name] = timeit(fn)
        print(name, res[name], flush=True)
    # CUDA graph of C
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        C()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            C()
    torch.cuda.synchronize()
    res["F_graph_of_C"] = timeit(g.replay)
    print("F_graph_of_C", res["F_graph_of_C"], flush=True)
    # isolated kernel (with idle gaps), like bench.py's roofline leg
    ts = []
    for _ in range(10):
        torch.cuda._sleep(3_000_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); A(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    res["G_isolated_moments"] = float(np.mean(ts))
    print("G_isolated_moments", res["G_isolated_moments"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"step_breakdown_{rows}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
