"""Time the kernel-design variants in tools/kexp.cu on the GPU (run under gpurun)."""
import ctypes as C
import json
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkexp.so")


def build():
    src = os.path.join(HERE, "kexp.cu")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
                        "-fPIC", "-shared", "-o", LIB, src], check=True)
    return LIB


def main():
    build()
    if not torch.cuda.is_available():
        print("built", LIB)
        return
    lib = C.CDLL(LIB)
    lib.kexp_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.kexp_name.restype = C.c_char_p
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
    out = []
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, device="cuda", generator=gen)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    VARIANTS = [int(v) for v in os.environ.get("KBENCH_VARIANTS", ",".join(map(str, range(26)))).split(",")]
    for G in [int(g) for g in os.environ.get("KBENCH_GROUPS", "10000,100,1000").split(",")]:
        by = torch.randint(0, G, (n,), device="cuda", generator=gen, dtype=torch.int64)
        acc = torch.zeros(G, dtype=torch.float64, device="cuda")
        cnt = torch.zeros(G, dtype=torch.int64, device="cuda")
        for variant in VARIANTS:
            for grid in (sms,):
                st = torch.cuda.current_stream().cuda_stream
                ts = []
                for rep in range(4):
                    acc.zero_(); cnt.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    rc = lib.kexp_run(variant, x.data_ptr(), by.data_ptr(), n, G, acc.data_ptr(), cnt.data_ptr(), grid, st)
                    b.record(); b.synchronize()
                    assert rc == 0, rc
                    ts.append(a.elapsed_time(b))
                ms = min(ts[1:])
                ok = None
                name = lib.kexp_name(variant).decode()
                if name.startswith(("full", "packed", "f32", "hybrid")):
                    t_, i_ = int(name.split(" t")[-1].split()[0]), int(name.split(" i")[-1])
                    ok = bool(int(cnt.sum()) == (n // (t_ * 4 * i_)) * (t_ * 4 * i_))
                row = dict(G=G, variant=variant, name=name, grid=grid, ms=ms, rows_per_s=n / ms * 1e3, gbs=12 * n / ms / 1e6, ok=ok)
                out.append(row)
                print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(HERE, "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(HERE, "..", "gpurun_out", "kbench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
