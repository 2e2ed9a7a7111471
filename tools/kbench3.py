"""Bisect the production two-level kernel against the bare pipelined loop (tools/kexp3.cu); run under gpurun."""
import ctypes as C
import json
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkexp3.so")


def build():
    src = os.path.join(HERE, "kexp3.cu")
    dep = os.path.join(HERE, "..", "flox_b200", "csrc", "common.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
                        "-fPIC", "-shared", "-o", LIB, src], check=True)
    return LIB


def main():
    build()
    if not torch.cuda.is_available():
        print("built", LIB)
        return
    lib = C.CDLL(LIB)
    lib.kexp3_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_void_p]
    lib.kexp3_name.restype = C.c_char_p
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 480_000_000
    nchunks = 48
    G = int(os.environ.get("KBENCH_G", "10000"))
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, device="cuda", generator=gen)
    by = torch.randint(0, G, (n,), device="cuda", generator=gen, dtype=torch.int64)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    acc = torch.zeros(G, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(G, dtype=torch.int64, device="cuda")
    ws = torch.empty(sms * (G * 12 + 64) // 8 + 8, dtype=torch.float64, device="cuda")
    variants = [int(v) for v in os.environ.get("KBENCH_VARIANTS", "0,6,8,10,11,12,13,14,15,16,17,18,19,20,21").split(",")]
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for rnd in range(2):  # two rounds: order effects / clock drift show up as disagreement
        for variant in variants:
            ts = []
            for rep in range(5):
                acc.zero_(); cnt.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = lib.kexp3_run(variant, x.data_ptr(), by.data_ptr(), n, nchunks, G, acc.data_ptr(), cnt.data_ptr(), ws.data_ptr(), sms, st)
                b.record(); b.synchronize()
                assert rc == 0, rc
                ts.append(a.elapsed_time(b))
            ms = min(ts[1:])
            name = lib.kexp3_name(variant).decode()
            walked = "walk" in name
            expect = n if walked else (n // 4096) * 4096
            probe = "probe" in name  # throughput-only variants: their counters are not meaningful
            row = dict(round=rnd, variant=variant, name=name, ms=ms, gbs=12 * n / ms / 1e6,
                       ok=None if probe else bool(int(cnt.sum()) == expect),
                       sum_ok=bool(abs(float(acc.sum()) - float(x.double().sum())) < 1e-6 * n) if (walked and not probe) else None)
            out.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(HERE, "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(HERE, "..", "gpurun_out", "kbench3.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
