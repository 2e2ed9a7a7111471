"""Time the L2-atomic layout variants in tools/kexp2.cu (BASELINE config-4 shape) on the GPU."""
import ctypes as C
import json
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkexp2.so")
NAMES = ["[B][G] val+cnt per elem (product)", "[G][B] val+cnt per elem", "[G][B] val per elem + row count", "[B][G] val per elem + row count",
         "loads only"]


def build():
    src = os.path.join(HERE, "kexp2.cu")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
                        "-fPIC", "-shared", "-o", LIB, src], check=True)


def main():
    build()
    if not torch.cuda.is_available():
        print("built", LIB)
        return
    lib = C.CDLL(LIB)
    lib.kexp2_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    gen = torch.Generator(device="cuda").manual_seed(3)
    n, B = 50_000_000, 8
    x = torch.randn((n, B), device="cuda", generator=gen, dtype=torch.float64)
    x.view(-1)[::7] = float("nan")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = []
    for G in (100_000, 1_000_000, 10_000):
        by = torch.randint(0, G, (n,), device="cuda", generator=gen)
        acc = torch.zeros(G * B, dtype=torch.float64, device="cuda")
        cnt, nan = torch.zeros(G * B, dtype=torch.int64, device="cuda"), torch.zeros(G * B, dtype=torch.int64, device="cuda")
        rows = torch.zeros(G, dtype=torch.int64, device="cuda")
        for variant in range(5):
            ts = []
            for _ in range(3):
                acc.zero_(); cnt.zero_(); nan.zero_(); rows.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = lib.kexp2_run(variant, x.data_ptr(), by.data_ptr(), n, B, G, acc.data_ptr(), cnt.data_ptr(), nan.data_ptr(),
                                   rows.data_ptr(), sms * 16, torch.cuda.current_stream().cuda_stream)
                b.record(); b.synchronize()
                assert rc == 0
                ts.append(a.elapsed_time(b))
            ms = min(ts[1:])
            row = dict(G=G, variant=variant, name=NAMES[variant], ms=ms, gbs=(n * B * 8 + n * 8) / ms / 1e6)
            out.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(HERE, "..", "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(HERE, "..", "gpurun_out", "kbench2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
