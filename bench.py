#!/usr/bin/env python
"""bench.py — grouped rows/s of the flox hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): 1-D float32, 1e9 rows in chunks of 1e7, int64 labels, 1e4 random groups,
func="mean", method="map-reduce".  One "step" = one full groupby_reduce over the job's rows through the public API.

  --scaling strong (default)  the metric as written: the problem is FIXED at 1e9 rows; N GPUs hold 1e9/N rows each
  --scaling weak              1e9 rows PER GPU
The line's `value` is the primary mode's rows/s; the other mode is measured in the same run and reported under
`other_scaling` (at N = 1 they are the same run).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the UNMODIFIED reference (baseline/_ref) on the host cores

One JSON line on rank 0:
  value        rows/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          the same through the public API with pinned HOST buffers (H2D + D2H inside the timed region)
  roofline     achieved GB/s of the dominant kernel (12 algorithmic bytes per row) over the measured HBM peak
  parity       the in-run parity gate: counts bit-exact against torch.bincount, sums / means within 1e-6 of an f64
               index_add_ on the device, and (N = 1) the GPU result against the reference's own chunk_reduce +
               simple_combine + _finalize_results on the sample the CPU baseline timed, same arrays
  cpu_baseline the reference's engine="flox" path (baseline/_ref, installed by oracle/install_reference.py) on the host
               cores, bounded sample; the oracle port is used only if the installed reference is missing (kind says which)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grouped rows/sec + %HBM-roofline, 1e9×f32 sum by 1e4 groups @1/2/4/8 GPU vs CPU ref"
UNIT = "rows/s"
ROWS, CHUNK, NGROUPS, FUNC = 1_000_000_000, 10_000_000, 10_000, "mean"
BYTES_PER_ROW = 12  # f32 value + i64 label, each read once


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--rows", type=int, default=ROWS, help="rows of the job (strong scaling) / per GPU (weak scaling)")
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--no-other-scaling", action="store_true", help="skip the secondary scaling mode at N > 1")
    p.add_argument("--chunk", type=int, default=CHUNK)
    p.add_argument("--groups", type=int, default=NGROUPS)
    p.add_argument("--func", default=FUNC)
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=20.0)
    p.add_argument("--profile-host", action="store_true", help="cProfile 20 steps of the host-side path (stderr)")
    return p.parse_args()


# --------------------------------------------------------------------------------------------
# CPU legs — the only place bench.py executes baseline/_ref (the reference) or oracle/ (its port)
# --------------------------------------------------------------------------------------------

def cpu_chunks(nchunks: int, chunk: int, ngroups: int, seed: int = 12345):
    # 8 distinct seeded chunks, cycled: generating one per host thread would cost more wall-clock
    # than the timed region on a 128-core box and does not change what is timed
    rng = np.random.default_rng(seed)
    distinct = [(rng.random(chunk, dtype=np.float32), rng.integers(0, ngroups, chunk)) for _ in range(min(nchunks, 8))]
    return [distinct[i % len(distinct)] for i in range(nchunks)]


_REF = None


def reference_fns():
    """The unmodified reference's own functions (baseline/_ref on the GPU box, /root/reference in the build
    container), or None when neither is present."""
    global _REF
    if _REF is None:
        try:
            from oracle import ref_loader

            ref_loader.load_reference(prefer="installed")
            import pandas as pd
            from flox.aggregations import _initialize_aggregation
            from flox.core import DUMMY_AXIS, _finalize_results, chunk_reduce
            from flox.reindex import ReindexStrategy

            _REF = dict(pd=pd, init=_initialize_aggregation, chunk_reduce=chunk_reduce, finalize=_finalize_results,
                        DUMMY_AXIS=DUMMY_AXIS, ReindexStrategy=ReindexStrategy, where=ref_loader.reference_location("installed"))
        except Exception as exc:  # noqa: BLE001 - any import problem means "no reference here"
            _REF = {"error": repr(exc)}
    return None if "error" in _REF else _REF


def cpu_mapreduce(chunks, ngroups: int, func: str, threads: int):
    """flox's map-reduce on the CPU, one chunk per worker thread (numpy's sort and reduceat release the GIL).

    kind "reference": the reference's OWN chunk_reduce(engine="flox") per chunk (core.py:214-394: one stable argsort per
    chunk, then reduceat per func), its agg.simple_combine over the stacked intermediates (dask.py:131-143) and its
    _finalize_results (core.py:410-475) — what a dask graph executes, minus the scheduler.
    kind "port": oracle/flox_oracle.py, the numpy restatement of the same algorithm (only when baseline/_ref is absent)."""
    from concurrent.futures import ThreadPoolExecutor

    R = reference_fns()
    if R is not None:
        agg = R["init"](func, None, np.dtype("float32"), None, 0, None)
        if agg.fill_value["user"] is None:
            agg.fill_value["user"] = agg.fill_value[agg.name]
        expected = R["pd"].RangeIndex(ngroups)
        D = R["DUMMY_AXIS"]

        def one(c):
            r = R["chunk_reduce"](c[0], c[1], func=agg.chunk, expected_groups=expected, axis=(0,),
                                  fill_value=agg.fill_value["intermediate"], dtype=agg.dtype["intermediate"], reindex=True,
                                  engine="flox", user_dtype=agg.dtype["user"])
            return [np.expand_dims(x, D) for x in r["intermediates"]]

        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(one, chunks))
        combined = []
        for i, comb in enumerate(agg.simple_combine):
            stacked = np.concatenate([p[i] for p in parts], axis=D)
            combined.append(comb(stacked, axis=(stacked.ndim - 2,), keepdims=True).squeeze(stacked.ndim - 2))
        fin = R["finalize"]({"groups": expected, "intermediates": combined}, agg, (0,), expected,
                            reindex=R["ReindexStrategy"](blockwise=True))
        return fin[agg.name]

    from oracle import flox_oracle as O

    agg = O.initialize(func, None, np.dtype("float32"), None, 0)
    agg.fill["user"] = agg.fill["final"]

    def one_port(c):
        return O.chunk_reduce(c[0], c[1], agg.chunk, ngroups=ngroups, fills=agg.fill["intermediate"],
                              dtypes=agg.dtype["intermediate"])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        per_chunk = list(ex.map(one_port, chunks))
    return O.finalize(O.simple_combine(per_chunk, agg), agg, do_finalize=True)


def cpu_kind() -> tuple[str, str]:
    R = reference_fns()
    if R is not None:
        return "reference", (f"the UNMODIFIED reference (flox @ {R['where']}): chunk_reduce(engine='flox') per chunk + "
                             f"agg.simple_combine + _finalize_results")
    return "port", "oracle port of engine='flox' (stable argsort once per chunk + reduceat) + simple_combine + finalize"


def best_cpu_threads(args) -> tuple[int, dict]:
    """The thread count at which the reference is FASTEST on this host (more threads than the container's
    CPU share only add contention: the label-pack probe ran 3x slower with 128 threads than with 16).  Calibrated on
    1e6-row chunks, one chunk per thread."""
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (8, 16, 32, 64, 128, ncpu) if t <= ncpu}) or [1]
    small = cpu_chunks(max(cands), min(1_000_000, args.chunk, args.rows), args.groups, seed=7)
    rates = {}
    for t in cands:
        cpu_mapreduce(small[:t], args.groups, args.func, t)
        t0 = time.perf_counter()
        cpu_mapreduce(small[:t], args.groups, args.func, t)
        rates[t] = t * small[0][0].size / (time.perf_counter() - t0)
    best = max(rates, key=rates.get)
    return best, {str(k): round(v) for k, v in rates.items()}


def time_cpu(args, budget_s: float, host_chunks=None):
    """rows/s of the reference on the host cores (at its best thread count), bounded to ~budget_s seconds.
    `host_chunks`: [(values, labels)] numpy views of the SAME arrays the GPU arm reduced (one per thread is used)."""
    threads, _ = best_cpu_threads(args)
    chunk = min(args.chunk, args.rows)
    if host_chunks is not None and len(host_chunks):
        chunks = [host_chunks[i % len(host_chunks)] for i in range(threads)]
        chunk = int(chunks[0][0].size)
    else:
        chunks = cpu_chunks(threads, chunk, args.groups)
    t0 = time.perf_counter()
    cpu_mapreduce(chunks[: max(1, threads // 4)], args.groups, args.func, threads)  # warm-up + cost probe
    probe = (time.perf_counter() - t0)
    rounds = max(1, min(5, int(budget_s / max(probe * 1.5, 1e-3))))
    best = None
    for _ in range(rounds):
        t0 = time.perf_counter()
        cpu_mapreduce(chunks, args.groups, args.func, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    rows = threads * chunk
    kind, how = cpu_kind()
    return dict(value=rows / best, unit=UNIT, cores=threads, kind=kind,
                sample=f"{threads} chunks x {chunk} rows of the same workload (one chunk per host thread, "
                       f"{'the arrays the GPU arm reduced' if host_chunks is not None else 'seeded numpy arrays'}), best of {rounds}; {how}"), rows, best


def run_reference(args):
    """--impl reference: the reference's own CPU path, same metric/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads, calib = best_cpu_threads(args)
    chunk = min(args.chunk, args.rows)
    chunks = cpu_chunks(threads, chunk, args.groups)
    for _ in range(min(args.warmup, 1)):
        cpu_mapreduce(chunks, args.groups, args.func, threads)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_mapreduce(chunks, args.groups, args.func, threads)
    dt = (time.perf_counter() - t0) / steps
    rows = threads * chunk
    value = rows / dt
    kind, how = cpu_kind()
    sample = (f"each step = {threads} chunks x {chunk} rows (bounded sample of the 1e9-row workload, one chunk per host "
              f"thread); {steps} timed steps; thread count = the fastest of a 1e6-row calibration {calib} rows/s; {how}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, cpu=True)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args, cpu: bool = False) -> str:
    """The SAME string for both arms (the driver compares configs)."""
    return (f"1D float32 len={args.rows:.0e} chunked {min(args.chunk, args.rows):.0e}, {args.groups} random int64-labelled "
            f"groups, func={args.func!r}, method='map-reduce'")


# --------------------------------------------------------------------------------------------
# GPU legs
# --------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clocks + throttle reasons DURING the timed region.  NVML in-process (a light query every 20 ms from
    a daemon thread); an `nvidia-smi -lms` child process — the B200_PROFILING.md recipe — is only the fallback,
    because its start-up and polling take driver locks that stalled launches inside short timed regions."""

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._thread, self.proc = index, [], False, None, None
        self.how = None

    def _nvml_loop(self, pynvml, handle):
        reasons = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                   "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                   "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                   "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}
        while not self._stop:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((float(sm), [k for k, bit in reasons.items() if mask & bit]))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES remapping
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self._thread.start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.sm_max = None
            self._thread = threading.Thread(target=self._smi_loop, daemon=True)
            self._thread.start()
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None

    def _smi_loop(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 6 and c[0].replace(".", "").isdigit():
                self.sm_max = float(c[1]) if c[1].replace(".", "").isdigit() else self.sm_max
                self.samples.append((float(c[0]), [n for n, v in zip(names, c[2:6]) if v.lower().startswith("active")]))

    @property
    def rows(self):
        return self.samples

    def mark(self):
        """Forget what was sampled so far (call right before the timed region)."""
        self.samples = []

    def stop(self):
        self._stop = True
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
        elif self._thread is not None:
            self._thread.join(timeout=1.0)
        sm = [s for s, _ in self.samples]
        reasons = sorted({r for _, rs in self.samples for r in rs})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "sm_max", None),
                "reasons": reasons, "samples": len(sm), "how": self.how}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes/launch of the dominant kernel from the committed ncu capture, if summarised."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import flox_b200
    from flox_b200 import ChunkedArray, kernels, parallel
    from flox_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_cpus = parallel.bind_to_gpu_numa(local)  # before any pinned allocation: staging buffers land on the GPU's node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k = kernels.get_kernels()
    G = args.groups
    eg = np.arange(G)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_inputs(rows_local: int):
        chunk = min(args.chunk, rows_local)
        gen = torch.Generator(device=dev).manual_seed(12345 + rank)
        x = torch.rand(rows_local, device=dev, dtype=torch.float32, generator=gen)
        by = torch.randint(0, G, (rows_local,), device=dev, dtype=torch.int64, generator=gen)
        return x, by, ChunkedArray.from_array(x, chunk), ChunkedArray.from_array(by, chunk)

    def make_step(arr, lab):
        def step():
            with parallel.distributed(sharding="local"):
                res, _ = flox_b200.groupby_reduce(arr, lab, func=args.func, expected_groups=eg, method="map-reduce")
            return res
        return step

    def timed(step, steps: int):
        """K steps between CUDA events, barrier + synchronize on both sides, max over ranks -> ms per step."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            res = step()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, res

    # ---- primary scaling mode ---------------------------------------------------------------------------------
    rows_local = args.rows // world if args.scaling == "strong" else args.rows
    job_rows = rows_local * world
    x, by, arr, lab = make_inputs(rows_local)
    step = make_step(arr, lab)

    # the clock sampler is started BEFORE warm-up and given time to settle: its start-up takes driver locks that
    # would otherwise stall launches inside the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 4.0:
            time.sleep(0.05)
    if args.profile_host and rank == 0:
        import cProfile
        import pstats

        prof = cProfile.Profile()
        prof.enable()
        for _ in range(20):
            step()
        torch.cuda.synchronize()
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumtime").print_stats(35)
    # Spin-up (untimed, in addition to the W warm-up steps): keep the GPU busy for ~0.3 s (100 steps) so that the timed
    # region starts in steady state.  Waiting for the clock sampler above idles the device, and the first
    # ~20 steps after an idle gap ran 1.6x slower than steady state (profiles/r01: 5.2 vs 3.2 ms/step).
    for _ in range(100):  # a FIXED count: every rank must issue the same number of collectives
        step()
    barrier()
    if rank == 0:
        sampler.mark()
    l0 = k.kernel_launches
    ms_per_step, res = timed(step, args.steps)
    launches = k.kernel_launches - l0
    value = job_rows / (ms_per_step * 1e-3)
    # host time per step (enqueue only, no sync inside): what the device must hide
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    host_ms = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize()

    # ---- parity gate on the timed result (every rank holds the merged result) -----------------------------------------
    def gate(res, x, by, rows_total):
        cnt_ref = torch.bincount(by, minlength=G)
        sum_ref = torch.zeros(G, dtype=torch.float64, device=dev).index_add_(0, by, x.to(torch.float64))
        if world > 1:
            dist.all_reduce(cnt_ref)
            dist.all_reduce(sum_ref)
        with parallel.distributed(sharding="local"):
            cnt, _ = flox_b200.groupby_reduce(arr, lab, func="count", expected_groups=eg, method="map-reduce")
            ssum, _ = flox_b200.groupby_reduce(arr, lab, func="sum", expected_groups=eg, method="map-reduce")
        want = (sum_ref / cnt_ref.to(torch.float64)) if args.func in ("mean", "nanmean") else sum_ref
        rel = ((res.to(torch.float64) - want).abs() / want.abs().clamp_min(1e-300)).max().item()
        rel_sum = ((ssum.to(torch.float64) - sum_ref).abs() / sum_ref.abs().clamp_min(1e-300)).max().item()
        out = {"count_bit_exact": bool(torch.equal(cnt.to(torch.int64), cnt_ref)), "count_total": int(cnt.sum().item()),
               "count_total_expected": int(rows_total), f"max_rel_err_{args.func}_vs_f64": rel, "max_rel_err_sum_vs_f64": rel_sum,
               "rtol": 1e-6}
        out["ok"] = bool(out["count_bit_exact"] and out["count_total"] == rows_total and rel <= 1e-6 and rel_sum <= 1e-6)
        return out

    parity = gate(res, x, by, job_rows)

    # ---- dominant kernel alone (roofline): floxb_moments over this rank's chunks, CUDA events
    # one launch covers a table of up to 128 chunks (csrc/common.cuh: kMaxChunksPerLaunch): the step's chunks
    # are exactly one launch of the dominant kernel; time that launch
    per_launch = k.max_chunks_per_launch()
    chunks = [(v.reshape(1, -1, 1), c) for v, c in zip(arr.chunks, lab.chunks)][:per_launch]
    launch_rows = sum(c.numel() for _, c in chunks)
    acc = torch.empty(G, dtype=torch.float64, device=dev)
    cnt = torch.empty(G, dtype=torch.int64, device=dev)
    nnan = torch.empty(G, dtype=torch.int64, device=dev)
    for _ in range(2):
        k.moments(chunks, 1, 1, G, L.OP_SUM, acc, cnt, nnan, init=False)
    torch.cuda.synchronize()
    reps = max(3, min(args.steps, 10))
    kt = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # ~1.5 ms of device-side spinning queued first: the host builds the chunk table and enqueues the
        # launch while it runs, so the interval between the two events is the kernel's duration and not the
        # host's enqueue latency on an idle GPU (~0.09 ms here, 7% of the kernel)
        torch.cuda._sleep(3_000_000)
        a.record()
        k.moments(chunks, 1, 1, G, L.OP_SUM, acc, cnt, nnan, init=False)
        b.record()
        b.synchronize()
        kt.append(a.elapsed_time(b))
    kernel_ms = float(np.mean(kt))
    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed steps and the kernel-only loop
    peak, peak_src = hbm_peak()
    achieved = BYTES_PER_ROW * launch_rows / (kernel_ms * 1e-3) / 1e9
    fx = k.force_path != 3
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(),
                "kernel": ("k_moments_smem_fx<long,1024,0,WS> (floxb_moments: fixed-point level 1, two native shared-memory atomics per row)"
                           if fx else "k_moments_smem_f32<long,SUM,1024,0,WS> (floxb_moments, f32 CAS-loop level 1)"),
                "peak_source": peak_src, "kernel_ms_per_launch": kernel_ms, "rows_per_launch": launch_rows,
                "launches_per_step": -(-len(arr.chunks) // per_launch), "algorithmic_bytes_per_launch": BYTES_PER_ROW * launch_rows,
                "non_kernel_ms_per_step": ms_per_step - kernel_ms * -(-len(arr.chunks) // per_launch),
                "host_enqueue_ms_per_step": host_ms}

    # ---- the other scaling mode, same run (N > 1 only: at N = 1 the two coincide) ----------------------------------
    other = None
    if world > 1 and not args.no_other_scaling:
        del x, by, arr, lab, chunks, step
        torch.cuda.empty_cache()
        o_mode = "weak" if args.scaling == "strong" else "strong"
        o_rows_local = args.rows if o_mode == "weak" else args.rows // world
        x, by, arr, lab = make_inputs(o_rows_local)
        o_step = make_step(arr, lab)
        barrier()  # ranks leave the kernel-only leg at different times (rank 0 also stops the clock sampler)
        for _ in range(max(args.warmup, 3) + 50):
            o_step()
        o_ms, o_res = timed(o_step, args.steps)
        o_par = gate(o_res, x, by, o_rows_local * world)
        other = {"scaling": o_mode, "value": o_rows_local * world / (o_ms * 1e-3), "unit": UNIT, "ms_per_step": o_ms,
                 "rows_per_gpu": o_rows_local, "parity_ok": o_par["ok"]}
        step = o_step

    # ---- end to end: pinned host buffers -> public API -> host result
    e2e, host_chunks = None, None
    if not args.no_e2e:
        hx = [torch.empty(c.shape, dtype=c.dtype, pin_memory=True) for c in arr.chunks]
        hb = [torch.empty(c.shape, dtype=c.dtype, pin_memory=True) for c in lab.chunks]
        for h, d in zip(hx + hb, list(arr.chunks) + list(lab.chunks)):
            h.copy_(d)
        torch.cuda.synchronize()
        e2e_rows_local = sum(int(c.numel()) for c in lab.chunks)
        gpu_sample = None
        if rank == 0 and world == 1 and not args.no_cpu:
            # GPU result on the sample the CPU baseline will reduce (first 8 chunks), for the reference leg of the gate
            n_s = min(8, len(hx))
            with parallel.distributed(sharding="local"):
                gpu_sample, _ = flox_b200.groupby_reduce(ChunkedArray(list(arr.chunks[:n_s])), ChunkedArray(list(lab.chunks[:n_s])),
                                                         func=args.func, expected_groups=eg, method="map-reduce")
            gpu_sample = gpu_sample.cpu().numpy()
            host_chunks = [(hx[i].numpy(), hb[i].numpy()) for i in range(n_s)]
        del x, by, arr, lab
        torch.cuda.empty_cache()
        harr, hlab = ChunkedArray(hx), ChunkedArray(hb)
        barrier()  # pinning 12 GB per rank takes different times on different ranks

        def e2e_step():
            with parallel.distributed(sharding="local"):
                r, _ = flox_b200.groupby_reduce(harr, hlab, func=args.func, expected_groups=eg, method="map-reduce")
            return r.cpu() if isinstance(r, torch.Tensor) else r

        from flox_b200.arrays import TransferStats

        e2e_step()
        if args.profile_host and rank == 0:
            import cProfile
            import pstats

            prof = cProfile.Profile()
            prof.enable()
            e2e_step()
            prof.disable()
            print("---- e2e step host profile ----", file=sys.stderr)
            pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(18)
        barrier()
        h2d0 = TransferStats.h2d_bytes
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            out = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        h2d = (TransferStats.h2d_bytes - h2d0) // args.e2e_steps  # counted from the tensors the package copies
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * e2e_rows_local / dt, "unit": UNIT, "h2d_bytes_per_step": world * h2d,
               "d2h_bytes_per_step": world * int(out.numel() * out.element_size()), "ms_per_step": dt * 1e3,
               "steps": args.e2e_steps, "host_bytes_per_step": world * e2e_rows_local * BYTES_PER_ROW,
               "rows_per_gpu": e2e_rows_local, "numa_local_cpus": (len(numa_cpus) if numa_cpus else None),
               "how": "pinned host chunks (f32 values + int64 labels) -> groupby_reduce(ChunkedArray, method='map-reduce') -> host "
                      "result; labels are range-checked and narrowed to int16 by host threads on the way to the pinned staging "
                      "buffers when that pays (see flox_b200/mapreduce.py:_stage_packed_codes)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if all_cpus is not None:
            os.sched_setaffinity(0, all_cpus)  # the reference gets every host core, not only the GPU's NUMA node
        cpu, _, _ = time_cpu(args, args.cpu_seconds, host_chunks)
        if host_chunks is not None:
            # reference leg of the parity gate: the reference's own map-reduce over the sample, same arrays
            want = np.asarray(cpu_mapreduce(host_chunks, G, args.func, min(len(host_chunks), 8)))
            rel = float(np.nanmax(np.abs(gpu_sample.astype(np.float64) - want.astype(np.float64)) / np.maximum(np.abs(want), 1e-300)))
            parity["max_rel_err_vs_reference_sample"] = rel
            parity["reference_sample"] = f"{len(host_chunks)} chunks x {host_chunks[0][0].size} rows, {cpu['kind']}"
            parity["ok"] = bool(parity["ok"] and rel <= 1e-6 and gpu_sample.dtype == want.dtype)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args)},
            "config_detail": {"engine": "b200", "rows_per_gpu": rows_local, "job_rows": job_rows,
                              "accumulate": "exact fixed-point level 1 -> f64 sums + i64 counts, rounded once to f32",
                              "l2": f"inputs ({rows_local * BYTES_PER_ROW / 1e9:.1f} GB/GPU) >> 126 MB L2, no flush needed",
                              "parallelism": f"dp{world}: rows sharded, [G] accumulators merged across ranks inside the step"},
            "roofline": roofline, "parity": parity, "other_scaling": other, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
