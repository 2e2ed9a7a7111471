cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "fixed_point or two_level or seeded or full_size or ragged" > gpurun_out/s1_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/s1_bench_fx.log 2>&1
FLOXB_FORCE_PATH=3 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/s1_bench_cas.log 2>&1
KBENCH_VARIANTS=18,20,21 timeout 600 python tools/kbench3.py > gpurun_out/s1_kbench3.log 2>&1
tail -3 gpurun_out/s1_pytest.log; tail -1 gpurun_out/s1_bench_fx.log; tail -1 gpurun_out/s1_bench_cas.log; tail -6 gpurun_out/s1_kbench3.log
