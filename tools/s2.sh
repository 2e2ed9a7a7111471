cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "fixed_point or two_level or seeded or full_size or ragged or small_trailing" > gpurun_out/s2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/s2_bench_fx.log 2>&1
timeout 600 python tools/configs_bench.py c4 > gpurun_out/s2_c4_new.log 2>&1
FLOXB_FORCE_PATH=4 timeout 600 python tools/configs_bench.py c4 > gpurun_out/s2_c4_old.log 2>&1
tail -5 gpurun_out/s2_pytest.log; tail -1 gpurun_out/s2_bench_fx.log; cat gpurun_out/s2_c4_new.log gpurun_out/s2_c4_old.log
