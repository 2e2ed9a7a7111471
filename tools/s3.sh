cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "fixed_point or two_level or full_size or mapreduce_golden" > gpurun_out/s3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s3_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/s3_bench_fx.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_moments_smem_fx -s 1 -c 1 -f -o gpurun_out/prof_fx python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu.log 2>&1
tail -5 gpurun_out/s3_pytest.log; tail -1 gpurun_out/s3_bench_fx.log | cut -c1-2200
