cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/s10_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s10_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s10_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s10_bench.log 2>&1; echo "rc=$?" >> gpurun_out/s10_bench.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/s10_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/s10_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s10_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bruns -s 2 -c 2 -f -o gpurun_out/prof_c4 python tools/configs_bench.py c4 > gpurun_out/s10_ncu_c4.log 2>&1
tail -2 gpurun_out/s10_smoke.log; tail -4 gpurun_out/s10_pytest.log; tail -2 gpurun_out/s10_bench.log | cut -c1-4000; tail -2 gpurun_out/s10_bench_ref.log | cut -c1-1500
