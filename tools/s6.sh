cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/s6_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s6_smoke.log
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s6_pytest.log
tail -3 gpurun_out/s6_smoke.log; tail -15 gpurun_out/s6_pytest.log
