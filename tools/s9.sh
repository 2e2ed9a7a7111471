cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py > gpurun_out/s9_dist_$NG.log 2>&1; echo "rc=$?" >> gpurun_out/s9_dist_$NG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/s9_bench_$NG.log 2>&1; echo "rc=$?" >> gpurun_out/s9_bench_$NG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/configs_multi.py ${S9_CONFIGS:-c3 c4 c5} > gpurun_out/s9_configs_$NG.log 2>&1; echo "rc=$?" >> gpurun_out/s9_configs_$NG.log
tail -4 gpurun_out/s9_dist_$NG.log; tail -2 gpurun_out/s9_bench_$NG.log | cut -c1-3800; grep "^{" gpurun_out/s9_configs_$NG.log | cut -c1-400; tail -3 gpurun_out/s9_configs_$NG.log | cut -c1-300
